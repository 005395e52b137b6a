"""CPU: the C-ABI library loads, exports every symbol include/witch_b200.h declares, and fails loudly (no fallback)
when no CUDA device is present. No compute calls here."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "witch_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(witch_[a-z0-9_]+)\s*\(", txt)))


@pytest.fixture(scope="module")
def lib():
    from witch_b200 import build
    build.build()
    from witch_b200 import _lib
    return _lib.load()


def test_every_declared_symbol_is_exported_and_bound(lib):
    from witch_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "library does not export " + n
        assert n in _lib.SYMBOLS, "python binding misses " + n
    assert sorted(_lib.SYMBOLS) == names


def test_version_and_device_count(lib):
    assert b"sm_100a" in lib.witch_version()
    assert lib.witch_device_count() >= 0


def test_io_errors_are_reported(lib, tmp_path):
    import witch_b200 as wb
    with pytest.raises(wb.WitchError, match="cannot open"):
        wb.EHMM([str(tmp_path / "missing.hmm")])
    bad = tmp_path / "bad.hmm"
    bad.write_text("not an hmm\n")
    with pytest.raises(wb.WitchError, match="HMMER3"):
        wb.EHMM([str(bad)])


def test_no_cpu_fallback_without_gpu(lib, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import witch_b200 as wb
    from golden_util import load_set
    _, _, paths = load_set("amino_small", str(tmp_path))
    with pytest.raises(wb.WitchError, match="no CUDA device"):
        wb.EHMM(paths)


def test_profile_cache_is_written_validated_and_invalidated(lib, tmp_path):
    """witch_ehmm_create_cached: the serialised profiles are written next to the HMM text on the first call and
    accepted on the next one only while every source file is unchanged. (Without a GPU the call itself still ends in
    'no CUDA device' -- the cache is host-side work that happens before the upload.)"""
    import ctypes
    import torch
    from golden_util import load_set
    _, _, paths = load_set("amino_small", str(tmp_path))
    cache = str(tmp_path / "witch_b200.profiles")
    arr = (ctypes.c_char_p * len(paths))(*[p.encode() for p in paths])

    def call():
        h, hit = ctypes.c_void_p(), ctypes.c_int(-1)
        rc = lib.witch_ehmm_create_cached(len(paths), arr, cache.encode(), ctypes.byref(hit), ctypes.byref(h))
        if rc == 0:
            lib.witch_ehmm_destroy(h)
        else:
            assert not torch.cuda.is_available() and b"no CUDA device" in lib.witch_last_error()
        return hit.value

    assert call() == 0 and os.path.getsize(cache) > 1000
    assert call() == 1
    st = os.stat(paths[0])
    os.utime(paths[0], ns=(st.st_atime_ns, st.st_mtime_ns + 1_000_000_000))   # a source file changed: parse again, rewrite
    assert call() == 0
    assert call() == 1
    with open(cache, "r+b") as f:   # a truncated / foreign file is ignored, never trusted
        f.truncate(200)
    assert call() == 0
    assert call() == 1
    assert lib.witch_ehmm_create_cached(len(paths), arr, None, None, ctypes.byref(ctypes.c_void_p())) == -1


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "witch_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inl")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.replace("Oracle", ""), f


def test_product_never_loads_the_simulation_build():
    """tools/sim (host SIMT simulation of the kernels) is developer/test tooling: the package must not know its library,
    and the only path the binding loads by default is the nvcc-built libwitch_b200.so."""
    pkg = os.path.join(ROOT, "witch_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "libwitch_sim" not in src and "tools/sim" not in src and "WITCH_HOST_SIM" not in src, f
    from witch_b200 import _lib
    assert _lib.LIB_PATH == os.path.join(pkg, "csrc", "libwitch_b200.so")
