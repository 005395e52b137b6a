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
