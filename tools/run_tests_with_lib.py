"""Runs GPU parity tests (tests/test_gpu_parity.py, -m gpu) against a VARIANT build of the CUDA library, e.g. the
two-items-per-warp envelope kernel: python tools/run_tests_with_lib.py tools/bin/libwitch_pair.so ["-k expression"].
Developer tooling for A/B work on a GPU box; the product binding keeps loading witch_b200/csrc/libwitch_b200.so."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from witch_b200 import _lib  # noqa: E402

_lib.LIB_PATH = os.path.join(ROOT, sys.argv[1])
args = [os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q"]
if len(sys.argv) > 2:
    args += ["-k", sys.argv[2]]
sys.exit(pytest.main(args))
