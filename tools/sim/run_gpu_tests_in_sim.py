"""Runs selected tests of tests/test_gpu_parity.py against the SIMULATION build (tools/sim/libwitch_sim.so) on the CPU --
developer tooling to dry-run the GPU parity tests when no GPU is at hand (slow: pick tests with -k). Not a parity claim.
usage: python tools/sim/run_gpu_tests_in_sim.py [lib name = sim] [pytest -k expression]"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from witch_b200 import _lib  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "sim"
kexpr = sys.argv[2] if len(sys.argv) > 2 else "edge_cases or merge_matches or graph_dp_matches"
os.environ["WITCH_SIM_DRYRUN"] = "1"   # tests/conftest.py: do not skip the gpu-marked tests in this process
_lib.LIB_PATH = os.path.join(ROOT, "tools", "sim", "libwitch_%s.so" % name)   # the SIMULATION build, explicitly
sys.exit(pytest.main([os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-x", "-q", "-k", kexpr]))
