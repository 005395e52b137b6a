"""CPU (-m "not gpu"): the library's own CUDA sources (kernels + C ABI), compiled with g++ against the SIMT simulator in
tools/sim (every CUDA thread a fiber; shuffles / barriers as rendez-vous points; TMA copies complete at issue), checked
against the oracle on small golden inputs. This is TEST TOOLING for developing kernels without a GPU: it checks the
arithmetic, indexing, ring / strip-boundary bookkeeping and work distribution of the kernels exactly as written. It is
not a fallback: the product binding (witch_b200/_lib.py) only ever loads the nvcc-built libwitch_b200.so, and the parity
claims rest on tests/test_gpu_parity.py on a B200."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sim", "sim_check.py")] + args, capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0 and "SIM CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.fixture(scope="module")
def simlib():
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "sim", "build_sim.sh"), "sim"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return "sim"


def test_simulated_kernels_match_oracle_single_strip_dna(simlib):
    out = _run([simlib, "dna_small", "3", "2"])      # 3 queries x 3 window profiles (M <= 249: one strip)
    assert "0 differ" in out


def test_simulated_kernels_match_oracle_multi_strip_dna(simlib):
    out = _run([simlib, "dna_sub8", "2", "1"])       # M = 1052: five strips -> TMA boundary ring, exponent blocks
    assert "0 differ" in out


def test_simulated_kernels_match_oracle_amino_lane_exponents(simlib):
    _run([simlib, "amino_small", "3", "2"])          # per-lane scaling exponents (LANE_EXP) and the C = 4 parser class
