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


def _run(args, env=None):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sim", "sim_check.py")] + args, capture_output=True,
                       text=True, timeout=900, env=dict(os.environ, **(env or {})))
    assert r.returncode == 0 and "SIM CHECK OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    return r.stdout


@pytest.fixture(scope="module")
def simlib():
    r = subprocess.run(["bash", os.path.join(ROOT, "tools", "sim", "build_sim.sh"), "sim", "-ffp-contract=off"], capture_output=True, text=True)
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


def test_simulated_packed_parser_classes(simlib):
    """The launch classes of the two-queries-per-CTA parser (parser2_kernel.cuh, witch_abi.cu:s_classes) as written: 13 columns x
    128 threads with the hybrid register / shared-memory parameter sets (M = 1,052; an odd number of queries, so the last one
    is paired with itself), 13 x 256 (M = 2,6xx), and the generation-5 / generation-1 fall-back classes on the same inputs."""
    out = _run([simlib, "dna_sub8", "3", "1"])
    assert "0 differ" in out
    _run([simlib, "dna_full", "2", "0"])
    _run([simlib, "dna_sub8", "3", "0"], env={"WITCH_PARSER": "5"})
    _run([simlib, "dna_sub8", "2", "0"], env={"WITCH_PARSER": "1"})


def test_simulated_multidomain_branch_matches_oracle(simlib):
    """md_kernel.cuh (HMMER's stochastic-trace clustering for regions that fail the single-domain test) exactly as
    written, against oracle/hmm_md.c, which is itself pinned to clusters captured from inside the reference binary:
    every pair of this run goes through the branch (sim_check.py asserts flags, reported sets and 1e-3-bit scores)."""
    out = _run([simlib, "dna_sub8", "6", "0"], env={"SIM_SHORTEST": "1"})
    assert "6 of 6 pairs through the multi-domain branch" in out
    out = _run([simlib, "amino_extreme", "3", "0"], env={"SIM_SHORTEST": "1"})   # low-complexity repeats: many clusters per region
    assert "pairs through the multi-domain branch" in out and " 0 of " not in out


def test_gpu_parity_tests_dry_run_in_simulation(simlib):
    """The GPU-marked parity tests themselves (tests/test_gpu_parity.py), executed against the simulation build: the full
    dna_small and amino_small golden sets (reported sets, 0.01-bit scores, hmmsearch's printed scores, weights, hmmalign's
    columns), the edge cases, the mirror interface, the graph-DP rows and the merged alignment of the reference's Python.
    A dry run of the test logic and of the kernels' arithmetic -- the parity claim itself is the same tests on a B200."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sim", "run_gpu_tests_in_sim.py"), simlib,
                        "scores_weights_columns and (dna_small or amino_small) or edge_cases or mirror_interface or "
                        "merge_matches or graph_dp_matches"], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "6 passed" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
