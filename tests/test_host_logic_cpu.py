"""CPU: host-side logic of the mirror interface (witch_b200/gcmm.py, sharding) and the synthetic generator."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

from oracle import oracle as O
from witch_b200 import gcmm, sharding


def test_adaptive_inclusion_counts_matches_reference_loop():
    rng = np.random.default_rng(3)
    n, k = 200, 10
    w = np.sort(rng.dirichlet(np.full(k, 0.2), n), axis=1)[:, ::-1].copy()
    cnt = rng.integers(0, k + 1, n).astype(np.int32)
    for q in range(n):
        w[q, cnt[q]:] = 0
    keep = gcmm.adaptive_inclusion_counts(w, cnt)
    for q in range(n):
        sw = [(j, w[q, j]) for j in range(cnt[q])]
        assert keep[q] == len(O.adaptive_inclusion(sw))


def test_write_weights_to_local_roundtrip(tmp_path):
    t2w = {"A": ((3, 0.75), (1, 0.25)), "B": ((0, 1.0),)}
    p = tmp_path / "weights.txt"
    gcmm.writeWeightsToLocal(t2w, str(p))
    back = {}
    for ln in open(p):  # the reference's readWeightsFromLocal (weighting.py:184-194)
        taxon, tw = ln.split(":")
        back[taxon] = eval(tw)
    assert back == t2w


def test_partition_balances_cells_and_covers_everything():
    rng = np.random.default_rng(0)
    lengths = rng.integers(80, 1600, 5000)
    for world in (1, 2, 3, 8):
        parts = [sharding.partition_queries(lengths, r, world) for r in range(world)]
        allq = np.sort(np.concatenate(parts))
        assert np.array_equal(allq, np.arange(len(lengths)))
        loads = np.array([lengths[p].sum() for p in parts], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.02


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, k = 101, 4
        lengths = np.arange(n) % 7 + 1
        mine = sharding.partition_queries(lengths, rank, world)
        idx = np.full((len(mine), k), -1, np.int32)
        w = np.zeros((len(mine), k))
        cnt = np.zeros(len(mine), np.int32)
        for j, qq in enumerate(mine):  # fake per-query records that encode the global query id
            idx[j, 0] = qq; w[j, 0] = qq / 1000.0; cnt[j] = 1
        gi, gw, gc = [t.numpy() for t in sharding.gather_topk(idx, w, cnt, mine, n, device="cpu")]
        ok = bool((gi[:, 0] == np.arange(n)).all() and np.allclose(gw[:, 0], np.arange(n) / 1000.0) and (gc == 1).all()
                  and (gi[:, 1:] == -1).all())
        # the sharded driver end to end with stand-ins for the device objects: every rank uploads exactly its own
        # queries' residues and ends with the same global table
        import ctypes
        import torch
        rng = np.random.default_rng(3)
        lens = rng.integers(1, 40, 57)
        off = np.zeros(len(lens) + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
        blob = rng.integers(65, 90, int(off[-1])).astype(np.uint8)

        class FakeQ:
            def __init__(self, ehmm, ptr, offsets):
                src = (ctypes.c_char * int(offsets[-1])).from_address(int(ptr))
                self.res = np.frombuffer(src, dtype=np.uint8).copy(); self.off = np.asarray(offsets); self.n = len(offsets) - 1

        class FakePipe:
            ehmm = None; device = torch.device("cpu"); k = 3

            def run(self, qq):   # record = (sum of residues, length) of each uploaded query
                sums = np.array([qq.res[qq.off[j]:qq.off[j + 1]].sum() for j in range(qq.n)], dtype=np.int32)
                idx = np.full((qq.n, 3), -1, np.int32); idx[:, 0] = sums
                w = np.zeros((qq.n, 3)); w[:, 0] = np.diff(qq.off)
                return dict(idx=torch.as_tensor(idx), w=torch.as_tensor(w), count=torch.ones(qq.n, dtype=torch.int32))

        out = sharding.run_sharded(FakePipe(), blob.ctypes.data, off, rank, world, queries_factory=FakeQ)
        want = np.array([blob[off[j]:off[j + 1]].sum() for j in range(len(lens))])
        ok = ok and bool((out["idx"][:, 0].numpy() == want).all() and (out["w"][:, 0].numpy() == lens).all()
                         and len(out["mine"]) in (28, 29) and out["queries"].n == len(out["mine"]))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gather_topk_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
    assert sorted(res) == [(0, True), (1, True)]


def test_synthetic_workload_is_seeded_and_consistent(tmp_path):
    import synth
    a = synth.make_workload(str(tmp_path), **synth.CONFIGS["tiny"])
    b = synth.make_workload(str(tmp_path), **synth.CONFIGS["tiny"])
    assert a["seqs"] == b["seqs"] and a["hmm_paths"] == b["hmm_paths"]
    prof = O.Profile(a["hmm_paths"][0])
    assert prof.M == len(a["retained_columns"][0]) == a["backbone_length"] or prof.M <= a["backbone_length"]
    assert prof.nseq == a["nseq"][0]
    r = O.score_pair(prof, prof.abc.digitize(a["seqs"][0]))
    assert r["reported"] and r["score"] > 10


def test_bench_slabs_weak_and_strong():
    """bench.py's batches: length-stratified slabs that partition the query set; weak scaling = `world` copies of a slab
    (copies > 0 mutated, lengths unchanged), strong scaling = the slab itself."""
    sys.path.insert(0, ROOT)
    import bench
    rng = np.random.default_rng(3)
    seqs = ["".join(rng.choice(list("ACGT"), size=int(n))) for n in rng.integers(50, 400, 101)]
    wl = {"seqs": seqs, "meta": {"alphabet": "dna"}}
    weak = bench.make_slabs(wl, 4, 3)
    strong = bench.make_slabs(wl, 4, 3, strong=True)
    ids = np.sort(np.concatenate([s[2] for s in weak]))
    assert np.array_equal(ids, np.arange(len(seqs)))                      # the slabs partition the set
    tot = [int(np.diff(s[1]).sum()) for s in strong]
    assert max(tot) - min(tot) < 0.15 * max(tot)                          # ... into batches of similar residue counts
    for (rw, ow, iw), (rs, os_, is_) in zip(weak, strong):
        assert np.array_equal(iw, is_) and len(ow) == 3 * len(is_) + 1 and len(os_) == len(is_) + 1
        assert len(rw) == 3 * len(rs) and np.array_equal(rw[:len(rs)], rs)     # copy 0 is the slab itself
        assert np.array_equal(np.diff(ow)[:len(is_)], np.diff(os_))
        m = rw[len(rs):2 * len(rs)] != rs
        assert 0 < m.mean() < 0.05                                            # copy 1: ~2 % point mutations
        assert "".join(seqs[i] for i in is_).encode() == rs.tobytes()


def test_bench_reference_arm_prints_one_json_line(tmp_path):
    """`bench.py --impl reference` (the reference's HMMER binaries from oracle/_ref on the host cores): one JSON line on
    stdout with the contract's keys; skipped when the binaries are not staged."""
    import json
    import subprocess
    from oracle.make_ref import have_ref
    if not have_ref():
        pytest.skip("oracle/_ref (the reference's HMMER binaries) is not staged")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "tiny", "--steps", "1",
                        "--warmup", "0", "--cpu-seconds", "1", "--workdir", str(tmp_path)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GCUPS" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["config"]["workload"] == "tiny"
