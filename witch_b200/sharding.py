"""Multi-GPU sharding of the query x HMM grid: queries are partitioned across ranks with the eHMM
replicated; every (query, HMM) pair is independent, so there is no data-path collective. ONE all-gather of the
fixed-shape packed per-query record {w[k] f64, id i64, idx[k] i32, count i32} assembles the top-k table on every rank
(NCCL over NVLink on GPUs; the same code runs over gloo on CPU in the tests). Alignment traces stay on the rank
that owns the query. Replaces the fork pool of the reference (witch_msa/gcmm/gcmm.py:110-111, 196-199)."""
import numpy as np


def partition_queries(lengths, rank, world):
    """Length-balanced deal: sort by length (descending) and deal round-robin in a serpentine order, so every rank gets
    the same number of queries (+-1) and nearly the same number of residues (DP cells are proportional to L)."""
    lengths = np.asarray(lengths)
    order = np.argsort(-lengths, kind="stable")
    pos = np.arange(len(order))
    rnd, slot = pos // world, pos % world
    owner = np.where(rnd % 2 == 0, slot, world - 1 - slot)
    return np.sort(order[owner == rank])


def record_bytes(k):
    """Bytes of one packed record: w[k] f64 | global id i64 | idx[k] i32 | count i32, padded to 8."""
    return (8 * k + 8 + 4 * k + 4 + 7) // 8 * 8


def pack_records(idx, w, cnt, ids, cap):
    """idx[m,k] i32, w[m,k] f64, cnt[m] i32, ids[m] i64 (torch tensors on one device) -> uint8 [cap, record_bytes(k)];
    rows >= m carry id -1."""
    import torch
    m, k = idx.shape
    rb = record_bytes(k)
    rec = torch.zeros((cap, rb), dtype=torch.uint8, device=idx.device)
    rec[:, 8 * k:8 * k + 8] = torch.full((cap, 1), -1, dtype=torch.int64, device=idx.device).view(torch.uint8)
    if m:
        rec[:m, :8 * k] = w.contiguous().view(torch.uint8).reshape(m, 8 * k)
        rec[:m, 8 * k:8 * k + 8] = ids.contiguous().view(torch.uint8).reshape(m, 8)
        rec[:m, 8 * k + 8:12 * k + 8] = idx.contiguous().view(torch.uint8).reshape(m, 4 * k)
        rec[:m, 12 * k + 8:12 * k + 12] = cnt.contiguous().view(torch.uint8).reshape(m, 4)
    return rec


def unpack_records(rec, k, n_total):
    """uint8 [R, record_bytes(k)] -> global tables (idx[n_total,k] i32 (-1), w[n_total,k] f64 (0), count[n_total] i32),
    ordered by global query id; rows with id < 0 are padding."""
    import torch
    R = rec.shape[0]
    w = rec[:, :8 * k].contiguous().view(torch.float64).reshape(R, k)
    ids = rec[:, 8 * k:8 * k + 8].contiguous().view(torch.int64).reshape(R)
    idx = rec[:, 8 * k + 8:12 * k + 8].contiguous().view(torch.int32).reshape(R, k)
    cnt = rec[:, 12 * k + 8:12 * k + 12].contiguous().view(torch.int32).reshape(R)
    valid = ids >= 0
    G_i = torch.full((n_total, k), -1, dtype=torch.int32, device=rec.device)
    G_w = torch.zeros((n_total, k), dtype=torch.float64, device=rec.device)
    G_c = torch.zeros((n_total,), dtype=torch.int32, device=rec.device)
    sel = ids[valid]
    G_i[sel] = idx[valid]
    G_w[sel] = w[valid]
    G_c[sel] = cnt[valid]
    return G_i, G_w, G_c


def gather_topk(idx, w, cnt, mine, n_total, device=None):
    """All-gather the per-query records of every rank into global tables ordered by global query id: one
    all_gather_into_tensor of the packed record. idx[len(mine), k] int32, w[len(mine), k] float64, cnt[len(mine)] int32
    (torch tensors or numpy arrays); `mine` = global ids owned by this rank. -> torch tensors on `device`."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    dev = device or (torch.device("cuda", torch.cuda.current_device()) if dist.is_initialized() and dist.get_backend() == "nccl"
                     else torch.device("cpu"))
    idx = torch.as_tensor(idx).to(dev)
    w = torch.as_tensor(w).to(dev)
    cnt = torch.as_tensor(cnt).to(dev)
    ids = torch.as_tensor(np.ascontiguousarray(mine, dtype=np.int64)).to(dev)
    k = idx.shape[1]
    cap = (n_total + world - 1) // world + 1  # fixed shape per rank
    rec = pack_records(idx, w, cnt, ids, cap)
    if world > 1:
        out = torch.empty((world * cap, rec.shape[1]), dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(out, rec)
    else:
        out = rec
    return unpack_records(out, k, n_total)


def run_sharded(pipe, residues_ptr, offsets, rank=0, world=1, owned=None, queries_factory=None, events=None):
    """The sharded hot path for one batch of queries given as a caller-owned HOST buffer (concatenated ASCII residues
    at integer address `residues_ptr`, `offsets` int64 [n_total+1]): partition_queries -> upload of the owned queries ->
    DevicePipeline.run (score, weights/top-k, adaptive inclusion, align) -> gather_topk. Every rank returns the global
    top-k tables; column lists stay with the owner. `owned` (optional) = precomputed partition of this rank;
    `queries_factory(ehmm, ptr, offsets)` defaults to api.Queries.from_buffer (the host-logic tests inject a stand-in);
    `events` = (start, end) torch.cuda.Event pair recorded on the current stream around the device work (after the
    upload, after the gather) -- bench.py's device-timed window.
    -> dict(mine, idx, w, count (global, device tensors), local=<DevicePipeline.run result>, queries)."""
    import ctypes
    if queries_factory is None:
        from . import api
        queries_factory = api.Queries.from_buffer
    offsets = np.ascontiguousarray(offsets, dtype=np.int64)
    lengths = np.diff(offsets)
    n_total = len(lengths)
    mine = partition_queries(lengths, rank, world) if owned is None else np.asarray(owned)
    if world == 1 and owned is None:
        q = queries_factory(pipe.ehmm, residues_ptr, offsets)
    else:  # compact the owned queries (host gather of their residues), then one upload
        src = (ctypes.c_char * int(offsets[-1])).from_address(int(residues_ptr))
        buf = np.frombuffer(src, dtype=np.uint8)
        loc_off = np.zeros(len(mine) + 1, dtype=np.int64)
        np.cumsum(lengths[mine], out=loc_off[1:])
        loc = np.empty(int(loc_off[-1]), dtype=np.uint8)
        for j, g in enumerate(mine):
            loc[loc_off[j]:loc_off[j + 1]] = buf[offsets[g]:offsets[g + 1]]
        q = queries_factory(pipe.ehmm, loc.ctypes.data, loc_off)
        q._keepalive = loc
    if events:
        events[0].record()
    res = pipe.run(q)
    G_i, G_w, G_c = gather_topk(res["idx"], res["w"], res["count"], mine, n_total, device=pipe.device)
    if events:
        events[1].record()
    return dict(mine=mine, idx=G_i, w=G_w, count=G_c, local=res, queries=q)
