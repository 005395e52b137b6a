"""Multi-GPU sharding of the query x HMM grid: queries are partitioned across ranks with the eHMM
replicated; every (query, HMM) pair is independent, so there is no data-path collective. One all-gather of the
fixed-shape per-query record {idx[k], w[k], count} assembles the top-k table on every rank (NCCL over NVLink on
GPUs; the same code runs over gloo on CPU in the tests). Alignment traces stay on the rank that owns the query."""
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


def gather_topk(idx, w, cnt, mine, n_total, device=None):
    """All-gather per-query records from every rank into global tables ordered by global query id.
    idx[len(mine), k] int32, w[len(mine), k] float64, cnt[len(mine)] int32; `mine` = global ids owned by this rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    k = idx.shape[1]
    cap = (n_total + world - 1) // world + 1  # fixed shape per rank
    dev = device or ("cuda" if dist.get_backend() == "nccl" else "cpu")

    def pad(a, fill, dtype):
        t = torch.full((cap,) + tuple(a.shape[1:]), fill, dtype=dtype)
        t[:len(a)] = torch.as_tensor(np.ascontiguousarray(a), dtype=dtype)
        return t.to(dev)

    ids = pad(np.asarray(mine, dtype=np.int64), -1, torch.int64)
    ti, tw, tc = pad(idx, -1, torch.int32), pad(w, 0.0, torch.float64), pad(cnt, 0, torch.int32)
    outs = []
    for t in (ids, ti, tw, tc):
        buf = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(buf, t)
        outs.append(torch.cat(buf).cpu().numpy())
    gids, gi, gw, gc = outs
    valid = gids >= 0
    G_i = np.full((n_total, k), -1, np.int32)
    G_w = np.zeros((n_total, k), np.float64)
    G_c = np.zeros(n_total, np.int32)
    G_i[gids[valid]] = gi[valid]
    G_w[gids[valid]] = gw[valid]
    G_c[gids[valid]] = gc[valid]
    return G_i, G_w, G_c
