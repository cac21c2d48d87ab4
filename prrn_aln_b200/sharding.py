"""Multi-GPU sharding of the all-vs-all distance step (one process per GPU).

The pairs of calcdist are independent (reference: CalcServer IM_EVRY, src/calcserv.h:847-859), so
rank r of N computes a contiguous range of the condensed index k = elem(i, j) and the ranks exchange
their shards with ONE all-gather (NCCL on GPUs, gloo in the CPU tests).  No other collective is on
the data path.
"""
import numpy as np


def shard_range(npair, world, rank):
    """Equal-count contiguous shard [k0, k1) of the condensed index; every rank gets `chunk`
    slots (the last ranks may be short or empty) so that all_gather_into_tensor applies."""
    chunk = (npair + world - 1) // world if world > 0 else npair
    k0 = min(rank * chunk, npair)
    k1 = min((rank + 1) * chunk, npair)
    return k0, k1, chunk


def gather_shards(shard, chunk, npair, world, dist=None):
    """all-gather equal-size (padded) shards and trim the result to npair entries.
    `shard` is a torch tensor of length `chunk` on the rank's device (padding beyond k1-k0 ignored)."""
    import torch
    if world == 1 or dist is None:
        return shard[:npair]
    full = torch.empty(chunk * world, dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(full, shard)
    return full[:npair]


def calcdist_sharded(ctx, seqs, prm, mtx, rank, world, dist=None, compute=None):
    """calcdist over `world` ranks.  `compute(k0, k1) -> numpy/torch vector` defaults to the CUDA
    path of `ctx`; the CPU tests inject a stand-in to exercise the sharding/gather logic only."""
    import torch
    npair = seqs.n * (seqs.n - 1) // 2
    k0, k1, chunk = shard_range(npair, world, rank)
    if compute is None:
        def compute(a, b):
            return ctx.calcdist(seqs, prm, mtx, a, b)
    part = compute(k0, k1)
    part = torch.as_tensor(np.asarray(part))
    shard = torch.zeros(chunk, dtype=part.dtype)
    shard[:k1 - k0] = part
    if ctx is not None and torch.cuda.is_available():
        shard = shard.cuda()
    return gather_shards(shard, chunk, npair, world, dist)


# ---- candidate partitions of the refinement step (Prrn::best_of_n, src/prrn5.cc:594-631) ---------
def shard_candidates(costs, world, rank):
    """The B candidate group pairs of one refinement step are independent: give rank r
    the candidates of a greedy longest-processing-time split of `costs` (DP cells per candidate), so
    every GPU gets about the same number of cells.  Deterministic; every candidate lands on one rank."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    load = [0] * world
    mine = []
    for i in order:
        r = min(range(world), key=lambda q: (load[q], q))
        load[r] += costs[i]
        if r == rank:
            mine.append(i)
    return sorted(mine)


def best_of_n_sharded(scores_of, costs, rank, world, dist=None):
    """Score the candidates of this rank with `scores_of(indices) -> list of float` (pg_align_groups on
    the rank's GPU), combine the B scores over the ranks and return (best index, best score, all scores).
    Ties go to the lowest index, as the reference's sequential arg-max does (prrn5.cc:618-626).  The only
    collective is one all-reduce(MAX) of B doubles (every slot is written by exactly one rank, the others
    hold -inf); the winner's corner list stays on the rank that owns it."""
    import torch
    mine = shard_candidates(costs, world, rank)
    vals = scores_of(mine) if mine else []
    n = len(costs)
    buf = torch.full((n,), float("-inf"), dtype=torch.float64)
    for i, v in zip(mine, vals):
        buf[i] = float(v)
    if world > 1 and dist is not None:
        dist.all_reduce(buf, op=dist.ReduceOp.MAX)      # each slot is written by exactly one rank
    best = int(torch.argmax(buf).item()) if n else -1
    return best, (float(buf[best]) if n else float("-inf")), buf
