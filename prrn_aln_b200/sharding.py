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


# ---- cost model: DP cells the reference visits per pair -----------------------------
def band(lq, ls, sh):
    """stripe() (reference src/aln2.cc:156-174) for two full windows of lq rows and ls columns, numpy-vectorised:
    returns (lw, up) in diagonal coordinates r = n - m."""
    lq = np.asarray(lq, dtype=np.int64)
    ls = np.asarray(ls, dtype=np.int64)
    shv = np.full(np.broadcast(lq, ls).shape, sh, dtype=np.int64)
    if sh < 0:
        shv = -sh * np.minimum(lq, ls) // 100
    hi = np.maximum(ls - lq, 0) + shv
    lo = np.minimum(ls - lq, 0) - shv
    return np.maximum(lo, -lq), np.minimum(hi, ls)


def _sum_clamped(a0, n, hi):
    """sum_{m=0}^{n-1} clamp(a0 + m, 0, hi), vectorised."""
    m0 = np.clip(-a0, 0, n)
    m1 = np.clip(hi - a0, m0, n)
    cnt = m1 - m0
    lin = cnt * a0 + (m0 + m1 - 1) * cnt // 2
    return lin + (n - m1) * hi


def band_cells(lq, ls, sh):
    """Cells (m, n) the reference's loops visit for a pair of lq x ls residues: rows 0..lq-1, columns
    max(m + lw, 0) .. min(m + up + 1, ls) - 1 (loop bounds src/fwd2d1.cc:136-146, src/fwd2c.h:364-374)."""
    lq = np.asarray(lq, dtype=np.int64)
    ls = np.asarray(ls, dtype=np.int64)
    lw, up = band(lq, ls, sh)
    c = _sum_clamped(up + 1, lq, ls) - _sum_clamped(lw, lq, ls)
    return np.where((lq > 0) & (ls > 0), c, 0)


def pair_costs(lens, a_idx, b_idx, sh):
    """band_cells for an explicit pair list: lengths have few distinct values, so the cells come from a table over
    (length of a, length of b) -- 320,000 pairs in a few ms instead of the closed form evaluated per pair."""
    lens = np.asarray(lens, dtype=np.int64)
    a_idx = np.asarray(a_idx, dtype=np.int64)
    b_idx = np.asarray(b_idx, dtype=np.int64)
    if len(a_idx) == 0:
        return np.zeros(0, np.int64)
    uniq, inv = np.unique(lens, return_inverse=True)
    if len(uniq) * len(uniq) > 4 * len(a_idx) + 4096:         # very ragged set: the table would be the larger job
        return band_cells(lens[a_idx], lens[b_idx], sh)
    table = band_cells(uniq[:, None], uniq[None, :], sh)
    return table[inv[a_idx], inv[b_idx]]


def row_costs(lens, sh):
    """cost[j] = cells of the pairs (i, j), i < j -- row j of the condensed triangle (a = i, b = j).  Lengths
    have few distinct values, so the cost is a table over (length of i, length of j) times running counts."""
    lens = np.asarray(lens, dtype=np.int64)
    n = len(lens)
    if n == 0:
        return np.zeros(0, np.int64), None, None
    uniq, inv = np.unique(lens, return_inverse=True)
    table = band_cells(uniq[:, None], uniq[None, :], sh)            # [length of a][length of b]
    counts = np.zeros((n + 1, len(uniq)), dtype=np.int64)
    np.add.at(counts, (np.arange(1, n + 1), inv), 1)
    counts = np.cumsum(counts, axis=0)                              # counts[j] = histogram of lens[:j]
    cost = np.einsum("ju,uj->j", counts[:n], table[:, inv])
    return cost, table, inv


def cost_balanced_ranges(lens, sh, world):
    """Contiguous ranges [k0, k1) of the condensed index k = j(j-1)/2 + i, one per rank, with about the same number
    of DP cells each (cost-balanced, not count-balanced: ragged sets put the long sequences'
    rows where they fall).  Returns [(k0, k1)] * world covering [0, n(n-1)/2)."""
    n = len(lens)
    npair = n * (n - 1) // 2
    if world <= 1 or npair == 0:
        return [(0, npair)] + [(npair, npair)] * (max(world, 1) - 1)
    cost, table, inv = row_costs(lens, sh)
    cum = np.concatenate([[0], np.cumsum(cost)])                    # cum[j] = cells of rows < j
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r // world
        j = int(np.searchsorted(cum, target, side="right")) - 1     # row holding the target
        j = min(max(j, 1), n - 1)
        within = np.cumsum(table[inv[:j], inv[j]])                  # cells of (0..i, j)
        i = int(np.searchsorted(within, target - int(cum[j]), side="left"))
        cuts.append(max(cuts[-1], min(j * (j - 1) // 2 + min(i, j), npair)))
    cuts.append(npair)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def gather_shards(shard, chunk, npair, world, dist=None):
    """all-gather equal-size (padded) shards and trim the result to npair entries.
    `shard` is a torch tensor of length `chunk` on the rank's device (padding beyond k1-k0 ignored)."""
    import torch
    if world == 1 or dist is None:
        return shard[:npair]
    full = torch.empty(chunk * world, dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(full, shard)
    return full[:npair]


def calcdist_sharded(ctx, seqs, prm, mtx, rank, world, dist=None, compute=None, sh=None):
    """calcdist over `world` ranks.  Rank r computes a contiguous range of the condensed index -- with `sh` given
    (the band shoulder, alprm.sh) the ranges carry equal numbers of DP cells (cost_balanced_ranges), else equal
    numbers of pairs -- and ONE all-gather of equal-size (padded) shards assembles the vector on every rank.
    `compute(k0, k1) -> numpy/torch vector` defaults to the CUDA path of `ctx`; the CPU tests inject a stand-in to
    exercise the sharding / gather logic only."""
    import torch
    npair = seqs.n * (seqs.n - 1) // 2
    if sh is None:
        ranges = [shard_range(npair, world, r)[:2] for r in range(world)]
    else:
        ranges = cost_balanced_ranges(np.asarray(seqs.lens), sh, world)
    chunk = max(max(b - a for a, b in ranges), 1)
    k0, k1 = ranges[rank]
    if compute is None:
        def compute(a, b):
            return ctx.calcdist(seqs, prm, mtx, a, b)
    part = compute(k0, k1)
    part = torch.as_tensor(np.asarray(part))
    shard = torch.zeros(chunk, dtype=part.dtype)
    shard[:k1 - k0] = part
    if ctx is not None and torch.cuda.is_available():
        shard = shard.cuda()
    if world == 1 or dist is None:
        return shard[:npair]
    full = torch.empty(chunk * world, dtype=shard.dtype, device=shard.device)
    dist.all_gather_into_tensor(full, shard)
    return torch.cat([full[r * chunk: r * chunk + (b - a)] for r, (a, b) in enumerate(ranges)])


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
    import math
    import torch
    mine = shard_candidates(costs, world, rank)
    vals = scores_of(mine) if mine else []
    n = len(costs)
    # NCCL reduces device tensors only: the buffer lives on this rank's GPU there, on the host under gloo
    on_gpu = world > 1 and dist is not None and dist.get_backend() == "nccl"
    buf = torch.full((n,), float("-inf"), dtype=torch.float64)
    for i, v in zip(mine, vals):
        v = float(v)
        buf[i] = v if not math.isnan(v) else float("-inf")     # the reference's strict `>` never picks a NaN (prrn5.cc:620)
    if on_gpu:
        buf = buf.cuda()
    if world > 1 and dist is not None:
        dist.all_reduce(buf, op=dist.ReduceOp.MAX)      # each slot is written by exactly one rank
    buf = buf.cpu()
    best = int(torch.argmax(buf).item()) if n else -1
    return best, (float(buf[best]) if n else float("-inf")), buf


# ---- candidate edges of the sparse distance graph (AdjacentMat::spaln_job, src/adjmat.cc:119-156) ----
def shard_queries(q_idx, costs, world, rank):
    """The candidate pairs (query, database sequence) of the SL-forest search are independent per query (SURVEY.md
    8(e), row 2): all edges of one query go to one rank (the query profile is built once), queries are dealt to
    the ranks by their DP cells.  Returns the indices of this rank's edges, ascending."""
    q_idx = np.asarray(q_idx, dtype=np.int64)
    costs = np.asarray(costs, dtype=np.int64)
    if len(q_idx) == 0:
        return np.zeros(0, np.int64)
    nq = int(q_idx.max()) + 1
    per_q = np.bincount(q_idx, weights=costs.astype(np.float64), minlength=nq)
    # thousands of queries of similar cost: deal them in descending order of cost, back and forth over the ranks
    # (0 1 .. w-1 w-1 .. 1 0 ...); loads then differ by less than one query.  Vectorised; deterministic (stable sort).
    order = np.argsort(-per_q, kind="stable")
    pos = np.arange(nq) % (2 * world)
    owner = np.empty(nq, dtype=np.int64)
    owner[order] = np.where(pos < world, pos, 2 * world - 1 - pos)
    return np.nonzero(owner[q_idx] == rank)[0]


def dist_edges_sharded(dist_of, q_idx, s_idx, lens, sh, rank, world, dist=None):
    """Distances of the candidate edges over `world` ranks: rank r evaluates `dist_of(edge indices) -> values`
    (pg_dist_pairs on its GPU) for its queries' edges; ONE all-reduce(SUM) of a vector that every rank fills only
    at its own edges (the others hold 0, so the sum is exact) puts all distances on every rank, where the caller
    applies the threshold and fills the graph (fillmat stays host code)."""
    import torch
    q_idx = np.asarray(q_idx, dtype=np.int64)
    s_idx = np.asarray(s_idx, dtype=np.int64)
    lens = np.asarray(lens, dtype=np.int64)
    costs = pair_costs(lens, q_idx, s_idx, sh)
    mine = shard_queries(q_idx, costs, world, rank)
    vals = np.asarray(dist_of(mine), dtype=np.float64) if len(mine) else np.zeros(0)
    buf = torch.zeros(len(q_idx), dtype=torch.float64)
    if len(mine):
        buf[torch.as_tensor(mine)] = torch.as_tensor(vals)
    if world > 1 and dist is not None:
        if dist.get_backend() == "nccl":
            buf = buf.cuda()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM)
    return buf.cpu(), mine
