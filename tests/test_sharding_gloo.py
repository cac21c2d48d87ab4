"""world_size-2 (and 3) gloo test of the N>1 host logic: shard ranges of the condensed index and
the single all-gather that assembles the distance vector.  The compute leg is a stand-in here (the
product has no CPU path); the GPU tests check that real shards concatenate to the full vector."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from prrn_aln_b200 import sharding
import prrn_aln_b200 as P


def test_shard_ranges_cover_everything():
    for npair in (0, 1, 5, 499500, 49995000):
        for world in (1, 2, 3, 4, 8):
            cov = []
            for r in range(world):
                k0, k1, chunk = sharding.shard_range(npair, world, r)
                assert 0 <= k0 <= k1 <= npair and k1 - k0 <= chunk
                cov.append((k0, k1))
            assert cov[0][0] == 0 and cov[-1][1] == npair
            assert all(cov[i][1] == cov[i + 1][0] for i in range(world - 1))


def test_cost_model_equals_the_library_cell_count_and_ranges_balance():
    """sharding.band_cells / row_costs (numpy) against pg_calcdist_cells (C, host-only) on ragged lengths and several
    band shoulders; cost-balanced shard ranges cover the condensed index and carry equal numbers of DP cells."""
    rng = np.random.default_rng(3)
    lens = np.concatenate([rng.integers(1, 40, size=30), rng.integers(300, 420, size=60), [0, 1, 2]])
    rng.shuffle(lens)
    enc = [rng.integers(3, 23, size=int(n)).astype(np.uint8) for n in lens]
    ss = P.SeqSet(enc)
    npair = len(lens) * (len(lens) - 1) // 2
    for sh in (-60, -20, 0, 3, 1000):
        prm = P.Params(P.ALPRM(sh=sh))
        cost, _, _ = sharding.row_costs(lens, sh)
        assert int(cost.sum()) == P.calcdist_cells(ss, prm), sh
        for world in (1, 2, 3, 8):
            rg = sharding.cost_balanced_ranges(lens, sh, world)
            assert rg[0][0] == 0 and rg[-1][1] == npair and all(rg[i][1] == rg[i + 1][0] for i in range(world - 1))
            cells = [P.calcdist_cells(ss, prm, a, b) for a, b in rg]
            assert sum(cells) == int(cost.sum())
            assert max(cells) - min(cells) <= 2 * int(sharding.band_cells(lens[:, None], lens[None, :], sh).max()) + 1, (sh, world, cells)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nseq, q, sh=None):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)
    enc = [rng.integers(3, 23, size=int(rng.integers(5, 30))).astype(np.uint8) for _ in range(nseq)]
    seqs = P.SeqSet(enc)

    def compute(k0, k1):            # stand-in "distance": a function of k only
        return np.arange(k0, k1, dtype=np.float64) * 0.5 + 1.0

    full = sharding.calcdist_sharded(None, seqs, None, None, rank, world, dist, compute, sh=sh)
    npair = nseq * (nseq - 1) // 2
    ok = bool(torch.equal(full, torch.arange(npair, dtype=torch.float64) * 0.5 + 1.0))
    q.put((rank, ok, int(full.numel())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,nseq,sh", [(2, 11, None), (3, 7, None), (2, 2, None), (2, 23, -60), (3, 17, 3)])
def test_all_gather_assembles_condensed_vector(world, nseq, sh):
    """sh None: equal-count shards; sh given: cost-balanced shards of unequal length, padded for the all-gather."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nseq, q, sh)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert all(n == nseq * (nseq - 1) // 2 for _, _, n in res)


def test_candidate_split_is_a_partition_and_balanced():
    rng = np.random.default_rng(4)
    for world in (1, 2, 3, 8):
        costs = [int(x) for x in rng.integers(1000, 400000, size=37)]
        parts = [sharding.shard_candidates(costs, world, r) for r in range(world)]
        assert sorted(i for p in parts for i in p) == list(range(len(costs)))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(costs)


def _bon_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [50, 400, 120, 400, 90, 300, 10]
    truth = [3.5, 9.25, -1.0, 9.25, 4.0, 8.0, 0.5]          # two tied maxima: the lower index must win

    def scores_of(idx):             # stand-in for pg_align_groups on this rank's GPU
        return [truth[i] for i in idx]

    best, val, allv = sharding.best_of_n_sharded(scores_of, costs, rank, world, dist)
    q.put((rank, best, val, allv.tolist() == truth))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_best_of_n_over_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bon_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(b == 1 and v == 9.25 and ok for _, b, v, ok in res), res


def test_pair_costs_table_equals_closed_form():
    rng = np.random.default_rng(6)
    for lens, npairs in ((rng.integers(270, 330, size=2000), 20000), (rng.integers(0, 5000, size=400), 300)):
        a, b = rng.integers(0, len(lens), size=npairs), rng.integers(0, len(lens), size=npairs)
        for sh in (-60, 0, 25):
            assert np.array_equal(sharding.pair_costs(lens, a, b, sh), sharding.band_cells(lens[a], lens[b], sh))
    assert len(sharding.pair_costs([5, 6], [], [], -60)) == 0


def test_query_split_keeps_a_query_on_one_rank():
    rng = np.random.default_rng(8)
    q = np.repeat(np.arange(40), rng.integers(1, 30, size=40))
    rng.shuffle(q)
    costs = rng.integers(1000, 90000, size=len(q))
    for world in (1, 2, 3, 8):
        parts = [sharding.shard_queries(q, costs, world, r) for r in range(world)]
        assert sorted(int(i) for p in parts for i in p) == list(range(len(q)))
        owner = {}
        for r, p in enumerate(parts):
            for i in p:
                assert owner.setdefault(int(q[i]), r) == r
        loads = [int(costs[p].sum()) for p in parts]
        per_q = np.bincount(q, weights=costs)
        assert max(loads) - min(loads) <= per_q.max()


def _edge_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(2)
    lens = rng.integers(50, 400, size=60)
    qi = np.repeat(np.arange(0, 60, 3), 7)
    si = rng.integers(0, 60, size=len(qi))
    truth = (qi * 1000 + si).astype(np.float64) / 8.0

    def dist_of(idx):               # stand-in for pg_dist_pairs on this rank's GPU
        return truth[idx]

    full, mine = sharding.dist_edges_sharded(dist_of, qi, si, lens, -60, rank, world, dist)
    q.put((rank, bool(np.array_equal(full.numpy(), truth)), len(mine)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_candidate_edges_over_ranks(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_edge_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res
    assert sum(n for _, _, n in res) == 140
