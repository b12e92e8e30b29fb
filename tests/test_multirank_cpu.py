"""world_size-2 gloo test of the multi-GPU host logic: frames sharded by global frame index,
one all-reduce of the counters; the totals equal the single-process result.  The per-rank compute is stood
in for by the CPU oracle fed with the Philox streams, which is what makes the result rank-count independent."""
import os
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT

N_FRAMES, N_SYM, SNRS, SEED = 600, 2, [3.0, 9.0], 99


def _rank_main(rank, world, port_file, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import __graft_entry__ as entry
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port_file
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_pkg()
    port = entry.load_oracle().Port()
    lo, hi = pkg.sweep.shard_range(N_FRAMES, rank, world)
    bits = port.philox_bits(SEED, lo, hi - lo, N_SYM)
    local = []
    for i, s in enumerate(SNRS):
        g = port.philox_normals(SEED, i, lo, hi - lo, 320)
        a = port.chain(bits, g, N_SYM, s)
        c = pkg.Counters()
        for k in ("bit_errors", "bits", "frames_in_error", "rail_errors", "frames", "sum_err2", "sum_ref2", "sum_evm_lin"):
            setattr(c, k, getattr(a, k))
        local.append(c)
    tot = pkg.sweep.allreduce_counters(local)
    if rank == 0:
        q.put([c.as_dict() for c in tot])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_sweep_matches_single_process(port, pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = str(29500 + os.getpid() % 1000)
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    bits = port.philox_bits(SEED, 0, N_FRAMES, N_SYM)
    for i, s in enumerate(SNRS):
        a = port.chain(bits, port.philox_normals(SEED, i, 0, N_FRAMES, 320), N_SYM, s)
        for k in ("bit_errors", "bits", "frames_in_error", "rail_errors", "frames"):
            assert got[i][k] == getattr(a, k), (s, k)
        assert abs(got[i]["sum_err2"] - a.sum_err2) <= 1e-9 * a.sum_err2


# ---- configs[3]'s stop rule ("until >= 100 bit errors or the bit budget") sharded over ranks by rounds -------------------
U_SNRS, U_TARGET, U_ROUND, U_BUDGET = [2.0, 7.0, 11.0], 60, 96, 5 * 96 * 192


def _oracle_round(port, points, lo, n):
    """this rank's share of a round, computed by the CPU oracle on the Philox streams (stream = the point's index in U_SNRS)"""
    ints = np.zeros((len(points), 5), np.int64)
    dbls = np.zeros((len(points), 3), np.float64)
    if n > 0:
        bits = port.philox_bits(SEED, lo, n, N_SYM)
        for j, pt in enumerate(points):
            a = port.chain(bits, port.philox_normals(SEED, pt, lo, n, 320), N_SYM, U_SNRS[pt])
            ints[j] = [a.bit_errors, a.bits, a.frames_in_error, a.rail_errors, a.frames]
            dbls[j] = [a.sum_err2, a.sum_ref2, a.sum_evm_lin]
    return ints, dbls


def _until_rank_main(rank, world, port_file, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = port_file
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pkg = entry.load_pkg()
    port = entry.load_oracle().Port()

    def allreduce(ints, dbls):
        ti, td = torch.from_numpy(ints.copy()), torch.from_numpy(dbls.copy())
        dist.all_reduce(ti); dist.all_reduce(td)
        return ti.numpy(), td.numpy()

    ints, dbls, rounds = pkg.sweep.until_loop(lambda pts, lo, n: _oracle_round(port, pts, lo, n), len(U_SNRS), U_TARGET, U_BUDGET, U_ROUND,
                                              rank, world, allreduce)
    if rank == 0:
        q.put((ints.tolist(), dbls.tolist(), rounds))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_until_rule_two_ranks_equals_one(port, pkg):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port_no = str(30600 + os.getpid() % 1000)
    procs = [ctx.Process(target=_until_rank_main, args=(r, 2, port_no, q)) for r in range(2)]
    for p in procs:
        p.start()
    ints2, dbls2, rounds2 = q.get(timeout=150)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ints1, dbls1, rounds1 = pkg.sweep.until_loop(lambda pts, lo, n: _oracle_round(port, pts, lo, n), len(U_SNRS), U_TARGET, U_BUDGET, U_ROUND)
    assert rounds1 == rounds2 and np.array_equal(np.array(ints2), ints1)
    assert np.allclose(np.array(dbls2), dbls1, rtol=1e-12)
    # the rule itself: every point stopped on errors or on the budget, at a round boundary; the low-SNR point after one round
    for row in ints1:
        assert row[0] >= U_TARGET or row[1] >= U_BUDGET
        assert row[4] % U_ROUND == 0
    assert ints1[0][4] == U_ROUND and ints1[-1][4] == 5 * U_ROUND
    # shard_range covers a round exactly, whatever the world size (ragged split)
    for world in (1, 2, 3, 8):
        parts = [pkg.sweep.shard_range(U_ROUND + 1, r, world) for r in range(world)]
        assert parts[0][0] == 0 and parts[-1][1] == U_ROUND + 1 and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
