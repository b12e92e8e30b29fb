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
