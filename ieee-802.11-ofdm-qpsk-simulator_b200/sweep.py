"""Host-side logic of a multi-GPU sweep: one process per GPU, frames sharded by global frame index,
one all-reduce of the per-SNR counters (the only exchange the path has -- SURVEY.md section 8(e)).
Works with any torch.distributed backend: nccl on the GPUs, gloo in the CPU tests."""
import ctypes as C

import numpy as np

from .binding import Counters, COUNTERS_BYTES

N_INT_FIELDS = 5          # bit_errors, bits, frames_in_error, rail_errors, frames ; then 3 doubles


def shard_range(n_frames, rank, world):
    """Contiguous shard [lo, hi) of the global frame range for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_frames, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def counters_to_arrays(counters):
    """list of Counters -> (int64 [n,5], float64 [n,3])"""
    ints = np.array([[c.bit_errors, c.bits, c.frames_in_error, c.rail_errors, c.frames] for c in counters], dtype=np.int64)
    dbls = np.array([[c.sum_err2, c.sum_ref2, c.sum_evm_lin] for c in counters], dtype=np.float64)
    return ints.reshape(-1, 5), dbls.reshape(-1, 3)


def arrays_to_counters(ints, dbls):
    out = []
    for i, d in zip(np.asarray(ints), np.asarray(dbls)):
        c = Counters()
        c.bit_errors, c.bits, c.frames_in_error, c.rail_errors, c.frames = (int(x) for x in i)
        c.sum_err2, c.sum_ref2, c.sum_evm_lin = (float(x) for x in d)
        out.append(c)
    return out


def allreduce_counters(counters, device=None, group=None):
    """Sum a list of Counters over all ranks of the process group (integers exactly; doubles in rank order
    of the backend's reduction).  No-op without an initialised group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return counters
    ints, dbls = counters_to_arrays(counters)
    ti = torch.from_numpy(ints)
    td = torch.from_numpy(dbls)
    if device is not None:
        ti, td = ti.to(device), td.to(device)
    dist.all_reduce(ti, group=group)
    dist.all_reduce(td, group=group)
    return arrays_to_counters(ti.cpu().numpy(), td.cpu().numpy())


def allreduce_counter_tensor(t, group=None):
    """In-place all-reduce of a device counters tensor [n_snr, 8] (int64 view of ofdm_counters)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return t
    ints = t[:, :N_INT_FIELDS].contiguous()
    dist.all_reduce(ints, group=group)
    dbls = t[:, N_INT_FIELDS:].contiguous().view(torch.float64)
    dist.all_reduce(dbls, group=group)
    t[:, :N_INT_FIELDS] = ints
    t[:, N_INT_FIELDS:] = dbls.view(torch.int64)
    return t


def mc_sweep_sharded(ofdm, seed, n_frames_total, n_sym, snr_db, mode, rank=0, world=1, group=None):
    """configs[3]: Philox Monte-Carlo sweep over ``n_frames_total`` frames split across ``world`` ranks by
    global frame index (so the totals do not depend on the GPU count), counters all-reduced.  Returns the
    global per-SNR Counters on every rank."""
    lo, hi = shard_range(n_frames_total, rank, world)
    cnt = ofdm.new_counters(len(snr_db))
    ofdm.mc_sweep_philox(seed, lo, hi - lo, n_sym, snr_db, mode, counters=cnt)
    allreduce_counter_tensor(cnt, group)
    return ofdm.read_counters(cnt)


def ber_table(snr_db, counters):
    rows = []
    for s, c in zip(snr_db, counters):
        ber = c.bit_errors / c.bits if c.bits else float("nan")
        evm = np.sqrt(c.sum_err2 / c.sum_ref2) if c.sum_ref2 else float("nan")
        rows.append((float(s), int(c.bit_errors), int(c.bits), ber, 20 * np.log10(evm) if evm > 0 else -np.inf, int(c.frames_in_error)))
    return rows
