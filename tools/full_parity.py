#!/usr/bin/env python
"""configs[1] pinned at its stated size against the compiled reference itself.

    python tools/full_parity.py [--frames 1000000] [--procs N] [--out gpurun_out/full_parity.json]

N frames x 2 data symbols x 21 SNR points (0..20 dB).  Payload bits are i.i.d. from a stated numpy seed; the
noise is the reference's own rand() / Box-Muller stream: every chunk of frames takes the draws that survive at
src/OFDM.c:651 from `srand(seed_chunk)` (oracle/ref_harness.c: ref_capture_gkeep), one per sample.

  CPU  oracle/_ref (the unmodified src/OFDM.c) `ref_chain_sweep` -- Transmitter once per frame, then
       Transmission_Over_Air's arithmetic with the injected draws + the Receiver stages per SNR point -- over all host
       cores (fork pool, one chunk per task).
  GPU  ofdm_sweep_inject_host (C-ABI, host buffers) in EXACT mode on the same bits and draws.

Pass = all five integer totals (bit errors, bits, frames in error, rail errors, frames) equal at all 21 SNR
points, and the EVM sums within 1e-5 relative.  Test infrastructure: this is the only place besides tests/ and
bench.py's CPU legs that loads oracle/.
"""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time
from multiprocessing import shared_memory

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402

N_SYM, LEN = 2, 320
SNRS = [float(s) for s in range(21)]
_W = {}


def _init(bits_name, g_name, n_frames):
    po = entry.load_oracle()
    _W["ref"] = po.Ref()
    _W["shm"] = (shared_memory.SharedMemory(name=bits_name), shared_memory.SharedMemory(name=g_name))
    _W["bits"] = np.ndarray((n_frames, 96 * N_SYM), np.uint8, buffer=_W["shm"][0].buf)
    _W["g"] = np.ndarray((n_frames, LEN), np.float32, buffer=_W["shm"][1].buf)


def _task(args):
    lo, hi, seed = args
    ref = _W["ref"]
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, (hi - lo, 96 * N_SYM), dtype=np.uint8)
    g = ref.capture_gkeep((hi - lo) * LEN, seed=seed & 0x7FFFFFFF).reshape(hi - lo, LEN)
    _W["bits"][lo:hi] = bits
    _W["g"][lo:hi] = g
    acc = ref.chain_sweep(bits, g, N_SYM, SNRS)
    return [[c.bit_errors, c.bits, c.frames_in_error, c.rail_errors, c.frames, c.sum_err2, c.sum_ref2, c.sum_evm_lin] for c in acc]


def run(n_frames, procs, chunk=2048, base_seed=20261018, mode=None, verbose=True):
    po = entry.load_oracle()
    po.build()
    if not po.have_ref():
        raise SystemExit("oracle/_ref/libofdm_ref.so is missing (built from /root/reference by oracle/Makefile)")
    pkg = entry.load_pkg()
    mode = pkg.MODE_EXACT if mode is None else mode
    shm_b = shared_memory.SharedMemory(create=True, size=n_frames * 96 * N_SYM)
    shm_g = shared_memory.SharedMemory(create=True, size=n_frames * LEN * 4)
    try:
        tasks = [(lo, min(lo + chunk, n_frames), base_seed + i) for i, lo in enumerate(range(0, n_frames, chunk))]
        t0 = time.perf_counter()
        with mp.get_context("fork").Pool(procs, initializer=_init, initargs=(shm_b.name, shm_g.name, n_frames)) as pool:
            parts = pool.map(_task, tasks, chunksize=1)
        cpu_s = time.perf_counter() - t0
        ints = np.zeros((len(SNRS), 5), np.int64)
        dbls = np.zeros((len(SNRS), 3), np.float64)
        for part in parts:
            a = np.array(part, dtype=np.float64)
            ints += np.array([[int(v) for v in row[:5]] for row in part], dtype=np.int64)
            dbls += a[:, 5:]
        bits = np.ndarray((n_frames, 96 * N_SYM), np.uint8, buffer=shm_b.buf)
        g = np.ndarray((n_frames, LEN), np.float32, buffer=shm_g.buf)
        o = pkg.Ofdm(0)
        packed = pkg.pack_bits_host(bits)
        t1 = time.perf_counter()
        got = o.sweep_inject_host(packed, g, n_frames, N_SYM, SNRS, mode)
        gpu_s = time.perf_counter() - t1
        replayed = o.replayed_frames()
        o.close()
        del bits, g
    finally:
        shm_b.close(); shm_b.unlink(); shm_g.close(); shm_g.unlink()
    rows, ok = [], True
    for i, (snr, c) in enumerate(zip(SNRS, got)):
        gi = [int(c.bit_errors), int(c.bits), int(c.frames_in_error), int(c.rail_errors), int(c.frames)]
        ri = [int(v) for v in ints[i]]
        evm_gpu = np.sqrt(c.sum_err2 / c.sum_ref2)
        evm_ref = np.sqrt(dbls[i, 0] / dbls[i, 1])
        rel_evm = abs(evm_gpu - evm_ref) / evm_ref
        rel_lin = abs(c.sum_evm_lin - dbls[i, 2]) / dbls[i, 2]
        same = gi == ri
        ok = ok and same and rel_evm <= 1e-5 and rel_lin <= 1e-5
        rows.append({"snr_db": snr, "ints_equal": same, "gpu": gi, "ref": ri, "ber": gi[0] / gi[1],
                     "evm_db_ref": float(20 * np.log10(evm_ref)), "evm_rel_diff": float(rel_evm), "evm_lin_sum_rel_diff": float(rel_lin)})
        if verbose:
            print("%5.1f dB  ints %s  bit_errors %d / %d  frames_in_error %d  rail %d  evm_rel %.2e  evm_lin_rel %.2e" %
                  (snr, "==" if same else "!=", gi[0], ri[0], gi[2], gi[3], rel_evm, rel_lin), file=sys.stderr)
    return {"pass": bool(ok), "frames": n_frames, "n_sym": N_SYM, "snr_points": len(SNRS), "mode": "exact" if mode == 0 else "fast",
            "noise": "reference rand()/Box-Muller stream (ref_capture_gkeep), srand(%d + chunk index) per %d-frame chunk" % (base_seed, chunk),
            "cpu": {"impl": "oracle/_ref ref_chain_sweep (unmodified src/OFDM.c)", "procs": procs, "seconds": cpu_s},
            "gpu": {"call": "ofdm_sweep_inject_host", "seconds_wall_incl_pageable_h2d": gpu_s, "frames_replayed_exactly": replayed},
            "points": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1_000_000)
    ap.add_argument("--procs", type=int, default=len(os.sched_getaffinity(0)))
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    res = run(args.frames, args.procs)
    text = json.dumps(res, indent=1)
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            f.write(text + "\n")
    print(json.dumps({k: v for k, v in res.items() if k != "points"}))
    return 0 if res["pass"] else 1


if __name__ == "__main__":
    sys.exit(main())
