"""Error behaviour of the C-ABI with a live context: bad arguments come back as OFDM_ERR_INVALID (never a crash, never
a silent fallback), the message is retrievable, and the context stays usable afterwards.  The reference has no error
codes (allocation failure -> exit(1), OFDM.c:150-153); this is the boundary's replacement for it."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

INVALID = 1


def test_bad_arguments_are_rejected(ofdm, pkg):
    lib, h = ofdm.lib, ofdm.h
    t = ofdm.torch
    bits = t.zeros((6,), dtype=t.int32, device=ofdm.device)
    frames = t.zeros((320, 2), dtype=t.float32, device=ofdm.device)
    cnt = ofdm.new_counters(1)
    snr = np.zeros(70, np.float32)
    assert lib.ofdm_ctx_set_option(h, b"no_such_option", 1) == INVALID
    assert b"unknown option" in lib.ofdm_last_error(h)
    assert lib.ofdm_ctx_set_option(h, b"multipath_path", 3) == INVALID
    assert lib.ofdm_ctx_replayed_frames(h, None, 0) == INVALID
    assert lib.ofdm_tx_frames(h, None, frames.data_ptr(), None, 1, 2, pkg.MODE_EXACT) == INVALID          # null bits
    assert lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), None, 1, 0, pkg.MODE_EXACT) == INVALID  # n_sym < 1
    assert lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), None, 1, 2, 7) == INVALID              # unknown mode
    assert lib.ofdm_tx_frames(h, bits.data_ptr(), frames.data_ptr(), None, -1, 2, pkg.MODE_EXACT) == INVALID
    assert lib.ofdm_rx_frames(h, frames.data_ptr(), None, 1, 2, pkg.MODE_EXACT, cnt.data_ptr(), None) == INVALID
    assert lib.ofdm_awgn_rx_inject(h, frames.data_ptr(), None, None, bits.data_ptr(), C.c_float(3.0), 1, 2, pkg.MODE_EXACT,
                                   cnt.data_ptr(), None) == INVALID                                        # injected noise without draws
    assert lib.ofdm_mc_sweep_philox_dev(h, 1, 0, 10, 2, snr.ctypes.data, 70, pkg.MODE_FAST, cnt.data_ptr()) == INVALID   # > 64 SNR points
    assert lib.ofdm_mc_sweep_multipath_dev(h, 1, 0, 10, 2, 17, snr.ctypes.data, 1, pkg.MODE_FAST, cnt.data_ptr()) == INVALID  # > 16 taps
    assert lib.ofdm_packet_detect(h, frames.data_ptr(), frames.data_ptr(), 1, 10) == INVALID               # capture shorter than a window
    # empty batches are fine and launch nothing
    before = ofdm.launch_count
    assert lib.ofdm_tx_frames(h, None, None, None, 0, 2, pkg.MODE_EXACT) == 0
    assert lib.ofdm_rx_frames(h, None, None, 0, 2, pkg.MODE_EXACT, None, None) == 0
    assert ofdm.launch_count == before
    # ... and the context still works
    packed = ofdm.random_bits(1, 0, 100, 2)
    fr = ofdm.tx_frames(packed, 2, pkg.MODE_EXACT, with_power=False)
    c, _ = ofdm.rx_frames(fr, packed, 2, pkg.MODE_EXACT)
    assert c.bit_errors == 0 and c.frames == 100


def test_null_context_everywhere(pkg, lib):
    """every entry point that takes a context rejects a null one"""
    assert lib.ofdm_ctx_sync(None) != 0
    assert lib.ofdm_ctx_set_option(None, b"force_replay", 1) != 0
    assert lib.ofdm_ctx_replayed_frames(None, None, 0) != 0
    assert lib.ofdm_tx_frames(None, None, None, None, 1, 2, 0) != 0
    assert lib.ofdm_rx_frames(None, None, None, 1, 2, 0, None, None) != 0
    assert lib.ofdm_mc_sweep_philox_dev(None, 1, 0, 1, 2, None, 1, 0, None) != 0
