"""ctypes front-ends for the two CPU checkers -- TEST INFRASTRUCTURE ONLY.

* ``Ref``    -> oracle/_ref/libofdm_ref.so   (the unmodified reference src/OFDM.c behind ref_harness.c)
* ``Port``   -> oracle/libofdm_oracle.so     (this repo's C restatement, ofdm_oracle.c)

Both expose the same numpy-level methods so tests can run either against the CUDA path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; nothing under the product package does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SO = os.path.join(HERE, "_ref", "libofdm_ref.so")
PORT_SO = os.path.join(HERE, "libofdm_oracle.so")


class RxStats(C.Structure):
    _fields_ = [("evm_lin", C.c_float), ("evm_db", C.c_float), ("evm_agc_lin", C.c_float),
                ("evm_agc_db", C.c_float), ("ber", C.c_float), ("bit_errors", C.c_int), ("rail_errors", C.c_int)]


class Counters(C.Structure):
    _fields_ = [("bit_errors", C.c_uint64), ("bits", C.c_uint64), ("frames_in_error", C.c_uint64),
                ("rail_errors", C.c_uint64), ("frames", C.c_uint64),
                ("sum_err2", C.c_double), ("sum_ref2", C.c_double), ("sum_evm_lin", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force=False):
    """(Re)build the checkers with oracle/Makefile.  Building the checker is not using it."""
    if force or not os.path.exists(PORT_SO):
        subprocess.check_call(["make", "-s", "-C", HERE, "libofdm_oracle.so"])
    if force or not os.path.exists(REF_SO):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, t=C.c_float):
    return a.ctypes.data_as(C.POINTER(t))


class _Base:
    prefix = ""

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.path = path
        self._f("init").restype = C.c_int
        self._f("init")()
        self._f("frame_power").restype = C.c_float

    def _f(self, name):
        return getattr(self.lib, self.prefix + name)

    @staticmethod
    def frame_len(n_sym):
        return 160 + 80 * n_sym

    def lts_freq(self):
        out = np.zeros((64, 2), np.float32)
        self._f("lts_freq")(_p(out))
        return out

    def lts_time(self):
        out = np.zeros((160, 2), np.float32)
        self._f("lts_time")(_p(out))
        return out

    def qpsk_mod(self, bits):
        bits = _u8(bits).reshape(-1, 96)
        out = np.zeros((bits.shape[0], 48, 2), np.float32)
        self._f("qpsk_mod")(_p(bits, C.c_uint8), C.c_int(bits.shape[0]), _p(out))
        return out

    def map_grid(self, mod):
        mod = _f32(mod).reshape(-1, 48, 2)
        out = np.zeros((mod.shape[0], 64, 2), np.float32)
        self._f("map_grid")(_p(mod), C.c_int(mod.shape[0]), _p(out))
        return out

    def ifft64(self, x):
        x = _f32(x).reshape(-1, 64, 2)
        out = np.zeros_like(x)
        for i in range(x.shape[0]):
            self._f("ifft64")(_p(x[i]), _p(out[i]))
        return out

    def fft64(self, x):
        x = _f32(x).reshape(-1, 64, 2)
        out = np.zeros_like(x)
        for i in range(x.shape[0]):
            self._f("fft64")(_p(x[i]), _p(out[i]))
        return out

    def tx_frames(self, bits, n_sym):
        bits = _u8(bits).reshape(-1, 96 * n_sym)
        out = np.zeros((bits.shape[0], self.frame_len(n_sym), 2), np.float32)
        for i in range(bits.shape[0]):
            self._f("tx_frame")(_p(bits[i], C.c_uint8), C.c_int(n_sym), _p(out[i]))
        return out

    def frame_power(self, tx):
        tx = _f32(tx).reshape(-1, 2)
        return float(self._f("frame_power")(_p(tx), C.c_int(tx.shape[0])))

    def awgn_inject(self, tx, g, snr_db):
        tx = _f32(tx)
        frames = tx.reshape(-1, tx.shape[-2], 2)
        g = _f32(g).reshape(frames.shape[0], frames.shape[1])
        out = np.zeros_like(frames)
        for i in range(frames.shape[0]):
            self._f("awgn_inject")(_p(frames[i]), _p(g[i]), _p(out[i]), C.c_float(snr_db), C.c_int(frames.shape[1]))
        return out.reshape(tx.shape)

    def rx_frames(self, ota, bits, n_sym):
        ota = _f32(ota).reshape(-1, self.frame_len(n_sym), 2)
        bits = _u8(bits).reshape(-1, 96 * n_sym)
        n = ota.shape[0]
        res = dict(H=np.zeros((n, 64, 2), np.float32), eq=np.zeros((n, n_sym * 48, 2), np.float32),
                   sliced=np.zeros((n, n_sym * 48, 2), np.float32), bits=np.zeros((n, n_sym * 96), np.uint8),
                   evm_lin=np.zeros(n, np.float32), evm_db=np.zeros(n, np.float32),
                   evm_agc_lin=np.zeros(n, np.float32), evm_agc_db=np.zeros(n, np.float32),
                   ber=np.zeros(n, np.float32), bit_errors=np.zeros(n, np.int32), rail_errors=np.zeros(n, np.int32))
        st = RxStats()
        for i in range(n):
            self._f("rx_frame")(_p(ota[i]), C.c_int(n_sym), _p(bits[i], C.c_uint8), _p(res["H"][i]), _p(res["eq"][i]),
                                _p(res["sliced"][i]), _p(res["bits"][i], C.c_uint8), C.byref(st))
            for k in ("evm_lin", "evm_db", "evm_agc_lin", "evm_agc_db", "ber", "bit_errors", "rail_errors"):
                res[k][i] = getattr(st, k)
        return res

    def rrc_taps(self):
        t = np.zeros(21, np.float32)
        self._f("rrc_taps")(_p(t))
        return t

    def rrc_tx(self, frames):
        frames = _f32(frames)
        fr = frames.reshape(-1, frames.shape[-2], 2)
        out = np.zeros((fr.shape[0], 2 * fr.shape[1] + 20, 2), np.float32)
        for i in range(fr.shape[0]):
            self._f("rrc_tx")(_p(fr[i]), C.c_int(fr.shape[1]), _p(out[i]))
        return out

    def rrc_rx(self, x, packet_idx, frame_len):
        x = _f32(x)
        xr = x.reshape(-1, x.shape[-2], 2)
        out = np.zeros((xr.shape[0], frame_len, 2), np.float32)
        for i in range(xr.shape[0]):
            self._f("rrc_rx")(_p(xr[i]), C.c_int(xr.shape[1]), C.c_int(packet_idx), C.c_int(frame_len), _p(out[i]))
        return out

    def sts_time(self):
        out = np.zeros((160, 2), np.float32)
        self._f("sts_time")(_p(out))
        return out

    def cfo_coarse(self, rx):
        rx = _f32(rx).reshape(-1, 2)
        out = np.zeros_like(rx)
        self._f("cfo_coarse")(_p(rx), C.c_int(rx.shape[0]), _p(out))
        return out

    def cfo_fine(self, rx):
        rx = _f32(rx).reshape(-1, 2)
        out = np.zeros_like(rx)
        self._f("cfo_fine")(_p(rx), C.c_int(rx.shape[0]), _p(out))
        return out

    def packet_detection(self, rx):
        rx = _f32(rx).reshape(-1, 2)
        out = np.zeros((rx.shape[0] - 47, 2), np.float32)
        self._f("packet_detection")(_p(rx), C.c_int(rx.shape[0]), _p(out))
        return out

    def packet_selection(self, corr):
        corr = _f32(corr).reshape(-1, 2)
        self._f("packet_selection").restype = C.c_int
        return int(self._f("packet_selection")(_p(corr), C.c_int(corr.shape[0])))

    def chain(self, bits, g, n_sym, snr_db, noise_mode=0, per_frame=False):
        bits = _u8(bits).reshape(-1, 96 * n_sym)
        n = bits.shape[0]
        gp = None
        if noise_mode == 0:
            g = _f32(g).reshape(n, self.frame_len(n_sym))
            gp = _p(g)
        acc = Counters()
        fe = np.zeros(n, np.int32) if per_frame else None
        fv = np.zeros(n, np.float32) if per_frame else None
        self._f("chain")(_p(bits, C.c_uint8), gp, C.c_long(n), C.c_int(n_sym), C.c_float(snr_db), C.c_int(noise_mode),
                         C.byref(acc), _p(fe, C.c_int) if per_frame else None, _p(fv) if per_frame else None)
        return (acc, fe, fv) if per_frame else acc

    def chain_sweep(self, bits, g, n_sym, snr_db, noise_mode=0):
        """Transmitter once per frame, channel + receiver per SNR point (the loop of main() :1191-1222); list of Counters."""
        bits = _u8(bits).reshape(-1, 96 * n_sym)
        n = bits.shape[0]
        gp = None
        if noise_mode == 0:
            g = _f32(g).reshape(n, self.frame_len(n_sym))
            gp = _p(g)
        snr = np.ascontiguousarray(snr_db, dtype=np.float32)
        acc = (Counters * len(snr))()
        self._f("chain_sweep")(_p(bits, C.c_uint8), gp, C.c_long(n), C.c_int(n_sym), _p(snr), C.c_int(len(snr)), C.c_int(noise_mode), acc)
        return list(acc)


class Ref(_Base):
    """The compiled, unmodified reference (oracle/_ref)."""
    prefix = "ref_"

    def __init__(self):
        super().__init__(REF_SO)

    def seed(self, s):
        self.lib.ref_seed(C.c_uint(s))

    def awgn(self, tx, snr_db, seed=None):
        tx = _f32(tx).reshape(-1, 2)
        out = np.zeros_like(tx)
        self.lib.ref_awgn(_p(tx), _p(out), C.c_float(snr_db), C.c_int(tx.shape[0]),
                          C.c_int(seed is not None), C.c_uint(seed or 0))
        return out

    def capture_gkeep(self, n, seed=None):
        g = np.zeros(n, np.float32)
        self.lib.ref_capture_gkeep(C.c_int(seed is not None), C.c_uint(seed or 0), C.c_long(n), _p(g))
        return g

    def main_default(self, seed, workdir, quiet=True):
        """cfg0: the reference's own main(); writes data/Output_*.txt under ``workdir``."""
        os.makedirs(os.path.join(workdir, "data"), exist_ok=True)
        cwd = os.getcwd()
        os.chdir(workdir)
        try:
            return self.lib.ref_main_default(C.c_uint(seed), C.c_int(quiet))
        finally:
            os.chdir(cwd)

    def transmit_full(self):
        """Transmitter(): the reference's own 9800-sample over-the-air waveform (STS, LTS, 2 symbols, x2, RRC, x10)."""
        out = np.zeros((12000, 2), np.float32)
        self.lib.ref_transmit_full.restype = C.c_int
        n = self.lib.ref_transmit_full(_p(out), C.c_int(12000))
        return out[:n]

    def full_point(self, seed_noise, seed_rx, snr_db):
        """One SNR point of the reference's main(): returns (ota [9800,2], Res[3] = EVM_dB, EVM_AGC_dB, BER, rx_start)."""
        ota = np.zeros((9800, 2), np.float32)
        res = np.zeros(3, np.float32)
        start = C.c_int(0)
        self.lib.ref_full_point(C.c_uint(seed_noise), C.c_uint(seed_rx), C.c_float(snr_db), _p(ota), _p(res), C.byref(start))
        return ota, res, start.value

    def write_complex(self, a, fname):
        a = _f32(a).reshape(-1, 2)
        self.lib.ref_write_complex(_p(a), C.c_int(a.shape[0]), fname.encode())

    def write_float(self, a, fname):
        a = _f32(a).reshape(-1)
        self.lib.ref_write_float(_p(a), C.c_int(a.shape[0]), fname.encode())


class Port(_Base):
    """This repo's C restatement (oracle/ofdm_oracle.c)."""
    prefix = "orc_"

    def __init__(self):
        super().__init__(PORT_SO)

    def twiddles(self):
        out = np.zeros((32, 2), np.float64)
        self.lib.orc_twiddles(_p(out, C.c_double))
        return out

    def philox(self, ctr, key):
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        self.lib.orc_philox4x32_10(c, k, o)
        return list(o)

    def philox_bits(self, seed, frame0, n_frames, n_sym):
        bits = np.zeros((n_frames, 96 * n_sym), np.uint8)
        self.lib.orc_philox_bits(C.c_uint32(seed), C.c_uint64(frame0), C.c_long(n_frames), C.c_int(n_sym), _p(bits, C.c_uint8))
        return bits

    def philox_normals(self, seed, stream, frame0, n_frames, length):
        g = np.zeros((n_frames, length), np.float32)
        self.lib.orc_philox_normals(C.c_uint32(seed), C.c_uint32(stream), C.c_uint64(frame0), C.c_long(n_frames),
                                    C.c_int(length), _p(g))
        return g

    def philox_taps(self, seed, frame0, n_frames, n_taps):
        t = np.zeros((n_frames, n_taps, 2), np.float32)
        self.lib.orc_philox_taps(C.c_uint32(seed), C.c_uint64(frame0), C.c_long(n_frames), C.c_int(n_taps), _p(t))
        return t

    def apply_taps(self, tx, taps):
        tx = _f32(tx)
        frames = tx.reshape(-1, tx.shape[-2], 2)
        taps = _f32(taps).reshape(frames.shape[0], -1, 2)
        out = np.zeros_like(frames)
        for i in range(frames.shape[0]):
            self.lib.orc_apply_taps(_p(frames[i]), _p(taps[i]), C.c_int(taps.shape[1]), _p(out[i]), C.c_int(frames.shape[1]))
        return out.reshape(tx.shape)

    def chain_multipath(self, bits, g, taps, n_sym, snr_db, per_frame=False):
        bits = _u8(bits).reshape(-1, 96 * n_sym)
        n = bits.shape[0]
        g = _f32(g).reshape(n, self.frame_len(n_sym))
        taps = _f32(taps).reshape(n, -1, 2)
        acc = Counters()
        fe = np.zeros(n, np.int32) if per_frame else None
        fv = np.zeros(n, np.float32) if per_frame else None
        self.lib.orc_chain_multipath(_p(bits, C.c_uint8), _p(g), _p(taps), C.c_int(taps.shape[1]), C.c_long(n), C.c_int(n_sym),
                                     C.c_float(snr_db), C.byref(acc), _p(fe, C.c_int) if per_frame else None,
                                     _p(fv) if per_frame else None)
        return (acc, fe, fv) if per_frame else acc


def have_ref():
    return os.path.exists(REF_SO)
