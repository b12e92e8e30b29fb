// ofdm_chain.cuh -- fused whole-chain kernels for the default frame shape (LTS + 2 data symbols,
// OFDM.c:20,439: the reference message gives data_frames_number = 2).
//
// A warp owns one frame at a time: its four 8-lane groups transform LTS half 1, LTS half 2, data
// symbol 0 and data symbol 1 in one pass (see ofdm_device.cuh).  After the transform the 96 data
// bins of the two symbols are dealt out evenly, three per lane, for equalise / slice / demod / EVM /
// BER ("items"), so no lane idles on null or pilot bins and nothing is selected per bin at run time:
// every per-item constant (which bin, which L sign, which payload word and bit) is fixed per lane.
//
//   k_mc_philox   Monte-Carlo sweep with nothing in HBM but the counters: payload bits and the noise
//                 come from Philox4x32-10 streams, the frame is built in shared memory once and reused
//                 for every SNR point of the sweep.
#pragma once
#include "ofdm_kernels.cuh"

namespace ofdm {

constexpr int kWin = 72;            // skewed pitch of a 64-sample window in shared memory (bank-conflict-free groups)
constexpr int kMaxSnr = 64;

struct ItemConst {
    int f_off[3];                   // where the item's FFT output sits in the F exchange tile: sym*kWin + bin
    int bin[3];                     // natural FFT bin (index into the LTS exchange tiles)
    float sc[3];                    // 0.5 * L[bin]
    int word[3];                    // payload word of the frame holding the item's bit pair (sym*3 + d/16)
    int shift[3];                   // position of the pair in that word
};

__device__ __forceinline__ ItemConst make_items(int lane)
{
    ItemConst c;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int t = lane + 32 * r, sym = t / 48, d = t - 48 * sym;
        const int bin = c_tab.data_bin[d];
        c.f_off[r] = sym * kWin + bin;
        c.bin[r] = bin;
        c.sc[r] = 0.5f * (float)c_tab.bin_lts[bin];
        c.word[r] = sym * 3 + (d >> 4);
        c.shift[r] = 2 * (d & 15);
    }
    return c;
}

// One data bin (see process_bin_hot in ofdm_kernels.cuh for the exactness argument): returns the
// rail errors packed as  I | Q << 8 | both << 16  and adds |E - tx|^2 to e2.
template <bool EXACT>
__device__ __forceinline__ uint32_t item_eval(float2 F, float2 Hh, float sc, uint32_t txp, float &e2)
{
    return process_bin_hot<EXACT>(F, Hh, sc, txp, true, e2);
}

struct McParams {
    uint32_t seed;
    uint64_t frame0;
    long n_frames;
    int n_snr;
    float snr_lin[kMaxSnr];         // (float)pow(10, snr/10), OFDM.c:645
    ofdm_counters *counters;        // [n_snr], accumulated into
};

// per-warp shared memory of the fused kernels
struct WarpShared {
    float2 tile[kWarpTile];         // transform transpose tile, then the F exchange tile (2 x kWin used)
    float2 lts[2][kWin];            // FFT of the two received LTS halves
    float2 body[2][kWin];           // the frame's two symbol bodies in time (skewed windows)
    uint32_t acc_u[kMaxSnr][4];     // per SNR point: I-rail, Q-rail, both-rail errors, frames in error
    float acc_f[kMaxSnr][2];        // per SNR point: sum |E-tx|^2, sum of per-frame EVM (flushed to double at the end)
};

template <bool EXACT>
__global__ void __launch_bounds__(kThreads, 2) k_mc_philox(McParams p)
{
    extern __shared__ __align__(16) unsigned char s_raw[];
    WarpShared *ws_all = reinterpret_cast<WarpShared *>(s_raw);
    float2 *s_ltsx = reinterpret_cast<float2 *>(s_raw + sizeof(WarpShared) * kWarpsPerBlock);   // [2][kWin] LTS halves, time
    double *s_terms = reinterpret_cast<double *>(s_ltsx + 2 * kWin);                            // [warps][160] (EXACT power)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, grp = lane >> 3, u = lane & 7;
    WarpShared &ws = ws_all[warp];
    float2 *tile = ws.tile + grp * kGroupPitch;
    Tw<EXACT> tw; tw.load(u);
    const ItemConst ic = make_items(lane);
    int dmap_tx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) dmap_tx[i] = c_tab.bin_data[8 * slot_m<EXACT>(i) + u];

    for (int i = threadIdx.x; i < 128; i += kThreads) s_ltsx[(i >> 6) * kWin + (i & 63)] = c_tab.lts_time[32 + i];
    for (int i = lane; i < kMaxSnr; i += 32) {
        ws.acc_u[i][0] = ws.acc_u[i][1] = ws.acc_u[i][2] = ws.acc_u[i][3] = 0u;
        ws.acc_f[i][0] = ws.acc_f[i][1] = 0.f;
    }
    __syncthreads();

    const float2 *src = grp < 2 ? s_ltsx + grp * kWin : ws.body[grp - 2];
    const int blk_base = (grp < 2 ? 8 + 16 * grp : 44 + 20 * (grp - 2)) + u;        // window_block_base(n0) + u
    const double q = (double)kQpsk;
    const float inv_ref2 = (float)(1.0 / (96.0 * (2.0 * q * q)));
    uint32_t n_done = 0;

    for (long f = (long)blockIdx.x * kWarpsPerBlock + warp; f < p.n_frames; f += (long)gridDim.x * kWarpsPerBlock) {
        const uint64_t fr = p.frame0 + (uint64_t)f;
        // ---- payload bits (Philox, one block per symbol) and Transmitter :500-565 for the two symbols
        const uint4 b0 = Philox::run(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), 0u, kDomainBits), p.seed, 0u);
        const uint4 b1 = Philox::run(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), 1u, kDomainBits), p.seed, 0u);
        {
            const bool s1 = (grp & 1) != 0;
            const uint32_t w0 = s1 ? b1.x : b0.x, w1 = s1 ? b1.y : b0.y, w2 = s1 ? b1.z : b0.z;
            float2 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int d = dmap_tx[i];
                float2 x = make_float2(0.f, -0.f);
                if (d >= 0) { x = qpsk_point(bit_pair(w0, w1, w2, d)); x.y = -x.y; }
                else if (d == -2) x.x = 1.f;
                else if (d == -3) x.x = -1.f;
                v[i] = x;
            }
            fft64<EXACT>(v, tw, tile, u);
            float pw = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int np = u + 8 * ((j + 4) & 7);
                const float2 y = make_float2(v[j].x * 0.015625f, -v[j].y * 0.015625f);
                if (grp >= 2) ws.body[grp - 2][np] = y;
                const float e = fmaf(y.x, y.x, y.y * y.y);
                pw += (grp >= 2) ? (np >= 48 ? 2.f * e : e) : 0.f;      // the CP repeats samples 48..63
            }
            __syncwarp();
            float P;
            if (EXACT) {
                // OFDM.c:637-643 on the 320-sample frame: the LTS prefix is a constant, the 160 data samples follow in order
                double *terms = s_terms + warp * 160;
                for (int i = lane; i < 160; i += 32) {
                    const int s = i / 80, k = i - 80 * s;
                    const float2 y = ws.body[s][k < 16 ? 48 + k : k - 16];
                    const double h = hypot_glibc((double)y.x, (double)y.y);
                    terms[i] = __dmul_rn(h, h);
                }
                __syncwarp();
                float acc = c_tab.lts_power_prefix;
                if (lane == 0) for (int i = 0; i < 160; ++i) acc = __double2float_rn(__dadd_rn((double)acc, terms[i]));
                P = __fdiv_rn(__shfl_sync(0xffffffffu, acc, 0), 320.f);
            } else {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pw += __shfl_xor_sync(0xffffffffu, pw, o);
                P = (pw + c_tab.lts_power_sum) * (1.f / 320.f);
            }
            // ---- SNR loop OFDM.c:1202: channel :635 + receiver :1018-1165 on the frame held in shared memory
            for (int si = 0; si < p.n_snr; ++si) {
                double sigma_d = 0.0; float sigma_f;
                if (EXACT) { sigma_d = __dsqrt_rn((double)__fdiv_rn(P, p.snr_lin[si])); sigma_f = (float)sigma_d; }
                else sigma_f = sqrtf(P / p.snr_lin[si]);
                float za[4], zb[4];
                philox_normals4(p.seed, (uint32_t)si, fr, (uint32_t)blk_base, kDomainNoise, za);
                philox_normals4(p.seed, (uint32_t)si, fr, (uint32_t)(blk_base + 8), kDomainNoise, zb);
                float2 r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int m = slot_m<EXACT>(i);
                    float2 s = src[u + 8 * m];
                    s.x = add_noise<EXACT>(s.x, m < 4 ? za[m & 3] : zb[m & 3], sigma_d, sigma_f);
                    r[i] = s;
                }
                fft64<EXACT>(r, tw, tile, u);
                // exchange: LTS groups publish A / B, data groups publish F (transform tile is free now)
                float2 *dst = grp < 2 ? ws.lts[grp] : ws.tile + (grp - 2) * kWin;
#pragma unroll
                for (int j = 0; j < 8; ++j) dst[u + 8 * j] = r[j];
                __syncwarp();
                float e2 = 0.f;
                uint32_t pk = 0;
#pragma unroll
                for (int t = 0; t < 3; ++t) {
                    const float2 A = ws.lts[0][ic.bin[t]], B = ws.lts[1][ic.bin[t]];
                    const float2 Hh = make_float2(__fmul_rn(__fadd_rn(A.x, B.x), ic.sc[t]), __fmul_rn(__fadd_rn(A.y, B.y), ic.sc[t]));
                    const int wsel = ic.word[t];
                    const uint32_t w = wsel == 0 ? b0.x : wsel == 1 ? b0.y : wsel == 2 ? b0.z : wsel == 3 ? b1.x : wsel == 4 ? b1.y : b1.z;
                    pk += item_eval<EXACT>(ws.tile[ic.f_off[t]], Hh, ic.sc[t], w >> ic.shift[t], e2);
                }
                __syncwarp();
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    pk += __shfl_xor_sync(0xffffffffu, pk, o);
                    e2 += __shfl_xor_sync(0xffffffffu, e2, o);
                }
                // totals of this frame at this SNR point: lanes 0..5 each own one accumulator column
                const uint32_t ti = pk & 0xFFu, tq = (pk >> 8) & 0xFFu, tb = pk >> 16;
                if (lane < 4) {
                    const uint32_t add = lane == 0 ? ti : lane == 1 ? tq : lane == 2 ? tb : (uint32_t)((ti + 2u * tq - 2u * tb) != 0u);
                    ws.acc_u[si][lane] += add;
                } else if (lane < 6) {
                    ws.acc_f[si][lane - 4] += lane == 4 ? e2 : sqrtf(e2 * inv_ref2);
                }
            }
        }
        n_done += 1;
        // the float EVM accumulators are flushed to the global double totals often enough to keep ~1e-6 relative accuracy
        if ((n_done & 63u) == 0u || f + (long)gridDim.x * kWarpsPerBlock >= p.n_frames) {
            __syncwarp();
            for (int si = lane; si < p.n_snr; si += 32) {
                atomicAdd(&p.counters[si].sum_err2, (double)ws.acc_f[si][0]);
                atomicAdd(&p.counters[si].sum_evm_lin, (double)ws.acc_f[si][1]);
                ws.acc_f[si][0] = 0.f; ws.acc_f[si][1] = 0.f;
            }
            __syncwarp();
        }
    }
    __syncwarp();
    const double ref2_frame = 96.0 * (2.0 * q * q);
    for (int si = lane; si < p.n_snr; si += 32) {
        if (n_done == 0) break;
        ofdm_counters *o = p.counters + si;
        const unsigned long long ti = ws.acc_u[si][0], tq = ws.acc_u[si][1], tb = ws.acc_u[si][2];
        atomicAdd(reinterpret_cast<unsigned long long *>(&o->bit_errors), ti + 2ull * tq - 2ull * tb);
        atomicAdd(reinterpret_cast<unsigned long long *>(&o->rail_errors), ti + tq);
        atomicAdd(reinterpret_cast<unsigned long long *>(&o->frames_in_error), (unsigned long long)ws.acc_u[si][3]);
        atomicAdd(reinterpret_cast<unsigned long long *>(&o->frames), (unsigned long long)n_done);
        atomicAdd(reinterpret_cast<unsigned long long *>(&o->bits), 192ull * n_done);
        atomicAdd(&o->sum_ref2, ref2_frame * (double)n_done);
    }
}

inline size_t mc_smem_bytes()
{
    return sizeof(WarpShared) * kWarpsPerBlock + 2 * kWin * sizeof(float2) + kWarpsPerBlock * 160 * sizeof(double);
}

// payload bits of the Philox bit stream, for callers that want the same frames in HBM (symbol s = block s)
__global__ void k_philox_bits(uint32_t seed, uint64_t frame0, long n_symbols, int n_sym, uint32_t *__restrict__ bits)
{
    long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_symbols) return;
    const long f = t / n_sym; const uint32_t s = (uint32_t)(t - f * n_sym);
    const uint64_t fr = frame0 + (uint64_t)f;
    const uint4 r = Philox::run(make_uint4((uint32_t)fr, (uint32_t)(fr >> 32), s, kDomainBits), seed, 0u);
    bits[t * 3] = r.x; bits[t * 3 + 1] = r.y; bits[t * 3 + 2] = r.z;
}

}  // namespace ofdm
