/*
 * oracle/ofdm_oracle.c -- TEST INFRASTRUCTURE ONLY (see ofdm_oracle.h).
 *
 * Plain-C restatement of the 802.11a QPSK stage chain of the reference
 * (/root/reference/src/OFDM.c), flat interleaved-float buffers, no globals beyond
 * read-only tables, no libc rand().  Each function cites the reference lines it follows.
 * Parity status: PINNED -- checked bit-for-bit against the compiled reference
 * (oracle/_ref) in tests/test_oracle_vs_ref.py and against tests/golden fixtures.
 *
 * Arithmetic notes (SURVEY.md section 7 "Hard parts"):
 *  - FFT: the reference's recursive radix-2 DIT (src/OFDM.c:282-312) has the dataflow of a
 *    bit-reversal followed by six in-place butterfly stages; twiddle = the reference's
 *    expression cexp(-I*2.0*PI*k/sz) in double, product in double (no FMA), rounded to
 *    float, add/sub in float.  W_sz^k == W_64^(k*64/sz) bit-for-bit (power-of-two scaling
 *    of the angle), so one 32-entry table serves every stage.
 *  - complex float division (:1050) is libgcc __divsc3 = double-widened quotient formula.
 */
#include "ofdm_oracle.h"
#include <complex.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define PI_REF 3.14159265358979323846   /* src/OFDM.c:13 */

typedef struct { float re, im; } cf32;

static double tw_re[32], tw_im[32];
static cf32 lts_freq[64];               /* centred order, DC at index 32 */
static cf32 lts_time[160];
static int inited = 0;

/* centred index c (0..63, DC at 32) -> data symbol index 0..47, -1 null, -2 pilot(+1), -3 pilot(-1)
 * (layout of src/OFDM.c:528-547 and its inverse :1063-1068) */
static int8_t data_index_of[64];

static const signed char Lk[53] = {1,1,-1,-1,1,1,-1,1,-1,1,1,1,1,1,1,-1,-1,1,1,-1,1,-1,1,1,1,1,0,
                                   1,-1,-1,1,1,-1,1,-1,1,-1,-1,-1,-1,-1,1,1,-1,-1,1,-1,1,-1,1,1,1,1};

static unsigned bitrev6(unsigned v)
{
    unsigned r = 0;
    for (int b = 0; b < 6; ++b) r |= ((v >> b) & 1u) << (5 - b);
    return r;
}

/* fft_Cooley :282-312 as an iterative transform; natural-order output (no shift). */
static void dit64(const cf32 *x, cf32 *y)
{
    for (unsigned i = 0; i < 64; ++i) y[i] = x[bitrev6(i)];
    for (int sz = 2; sz <= 64; sz <<= 1) {
        int half = sz >> 1, step = 64 / sz;
        for (int base = 0; base < 64; base += sz)
            for (int k = 0; k < half; ++k) {
                double wr = tw_re[k * step], wi = tw_im[k * step];
                cf32 o = y[base + k + half], e = y[base + k];
                /* double complex * (float complex promoted): (wr*c - wi*d) + i(wr*d + wi*c), :303 */
                double pr = wr * (double)o.re - wi * (double)o.im;
                double pi = wr * (double)o.im + wi * (double)o.re;
                float wre = (float)pr, wim = (float)pi;
                y[base + k].re = e.re + wre;        y[base + k].im = e.im + wim;          /* :304 */
                y[base + k + half].re = e.re - wre; y[base + k + half].im = e.im - wim;   /* :305 */
            }
    }
}

/* fft :314-318 = Cooley + fft_shift :227 (bin p lands at centred index (p+32)%64) */
static void fft_centred(const cf32 *x, cf32 *y)
{
    cf32 t[64];
    dit64(x, t);
    for (int p = 0; p < 64; ++p) y[(p + 32) & 63] = t[p];
}

/* ifft :320-339 = ifft_shift :208, conj, fft (including ITS fft_shift -> 32-sample rotation,
 * SURVEY.md Q4), conj, /64 */
static void ifft_centred(const cf32 *X, cf32 *y)
{
    cf32 z[64], t[64];
    for (int i = 0; i < 64; ++i) { cf32 v = X[(i + 32) & 63]; z[i].re = v.re; z[i].im = -v.im; }
    fft_centred(z, t);
    for (int i = 0; i < 64; ++i) { y[i].re = t[i].re / 64.0f; y[i].im = -t[i].im / 64.0f; }
}

int orc_init(void)
{
    if (inited) return 0;
    for (int k = 0; k < 32; ++k) {
        double complex w = cexp(-I * 2.0 * PI_REF * k / 64);   /* expression of :303 with sz = 64 */
        tw_re[k] = creal(w); tw_im[k] = cimag(w);
    }
    for (int c = 0; c < 64; ++c) data_index_of[c] = -1;
    {
        /* runs of data bins: c=6..10, 12..24, 26..31, 33..38, 40..52, 54..58 */
        static const int lo[6] = {6, 12, 26, 33, 40, 54}, hi[6] = {10, 24, 31, 38, 52, 58};
        int d = 0;
        for (int r = 0; r < 6; ++r) for (int c = lo[r]; c <= hi[r]; ++c) data_index_of[c] = (int8_t)d++;
        data_index_of[11] = -2; data_index_of[25] = -2; data_index_of[39] = -2; data_index_of[53] = -3;
    }
    /* LTS: Preamble_Generator(type 1) :368-399 with L_k :494 at c = 6..58 */
    memset(lts_freq, 0, sizeof lts_freq);
    for (int i = 0; i < 53; ++i) lts_freq[6 + i].re = (float)Lk[i];
    cf32 t[64];
    ifft_centred(lts_freq, t);
    for (int i = 0; i < 32; ++i) lts_time[i] = t[32 + i];                    /* :396 */
    for (int r = 0; r < 2; ++r) for (int i = 0; i < 64; ++i) lts_time[32 + 64 * r + i] = t[i];  /* :397 */
    inited = 1;
    return 0;
}

void orc_twiddles(double *out) { orc_init(); for (int k = 0; k < 32; ++k) { out[2*k] = tw_re[k]; out[2*k+1] = tw_im[k]; } }
void orc_lts_freq(float *out) { orc_init(); memcpy(out, lts_freq, sizeof lts_freq); }
void orc_lts_time(float *out) { orc_init(); memcpy(out, lts_time, sizeof lts_time); }

/* QPSK_Modulator :415-433: 00->(+,+) 01->(-,+) 10->(-,-) 11->(+,-), each (+-1)/sqrt(2) in double -> float */
static cf32 qpsk_point(int a, int b)
{
    float p = (float)(1.0 / sqrt(2.0)), m = (float)(-1.0 / sqrt(2.0));
    cf32 v;
    v.re = (a ^ b) ? m : p;     /* I rail negative for 01 and 10 */
    v.im = a ? m : p;           /* Q rail negative for 10 and 11 */
    return v;
}
void orc_qpsk_mod(const uint8_t *bits, int n_sym, float *out)
{
    cf32 *o = (cf32 *)out;
    for (int i = 0; i < n_sym * 48; ++i) o[i] = qpsk_point(bits[2*i], bits[2*i+1]);
}

static void map_symbol(const cf32 *mod48, cf32 *grid)
{
    for (int c = 0; c < 64; ++c) {
        int d = data_index_of[c];
        cf32 v = {0.0f, 0.0f};
        if (d >= 0) v = mod48[d];
        else if (d == -2) v.re = 1.0f;
        else if (d == -3) v.re = -1.0f;
        grid[c] = v;
    }
}
void orc_map_grid(const float *mod, int n_sym, float *grid)
{
    orc_init();
    for (int s = 0; s < n_sym; ++s) map_symbol((const cf32 *)mod + 48*s, (cf32 *)grid + 64*s);
}

void orc_ifft64(const float *in, float *out) { orc_init(); ifft_centred((const cf32 *)in, (cf32 *)out); }
void orc_fft64(const float *in, float *out)  { orc_init(); fft_centred((const cf32 *)in, (cf32 *)out); }

/* Transmitter :500-581 without the STS slot: LTS(160) || n_sym x (16-sample CP + 64) */
static void tx_frame(const uint8_t *bits, int n_sym, cf32 *frame, cf32 *mod_keep)
{
    memcpy(frame, lts_time, sizeof lts_time);
    for (int s = 0; s < n_sym; ++s) {
        cf32 mod[48], grid[64], t[64];
        for (int j = 0; j < 48; ++j) mod[j] = qpsk_point(bits[96*s + 2*j], bits[96*s + 2*j + 1]);
        if (mod_keep) memcpy(mod_keep + 48*s, mod, sizeof mod);
        map_symbol(mod, grid);
        ifft_centred(grid, t);
        cf32 *dst = frame + 160 + 80*s;
        memcpy(dst, t + 48, 16 * sizeof(cf32));      /* :563 */
        memcpy(dst + 16, t, 64 * sizeof(cf32));      /* :564 */
    }
}
void orc_tx_frame(const uint8_t *bits, int n_sym, float *frame) { orc_init(); tx_frame(bits, n_sym, (cf32 *)frame, NULL); }

/* Transmission_Over_Air :637-643: float accumulator, double terms cabs*cabs, sequential */
static float frame_power(const cf32 *x, int len)
{
    float p = 0.0f;
    for (int i = 0; i < len; ++i) {
        double h = hypot((double)x[i].re, (double)x[i].im);     /* cabs of the promoted value */
        p = (float)((double)p + h * h);
    }
    return p / (float)len;
}
float orc_frame_power(const float *tx, int len) { return frame_power((const cf32 *)tx, len); }

/* :645-653 with the surviving draw injected: real rail only */
static void awgn_inject(const cf32 *x, const float *g, cf32 *y, float snr_db, int len)
{
    float p = frame_power(x, len);
    float snr_lin = (float)pow(10.0, (double)(snr_db / 10));
    float npow = p / snr_lin;
    double sigma = sqrt((double)npow);
    for (int i = 0; i < len; ++i) {
        float n = (float)(sigma * (double)g[i]);
        y[i].re = x[i].re + n;
        y[i].im = x[i].im;
    }
}
void orc_awgn_inject(const float *tx, const float *g, float *out, float snr_db, int len)
{ awgn_inject((const cf32 *)tx, g, (cf32 *)out, snr_db, len); }

/* libgcc __divsc3 as built by gcc 13 (double-widened), incl. its zero-denominator recovery */
static cf32 div_cf32(cf32 n, cf32 h)
{
    double a = n.re, b = n.im, c = h.re, d = h.im;
    double den = c * c + d * d;
    double x = (a * c + b * d) / den, y = (b * c - a * d) / den;
    if (isnan(x) && isnan(y)) {
        if (den == 0.0 && (!isnan(a) || !isnan(b))) {
            x = copysign(INFINITY, c) * a;
            y = copysign(INFINITY, c) * b;
        }
    }
    cf32 r = {(float)x, (float)y};
    return r;
}

/* Receiver :1018-1165 on an LTS||data frame */
static void rx_frame(const cf32 *ota, int n_sym, const uint8_t *tx_bits, const cf32 *tx_mod,
                     cf32 *H_out, cf32 *eq_out, cf32 *sliced_out, uint8_t *bits_out, orc_rx_stats *st)
{
    cf32 A[64], B[64], H[64];
    fft_centred(ota + 32, A);                         /* Channel_Estimation :837-844 (offsets 192, 256 minus the STS) */
    fft_centred(ota + 96, B);
    for (int c = 0; c < 64; ++c) {                    /* :848  0.5*(A+B)*conj(L) in double, L real */
        float sr = A[c].re + B[c].re, si = A[c].im + B[c].im;
        double hr = 0.5 * (double)sr, hi = 0.5 * (double)si;
        double lr = lts_freq[c].re, li = -(double)lts_freq[c].im;
        H[c].re = (float)(hr * lr - hi * li);
        H[c].im = (float)(hr * li + hi * lr);
    }
    if (H_out) memcpy(H_out, H, sizeof H);

    float p = (float)(1.0 / sqrt(2.0)), m = (float)(-1.0 / sqrt(2.0));
    int n = n_sym * 48;
    cf32 *eqs = (cf32 *)malloc(sizeof(cf32) * (size_t)n), *sl = (cf32 *)malloc(sizeof(cf32) * (size_t)n);
    int bit_errors = 0;
    for (int s = 0; s < n_sym; ++s) {
        cf32 F[64];
        fft_centred(ota + 160 + 80*s + 16, F);        /* CP strip :1028-1030, fft :1039 */
        for (int c = 0; c < 64; ++c) {
            int d = data_index_of[c];
            if (d < 0) continue;                      /* demap :1063-1068 keeps data bins only */
            cf32 e = div_cf32(F[c], H[c]);            /* :1050 */
            cf32 q;                                   /* AGC_Receiver :860-868 */
            q.re = e.re > 0 ? p : m;
            q.im = e.im > 0 ? p : m;
            int b0, b1;                               /* QPSK_Demodulator :883-902 */
            if (q.re > 0 && q.im > 0)      { b0 = 0; b1 = 0; }
            else if (q.re < 0 && q.im > 0) { b0 = 0; b1 = 1; }
            else if (q.re < 0 && q.im < 0) { b0 = 1; b1 = 0; }
            else                           { b0 = 1; b1 = 1; }
            int i = 48*s + d;
            eqs[i] = e; sl[i] = q;
            if (bits_out) { bits_out[2*i] = (uint8_t)b0; bits_out[2*i+1] = (uint8_t)b1; }
            bit_errors += abs((int)tx_bits[2*i] - b0) + abs((int)tx_bits[2*i+1] - b1);
        }
    }
    if (eq_out) memcpy(eq_out, eqs, sizeof(cf32) * (size_t)n);
    if (sliced_out) memcpy(sliced_out, sl, sizeof(cf32) * (size_t)n);
    if (st) {
        /* EVM loops :1110-1118 and :1134-1142: data order (symbol-major), float accumulators fed double terms */
        float err_sum = 0.0f, ref_sum = 0.0f, err_sum_agc = 0.0f, ref_sum_agc = 0.0f;
        int rail_errors = 0;
        for (int i = 0; i < n; ++i) {
            cf32 t = tx_mod[i];
            float er = eqs[i].re - t.re, ei = eqs[i].im - t.im;
            float qr = sl[i].re - t.re, qi = sl[i].im - t.im;
            double t2 = pow(hypot((double)t.re, (double)t.im), 2);
            err_sum = (float)((double)err_sum + pow(hypot((double)er, (double)ei), 2));
            ref_sum = (float)((double)ref_sum + t2);
            err_sum_agc = (float)((double)err_sum_agc + pow(hypot((double)qr, (double)qi), 2));
            ref_sum_agc = (float)((double)ref_sum_agc + t2);
            rail_errors += (qr != 0) + (qi != 0);
        }
        float evm = (float)(sqrt((double)(err_sum / n)) / sqrt((double)(ref_sum / n)));              /* :1124 */
        float evm_agc = (float)(sqrt((double)(err_sum_agc / n)) / sqrt((double)(ref_sum_agc / n)));  /* :1148 */
        st->evm_lin = evm;          st->evm_db = (float)(20 * log10((double)evm));                   /* :1126 */
        st->evm_agc_lin = evm_agc;  st->evm_agc_db = (float)(20 * log10((double)evm_agc));           /* :1150 */
        st->bit_errors = bit_errors; st->rail_errors = rail_errors;
        st->ber = (float)bit_errors / (float)(n_sym * 96);                                           /* :1161 */
    }
    free(eqs); free(sl);
}

void orc_rx_frame(const float *ota, int n_sym, const uint8_t *tx_bits, float *H_out, float *eq_out,
                  float *sliced_out, uint8_t *bits_out, orc_rx_stats *st)
{
    orc_init();
    cf32 *mod = (cf32 *)malloc(sizeof(cf32) * 48 * (size_t)n_sym);
    orc_qpsk_mod(tx_bits, n_sym, (float *)mod);
    rx_frame((const cf32 *)ota, n_sym, tx_bits, mod, (cf32 *)H_out, (cf32 *)eq_out, (cf32 *)sliced_out, bits_out, st);
    free(mod);
}

static void accumulate(orc_counters *acc, const orc_rx_stats *st, int n_sym)
{
    acc->bit_errors += (uint64_t)st->bit_errors;
    acc->bits += 96u * (uint64_t)n_sym;
    acc->frames_in_error += st->bit_errors > 0;
    acc->rail_errors += (uint64_t)st->rail_errors;
    acc->frames += 1;
    double e = (double)st->evm_lin;
    acc->sum_evm_lin += e;
    acc->sum_err2 += e * e * 48.0 * n_sym;
    acc->sum_ref2 += 48.0 * n_sym;
}

/* noise_mode 0: injected g; 2: none (mode 1, libc rand(), exists only in the _ref harness) */
void orc_chain(const uint8_t *bits, const float *g, long n_frames, int n_sym, float snr_db, int noise_mode,
               orc_counters *acc, int *frame_bit_errors, float *frame_evm_lin)
{
    orc_init();
    int len = 160 + 80 * n_sym;
    cf32 *tx = (cf32 *)malloc(sizeof(cf32) * (size_t)len), *ota = (cf32 *)malloc(sizeof(cf32) * (size_t)len);
    cf32 *mod = (cf32 *)malloc(sizeof(cf32) * 48 * (size_t)n_sym);
    for (long f = 0; f < n_frames; ++f) {
        const uint8_t *b = bits + f * 96 * n_sym;
        tx_frame(b, n_sym, tx, mod);
        if (noise_mode == 0) awgn_inject(tx, g + f * len, ota, snr_db, len);
        else memcpy(ota, tx, sizeof(cf32) * (size_t)len);
        orc_rx_stats st;
        rx_frame(ota, n_sym, b, mod, NULL, NULL, NULL, NULL, &st);
        accumulate(acc, &st, n_sym);
        if (frame_bit_errors) frame_bit_errors[f] = st.bit_errors;
        if (frame_evm_lin) frame_evm_lin[f] = st.evm_lin;
    }
    free(tx); free(ota); free(mod);
}

/* sweep form: transmitter once per frame, channel + receiver per SNR point (main() :1191-1222); acc[n_snr] */
void orc_chain_sweep(const uint8_t *bits, const float *g, long n_frames, int n_sym, const float *snr_db, int n_snr,
                     int noise_mode, orc_counters *acc)
{
    orc_init();
    int len = 160 + 80 * n_sym;
    cf32 *tx = (cf32 *)malloc(sizeof(cf32) * (size_t)len), *ota = (cf32 *)malloc(sizeof(cf32) * (size_t)len);
    cf32 *mod = (cf32 *)malloc(sizeof(cf32) * 48 * (size_t)n_sym);
    for (long f = 0; f < n_frames; ++f) {
        const uint8_t *b = bits + f * 96 * n_sym;
        tx_frame(b, n_sym, tx, mod);
        for (int s = 0; s < n_snr; ++s) {
            if (noise_mode == 0) awgn_inject(tx, g + f * len, ota, snr_db[s], len);
            else memcpy(ota, tx, sizeof(cf32) * (size_t)len);
            orc_rx_stats st;
            rx_frame(ota, n_sym, b, mod, NULL, NULL, NULL, NULL, &st);
            accumulate(acc + s, &st, n_sym);
        }
    }
    free(tx); free(ota); free(mod);
}

/* ---------------- counter-based streams (new-build definition, mirrored by csrc/philox.cuh) ---------------- */

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum { DOMAIN_NOISE = 0, DOMAIN_BITS = 1, DOMAIN_TAPS = 2 };

/* payload bits: symbol s of a frame takes words r0..r2 (96 bits, LSB first) of block s; stream id 0 */
void orc_philox_bits(uint32_t seed, uint64_t frame0, long n_frames, int n_sym, uint8_t *bits)
{
    for (long f = 0; f < n_frames; ++f) {
        uint64_t fr = frame0 + (uint64_t)f;
        for (int s = 0; s < n_sym; ++s) {
            uint32_t ctr[4] = {(uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)s, DOMAIN_BITS}, key[2] = {seed, 0u}, r[4];
            orc_philox4x32_10(ctr, key, r);
            for (int j = 0; j < 96; ++j)
                bits[(f * n_sym + s) * 96 + j] = (uint8_t)((r[j >> 5] >> (j & 31)) & 1u);
        }
    }
}

/* one Philox block -> four standard normals (two Box-Muller pairs, both branches used) */
static void block_normals(const uint32_t r[4], float z[4])
{
    for (int h = 0; h < 2; ++h) {
        float u1 = fmaf((float)r[2*h], 0x1p-32f, 0x1p-33f);
        float u2 = (float)r[2*h + 1] * 0x1p-32f;
        float rad = sqrtf(-2.0f * logf(u1));
        float ang = 6.283185307179586f * u2;
        z[2*h] = rad * cosf(ang);
        z[2*h + 1] = rad * sinf(ang);
    }
}

/* Which Philox block / which of its four normals feeds frame sample n.  The 64-sample FFT
 * windows (LTS halves, symbol bodies) use a layout in which samples l, l+8, l+16, l+24 share
 * a block, so that one GPU thread (which owns samples l+8m of a window) consumes whole blocks;
 * guard/CP samples use consecutive quadruples.  Frame = GI2(32) LTS1(64) LTS2(64) then per
 * symbol CP(16) body(64). */
static void noise_slot(int n, int *blk, int *j)
{
    int base, l;
    if (n < 32) { *blk = n >> 2; *j = n & 3; return; }
    if (n < 160) { l = (n - 32) & 63; base = n < 96 ? 8 : 24; }
    else {
        int s = (n - 160) / 80; l = (n - 160) - 80 * s;
        if (l < 16) { *blk = 40 + 20 * s + (l >> 2); *j = l & 3; return; }
        l -= 16; base = 44 + 20 * s;
    }
    *blk = base + (l & 7) + 8 * (l >> 5);
    *j = (l >> 3) & 3;
}

/* key = (seed, stream); counter = (frame_lo, frame_hi, block, DOMAIN_NOISE) */
void orc_philox_normals(uint32_t seed, uint32_t stream, uint64_t frame0, long n_frames, int len, float *g)
{
    for (long f = 0; f < n_frames; ++f) {
        uint64_t fr = frame0 + (uint64_t)f;
        for (int n = 0; n < len; ++n) {
            int blk, j;
            noise_slot(n, &blk, &j);
            uint32_t ctr[4] = {(uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)blk, DOMAIN_NOISE}, key[2] = {seed, stream}, r[4];
            float z[4];
            orc_philox4x32_10(ctr, key, r);
            block_normals(r, z);
            g[f * len + n] = z[j];
        }
    }
}

/* cfg4 extension (no counterpart in the reference): per-frame taps, i.i.d.
 * complex Gaussian with E|h_l|^2 = 1/n_taps; tap l = normals (2l, 2l+1) of the TAPS domain */
void orc_philox_taps(uint32_t seed, uint64_t frame0, long n_frames, int n_taps, float *taps)
{
    float scale = sqrtf(0.5f / (float)n_taps);
    for (long f = 0; f < n_frames; ++f) {
        uint64_t fr = frame0 + (uint64_t)f;
        for (int blk = 0; blk * 2 < n_taps; ++blk) {
            uint32_t ctr[4] = {(uint32_t)fr, (uint32_t)(fr >> 32), (uint32_t)blk, DOMAIN_TAPS}, key[2] = {seed, 0u}, r[4];
            float z[4];
            orc_philox4x32_10(ctr, key, r);
            block_normals(r, z);
            for (int j = 0; j < 2 && blk * 2 + j < n_taps; ++j) {
                taps[(f * n_taps + blk * 2 + j) * 2] = scale * z[2*j];
                taps[(f * n_taps + blk * 2 + j) * 2 + 1] = scale * z[2*j + 1];
            }
        }
    }
}

/* y[n] = sum_l h[l] x[n-l], x[n<0] = 0 (same sum order as the reference's Convolution :353-359
 * would give per output sample: ascending input index, i.e. descending l) */
void orc_apply_taps(const float *tx, const float *taps, int n_taps, float *out, int len)
{
    const cf32 *x = (const cf32 *)tx, *h = (const cf32 *)taps;
    cf32 *y = (cf32 *)out;
    for (int n = 0; n < len; ++n) {
        float ar = 0.0f, ai = 0.0f;
        for (int l = (n_taps - 1 < n ? n_taps - 1 : n); l >= 0; --l) {
            cf32 a = x[n - l], b = h[l];
            ar += a.re * b.re - a.im * b.im;
            ai += a.re * b.im + a.im * b.re;
        }
        y[n].re = ar; y[n].im = ai;
    }
}

void orc_chain_multipath(const uint8_t *bits, const float *g, const float *taps, int n_taps, long n_frames,
                         int n_sym, float snr_db, orc_counters *acc, int *frame_bit_errors, float *frame_evm_lin)
{
    orc_init();
    int len = 160 + 80 * n_sym;
    cf32 *tx = (cf32 *)malloc(sizeof(cf32) * (size_t)len), *ch = (cf32 *)malloc(sizeof(cf32) * (size_t)len);
    cf32 *ota = (cf32 *)malloc(sizeof(cf32) * (size_t)len), *mod = (cf32 *)malloc(sizeof(cf32) * 48 * (size_t)n_sym);
    for (long f = 0; f < n_frames; ++f) {
        const uint8_t *b = bits + f * 96 * n_sym;
        tx_frame(b, n_sym, tx, mod);
        orc_apply_taps((const float *)tx, taps + f * n_taps * 2, n_taps, (float *)ch, len);
        awgn_inject(ch, g + f * len, ota, snr_db, len);
        orc_rx_stats st;
        rx_frame(ota, n_sym, b, mod, NULL, NULL, NULL, NULL, &st);
        accumulate(acc, &st, n_sym);
        if (frame_bit_errors) frame_bit_errors[f] = st.bit_errors;
        if (frame_evm_lin) frame_evm_lin[f] = st.evm_lin;
    }
    free(tx); free(ch); free(ota); free(mod);
}

/* ---------------- section 8(f) rank 1: x2 oversampling + RRC pulse shaping (OFDM.c:32, 342-364, 587-605, 959-996) ----------------
 * Convolution() accumulates Out[i+j] += Inp[i]*H[j] in float, i ascending; H is real (imaginary part 0), so each
 * product is (a*h, b*h) rounded to float.  Per output sample k that is: for i ascending, j = k - i in [0, 20]. */
static const double rrc_taps_d[21] = {-0.000454720514876223, 0.00353689555574986, -0.00714560809091226, 0.00757906190517828,
    0.00214368242727367, -0.0106106866672496, 0.0300115539818315, -0.0530534333362480, -0.0750288849545787, 0.409168714634052,
    0.803738600397980, 0.409168714634052, -0.0750288849545787, -0.0530534333362480, 0.0300115539818315, -0.0106106866672496,
    0.00214368242727367, 0.00757906190517828, -0.00714560809091226, 0.00353689555574986, -0.000454720514876223};   /* RRC_Filter_Tx :32 */

void orc_rrc_taps(float *out21) { for (int i = 0; i < 21; ++i) out21[i] = (float)rrc_taps_d[i]; }

static void conv21(const cf32 *in, int n_in, cf32 *out)
{
    float h[21];
    orc_rrc_taps(h);
    for (int k = 0; k < n_in + 20; ++k) {
        float ar = 0.0f, ai = 0.0f;
        int lo = k - 20 < 0 ? 0 : k - 20, hi = k < n_in - 1 ? k : n_in - 1;
        for (int i = lo; i <= hi; ++i) { ar += in[i].re * h[k - i]; ai += in[i].im * h[k - i]; }
        out[k].re = ar; out[k].im = ai;
    }
}
void orc_rrc_tx(const float *frame, int len, float *out)
{
    const cf32 *x = (const cf32 *)frame;
    cf32 *up = (cf32 *)calloc((size_t)(2 * len), sizeof(cf32));
    for (int i = 0; i < len; ++i) up[2 * i] = x[i];
    conv21(up, 2 * len, (cf32 *)out);
    free(up);
}
void orc_rrc_rx(const float *in, int in_len, int packet_idx, int frame_len, float *out)
{
    cf32 *f = (cf32 *)malloc(sizeof(cf32) * (size_t)(in_len + 20));
    conv21((const cf32 *)in, in_len, f);
    cf32 *r = (cf32 *)out;
    int index = 0;
    for (int i = packet_idx; i < 2 * frame_len + packet_idx - 1; i += 2) r[index++] = f[i];
    free(f);
}

/* ---------------- section 8(f) rank 2: packet detection / selection (OFDM.c:659-771) ----------------
 * Packet_Detection: delay 16, window 32, no conjugate (as the reference and its MATLAB model): per lag i
 *   corr  = sum_k r[i+k]*r[i+k+16]          float complex, sequential, product (ac-bd, ad+bc) in float
 *   peak  = sum_k cabs(r[i+k+16])^2          double terms added into a float accumulator
 *   out_i = (float)( cabs(corr)^2 / (double)(float)(peak*peak) )      (libgcc __divdc3 with zero imaginary parts = a/c)
 * Packet_Selection: threshold 0.75, gaps > 300 open a packet, the correlation 230 lags later must still be above
 * the threshold, result = front + len_RRC_rx + 1 (= +11), 0 when nothing qualifies.  The reference reads
 * Corr_Out[front+230] without a bound check; here (and on the GPU) lags beyond the array count as below threshold. */
void orc_packet_detection(const float *rx, int len, float *corr_out)
{
    const cf32 *r = (const cf32 *)rx;
    int n = len - 47;
    for (int i = 0; i < n; ++i) {
        float cr = 0.0f, ci = 0.0f, peak = 0.0f;
        for (int k = 0; k < 32; ++k) {
            cf32 a = r[i + k], b = r[i + k + 16];
            float pr = a.re * b.re - a.im * b.im, pi = a.re * b.im + a.im * b.re;
            cr += pr; ci += pi;
            double h = hypot((double)b.re, (double)b.im);
            peak = (float)((double)peak + h * h);
        }
        double hc = hypot((double)cr, (double)ci);
        float p2 = peak * peak;
        corr_out[2 * i] = (float)((hc * hc) / (double)p2);
        corr_out[2 * i + 1] = (float)(0.0 / (double)p2);      /* 0/0 = NaN for an all-zero window, as __divdc3 gives */
    }
}
int orc_packet_selection(const float *corr, int len_corr)
{
    int *idx = (int *)malloc(sizeof(int) * (size_t)(len_corr + 1)), count = 0;
    for (int i = 0; i < len_corr; ++i) if (fabs((double)corr[2 * i]) > 0.75f && corr[2 * i + 1] == 0.0f) idx[count++] = i;
    int result = 0, fronts = 0, prev_front = -1;
    /* fronts: positions j in 0..count with idx[j] - idx[j-1] > 300 (idx[-1] = -1); the loop at :754 skips the last one */
    int n_fronts = 0;
    for (int j = 0; j < count; ++j) { int b = j ? idx[j - 1] : -1; if (idx[j] - b > 300) ++n_fronts; }
    for (int j = 0; j < count && result == 0; ++j) {
        int b = j ? idx[j - 1] : -1;
        if (idx[j] - b <= 300) continue;
        ++fronts;
        if (fronts > n_fronts - 1) break;                      /* x < packet_front_count - 1 */
        int look = idx[j] + 230;
        if (look < len_corr && fabs((double)corr[2 * look]) > 0.75f) result = idx[j] + 10 + 1;
    }
    (void)prev_front;
    free(idx);
    return result;
}

/* ---------------- section 8(f) ranks 3/4: short preamble and CFO stages (OFDM.c:483-492, 773-828) ---------------- */
void orc_sts_time(float *out160)
{
    orc_init();
    /* S_k :483-490 scaled by (float)sqrt(13/6) :479, placed at c = 6..58, ifft, first 16 samples x 10 (:393) */
    static const signed char Sk[53] = {0,0,1,0,0,0,-1,0,0,0, 1,0,0,0,-1,0,0,0,-1,0,0,0, 1,0,0,0,0,0,0,0,-1,0,0,0, -1,0,0,0,1,0,0,0,1,0,0,0, 1,0,0,0,1,0,0};
    float scale = (float)sqrt(13.0 / 6.0);
    cf32 grid[64], t[64];
    memset(grid, 0, sizeof grid);
    for (int i = 0; i < 53; ++i) { grid[6 + i].re = (float)Sk[i] * scale; grid[6 + i].im = (float)Sk[i] * scale; }
    ifft_centred(grid, t);
    cf32 *o = (cf32 *)out160;
    for (int r = 0; r < 10; ++r) for (int i = 0; i < 16; ++i) o[16 * r + i] = t[i];
}

static const double TS_SEC = 1 / 20e6;      /* ts_sec :17 */

/* sum_i a[i]*conj(b[i]) (:793-796, :816-819).  conj() is the DOUBLE complex function, so each product is formed in
 * double (float*float products are exact there; one rounding for the sum) and the float complex accumulator takes
 * (float)((double)acc + product), sequentially. */
static cf32 conj_dot(const cf32 *a, const cf32 *b, int n)
{
    cf32 acc = {0.0f, 0.0f};
    for (int i = 0; i < n; ++i) {
        double ar = a[i].re, ai = a[i].im, br = b[i].re, nbi = -(double)b[i].im;
        double pr = ar * br - ai * nbi, pi = ar * nbi + ai * br;
        acc.re = (float)((double)acc.re + pr); acc.im = (float)((double)acc.im + pi);
    }
    return acc;
}
/* coarse (:802): rx[i] * cexp(...) evaluated in double, rounded to float once.
 * fine (:825-826): the rotator is first stored in a FLOAT complex (exp_term), then multiplied in float. */
static void derotate(const cf32 *x, cf32 *y, int len, float freq, int rotator_in_float)
{
    for (int i = 0; i < len; ++i) {
        double ang = -2.0 * PI_REF * (double)freq * TS_SEC * i;     /* imaginary part of -I*2*PI*f*ts*i, same association */
        double c = cos(ang), s = sin(ang);
        if (rotator_in_float) {
            float cf = (float)c, sf = (float)s;
            y[i].re = x[i].re * cf - x[i].im * sf;
            y[i].im = x[i].re * sf + x[i].im * cf;
        } else {
            double xr = x[i].re, xi = x[i].im;
            y[i].re = (float)(xr * c - xi * s);
            y[i].im = (float)(xr * s + xi * c);
        }
    }
}
void orc_cfo_coarse(const float *rx, int len, float *out)
{
    const cf32 *x = (const cf32 *)rx;
    cf32 p = conj_dot(x + 80, x + 96, 16);                                                     /* STS slots 5 and 6 */
    float f = (float)((-1.0 / (2 * PI_REF * 16 * TS_SEC)) * atan2((double)p.im, (double)p.re));   /* :798 */
    derotate(x, (cf32 *)out, len, f, 0);
}
void orc_cfo_fine(const float *rx, int len, float *out)
{
    const cf32 *x = (const cf32 *)rx;
    cf32 p = conj_dot(x + 192, x + 256, 64);                                                   /* the two LTS halves */
    float f = (float)((-1.0 / (2 * PI_REF * 64 * TS_SEC)) * atan2((double)p.im, (double)p.re));   /* :821 */
    derotate(x, (cf32 *)out, len, f, 1);
}
