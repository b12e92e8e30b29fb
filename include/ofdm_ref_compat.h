/*
 * ofdm_ref_compat.h -- the reference's OWN function names and signatures (src/OFDM.c) on host pointers, served by
 * libofdm_b200.so.  Optional convenience for a maintainer who wants to move call sites over one at a time: include this
 * header instead of defining the functions, link -lofdm_b200, and the existing `main()` body keeps compiling.  Every
 * function copies its (small, per-frame) arguments to the GPU, runs the batched stage of include/ofdm_b200.h in
 * OFDM_MODE_EXACT on a batch of one, and copies the result back -- bit-identical to the reference's function, slow by
 * construction (one launch per call); the batched entry points are the product.  There is no CPU fallback: without a
 * CUDA device the first call prints the error and exit(1)s, as the reference does on allocation failure (:150-153).
 *
 *   reference function (src/OFDM.c)                                        served by
 *   void QPSK_Modulator(float complex **in, float complex **out, int n)    :415   ofdm_pack_bits + ofdm_qpsk_modulate
 *   void ifft(float complex *X, float complex *Y, int sz)                  :320   ofdm_ifft64   (sz must be 64; X is left
 *                                                                                  ifft_shift'ed in place, as the reference leaves it)
 *   void fft(float complex *X, float complex *Y, int sz)                   :314   ofdm_fft64
 *   void Transmission_Over_Air(float complex *tx, float complex *out,
 *                              float snr, int len)                         :635   ofdm_awgn_inject_len; the draws come from the
 *                                                                                  host's rand() exactly as gaussian_noise() :622 takes
 *                                                                                  them (two calls per sample, the second one survives
 *                                                                                  in the real part -- SURVEY.md Q1/Q2)
 *   void Channel_Estimation(float complex *frame, float complex *H_est,
 *                           int frame_size)                                :830   ofdm_channel_estimate (lts_off 160: the reference's
 *                                                                                  frame starts with the STS slot)
 *   void AGC_Receiver(float complex **in, float complex **out)             :852   ofdm_agc_slicer  (reads data_frames_number)
 *   void QPSK_Demodulator(float complex **in, float complex **out, int n)  :873   ofdm_qpsk_demodulate + ofdm_unpack_bits
 *
 * Define OFDM_COMPAT_NO_GLOBALS before including if the translation unit already defines `data_frames_number` (:26).
 */
#ifndef OFDM_REF_COMPAT_H
#define OFDM_REF_COMPAT_H

#include <complex.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ofdm_b200.h"

#ifndef OFDM_COMPAT_NO_GLOBALS
static int data_frames_number = 2;       /* OFDM.c:26, set by Data_Generator :439 */
#endif

static ofdm_ctx *ofdm_compat_ctx_(void)
{
    static ofdm_ctx *ctx = NULL;
    if (!ctx) {
        const char *dev = getenv("OFDM_COMPAT_DEVICE");
        int st = ofdm_ctx_create(&ctx, dev ? atoi(dev) : 0);
        if (st != OFDM_OK) { fprintf(stderr, "ofdm_ref_compat: %s\n", ofdm_strerror(st)); exit(1); }
    }
    return ctx;
}

#define OFDM_COMPAT_CHECK_(call)                                                                                  \
    do {                                                                                                          \
        int st_ = (call);                                                                                         \
        if (st_ != OFDM_OK) {                                                                                     \
            fprintf(stderr, "ofdm_ref_compat: %s: %s (%s)\n", #call, ofdm_strerror(st_), ofdm_last_error(ofdm_compat_ctx_())); \
            exit(1);                                                                                              \
        }                                                                                                         \
    } while (0)

/* one stage on host buffers: in (in_bytes) -> device -> fn -> device -> out (out_bytes) */
typedef int (*ofdm_compat_stage_)(ofdm_ctx *, const void *, void *, void *);
static void ofdm_compat_run_(const void *in, size_t in_bytes, void *out, size_t out_bytes, ofdm_compat_stage_ fn, void *arg)
{
    ofdm_ctx *c = ofdm_compat_ctx_();
    void *d_in = NULL, *d_out = NULL;
    OFDM_COMPAT_CHECK_(ofdm_dev_alloc(c, &d_in, in_bytes));
    OFDM_COMPAT_CHECK_(ofdm_dev_alloc(c, &d_out, out_bytes));
    OFDM_COMPAT_CHECK_(ofdm_memcpy_h2d(c, d_in, in, in_bytes));
    OFDM_COMPAT_CHECK_(fn(c, d_in, d_out, arg));
    OFDM_COMPAT_CHECK_(ofdm_memcpy_d2h(c, out, d_out, out_bytes));
    OFDM_COMPAT_CHECK_(ofdm_ctx_sync(c));
    ofdm_dev_free(c, d_in); ofdm_dev_free(c, d_out);
}

/* ---- QPSK_Modulator :415 ---------------------------------------------------------------------------------------- */
static int ofdm_compat_qpsk_(ofdm_ctx *c, const void *bits_u8, void *mod, void *arg)
{
    const long n = *(const long *)arg;
    void *packed = NULL;
    int st = ofdm_dev_alloc(c, &packed, (size_t)n * 12);
    if (st == OFDM_OK) st = ofdm_pack_bits(c, (const uint8_t *)bits_u8, (uint32_t *)packed, n);
    if (st == OFDM_OK) st = ofdm_qpsk_modulate(c, (const uint32_t *)packed, (float *)mod, n);
    if (st == OFDM_OK) st = ofdm_ctx_sync(c);
    if (packed) ofdm_dev_free(c, packed);
    return st;
}
static void QPSK_Modulator(float complex **Data_Payload, float complex **Data_Payload_Mod, int n_frames)
{
    long n = n_frames;
    uint8_t *bits = (uint8_t *)malloc((size_t)n * 96);
    float complex *mod = (float complex *)malloc((size_t)n * 48 * sizeof(float complex));
    if (!bits || !mod) { fprintf(stderr, "Memory allocation failed\n"); exit(1); }
    for (long i = 0; i < n; ++i) for (int j = 0; j < 96; ++j) bits[i * 96 + j] = (uint8_t)((int)crealf(Data_Payload[i][j]) & 1);
    ofdm_compat_run_(bits, (size_t)n * 96, mod, (size_t)n * 48 * sizeof(float complex), ofdm_compat_qpsk_, &n);
    for (long i = 0; i < n; ++i) memcpy(Data_Payload_Mod[i], mod + i * 48, 48 * sizeof(float complex));
    free(bits); free(mod);
}

/* ---- fft :314, ifft :320 ---------------------------------------------------------------------------------------- */
static int ofdm_compat_fft_(ofdm_ctx *c, const void *in, void *out, void *arg)
{
    return *(const int *)arg ? ofdm_ifft64(c, (const float *)in, (float *)out, 1, OFDM_MODE_EXACT) : ofdm_fft64(c, (const float *)in, (float *)out, 1, OFDM_MODE_EXACT);
}
static void fft(float complex *X, float complex *Y, int sz)
{
    int inverse = 0;
    if (sz != 64) { fprintf(stderr, "ofdm_ref_compat: fft size %d (only N_FFT = 64, OFDM.c:11)\n", sz); exit(1); }
    ofdm_compat_run_(X, 64 * sizeof(float complex), Y, 64 * sizeof(float complex), ofdm_compat_fft_, &inverse);
}
static void ifft(float complex *X, float complex *Y, int sz)
{
    int inverse = 1;
    if (sz != 64) { fprintf(stderr, "ofdm_ref_compat: ifft size %d (only N_FFT = 64, OFDM.c:11)\n", sz); exit(1); }
    ofdm_compat_run_(X, 64 * sizeof(float complex), Y, 64 * sizeof(float complex), ofdm_compat_fft_, &inverse);
    for (int i = 0; i < 32; ++i) { float complex t = X[i]; X[i] = X[i + 32]; X[i + 32] = t; }      /* the reference ifft_shift()s X in place (:322) */
}

/* ---- Transmission_Over_Air :635 --------------------------------------------------------------------------------- */
static float ofdm_compat_gaussian_noise_(void)         /* gaussian_noise(0, 1), OFDM.c:622-632, on the host's libc stream */
{
    float u1 = ((float)rand() + 1.0) / ((float)RAND_MAX + 1.0);
    float u2 = ((float)rand() + 1.0) / ((float)RAND_MAX + 1.0);
    float z0 = sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159 * u2);
    return 0 + sqrt((float)1) * z0;
}
static void Transmission_Over_Air(float complex *TX_signal, float complex *TX_OTA_signal, float snr, int len)
{
    ofdm_ctx *c = ofdm_compat_ctx_();
    float *g = (float *)malloc((size_t)len * sizeof(float));
    void *d_tx = NULL, *d_g = NULL, *d_out = NULL;
    if (!g) { fprintf(stderr, "Memory allocation failed\n"); exit(1); }
    for (int i = 0; i < len; ++i) {                       /* :651: two draws per sample, the second survives (gcc evaluation order, Q2) */
        float first = ofdm_compat_gaussian_noise_();
        float second = ofdm_compat_gaussian_noise_();
        (void)first;
        g[i] = second;
    }
    OFDM_COMPAT_CHECK_(ofdm_dev_alloc(c, &d_tx, (size_t)len * 8));
    OFDM_COMPAT_CHECK_(ofdm_dev_alloc(c, &d_g, (size_t)len * 4));
    OFDM_COMPAT_CHECK_(ofdm_dev_alloc(c, &d_out, (size_t)len * 8));
    OFDM_COMPAT_CHECK_(ofdm_memcpy_h2d(c, d_tx, TX_signal, (size_t)len * 8));
    OFDM_COMPAT_CHECK_(ofdm_memcpy_h2d(c, d_g, g, (size_t)len * 4));
    OFDM_COMPAT_CHECK_(ofdm_awgn_inject_len(c, (const float *)d_tx, (const float *)d_g, NULL, snr, (float *)d_out, 1, len, OFDM_MODE_EXACT));
    OFDM_COMPAT_CHECK_(ofdm_memcpy_d2h(c, TX_OTA_signal, d_out, (size_t)len * 8));
    OFDM_COMPAT_CHECK_(ofdm_ctx_sync(c));
    ofdm_dev_free(c, d_tx); ofdm_dev_free(c, d_g); ofdm_dev_free(c, d_out);
    free(g);
}

/* ---- Channel_Estimation :830 ------------------------------------------------------------------------------------ */
static int ofdm_compat_chest_(ofdm_ctx *c, const void *frame, void *H, void *arg)
{
    return ofdm_channel_estimate(c, (const float *)frame, (float *)H, 1, *(const int *)arg, 160, OFDM_MODE_EXACT);
}
static void Channel_Estimation(float complex *rx_frame_after_fine, float complex *H_est, int rx_frame_size)
{
    int len = rx_frame_size < 320 ? 320 : rx_frame_size;           /* the reference ignores its third argument and reads samples 192..319 */
    ofdm_compat_run_(rx_frame_after_fine, 320 * sizeof(float complex), H_est, 64 * sizeof(float complex), ofdm_compat_chest_, &len);
}

/* ---- AGC_Receiver :852, QPSK_Demodulator :873 ------------------------------------------------------------------- */
static int ofdm_compat_agc_(ofdm_ctx *c, const void *in, void *out, void *arg) { return ofdm_agc_slicer(c, (const float *)in, (float *)out, *(const long *)arg); }
static void AGC_Receiver(float complex **Rx_Payload_No_Pilot, float complex **Rx_Payload_Final)
{
    long n = data_frames_number;
    float complex *a = (float complex *)malloc((size_t)n * 48 * sizeof(float complex)), *b = (float complex *)malloc((size_t)n * 48 * sizeof(float complex));
    if (!a || !b) { fprintf(stderr, "Memory allocation failed\n"); exit(1); }
    for (long i = 0; i < n; ++i) memcpy(a + i * 48, Rx_Payload_No_Pilot[i], 48 * sizeof(float complex));
    ofdm_compat_run_(a, (size_t)n * 48 * sizeof(float complex), b, (size_t)n * 48 * sizeof(float complex), ofdm_compat_agc_, &n);
    for (long i = 0; i < n; ++i) memcpy(Rx_Payload_Final[i], b + i * 48, 48 * sizeof(float complex));
    free(a); free(b);
}
static int ofdm_compat_demod_(ofdm_ctx *c, const void *in, void *bits_u8, void *arg)
{
    const long n = *(const long *)arg;
    void *packed = NULL;
    int st = ofdm_dev_alloc(c, &packed, (size_t)n * 12);
    if (st == OFDM_OK) st = ofdm_qpsk_demodulate(c, (const float *)in, (uint32_t *)packed, n);
    if (st == OFDM_OK) st = ofdm_unpack_bits(c, (const uint32_t *)packed, (uint8_t *)bits_u8, n);
    if (st == OFDM_OK) st = ofdm_ctx_sync(c);
    if (packed) ofdm_dev_free(c, packed);
    return st;
}
static void QPSK_Demodulator(float complex **Rx_Payload, float complex **Data_Payload_Demod, int n_frames)
{
    long n = n_frames;
    float complex *a = (float complex *)malloc((size_t)n * 48 * sizeof(float complex));
    uint8_t *bits = (uint8_t *)malloc((size_t)n * 96);
    if (!a || !bits) { fprintf(stderr, "Memory allocation failed\n"); exit(1); }
    for (long i = 0; i < n; ++i) memcpy(a + i * 48, Rx_Payload[i], 48 * sizeof(float complex));
    ofdm_compat_run_(a, (size_t)n * 48 * sizeof(float complex), bits, (size_t)n * 96, ofdm_compat_demod_, &n);
    for (long i = 0; i < n; ++i) for (int j = 0; j < 96; ++j) Data_Payload_Demod[i][j] = bits[i * 96 + j];
    free(a); free(bits);
}

#endif /* OFDM_REF_COMPAT_H */
