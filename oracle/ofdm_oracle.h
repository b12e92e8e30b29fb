/*
 * oracle/ofdm_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement ("port") of the reference's stage chain, written from the behaviour
 * described in SURVEY.md section 8(a) with the reference file:line cited per function in
 * ofdm_oracle.c.  Pinned against the compiled reference itself (oracle/_ref, built from
 * /root/reference/src/OFDM.c) by tests/test_oracle_vs_ref.py and against the committed
 * fixtures in tests/golden/ (generated from oracle/_ref by tests/golden/make_golden.py).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library; the product (libofdm_b200.so) never does.
 *
 * All IQ buffers are interleaved (re, im) float pairs -- the memory layout of C's
 * `float complex`; bits are one byte per bit (0/1).
 */
#ifndef OFDM_ORACLE_H
#define OFDM_ORACLE_H
#include <stdint.h>

typedef struct {            /* same layout as ref_rx_stats in ref_harness.c */
    float evm_lin, evm_db, evm_agc_lin, evm_agc_db, ber;
    int bit_errors, rail_errors;
} orc_rx_stats;

typedef struct {            /* same layout as ref_counters in ref_harness.c */
    uint64_t bit_errors, bits, frames_in_error, rail_errors, frames;
    double sum_err2, sum_ref2, sum_evm_lin;
} orc_counters;

int   orc_init(void);
void  orc_twiddles(double *out /* 32 x (re,im) */);
void  orc_lts_freq(float *out /* 64 x 2 */);
void  orc_lts_time(float *out /* 160 x 2 */);
void  orc_qpsk_mod(const uint8_t *bits, int n_sym, float *out /* n_sym*48*2 */);
void  orc_map_grid(const float *mod, int n_sym, float *grid /* n_sym*64*2 */);
void  orc_ifft64(const float *in, float *out);
void  orc_fft64(const float *in, float *out);
void  orc_tx_frame(const uint8_t *bits, int n_sym, float *frame /* (160+80*n_sym)*2 */);
float orc_frame_power(const float *tx, int len);
void  orc_awgn_inject(const float *tx, const float *g, float *out, float snr_db, int len);
void  orc_rx_frame(const float *ota, int n_sym, const uint8_t *tx_bits, float *H_out, float *eq_out,
                   float *sliced_out, uint8_t *bits_out, orc_rx_stats *st);
void  orc_chain(const uint8_t *bits, const float *g, long n_frames, int n_sym, float snr_db, int noise_mode,
                orc_counters *acc, int *frame_bit_errors, float *frame_evm_lin);

void  orc_chain_sweep(const uint8_t *bits, const float *g, long n_frames, int n_sym, const float *snr_db, int n_snr,
                      int noise_mode, orc_counters *acc /* [n_snr] */);

/* Counter-based streams shared with the CUDA Monte-Carlo kernels (new-build definition,
 * see DESIGN.md "Philox streams"): Philox4x32-10, key = (seed, stream),
 * counter = (frame_lo, frame_hi, block, domain). */
void  orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
void  orc_philox_bits(uint32_t seed, uint64_t frame0, long n_frames, int n_sym, uint8_t *bits);
void  orc_philox_normals(uint32_t seed, uint32_t stream, uint64_t frame0, long n_frames, int len, float *g);
void  orc_philox_taps(uint32_t seed, uint64_t frame0, long n_frames, int n_taps, float *taps /* n_frames*n_taps*2 */);
void  orc_apply_taps(const float *tx, const float *taps, int n_taps, float *out, int len);
void  orc_chain_multipath(const uint8_t *bits, const float *g, const float *taps, int n_taps, long n_frames,
                          int n_sym, float snr_db, orc_counters *acc, int *frame_bit_errors, float *frame_evm_lin);
void  orc_rrc_taps(float *out21);
void  orc_rrc_tx(const float *frame, int len, float *out /* (2*len+20)*2 */);
void  orc_rrc_rx(const float *in, int in_len, int packet_idx, int frame_len, float *out /* frame_len*2 */);
void  orc_packet_detection(const float *rx, int len, float *corr_out /* (len-47)*2 */);
int   orc_packet_selection(const float *corr, int len_corr);
void  orc_sts_time(float *out160);
void  orc_cfo_coarse(const float *rx, int len, float *out);
void  orc_cfo_fine(const float *rx, int len, float *out);
#endif
