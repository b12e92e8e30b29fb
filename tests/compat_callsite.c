/* A call site written the way the reference's main() / Transmitter() / Receiver() call their stage functions
 * (row-pointer float complex arrays, globals, srand + rand), compiled against include/ofdm_ref_compat.h instead of
 * src/OFDM.c.  It writes every stage output to a binary file which tests/test_ref_compat.py compares with the oracle.
 * usage: compat_callsite <out.bin> <seed> <snr_db> */
#include "ofdm_ref_compat.h"

static float complex **alloc2(int r, int c)
{
    float complex **a = (float complex **)calloc((size_t)r, sizeof(float complex *));
    for (int i = 0; i < r; ++i) a[i] = (float complex *)calloc((size_t)c, sizeof(float complex));
    return a;
}

int main(int argc, char **argv)
{
    if (argc < 4) return 2;
    FILE *f = fopen(argv[1], "wb");
    unsigned seed = (unsigned)strtoul(argv[2], NULL, 10);
    float snr = (float)atof(argv[3]);
    if (!f) return 1;
    data_frames_number = 2;
    float complex **Data = alloc2(2, 96), **Mod = alloc2(2, 48);
    srand(seed);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 96; ++j) Data[i][j] = rand() & 1;
    QPSK_Modulator(Data, Mod, data_frames_number);                                   /* OFDM.c:517 */
    for (int i = 0; i < 2; ++i) { fwrite(Data[i], sizeof(float complex), 96, f); }
    for (int i = 0; i < 2; ++i) { fwrite(Mod[i], sizeof(float complex), 48, f); }

    float complex X[64], Y[64], Z[64];
    for (int i = 0; i < 64; ++i) X[i] = (float)(rand() % 2001 - 1000) / 1000.0f + I * ((float)(rand() % 2001 - 1000) / 1000.0f);
    fwrite(X, sizeof(float complex), 64, f);
    fft(X, Y, 64);                                                                   /* :314 */
    fwrite(Y, sizeof(float complex), 64, f);
    ifft(X, Z, 64);                                                                  /* :320 (X comes back ifft_shift'ed) */
    fwrite(Z, sizeof(float complex), 64, f);
    fwrite(X, sizeof(float complex), 64, f);

    enum { LEN = 480 };
    float complex tx[LEN], ota[LEN], H[64];
    for (int i = 0; i < LEN; ++i) tx[i] = (float)(rand() % 2001 - 1000) / 3000.0f + I * ((float)(rand() % 2001 - 1000) / 3000.0f);
    fwrite(tx, sizeof(float complex), LEN, f);
    srand(seed + 1);
    Transmission_Over_Air(tx, ota, snr, LEN);                                        /* :1208 */
    fwrite(ota, sizeof(float complex), LEN, f);
    Channel_Estimation(ota, H, LEN);                                                 /* :1020 */
    fwrite(H, sizeof(float complex), 64, f);

    float complex **NoPilot = alloc2(2, 48), **Final = alloc2(2, 48), **Demod = alloc2(2, 96);
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 48; ++j) NoPilot[i][j] = ota[100 + 48 * i + j] * 3.0f;
    NoPilot[0][0] = 0; NoPilot[1][5] = -0.25f;                                       /* a zero rail goes negative (:860-868) */
    AGC_Receiver(NoPilot, Final);                                                    /* :1077 */
    QPSK_Demodulator(Final, Demod, data_frames_number);                              /* :1083 */
    for (int i = 0; i < 2; ++i) fwrite(NoPilot[i], sizeof(float complex), 48, f);
    for (int i = 0; i < 2; ++i) fwrite(Final[i], sizeof(float complex), 48, f);
    for (int i = 0; i < 2; ++i) fwrite(Demod[i], sizeof(float complex), 96, f);
    fclose(f);
    printf("compat call site ok\n");
    return 0;
}
