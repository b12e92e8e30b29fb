/*
 * ofdm_io.c -- host-side result / dump writers of libofdm_b200.so (plain C, no CUDA).
 * File formats are the contract with the reference's scripts (scripts/OFDM_Plotting.py,
 * compare_double.py, compare_complex.py), which must keep working unchanged.
 */
#include <math.h>
#include <stdio.h>

#include "../../include/ofdm_b200.h"

/* write_float_array_to_file, OFDM.c:123-143: "%.2e" values, tab-separated, one trailing newline */
int ofdm_write_float_array_to_file(const float *a, int n, const char *fname)
{
    if (!fname || n < 0 || (n > 0 && !a)) return OFDM_ERR_INVALID;
    FILE *fp = fopen(fname, "w");
    if (!fp) return OFDM_ERR_IO;                          /* reference: perror + return (:126-129) */
    for (int i = 0; i < n; ++i) {
        fprintf(fp, "%.2e", a[i]);
        if (i < n - 1) fputc('\t', fp);
    }
    fputc('\n', fp);
    return fclose(fp) == 0 ? OFDM_OK : OFDM_ERR_IO;
}

/* write_complex_array_to_file, OFDM.c:94-121.  Elements with an infinite part are skipped together
 * with their separator, as the reference does (:103, :113-116). */
int ofdm_write_complex_array_to_file(const float *a_iq, int n, const char *fname, int format)
{
    if (!fname || n < 0 || (n > 0 && !a_iq) || (format != 0 && format != 1)) return OFDM_ERR_INVALID;
    FILE *fp = fopen(fname, "w");
    if (!fp) return OFDM_ERR_IO;
    for (int i = 0; i < n; ++i) {
        double re = a_iq[2 * i], im = a_iq[2 * i + 1];
        if (isinf(re) || isinf(im)) continue;
        if (format == 0) fprintf(fp, "%.15e", re);                   /* live line :107 */
        else fprintf(fp, "%.15e + %.15ei", re, im);                  /* commented-out line :105 */
        if (i < n - 1) fputc('\t', fp);
    }
    fputc('\n', fp);
    return fclose(fp) == 0 ? OFDM_OK : OFDM_ERR_IO;
}
