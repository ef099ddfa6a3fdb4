/* Plain-C float64 restatement of the reference's IIR recursions -- TEST INFRASTRUCTURE
 * ONLY (see oracle/__init__.py).  Built by oracle/Makefile into oracle/_build/.
 *
 * frt_oracle_lfilter  : direct-form-II-transposed loop, friture/signal/lfilter.py:131-139,
 *                       same operation order ((z[n+1] + x*b[n+1]) - y*a[n+1]); compiled with
 *                       -ffp-contract=off so no FMA changes the rounding.
 * frt_oracle_bank     : multi-rate 1/N-octave bank with decimation, friture/filter.py:86-118
 *                       (+ friture/signal/decimate.py:39-41 for the [::2]), with the widget's
 *                       x^2 -> exponential smoothing (friture/signal/exp_smoothing.py:11-56 in
 *                       its recursive form s <- alpha*x + (1-alpha)*s) fused behind it, for
 *                       many channels (used as the fast multi-core CPU baseline).
 */
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

void frt_oracle_lfilter(const double *b, const double *a, int nb,
                        const double *x, long n, double *z, double *y)
{
    for (long k = 0; k < n; k++) {
        double xk = x[k];
        double yk = z[0] + b[0] * xk;
        y[k] = yk;
        for (int i = 0; i < nb - 2; i++)
            z[i] = z[i + 1] + xk * b[i + 1] - yk * a[i + 1];
        z[nb - 2] = xk * b[nb - 1] - yk * a[nb - 1];
    }
}

/* One channel, one block.  State layout `zis` = per stage: band bpo-1 .. band 0
 * (nb_band-1 doubles each) then decimator (nb_dec-1 doubles): friture/filter.py:121-133.
 * y_out (optional) receives the ragged band outputs concatenated in band order k=0..nbands-1
 * (k = (noctave-1-j)*bpo + i, length n>>j); energies[nbands] are the smoothed squares
 * (in/out: previous value in, new value out), alphas[nbands] per band.  */
int frt_oracle_bank_block(const double *bdec, const double *adec, int nb_dec,
                          const double *boct, const double *aoct, int nb_band, int bpo,
                          int noctave, const double *x, long n, double *zis,
                          const double *alphas, double *energies, double *y_out)
{
    if (n % (1L << (noctave - 1)) != 0) return -1;
    double *cur = (double *)malloc(sizeof(double) * (size_t)n);
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    if (!cur || !tmp) { free(cur); free(tmp); return -2; }
    memcpy(cur, x, sizeof(double) * (size_t)n);
    int nbands = noctave * bpo;
    /* offsets of each band in the concatenated y_out */
    long len = n;
    double *z = zis;
    for (int j = 0; j < noctave; j++) {
        for (int i = bpo - 1; i >= 0; i--) {
            int k = (noctave - 1 - j) * bpo + i;
            frt_oracle_lfilter(boct + (size_t)i * nb_band, aoct + (size_t)i * nb_band, nb_band,
                               cur, len, z, tmp);
            z += nb_band - 1;
            double alpha = alphas[k], e = energies[k], om = 1.0 - alpha;
            for (long t = 0; t < len; t++) e = alpha * (tmp[t] * tmp[t]) + om * e;
            energies[k] = e;
            if (y_out) {
                long off = 0;
                for (int kk = 0; kk < k; kk++) off += n >> (noctave - 1 - kk / bpo);
                memcpy(y_out + off, tmp, sizeof(double) * (size_t)len);
            }
        }
        frt_oracle_lfilter(bdec, adec, nb_dec, cur, len, z, tmp);
        z += nb_dec - 1;
        for (long t = 0; t < len / 2; t++) cur[t] = tmp[2 * t];
        len /= 2;
    }
    (void)nbands;
    free(cur); free(tmp);
    return 0;
}

/* Many channels, many blocks: x[c*x_stride + blk*n ...]; zis[c][...]; energies[c][nbands]
 * carried across blocks; energies_out[c][nblocks][nbands] gets the value after each block. */
int frt_oracle_bank_stream(const double *bdec, const double *adec, int nb_dec,
                           const double *boct, const double *aoct, int nb_band, int bpo,
                           int noctave, const float *x, long x_stride, int c0, int c1,
                           long n, int nblocks, double *zis, long zis_stride,
                           const double *alphas, double *energies, double *energies_out)
{
    int nbands = noctave * bpo;
    double *xb = (double *)malloc(sizeof(double) * (size_t)n);
    if (!xb) return -2;
    for (int c = c0; c < c1; c++) {
        for (int blk = 0; blk < nblocks; blk++) {
            const float *xs = x + (size_t)c * x_stride + (size_t)blk * n;
            for (long t = 0; t < n; t++) xb[t] = (double)xs[t];
            int rc = frt_oracle_bank_block(bdec, adec, nb_dec, boct, aoct, nb_band, bpo, noctave,
                                           xb, n, zis + (size_t)c * zis_stride, alphas,
                                           energies + (size_t)c * nbands, NULL);
            if (rc) { free(xb); return rc; }
            if (energies_out)
                memcpy(energies_out + ((size_t)c * nblocks + blk) * nbands,
                       energies + (size_t)c * nbands, sizeof(double) * nbands);
        }
    }
    free(xb);
    return 0;
}
