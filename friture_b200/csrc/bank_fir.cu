// The reference's LIVE filterbank numerics for sm_100a: 512-tap minimum-phase FIR approximations of
// the fractional-octave and decimation filters, which Octave_Filters.filter() runs by FFT
// overlap-add (friture/octavefilters.py:49-58,123-158; friture/filter.py:136-247).  Overlap-add
// with a tail buffer IS an exact linear convolution, so the kernel computes the convolutions
// directly, carrying the last L-1 input samples of every stage instead of the L-1 pending output
// samples of every filter:
//     stage j:  y_i = (x_j * h_i)[0:N_j]  (i = bpo-1 .. 0),   x_{j+1} = (x_j * h_dec)[0:N_j:2]
// This is the parity mode for callers that want the live path's exact output (it differs from
// the IIR designs it approximates by ~5e-4 on band energy); the throughput path is the IIR bank
// of bank.cu / bank_pipe.cu.  One CTA per channel, the taps and the stage signal in shared memory,
// double-precision accumulation (512-term sums with heavy cancellation in the stop bands).
#include <cmath>

#include "frt_internal.cuh"

namespace {

constexpr int FIR_MAX_OCT = 10;
constexpr int FIR_THREADS = 256;
constexpr int FIR_MAX_BLOCK = 4096;

struct FirArgs {
    const float *x;
    long long x_stride;
    int n_samples;           // per channel in this call
    int L;                   // taps
    int bpo, n_oct;
    const double *taps;      // [bpo + 1][L]: bands 0..bpo-1, then the decimator
    float *hist;             // [C][n_oct][L-1] last inputs of every stage
    float *y;                // ragged band outputs: channel c at y + c*y_stride, bands k = 0..nbands-1
    long long y_stride;      //   concatenated, band k = (n_oct-1-j)*bpo + i holding n_samples >> j samples
};

__global__ void __launch_bounds__(FIR_THREADS) fir_bank_kernel(const FirArgs a) {
    extern __shared__ double smem_d[];
    const int L = a.L, Lm1 = a.L - 1, nf = a.bpo + 1;
    double *s_taps = smem_d;                                        // [nf][L]
    float *s_x = reinterpret_cast<float *>(s_taps + (size_t)nf * L);   // [Lm1 + n_samples]
    float *s_next = s_x + Lm1 + a.n_samples;                        // [Lm1 + n_samples/2]
    const int c = blockIdx.x, tid = threadIdx.x;
    for (int i = tid; i < nf * L; i += FIR_THREADS) s_taps[i] = a.taps[i];
    float *hist = a.hist + (size_t)c * a.n_oct * Lm1;
    for (int i = tid; i < Lm1; i += FIR_THREADS) s_x[i] = hist[i];
    const float *xc = a.x + (size_t)c * a.x_stride;
    for (int i = tid; i < a.n_samples; i += FIR_THREADS) s_x[Lm1 + i] = xc[i];
    __syncthreads();
    float *yc = a.y + (size_t)c * a.y_stride;
    int N = a.n_samples;
    for (int j = 0; j < a.n_oct; j++) {
        // offset of band (j, i = 0) in the ragged layout: stages j' > j first (they are shorter)
        long long off = 0;
        for (int jj = a.n_oct - 1; jj > j; jj--) off += (long long)a.bpo * (a.n_samples >> jj);
        const bool last = (j + 1 == a.n_oct);
        if (!last)
            for (int i = tid; i < Lm1; i += FIR_THREADS) s_next[i] = hist[(size_t)(j + 1) * Lm1 + i];
        for (int n = tid; n < N; n += FIR_THREADS) {
            const float *xp = s_x + Lm1 + n;          // xp[-k] = x_j[n - k]
            for (int f0 = 0; f0 < nf; f0 += 4) {
                double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
                const double *t0 = s_taps + (size_t)f0 * L;
                const int nv = nf - f0;
                for (int k = 0; k < L; k++) {
                    const double xv = (double)xp[-k];
                    acc0 = fma(t0[k], xv, acc0);
                    if (nv > 1) acc1 = fma(t0[L + k], xv, acc1);
                    if (nv > 2) acc2 = fma(t0[2 * L + k], xv, acc2);
                    if (nv > 3) acc3 = fma(t0[3 * L + k], xv, acc3);
                }
                const double accs[4] = {acc0, acc1, acc2, acc3};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const int f = f0 + q;
                    if (f < a.bpo) yc[off + (long long)f * N + n] = (float)accs[q];
                    else if (f == a.bpo && !last && !(n & 1)) s_next[Lm1 + (n >> 1)] = (float)accs[q];
                }
            }
        }
        __syncthreads();
        // this stage's new history = its last L-1 inputs
        for (int i = tid; i < Lm1; i += FIR_THREADS) hist[(size_t)j * Lm1 + i] = s_x[N + i];
        __syncthreads();
        if (last) break;
        float *t = s_x;
        s_x = s_next;
        s_next = t;
        N >>= 1;
    }
}

}   // namespace

struct FirPlan {
    int n_channels = 0, bpo = 0, n_oct = 0, L = 0;
    double *taps = nullptr;
    float *hist = nullptr;
};

void frt_fir_release(frt_ctx *h) {
    FirPlan *pl = reinterpret_cast<FirPlan *>(h->fir);
    if (!pl) return;
    if (pl->taps) cudaFree(pl->taps);
    if (pl->hist) cudaFree(pl->hist);
    delete pl;
    h->fir = nullptr;
}

extern "C" int frt_firbank_plan(frt_handle h, int n_channels, int bands_per_octave, int n_octaves,
                                int n_taps, const double *fir_band, const double *fir_dec) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_channels >= 1, "n_channels must be >= 1");
    FRT_CHECK_ARG(h, bands_per_octave >= 1 && bands_per_octave <= 24, "bands_per_octave must be in [1, 24]");
    FRT_CHECK_ARG(h, n_octaves >= 1 && n_octaves <= FIR_MAX_OCT, "n_octaves must be in [1, 10]");
    FRT_CHECK_ARG(h, n_taps >= 2 && n_taps <= 1024, "n_taps must be in [2, 1024]");
    FRT_CHECK_ARG(h, fir_band && fir_dec, "NULL tap table");
    frt_fir_release(h);
    FirPlan *pl = new (std::nothrow) FirPlan();
    if (!pl) return frt_fail(h, FRT_ENOMEM, "out of host memory");
    pl->n_channels = n_channels;
    pl->bpo = bands_per_octave;
    pl->n_oct = n_octaves;
    pl->L = n_taps;
    const size_t nt = (size_t)(bands_per_octave + 1) * n_taps;
    const size_t nh = (size_t)n_channels * n_octaves * (n_taps - 1);
    cudaError_t e = cudaMalloc(&pl->taps, sizeof(double) * nt);
    if (e == cudaSuccess) e = cudaMalloc(&pl->hist, sizeof(float) * nh);
    if (e == cudaSuccess)
        e = cudaMemcpy(pl->taps, fir_band, sizeof(double) * (size_t)bands_per_octave * n_taps, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(pl->taps + (size_t)bands_per_octave * n_taps, fir_dec, sizeof(double) * n_taps,
                       cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemset(pl->hist, 0, sizeof(float) * nh);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (pl->taps) cudaFree(pl->taps);
        if (pl->hist) cudaFree(pl->hist);
        delete pl;
        return frt_fail(h, FRT_ECUDA, "frt_firbank_plan: %s", cudaGetErrorString(e));
    }
    h->fir = pl;
    return FRT_OK;
}

extern "C" int frt_firbank_reset(frt_handle h) {
    if (!h) return FRT_EINVAL;
    FirPlan *pl = reinterpret_cast<FirPlan *>(h->fir);
    if (!pl) return frt_fail(h, FRT_ESTATE, "frt_firbank_reset: no plan");
    DeviceGuard g(h->device);
    FRT_CUDA(h, cudaMemset(pl->hist, 0, sizeof(float) * (size_t)pl->n_channels * pl->n_oct * (pl->L - 1)));
    FRT_CUDA(h, cudaDeviceSynchronize());
    return FRT_OK;
}

extern "C" int frt_firbank_process(frt_handle h, const float *x_dev, int64_t x_stride, int n_samples,
                                   float *y_dev, int64_t y_stride, void *stream) {
    if (!h) return FRT_EINVAL;
    FirPlan *pl = reinterpret_cast<FirPlan *>(h->fir);
    if (!pl) return frt_fail(h, FRT_ESTATE, "frt_firbank_process: call frt_firbank_plan first");
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_samples >= 0, "negative length");
    if (n_samples == 0) return FRT_OK;
    FRT_CHECK_ARG(h, n_samples <= FIR_MAX_BLOCK, "n_samples must be <= 4096");
    FRT_CHECK_ARG(h, n_samples % (1 << (pl->n_oct - 1)) == 0,
                  "n_samples must be a multiple of 2^(n_octaves-1) (every stage keeps its even samples)");
    FRT_CHECK_ARG(h, x_dev && y_dev, "NULL buffer");
    FRT_CHECK_ARG(h, x_stride >= n_samples, "x_stride smaller than n_samples");
    long long need = 0;
    for (int j = 0; j < pl->n_oct; j++) need += (long long)pl->bpo * (n_samples >> j);
    FRT_CHECK_ARG(h, y_stride >= need, "y_stride smaller than the ragged output size");
    FirArgs a;
    a.x = x_dev;
    a.x_stride = x_stride;
    a.n_samples = n_samples;
    a.L = pl->L;
    a.bpo = pl->bpo;
    a.n_oct = pl->n_oct;
    a.taps = pl->taps;
    a.hist = pl->hist;
    a.y = y_dev;
    a.y_stride = y_stride;
    const size_t smem = sizeof(double) * (size_t)(pl->bpo + 1) * pl->L +
                        sizeof(float) * (2 * (size_t)(pl->L - 1) + n_samples + n_samples / 2 + 8);
    FRT_CUDA(h, cudaFuncSetAttribute(fir_bank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    fir_bank_kernel<<<pl->n_channels, FIR_THREADS, smem, (cudaStream_t)stream>>>(a);
    h->launches++;
    FRT_CUDA(h, cudaGetLastError());
    return FRT_OK;
}
