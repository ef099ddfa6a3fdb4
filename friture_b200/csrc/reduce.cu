// Widget-level reductions behind the spectrum widget: exponential smoothing across
// the frames of a tick, weighting, dB, arg-max and harmonic product spectrum, fused in one pass
// over the power columns.  Stands behind friture/spectrum.py:158-181 (exp_smoothed_value_2d,
// friture/signal/exp_smoothing.py:59-107; log_spectrogram spectrum.py:95-101;
// harmonic_product_spectrum spectrum.py:103-123).
#include "frt_internal.cuh"

namespace {

__device__ __forceinline__ float lg2_fast(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ void argmax_merge(float &best, int &besti, float ob, int oi) {
    if (ob > best || (ob == best && oi < besti)) {
        best = ob;
        besti = oi;
    }
}

__device__ int block_argmax(float best, int besti, float *s_val, int *s_idx) {
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
        argmax_merge(best, besti, ob, oi);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) {
        s_val[w] = best;
        s_idx[w] = besti;
    }
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        best = l < nw ? s_val[l] : -INFINITY;
        besti = l < nw ? s_idx[l] : 0x7fffffff;
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            argmax_merge(best, besti, ob, oi);
        }
        if (l == 0) s_idx[32] = besti;
    }
    __syncthreads();
    return s_idx[32];
}

__global__ void __launch_bounds__(256)
spectrum_reduce_kernel(const float *__restrict__ power, long long stride_c, long long stride_f,
                       int n_frames, int nbins, float alpha, float *__restrict__ disp,
                       const float *__restrict__ weight, float *__restrict__ db,
                       int *__restrict__ fmax_idx, int *__restrict__ pitch_idx) {
    extern __shared__ float s_sp[];   // [nbins] smoothed power of this channel
    __shared__ float s_val[32];
    __shared__ int s_idx[33];
    const int c = blockIdx.x;
    const float *p = power + (size_t)c * stride_c;
    float *d = disp + (size_t)c * nbins;
    const float om = 1.0f - alpha;
    float best = -INFINITY;
    int besti = 0x7fffffff;
    for (int k = threadIdx.x; k < nbins; k += blockDim.x) {
        float s = d[k];
        for (int f = 0; f < n_frames; f++)   // s <- alpha x + (1-alpha) s  (exp_smoothing.py:59-107)
            s = fmaf(alpha, __ldg(p + (size_t)f * stride_f + k), om * s);
        d[k] = s;
        s_sp[k] = s;
        float v = 3.01029995663981195f * lg2_fast(s + 1e-30f);   // spectrum.py:95-101
        if (weight) v += __ldg(weight + k);
        db[(size_t)c * nbins + k] = v;
        if (v > best) {
            best = v;
            besti = k;
        }
    }
    const int imax = block_argmax(best, besti, s_val, s_idx);      // spectrum.py:175
    // harmonic product spectrum, 3 harmonics (spectrum.py:103-123)
    const int h = nbins / 3;
    best = -INFINITY;
    besti = 0x7fffffff;
    for (int k = threadIdx.x; k < h; k += blockDim.x) {
        const float r = s_sp[k] * s_sp[2 * k] * s_sp[3 * k];
        if (r > best) {
            best = r;
            besti = k;
        }
    }
    const int ipitch = block_argmax(best, besti, s_val, s_idx);    // spectrum.py:180
    if (threadIdx.x == 0) {
        fmax_idx[c] = imax;
        pitch_idx[c] = ipitch;
    }
}

}   // namespace

extern "C" int frt_spectrum_reduce(frt_handle h, const float *power_dev, int64_t stride_c,
                                   int64_t stride_f, int n_channels, int n_frames, int nbins,
                                   float alpha, float *disp_dev, const float *weight_dev,
                                   float *db_dev, int *fmax_idx_dev, int *pitch_idx_dev,
                                   void *stream) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_channels >= 0 && n_frames >= 0 && nbins >= 3, "bad shape");
    FRT_CHECK_ARG(h, n_frames <= 8192, "more frames than the reference's 8192-tap kernel");
    FRT_CHECK_ARG(h, alpha > 0.f && alpha <= 1.f, "alpha must be in (0, 1]");
    if (n_channels == 0) return FRT_OK;
    FRT_CHECK_ARG(h, (n_frames == 0 || power_dev) && disp_dev && db_dev && fmax_idx_dev &&
                         pitch_idx_dev, "NULL buffer");
    FRT_CHECK_ARG(h, sizeof(float) * (size_t)nbins <= 200 * 1024, "nbins too large");
    const size_t smem = sizeof(float) * (size_t)nbins;
    if (smem > 48 * 1024)
        FRT_CUDA(h, cudaFuncSetAttribute(spectrum_reduce_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    spectrum_reduce_kernel<<<n_channels, 256, smem, (cudaStream_t)stream>>>(
        power_dev, stride_c, stride_f, n_frames, nbins, alpha, disp_dev, weight_dev, db_dev,
        fmax_idx_dev, pitch_idx_dev);
    h->launches++;
    FRT_CUDA(h, cudaGetLastError());
    return FRT_OK;
}
