// GCC-PHAT delay estimation for sm_100a: one CTA per channel pair, everything for a pair stays in
// shared memory.  Stands behind generalized_cross_correlation (friture/signal/correlation.py:24-43)
// and the smoothing + peak pick of the delay estimator (friture/delay_estimator.py:129-142).
//
//   a = (d0 - mean(d0)) * hanning(L),  b = (d1 - mean(d1)) * hanning(L)        correlation.py:27-31
//   Z = FFT_L(a + j b)  ->  D0[k] = (Z[k] + conj Z[L-k]) / 2,  D1[k] = (Z[k] - conj Z[L-k]) / 2j
//   G = conj(D0) D1,  W = 1 / (1e-10 max|G| + |G|),  X = irfft(W G)            correlation.py:34-41
//   Xs = 0.3 X + 0.7 Xs_prev (when a previous frame exists), i = argmax |Xs|   delay_estimator.py:134-142
//
// The two real transforms share ONE complex FFT of length L (L = 24 000 = 2^6 * 3 * 5^3 for the
// default one-second range at 12 kHz, delay_estimator.py:53-54,114-117); the inverse real
// transform is a second complex FFT of the conjugated Hermitian spectrum.  The FFT is an in-place
// mixed-radix (4/2/3/5) decimation-in-frequency over the 8*L-byte shared-memory buffer; results
// are addressed through the mixed-radix digit reversal instead of being permuted.
#include <cmath>

#include "frt_internal.cuh"

namespace {

constexpr int GCC_THREADS = 512;
constexpr int GCC_MAX_PASSES = 16;
constexpr int GCC_MAX_PER_THREAD = 28;   // (L/2+1) / GCC_THREADS rounded up must not exceed this

struct GccRadices {
    int n_passes;
    int radix[GCC_MAX_PASSES];
    unsigned magic[GCC_MAX_PASSES];   // ceil(2^32 / n_next) of the pass: idx / n_next == umulhi(idx, magic)
};

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    return make_float2(fmaf(a.x, w.x, -a.y * w.y), fmaf(a.x, w.y, a.y * w.x));
}
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }   // * (-j)
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }

template <int R> __device__ __forceinline__ void dft(float2 (&v)[R]);

template <> __device__ __forceinline__ void dft<2>(float2 (&v)[2]) {
    const float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
}
template <> __device__ __forceinline__ void dft<4>(float2 (&v)[4]) {
    const float2 s02 = cadd(v[0], v[2]), d02 = csub(v[0], v[2]);
    const float2 s13 = cadd(v[1], v[3]), d13 = mul_mj(csub(v[1], v[3]));
    v[0] = cadd(s02, s13);
    v[1] = cadd(d02, d13);
    v[2] = csub(s02, s13);
    v[3] = csub(d02, d13);
}
template <> __device__ __forceinline__ void dft<3>(float2 (&v)[3]) {
    const float s = 0.86602540378443864676f;
    const float2 t = cadd(v[1], v[2]);
    const float2 m1 = make_float2(fmaf(-0.5f, t.x, v[0].x), fmaf(-0.5f, t.y, v[0].y));
    const float2 m2 = cscale(mul_mj(csub(v[1], v[2])), s);
    v[0] = cadd(v[0], t);
    v[1] = cadd(m1, m2);
    v[2] = csub(m1, m2);
}
template <> __device__ __forceinline__ void dft<5>(float2 (&v)[5]) {
    const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;
    const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;
    const float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    const float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    const float2 t1 = make_float2(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
    const float2 t2 = make_float2(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
    const float2 u1 = mul_mj(make_float2(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y));
    const float2 u2 = mul_mj(make_float2(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y));
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(t1, u1);
    v[4] = csub(t1, u1);
    v[2] = cadd(t2, u2);
    v[3] = csub(t2, u2);
}

#include "unit_roots.inc"

// Composite in-register DFT of R = R1*R2 points: R2 DFTs of size R1 over the stride-R2 sub-
// sequences, twiddles W_R^(q2 k1) (compile-time constants), R1 DFTs of size R2;
// V[k1 + R1 k2] = sum_q v[q] W_R^(q k), q = q1 R2 + q2.
template <int R1, int R2>
__device__ __forceinline__ void dft_comp(float2 (&v)[R1 * R2]) {
    constexpr int R = R1 * R2;
#pragma unroll
    for (int q2 = 0; q2 < R2; q2++) {
        float2 t[R1];
#pragma unroll
        for (int q1 = 0; q1 < R1; q1++) t[q1] = v[q1 * R2 + q2];
        dft<R1>(t);
#pragma unroll
        for (int k1 = 0; k1 < R1; k1++) {
            const int m = q2 * k1;
            v[k1 * R2 + q2] = (m == 0) ? t[k1]
                                       : cmul(t[k1], make_float2(UnitRoots<R>::c(m), -UnitRoots<R>::s(m)));
        }
    }
    float2 o[R];
#pragma unroll
    for (int k1 = 0; k1 < R1; k1++) {
        float2 t[R2];
#pragma unroll
        for (int q2 = 0; q2 < R2; q2++) t[q2] = v[k1 * R2 + q2];
        dft<R2>(t);
#pragma unroll
        for (int k2 = 0; k2 < R2; k2++) o[k1 + R1 * k2] = t[k2];
    }
#pragma unroll
    for (int k = 0; k < R; k++) v[k] = o[k];
}
template <> __device__ __forceinline__ void dft<16>(float2 (&v)[16]) { dft_comp<4, 4>(v); }
template <> __device__ __forceinline__ void dft<25>(float2 (&v)[25]) { dft_comp<5, 5>(v); }
template <> __device__ __forceinline__ void dft<15>(float2 (&v)[15]) { dft_comp<3, 5>(v); }

// One in-place DIF pass of radix R on sub-transforms of length n (n_next = n / R).
// FIRST: the pass also applies (x - mean) * hanning on its loads (correlation.py:27-31), so the
// windowing costs no extra sweep over shared memory.
template <int R, bool FIRST>
__device__ __forceinline__ void dif_pass(float2 *z, int L, int n, unsigned magic,
                                         const float2 *__restrict__ tw,
                                         const float *__restrict__ window, float m0, float m1) {
    const int n_next = n / R;
    const int tw_step = L / n;
    const int total = L / R;
    for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
        const int blk = (n_next == 1) ? idx : (int)__umulhi((unsigned)idx, magic);
        const int j = idx - blk * n_next;
        float2 *p = z + blk * n + j;
        float2 v[R];
#pragma unroll
        for (int q = 0; q < R; q++) {
            v[q] = p[q * n_next];
            if (FIRST) {
                const float w = __ldg(window + blk * n + j + q * n_next);
                v[q] = make_float2((v[q].x - m0) * w, (v[q].y - m1) * w);
            }
        }
        dft<R>(v);
        p[0] = v[0];
        // W_n^(j q), q = 1..R-1: one table read, the rest by successive products
        const float2 w1 = __ldg(tw + (size_t)j * tw_step);
        float2 wq = w1;
        p[n_next] = cmul(v[1], w1);
#pragma unroll
        for (int q = 2; q < R; q++) {
            wq = cmul(wq, w1);
            p[q * n_next] = cmul(v[q], wq);
        }
    }
}

template <bool FIRST>
__device__ __forceinline__ void run_pass(int r, float2 *z, int L, int n, unsigned magic,
                                         const float2 *__restrict__ tw,
                                         const float *__restrict__ window, float m0, float m1) {
    if (r == 16) dif_pass<16, FIRST>(z, L, n, magic, tw, window, m0, m1);
    else if (r == 25) dif_pass<25, FIRST>(z, L, n, magic, tw, window, m0, m1);
    else if (r == 15) dif_pass<15, FIRST>(z, L, n, magic, tw, window, m0, m1);
    else if (r == 4) dif_pass<4, FIRST>(z, L, n, magic, tw, window, m0, m1);
    else if (r == 5) dif_pass<5, FIRST>(z, L, n, magic, tw, window, m0, m1);
    else if (r == 3) dif_pass<3, FIRST>(z, L, n, magic, tw, window, m0, m1);
    else dif_pass<2, FIRST>(z, L, n, magic, tw, window, m0, m1);
}

// In-place FFT; when `window` is given the first pass also removes the means and applies it.
__device__ __noinline__ void fft_inplace(float2 *z, int L, const GccRadices &rd,
                                         const float2 *__restrict__ tw,
                                         const float *__restrict__ window, float m0, float m1) {
    int n = L;
    for (int i = 0; i < rd.n_passes; i++) {
        const int r = rd.radix[i];
        if (i == 0 && window) run_pass<true>(r, z, L, n, rd.magic[i], tw, window, m0, m1);
        else run_pass<false>(r, z, L, n, rd.magic[i], tw, window, m0, m1);
        n /= r;
        __syncthreads();
    }
}

// storage position of frequency (or time) index k after the in-place DIF passes
static int digit_reverse(int k, int L, const GccRadices &rd) {
    int pos = 0, n = L;
    for (int i = 0; i < rd.n_passes; i++) {
        const int r = rd.radix[i];
        n /= r;
        const int kq = k / r;
        pos += (k - kq * r) * n;
        k = kq;
    }
    return pos;
}

__device__ __forceinline__ float block_sum(float v, float *s_red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s_red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.f;
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}

__device__ __forceinline__ float block_max(float v, float *s_red) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s_red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : -INFINITY;
    if (w == 0) {
        for (int o = 16; o > 0; o >>= 1) t = fmaxf(t, __shfl_xor_sync(0xffffffffu, t, o));
        if (l == 0) s_red[32] = t;
    }
    __syncthreads();
    return s_red[32];
}

struct GccArgs {
    const float *d0, *d1;
    long long stride;
    int n_pairs, L;
    const float *window;       // [L] np.hanning(L)
    const float2 *tw;          // [L] W_L^i
    const int *perm;           // [L] storage position of index k after the in-place DIF passes
    float *xcorr;              // [n_pairs][L] or NULL: raw Xcorr of this frame
    float *smoothed;           // [n_pairs][L] or NULL: in (if have_prev) / out smoothed Xcorr
    int have_prev;
    int *idx;                  // [n_pairs] argmax |Xs|
    float *val;                // [n_pairs] Xs[argmax]
    GccRadices rd;
};

__global__ void __launch_bounds__(GCC_THREADS, 1) gcc_phat_kernel(const GccArgs a) {
    extern __shared__ float2 z[];
    __shared__ float s_red[40];
    __shared__ int s_idx[33];
    const int L = a.L, half = L / 2;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int pair = blockIdx.x; pair < a.n_pairs; pair += gridDim.x) {
        const float *d0 = a.d0 + (size_t)pair * a.stride;
        const float *d1 = a.d1 + (size_t)pair * a.stride;
        // means; constant inputs (std == 0) are skipped by the reference (delay_estimator.py:129-131)
        float s0 = 0.f, s1 = 0.f, mn0 = INFINITY, mx0 = -INFINITY, mn1 = INFINITY, mx1 = -INFINITY;
        for (int n = tid; n < L; n += nt) {
            const float u = __ldg(d0 + n), v = __ldg(d1 + n);
            z[n] = make_float2(u, v);
            s0 += u;
            s1 += v;
            mn0 = fminf(mn0, u); mx0 = fmaxf(mx0, u);
            mn1 = fminf(mn1, v); mx1 = fmaxf(mx1, v);
        }
        const float m0 = block_sum(s0, s_red) / (float)L;
        const float m1 = block_sum(s1, s_red) / (float)L;
        const float span0 = block_max(mx0, s_red) + block_max(-mn0, s_red);
        const float span1 = block_max(mx1, s_red) + block_max(-mn1, s_red);
        if (!(span0 > 0.f) || !(span1 > 0.f)) {
            if (tid == 0) {
                a.idx[pair] = 0;
                a.val[pair] = 0.f;
            }
            if (a.xcorr)
                for (int n = tid; n < L; n += nt) a.xcorr[(size_t)pair * L + n] = 0.f;
            // the reference leaves old_Xcorr alone on a silent frame (delay_estimator.py:129-131,162-165);
            // on the very first call there is none yet: mark the pair "no previous" (NaN in slot 0)
            if (a.smoothed && a.have_prev == 0 && tid == 0) a.smoothed[(size_t)pair * L] = nanf("");
            __syncthreads();
            continue;
        }
        fft_inplace(z, L, a.rd, a.tw, a.window, m0, m1);   // mean removal + window fused in

        // G[k] = conj(D0[k]) D1[k], k = 0..L/2, kept in registers
        float2 g[GCC_MAX_PER_THREAD];
        float gmax = 0.f;
#pragma unroll
        for (int i = 0; i < GCC_MAX_PER_THREAD; i++) {
            const int k = tid + i * nt;
            g[i] = make_float2(0.f, 0.f);
            if (k <= half) {
                const float2 zk = z[__ldg(a.perm + k)];
                const float2 zm = z[__ldg(a.perm + (k == 0 ? 0 : L - k))];
                // D0 = (Zk + conj Zm)/2 ; D1 = (Zk - conj Zm)/(2j)
                const float2 D0 = make_float2(0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y));
                const float2 D1 = make_float2(0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x));
                g[i] = make_float2(D0.x * D1.x + D0.y * D1.y, D0.x * D1.y - D0.y * D1.x);
                gmax = fmaxf(gmax, sqrtf(g[i].x * g[i].x + g[i].y * g[i].y));
            }
        }
        const float m = block_max(gmax, s_red);   // also orders the reads above before the writes below
        const float floor_ = 1e-10f * m;
#pragma unroll
        for (int i = 0; i < GCC_MAX_PER_THREAD; i++) {
            const int k = tid + i * nt;
            if (k <= half) {
                const float ag = sqrtf(g[i].x * g[i].x + g[i].y * g[i].y);
                const float w = 1.0f / (floor_ + ag);
                float2 y = make_float2(g[i].x * w, g[i].y * w);
                if (k == 0 || k == half) y.y = 0.f;          // irfft ignores these imaginary parts
                // second FFT runs on conj(Y): x = Re(FFT(conj Y)) / L
                z[k] = make_float2(y.x, -y.y);
                if (k != 0 && k != half) z[L - k] = y;        // Y[L-k] = conj(Y[k])
            }
        }
        __syncthreads();
        fft_inplace(z, L, a.rd, a.tw, nullptr, 0.f, 0.f);

        // smoothing (delay_estimator.py:134-139) and arg-max of |Xs| (first maximum wins)
        const float inv = 1.0f / (float)L;
        float best = -1.f;
        int besti = 0x7fffffff;
        float *sm = a.smoothed ? a.smoothed + (size_t)pair * L : nullptr;
        float *xc = a.xcorr ? a.xcorr + (size_t)pair * L : nullptr;
        // have_prev: 1 = every pair has a previous smoothed frame; 2 = per pair, a NaN in slot 0
        // marks a pair that has had only silent frames so far (old_Xcorr is None there)
        bool prev = a.have_prev == 1;
        if (sm && a.have_prev == 2) prev = !isnan(sm[0]);
        __syncthreads();     // everyone has read the marker before slot 0 is rewritten
        for (int n = tid; n < L; n += nt) {
            float v = z[__ldg(a.perm + n)].x * inv;
            if (xc) xc[n] = v;
            if (sm) {
                if (prev) v = 0.3f * v + 0.7f * sm[n];
                sm[n] = v;
            }
            const float av = fabsf(v);
            if (av > best) {
                best = av;
                besti = n;
            }
        }
        // block arg-max
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
            if (ob > best || (ob == best && oi < besti)) {
                best = ob;
                besti = oi;
            }
        }
        __syncthreads();
        if ((tid & 31) == 0) {
            s_red[tid >> 5] = best;
            s_idx[tid >> 5] = besti;
        }
        __syncthreads();
        if (tid < 32) {
            best = (tid < (nt >> 5)) ? s_red[tid] : -1.f;
            besti = (tid < (nt >> 5)) ? s_idx[tid] : 0x7fffffff;
            for (int o = 16; o > 0; o >>= 1) {
                const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
                if (ob > best || (ob == best && oi < besti)) {
                    best = ob;
                    besti = oi;
                }
            }
            if (tid == 0) s_idx[32] = besti;
        }
        __syncthreads();
        if (tid == 0) {
            const int bi = s_idx[32];
            float v = z[__ldg(a.perm + bi)].x * inv;
            if (sm) v = sm[bi];
            a.idx[pair] = bi;
            a.val[pair] = v;
        }
        __syncthreads();
    }
}

}   // namespace

struct GccPlan {
    int L = 0;
    float *window = nullptr;
    float2 *tw = nullptr;
    int *perm = nullptr;
    GccRadices rd;
};

void frt_gcc_release(frt_ctx *h) {
    if (!h->gcc) return;
    if (h->gcc->window) cudaFree(h->gcc->window);
    if (h->gcc->tw) cudaFree(h->gcc->tw);
    if (h->gcc->perm) cudaFree(h->gcc->perm);
    delete h->gcc;
    h->gcc = nullptr;
}

extern "C" int frt_gcc_plan(frt_handle h, int length) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, length >= 4 && (length % 2) == 0, "length must be even and >= 4");
    FRT_CHECK_ARG(h, (size_t)length * sizeof(float2) <= 220 * 1024,
                  "length too large for the shared-memory FFT (max 28160)");
    FRT_CHECK_ARG(h, (length / 2 + 1 + GCC_THREADS - 1) / GCC_THREADS <= GCC_MAX_PER_THREAD,
                  "length too large");
    if (h->gcc && h->gcc->L == length) return FRT_OK;
    GccRadices rd;
    rd.n_passes = 0;
    int n = length;
    // large composite radices first (16 = 4x4, 25 = 5x5, 15 = 3x5 run in registers): every pass
    // is a full sweep over the 8*L-byte shared-memory buffer
    while (n % 16 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 16; n /= 16; }
    while (n % 4 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 4; n /= 4; }
    while (n % 2 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 2; n /= 2; }
    while (n % 25 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 25; n /= 25; }
    while (n % 15 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 15; n /= 15; }
    while (n % 5 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 5; n /= 5; }
    while (n % 3 == 0 && rd.n_passes < GCC_MAX_PASSES) { rd.radix[rd.n_passes++] = 3; n /= 3; }
    if (n != 1)
        return frt_fail(h, FRT_EINVAL, "length %d is not of the form 2^a 3^b 5^c", length);
    {
        int nn = length;
        for (int i = 0; i < rd.n_passes; i++) {
            const int n_next = nn / rd.radix[i];
            rd.magic[i] = n_next > 1 ? (unsigned)((0x100000000ULL + n_next - 1) / n_next) : 0u;
            nn = n_next;
        }
    }
    frt_gcc_release(h);
    GccPlan *pl = new (std::nothrow) GccPlan();
    if (!pl) return frt_fail(h, FRT_ENOMEM, "out of host memory");
    pl->L = length;
    pl->rd = rd;
    const double PI = 3.14159265358979323846;
    std::vector<float> win(length);
    std::vector<float2> tw(length);
    for (int i = 0; i < length; i++) {
        // numpy.hanning(L): 0.5 - 0.5 cos(2 pi n / (L-1))   (correlation.py:31)
        win[i] = (float)(0.5 - 0.5 * cos(2.0 * PI * i / (double)(length - 1)));
        const double ang = -2.0 * PI * (double)i / (double)length;
        tw[i] = make_float2((float)cos(ang), (float)sin(ang));
    }
    std::vector<int> perm(length);
    for (int i = 0; i < length; i++) perm[i] = digit_reverse(i, length, rd);
    cudaError_t e = cudaMalloc(&pl->window, sizeof(float) * length);
    if (e == cudaSuccess) e = cudaMalloc(&pl->perm, sizeof(int) * length);
    if (e == cudaSuccess)
        e = cudaMemcpy(pl->perm, perm.data(), sizeof(int) * length, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&pl->tw, sizeof(float2) * length);
    if (e == cudaSuccess)
        e = cudaMemcpy(pl->window, win.data(), sizeof(float) * length, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaMemcpy(pl->tw, tw.data(), sizeof(float2) * length, cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(gcc_phat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(sizeof(float2) * length));
    if (e != cudaSuccess) {
        if (pl->window) cudaFree(pl->window);
        if (pl->tw) cudaFree(pl->tw);
        if (pl->perm) cudaFree(pl->perm);
        delete pl;
        return frt_fail(h, FRT_ECUDA, "frt_gcc_plan: %s", cudaGetErrorString(e));
    }
    h->gcc = pl;
    return FRT_OK;
}

extern "C" int frt_gcc_phat(frt_handle h, const float *d0_dev, const float *d1_dev, int64_t stride,
                            int n_pairs, float *xcorr_dev, float *smoothed_dev, int have_prev,
                            int *idx_dev, float *val_dev, void *stream) {
    if (!h) return FRT_EINVAL;
    if (!h->gcc) return frt_fail(h, FRT_ESTATE, "frt_gcc_phat: call frt_gcc_plan first");
    DeviceGuard g(h->device);
    GccPlan *pl = h->gcc;
    FRT_CHECK_ARG(h, n_pairs >= 0, "negative n_pairs");
    if (n_pairs == 0) return FRT_OK;
    FRT_CHECK_ARG(h, d0_dev && d1_dev && idx_dev && val_dev, "NULL buffer");
    FRT_CHECK_ARG(h, stride >= pl->L, "stride smaller than the frame length");
    FRT_CHECK_ARG(h, have_prev >= 0 && have_prev <= 2, "have_prev must be 0, 1 or 2");
    FRT_CHECK_ARG(h, !have_prev || smoothed_dev, "have_prev needs the smoothed buffer");
    GccArgs a;
    a.d0 = d0_dev;
    a.d1 = d1_dev;
    a.stride = stride;
    a.n_pairs = n_pairs;
    a.L = pl->L;
    a.window = pl->window;
    a.tw = pl->tw;
    a.perm = pl->perm;
    a.xcorr = xcorr_dev;
    a.smoothed = smoothed_dev;
    a.have_prev = have_prev;
    a.idx = idx_dev;
    a.val = val_dev;
    a.rd = pl->rd;
    int blocks = n_pairs < h->sm_count ? n_pairs : h->sm_count;
    // the dynamic shared-memory limit is a per-device attribute of the kernel: another handle that
    // planned a shorter frame may have lowered it since this plan was built
    FRT_CUDA(h, cudaFuncSetAttribute(gcc_phat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(sizeof(float2) * pl->L)));
    gcc_phat_kernel<<<blocks, GCC_THREADS, sizeof(float2) * pl->L, (cudaStream_t)stream>>>(a);
    h->launches++;
    FRT_CUDA(h, cudaGetLastError());
    return FRT_OK;
}
