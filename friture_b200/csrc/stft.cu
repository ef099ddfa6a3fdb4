// Fused STFT for sm_100a: Hann window -> real FFT -> |X|^2/N^2 -> (10*log10(.+1e-30)), one pass
// HBM -> HBM, no cuFFT.  Stands behind audioproc.analyzelive (friture/audioproc.py:42-50), the
// framing loops of friture/spectrogram.py:131-159 / friture/spectrum.py:125-155 and
// log_spectrogram (friture/spectrogram.py:119-125).
//
// A real FFT of N points is computed as a complex FFT of M = N/2 points on
// z[n] = x[2n] + j*x[2n+1], followed by the split step
//     X[k]   = (E + T)/2,  X[M-k] = conj(E - T)/2,
//     E = Z[k] + conj(Z[M-k]),  O = Z[k] - conj(Z[M-k]),  T = (-j*W_N^k) * O.
//
// Fast path (N = 2048, the benchmark's size): ONE WARP PER FRAME, M = 1024 = 32 x 32.
//   lane t holds z[32*n1 + t] (n1 = 0..31) in registers -> radix-32 DFT over n1 in registers
//   -> twiddle W_M^(k1*t) -> 32x32 transpose through a warp-private padded smem tile
//   -> radix-32 DFT over n2 -> lane t holds X[t + 32*k2]; the split partner X[M-k] lives in
//   lane (32-t)%32 and is fetched with warp shuffles.  Only __syncwarp(), no block barrier.
//   Complex arithmetic is written on float2 so ptxas emits packed FADD2/FMUL2/FFMA2 (one issue
//   slot per complex add, two per complex multiply).
// N = 64 .. 1024: one warp per 32/Q frames (N = 64 Q), the same two register passes (stft_small_kernel).
// N = 4096 / 8192: R = 2 / 4 warps per frame, each a 1024-point FFT of a decimated sequence, radix-R
//   combine through shared memory (stft_large_kernel; stft_multi_kernel for unaligned input and 16384).
// Generic path (N = 32): one CTA per frame, Stockham radix-4 (+ one radix-2 pass when log2(M) is
//   odd) ping-ponging between two shared-memory buffers.
#include <cmath>
#include <cstdlib>

#include "frt_internal.cuh"

namespace {

constexpr float kLog10Scale = 3.01029995663981195f;   // 10 / log2(10)
constexpr float kEps = 1e-30f;                        // friture/spectrogram.py:124

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    return __fadd2_rn(a, make_float2(-b.x, -b.y));
}
// a * w
__device__ __forceinline__ float2 cmul(float2 a, float2 w) {
    float2 t = __fmul2_rn(a, make_float2(w.x, w.x));
    return __ffma2_rn(make_float2(-a.y, a.x), make_float2(w.y, w.y), t);
}
// a * (-j)
__device__ __forceinline__ float2 mul_mj(float2 a) { return make_float2(a.y, -a.x); }

// cos/sin(2*pi*i/32), i = 0..15
__device__ __forceinline__ constexpr float cos32(int i) {
    constexpr float t[16] = {1.0f,
                             0.98078528040323044913f,
                             0.92387953251128675613f,
                             0.83146961230254523708f,
                             0.70710678118654752440f,
                             0.55557023301960222474f,
                             0.38268343236508977173f,
                             0.19509032201612826785f,
                             0.0f,
                             -0.19509032201612826785f,
                             -0.38268343236508977173f,
                             -0.55557023301960222474f,
                             -0.70710678118654752440f,
                             -0.83146961230254523708f,
                             -0.92387953251128675613f,
                             -0.98078528040323044913f};
    return t[i];
}
__device__ __forceinline__ constexpr float sin32(int i) {
    constexpr float t[16] = {0.0f,
                             0.19509032201612826785f,
                             0.38268343236508977173f,
                             0.55557023301960222474f,
                             0.70710678118654752440f,
                             0.83146961230254523708f,
                             0.92387953251128675613f,
                             0.98078528040323044913f,
                             1.0f,
                             0.98078528040323044913f,
                             0.92387953251128675613f,
                             0.83146961230254523708f,
                             0.70710678118654752440f,
                             0.55557023301960222474f,
                             0.38268343236508977173f,
                             0.19509032201612826785f};
    return t[i];
}

// cos/sin(2*pi*m/64), m = 0..15 (split-step twiddle steps W_2048^(32 m))
[[maybe_unused]] __device__ __forceinline__ constexpr float cos64(int m) {
    constexpr float t[16] = {1.0f, 0.99518472667219688624f, 0.98078528040323044913f,
                             0.95694033573220886494f, 0.92387953251128675613f,
                             0.88192126434835502971f, 0.83146961230254523708f,
                             0.77301045336273696081f, 0.70710678118654752440f,
                             0.63439328416364549822f, 0.55557023301960222474f,
                             0.47139673682599764856f, 0.38268343236508977173f,
                             0.29028467725446236764f, 0.19509032201612826785f,
                             0.09801714032956060199f};
    return t[m];
}
[[maybe_unused]] __device__ __forceinline__ constexpr float sin64(int m) {
    constexpr float t[16] = {0.0f, 0.09801714032956060199f, 0.19509032201612826785f,
                             0.29028467725446236764f, 0.38268343236508977173f,
                             0.47139673682599764856f, 0.55557023301960222474f,
                             0.63439328416364549822f, 0.70710678118654752440f,
                             0.77301045336273696081f, 0.83146961230254523708f,
                             0.88192126434835502971f, 0.92387953251128675613f,
                             0.95694033573220886494f, 0.98078528040323044913f,
                             0.99518472667219688624f};
    return t[m];
}

// Tuning switches (round 1 measurements on B200, config #2, % of the measured HBM peak, 20-launch /
// 300-launch runs): tables from shared memory 65.2 / 59.4 (default); split twiddles derived from one
// per-lane value 64.3 / 58.7; Hann window computed on the fly from per-lane phase factors (needs 12
// warps per CTA to avoid spills) 61.1 / 57.8; both with 16 warps (spilling) 59.6 / 55.4.  Moving
// table traffic from the LSU pipe to the FMA pipe does not pay: both are close to their limits.
#ifndef FRT_STFT_WINDOW_ON_THE_FLY
#define FRT_STFT_WINDOW_ON_THE_FLY 0
#endif
#ifndef FRT_STFT_DERIVE_POST
#define FRT_STFT_DERIVE_POST 0
#endif

// cos/sin(64 * n1 * 2*pi/2047), n1 = 0..31: the per-register part of the Hann phase (N = 2048)
#include "hann2048_tables.inc"

__host__ __device__ constexpr int brev5(int r) {
    return ((r & 1) << 4) | ((r & 2) << 2) | (r & 4) | ((r & 8) >> 2) | ((r & 16) >> 4);
}

// In-register radix-32 DFT (decimation in frequency, fully unrolled): v[r] <- bin brev5(r).
__device__ __forceinline__ void dft32(float2 (&v)[32]) {
#pragma unroll
    for (int len = 32; len >= 2; len >>= 1) {
        const int half = len >> 1;
        const int tstep = 32 / len;
#pragma unroll
        for (int base = 0; base < 32; base += len) {
#pragma unroll
            for (int i = 0; i < half; i++) {
                const float2 a = v[base + i], b = v[base + i + half];
                v[base + i] = cadd(a, b);
                const float2 d = csub(a, b);
                const int ti = i * tstep;
                if (ti == 0)
                    v[base + i + half] = d;
                else if (ti == 8)
                    v[base + i + half] = mul_mj(d);
                else
                    v[base + i + half] = cmul(d, make_float2(cos32(ti), -sin32(ti)));
            }
        }
    }
}

// 32/LEN independent LEN-point DFTs on consecutive register groups (same butterflies as dft32,
// started at span LEN): v[g*LEN + r] <- bin brevq<LEN>(r) of group g.
template <int LEN>
__device__ __forceinline__ void dft_groups(float2 (&v)[32]) {
#pragma unroll
    for (int len = LEN; len >= 2; len >>= 1) {
        const int half = len >> 1;
        const int tstep = 32 / len;
#pragma unroll
        for (int base = 0; base < 32; base += len) {
#pragma unroll
            for (int i = 0; i < half; i++) {
                const float2 a = v[base + i], b = v[base + i + half];
                v[base + i] = cadd(a, b);
                const float2 d = csub(a, b);
                const int ti = i * tstep;
                if (ti == 0)
                    v[base + i + half] = d;
                else if (ti == 8)
                    v[base + i + half] = mul_mj(d);
                else
                    v[base + i + half] = cmul(d, make_float2(cos32(ti), -sin32(ti)));
            }
        }
    }
}
template <int Q>
__host__ __device__ constexpr int brevq(int r) {
    int o = 0;
    for (int b = 1, c = Q >> 1; b < Q; b <<= 1, c >>= 1)
        if (r & b) o |= c;
    return o;
}

__device__ __forceinline__ float lg2_fast(float v) {
    // MUFU.LG2; the argument is always >= 1e-30 (normal), so the denormal fix-up of
    // __log2f is not needed
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

template <int MODE>
__device__ __forceinline__ float finish(float mag2, float scale) {
    // mag2*scale = |X|^2/N^2 (audioproc.py:49-50); log mode: spectrogram.py:119-125
    if (MODE == FRT_STFT_POWER) return mag2 * scale;
    return kLog10Scale * lg2_fast(fmaf(mag2, scale, kEps));
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers ------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy by the TMA engine, completion signalled on the mbarrier
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, unsigned bytes,
                                         unsigned long long *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// Fast path, N = 2048.
constexpr int FAST_N = 2048;
constexpr int FAST_M = 1024;
#ifndef FRT_STFT_WARPS
#define FRT_STFT_WARPS 16
#endif
constexpr int FAST_WARPS = FRT_STFT_WARPS;
constexpr int FAST_TILE = 32 * 33;   // padded 32x32 complex tile per warp
constexpr int FAST_POST = 17 * 32;
constexpr size_t FAST_SMEM =
    sizeof(float2) * (FAST_M + FAST_M + FAST_POST + (size_t)FAST_WARPS * FAST_TILE) +
    sizeof(unsigned long long) * FAST_WARPS;

template <int MODE, int VEC>
__global__ void __launch_bounds__(FAST_WARPS * 32, 1)
stft2048_kernel(const float *__restrict__ x, long long x_stride, long long n_frames, int hop,
                float *__restrict__ out, long long out_stride_c, long long out_stride_f,
                const float2 *__restrict__ win2, const float2 *__restrict__ tw,
                const float2 *__restrict__ post, const float2 *__restrict__ wlane,
                long long total_items) {
    extern __shared__ float2 smem[];
    float2 *s_win = smem;                 // [1024]  (w[2n], w[2n+1])
    float2 *s_tw = s_win + FAST_M;        // [32][32] W_1024^(k1*t)
    float2 *s_post = s_tw + FAST_M;       // [17][32] U[t + 32 m] = -j W_2048^(t+32m)
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    float2 *s_x = s_post + FAST_POST + warp * FAST_TILE;   // 16-byte aligned: FAST_TILE*8 % 16 == 0
    unsigned long long *s_bar =
        reinterpret_cast<unsigned long long *>(s_post + FAST_POST + FAST_WARPS * FAST_TILE);

    for (int i = threadIdx.x; i < FAST_M; i += blockDim.x) {
        s_win[i] = win2[i];
        s_tw[i] = tw[i];
    }
    for (int i = threadIdx.x; i < FAST_POST; i += blockDim.x) s_post[i] = post[i];
    __syncthreads();

    // contiguous range of (channel, frame) items per warp: consecutive frames of a channel
    // stay on one SM so the overlapping half of each frame is re-read from L1/L2, not HBM
    const long long warps_total = (long long)gridDim.x * FAST_WARPS;
    const long long gw = (long long)blockIdx.x * FAST_WARPS + warp;
    const long long per = (total_items + warps_total - 1) / warps_total;
    long long item = gw * per;
    long long item_end = item + per;
    if (item_end > total_items) item_end = total_items;
    const float scale = 1.0f / (4.0f * (float)FAST_N * (float)FAST_N);
    const int src_lane = (32 - lane) & 31;
    // VEC == 2: the next frame's 8 KB of samples are fetched by the TMA engine (cp.async.bulk)
    // into this warp's exchange tile while the current frame is in its second FFT pass and split
    // step, so no warp ever waits on a global load with its registers tied up.
    unsigned long long *bar = s_bar + warp;
    unsigned parity = 0;
    if (VEC == 2) {
        if (lane == 0) mbar_init(bar, 1);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0 && item < item_end) {
            mbar_expect_tx(bar, FAST_N * 4);
            bulk_g2s(s_x, x + (item / n_frames) * x_stride + (item % n_frames) * hop, FAST_N * 4,
                     bar);
        }
    }
#if FRT_STFT_WINDOW_ON_THE_FLY
    // Hann window w[n] = 0.5 - 0.5 cos(theta n), theta = 2 pi/(N-1), n = 64 n1 + 2 t + e:
    // cos(theta n) = cos(A) cos(B) - sin(A) sin(B) with A = 64 theta n1 (compile-time constants in
    // the unrolled loop) and B = theta (2t + e) (4 per-lane registers) -- no table traffic.
    const float2 wcb = wlane[lane], wsb = wlane[32 + lane];   // (cos B_e), (sin B_e), e = 0, 1
#endif
#if FRT_STFT_DERIVE_POST
    const float2 u_lane = s_post[lane];   // U[t] = -j W_2048^t; U[t + 32 m] = U[t] * W_64^m
#endif

    long long c = (item < item_end) ? item / n_frames : 0;
    long long f = item - c * n_frames;
    for (; item < item_end; item++) {
        const float *p = x + c * x_stride + f * hop;
        float2 v[32];
        if (VEC == 2) {
            mbar_wait(bar, parity);
            parity ^= 1;
#pragma unroll
            for (int n1 = 0; n1 < 32; n1++) v[n1] = s_x[32 * n1 + lane];
            __syncwarp();   // every lane has its samples before the tile is reused for the exchange
        } else if (VEC) {
            const float2 *p2 = reinterpret_cast<const float2 *>(p);
#pragma unroll
            for (int n1 = 0; n1 < 32; n1++) v[n1] = __ldg(p2 + 32 * n1 + lane);
        } else {
#pragma unroll
            for (int n1 = 0; n1 < 32; n1++) {
                v[n1].x = __ldg(p + 64 * n1 + 2 * lane);
                v[n1].y = __ldg(p + 64 * n1 + 2 * lane + 1);
            }
        }
#if FRT_STFT_WINDOW_ON_THE_FLY
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {
            const float ca = 0.5f * hann_cos_a(n1), sa = 0.5f * hann_sin_a(n1);
            float2 wv = __ffma2_rn(make_float2(-ca, -ca), wcb, make_float2(0.5f, 0.5f));
            wv = __ffma2_rn(make_float2(sa, sa), wsb, wv);
            v[n1] = __fmul2_rn(v[n1], wv);
        }
#else
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) v[n1] = __fmul2_rn(v[n1], s_win[32 * n1 + lane]);
#endif

        dft32(v);   // v[r] = sum_n1 z[32 n1 + t] W_32^(n1 k1), k1 = brev5(r)
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const int k1 = brev5(r);
            const float2 val = (k1 == 0) ? v[r] : cmul(v[r], s_tw[k1 * 32 + lane]);
            s_x[k1 * 33 + lane] = val;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = s_x[lane * 33 + n2];
        __syncwarp();
        if (VEC == 2 && item + 1 < item_end) {   // tile is free: prefetch the next frame into it
            long long cn = c, fn = f + 1;
            if (fn == n_frames) {
                fn = 0;
                cn++;
            }
            fence_proxy_async();
            if (lane == 0) {
                mbar_expect_tx(bar, FAST_N * 4);
                bulk_g2s(s_x, x + cn * x_stride + fn * hop, FAST_N * 4, bar);
            }
        }
        dft32(v);   // v[r] = Z[lane + 32*brev5(r)]

        float *o = out + c * out_stride_c + f * out_stride_f;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const float2 z = v[brev5(m)];
            const float2 other = v[brev5(31 - m)];
            float2 zp;
            zp.x = __shfl_sync(0xffffffffu, other.x, src_lane);
            zp.y = __shfl_sync(0xffffffffu, other.y, src_lane);
            if (lane == 0) zp = (m == 0) ? v[0] : v[brev5(32 - m)];
            const float2 E = make_float2(z.x + zp.x, z.y - zp.y);
            const float2 O = make_float2(z.x - zp.x, z.y + zp.y);
#if FRT_STFT_DERIVE_POST
            const float2 T = (m == 0) ? cmul(O, u_lane)
                                      : cmul(cmul(O, make_float2(cos64(m), -sin64(m))), u_lane);
#else
            const float2 T = cmul(O, s_post[m * 32 + lane]);
#endif
            const float2 X = cadd(E, T);
            const float2 Y = csub(E, T);
            const float p1 = fmaf(X.x, X.x, X.y * X.y);
            const float p2 = fmaf(Y.x, Y.x, Y.y * Y.y);
            const int k = lane + 32 * m;
            o[k] = finish<MODE>(p1, scale);
            o[FAST_M - k] = finish<MODE>(p2, scale);
        }
        if (lane == 0) {
            const float2 z = v[brev5(16)];   // bin M/2: X = conj(Z)
            o[FAST_M / 2] = finish<MODE>(4.0f * fmaf(z.x, z.x, z.y * z.y), scale);
        }
        if (++f == n_frames) {
            f = 0;
            c++;
        }
    }
}

// ------------------------------------------------------------------------------------------
// N = 4096 / 8192 (the spectrogram's and the spectrum's default sizes) and 16384: R = 2 / 4 / 8
// warps per frame.
// Warp w runs the 1024-point complex FFT of the decimated sequence z[R n + w] exactly like the
// N = 2048 kernel (radix-32 x 32 in registers, one transpose through its smem tile), leaves
// Z_w[k0] in natural order in the tile, and after one block barrier the CTA combines
//   Z[k0 + 1024 q] = sum_w W_R^(w q) W_M^(w k0) Z_w[k0]          (M = 1024 R)
// and applies the real-FFT split step.  The partner of bin k0 + 1024 q is (1024 - k0) + 1024 (R-1-q),
// so a thread that builds the groups of k0 and 1024 - k0 has all 2R bins it needs in registers.
template <int R> __device__ __forceinline__ void combine(float2 (&t)[R]);
template <> __device__ __forceinline__ void combine<2>(float2 (&t)[2]) {
    const float2 a = t[0], b = t[1];
    t[0] = cadd(a, b);
    t[1] = csub(a, b);
}
template <> __device__ __forceinline__ void combine<4>(float2 (&t)[4]) {
    const float2 s02 = cadd(t[0], t[2]), d02 = csub(t[0], t[2]);
    const float2 s13 = cadd(t[1], t[3]), d13 = mul_mj(csub(t[1], t[3]));
    t[0] = cadd(s02, s13);
    t[1] = cadd(d02, d13);
    t[2] = csub(s02, s13);
    t[3] = csub(d02, d13);
}

template <> __device__ __forceinline__ void combine<8>(float2 (&t)[8]) {
    float2 e[4] = {t[0], t[2], t[4], t[6]}, o[4] = {t[1], t[3], t[5], t[7]};
    combine<4>(e);
    combine<4>(o);
    const float h = 0.70710678118654752440f;
    // W_8^q o_q: q = 1: (1 - j)/sqrt2, q = 2: -j, q = 3: (-1 - j)/sqrt2
    const float2 o1 = make_float2(h * (o[1].x + o[1].y), h * (o[1].y - o[1].x));
    const float2 o2 = mul_mj(o[2]);
    const float2 o3 = make_float2(h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y));
    t[0] = cadd(e[0], o[0]); t[4] = csub(e[0], o[0]);
    t[1] = cadd(e[1], o1);   t[5] = csub(e[1], o1);
    t[2] = cadd(e[2], o2);   t[6] = csub(e[2], o2);
    t[3] = cadd(e[3], o3);   t[7] = csub(e[3], o3);
}

template <int R>
__device__ __forceinline__ void load_group(const float2 *__restrict__ tiles,
                                           const float2 *__restrict__ comb, int k0,
                                           float2 (&t)[R]) {
    t[0] = tiles[k0];
#pragma unroll
    for (int w = 1; w < R; w++)
        t[w] = cmul(tiles[w * FAST_TILE + k0], __ldg(comb + (w - 1) * 1024 + k0));
    combine<R>(t);
}

template <int MODE>
__device__ __forceinline__ void split_store(float *__restrict__ o, int kk, int M, float2 z, float2 zp,
                                            const float2 *__restrict__ post, float scale) {
    const float2 E = make_float2(z.x + zp.x, z.y - zp.y);
    const float2 O = make_float2(z.x - zp.x, z.y + zp.y);
    const float2 T = cmul(O, __ldg(post + kk));
    const float2 X = cadd(E, T);
    const float2 Y = csub(E, T);
    o[kk] = finish<MODE>(fmaf(X.x, X.x, X.y * X.y), scale);
    o[M - kk] = finish<MODE>(fmaf(Y.x, Y.x, Y.y * Y.y), scale);
}

template <int R, int MODE>
__global__ void __launch_bounds__(32 * R)
stft_multi_kernel(const float *__restrict__ x, long long x_stride, long long n_frames, int hop,
                  float *__restrict__ out, long long out_stride_c, long long out_stride_f,
                  const float2 *__restrict__ win2, const float2 *__restrict__ tw32,
                  const float2 *__restrict__ comb, const float2 *__restrict__ post,
                  long long total_items, int vec_ok) {
    constexpr int M = 1024 * R, N = 2 * M;
    extern __shared__ float2 smem[];
    float2 *s_tw = smem;                       // [32][32] W_1024^(k1 t)
    float2 *s_tiles = s_tw + FAST_M;           // R padded tiles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, tid = threadIdx.x;
    float2 *s_x = s_tiles + warp * FAST_TILE;
    for (int i = tid; i < FAST_M; i += blockDim.x) s_tw[i] = tw32[i];
    __syncthreads();
    const float scale = 1.0f / (4.0f * (float)N * (float)N);
    for (long long item = blockIdx.x; item < total_items; item += gridDim.x) {
        const long long c = item / n_frames;
        const long long f = item - c * n_frames;
        const float *p = x + c * x_stride + f * hop;
        float2 v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {
            const int idx = R * (32 * n1 + lane) + warp;      // z[idx] = x[2 idx] + j x[2 idx + 1]
            float2 xv;
            if (vec_ok) {
                xv = __ldg(reinterpret_cast<const float2 *>(p) + idx);
            } else {
                xv.x = __ldg(p + 2 * idx);
                xv.y = __ldg(p + 2 * idx + 1);
            }
            v[n1] = __fmul2_rn(xv, __ldg(win2 + idx));
        }
        dft32(v);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const int k1 = brev5(r);
            const float2 val = (k1 == 0) ? v[r] : cmul(v[r], s_tw[k1 * 32 + lane]);
            s_x[k1 * 33 + lane] = val;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = s_x[lane * 33 + n2];
        __syncwarp();
        dft32(v);   // v[r] = Z_w[lane + 32*brev5(r)]
#pragma unroll
        for (int r = 0; r < 32; r++) s_x[lane + 32 * brev5(r)] = v[r];   // natural order
        __syncthreads();

        float *o = out + c * out_stride_c + f * out_stride_f;
        for (int k0 = 1 + tid; k0 < 512; k0 += 32 * R) {
            float2 A[R], B[R];
            load_group<R>(s_tiles, comb, k0, A);
            load_group<R>(s_tiles, comb, 1024 - k0, B);
#pragma unroll
            for (int q = 0; q < R; q++)
                split_store<MODE>(o, k0 + 1024 * q, M, A[q], B[R - 1 - q], post, scale);
        }
        if (tid == 0) {          // k0 = 0: partners inside the group, q <-> R - q
            float2 A[R];
            load_group<R>(s_tiles, comb, 0, A);
            split_store<MODE>(o, 0, M, A[0], A[0], post, scale);
#pragma unroll
            for (int q = 1; q <= R / 2; q++) split_store<MODE>(o, 1024 * q, M, A[q], A[R - q], post, scale);
        }
        if (tid == 32) {         // k0 = 512: partners inside the group, q <-> R - 1 - q
            float2 A[R];
            load_group<R>(s_tiles, comb, 512, A);
#pragma unroll
            for (int q = 0; q < R / 2; q++)
                split_store<MODE>(o, 512 + 1024 * q, M, A[q], A[R - 1 - q], post, scale);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// N = 64 Q, Q in {1, 2, 4, 8, 16} (64 .. 1024 points): ONE WARP PER G = 32/Q FRAMES.
// M = N/2 = 32 Q.  Lane t holds z_g[32 n1 + t] (n1 < Q) of each of the G frames in register g*Q + n1,
// so the register file is as full as in the N = 2048 kernel: G radix-Q DFTs in registers -> twiddle
// W_M^(k1 t) -> ONE 32x32 transpose through the warp's tile, after which lane g*Q + k1 holds the 32
// values (over t) of frame g, residue k1 -> radix-32 DFT -> that lane owns Z_g[k1 + Q k2], k2 = 0..31.
// The split partner M - (k1 + Q m) = (Q - k1) + Q (31 - m) lives in lane g*Q + (Q - k1) % Q (the
// lane itself when k1 = 0), one shuffle away.  Window, twiddle and split tables are read once per
// G frames.
constexpr int SMALL_WARPS = 8;
template <int Q>
constexpr size_t small_smem() {
    return sizeof(float2) * (32 * Q + 32 * Q + 16 * Q + (size_t)SMALL_WARPS * FAST_TILE);
}

template <int Q, int MODE>
__global__ void __launch_bounds__(SMALL_WARPS * 32, 2)
stft_small_kernel(const float *__restrict__ x, long long x_stride, long long n_frames, int hop,
                  float *__restrict__ out, long long out_stride_c, long long out_stride_f,
                  const float2 *__restrict__ win2, const float2 *__restrict__ tw,
                  const float2 *__restrict__ post, long long total_items, int vec_ok) {
    constexpr int G = 32 / Q, M = 32 * Q, N = 2 * M;
    extern __shared__ float2 smem[];
    float2 *s_win = smem;              // [M]     (w[2n], w[2n+1])
    float2 *s_tw = s_win + M;          // [Q][32] W_M^(k1 t)
    float2 *s_post = s_tw + M;         // [M/2]   U[k] = -j W_N^k
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 *s_x = s_post + 16 * Q + warp * FAST_TILE;
    for (int i = threadIdx.x; i < M; i += blockDim.x) {
        s_win[i] = win2[i];
        s_tw[i] = tw[i];
    }
    for (int i = threadIdx.x; i < 16 * Q; i += blockDim.x) s_post[i] = post[i];
    __syncthreads();

    // contiguous range of frames per warp, a multiple of G long
    const long long warps_total = (long long)gridDim.x * SMALL_WARPS;
    const long long gw = (long long)blockIdx.x * SMALL_WARPS + warp;
    long long per = (total_items + warps_total - 1) / warps_total;
    per = (per + G - 1) / G * G;
    long long item = gw * per;
    long long item_end = item + per;
    if (item_end > total_items) item_end = total_items;
    const float scale = 1.0f / (4.0f * (float)N * (float)N);
    const int k1 = lane & (Q - 1), gsel = lane / Q;
    const int src_lane = (lane - k1) + ((Q - k1) & (Q - 1));

    long long c = (item < item_end) ? item / n_frames : 0;
    long long f = item - c * n_frames;
    for (; item < item_end; item += G) {
        float2 v[32];
        long long myc = 0, myf = 0;
        bool myok = false;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const bool ok = item + g < item_end;
            const float *p = x + c * x_stride + f * hop;
            if (g == gsel) {
                myc = c;
                myf = f;
                myok = ok;
            }
            if (ok && vec_ok) {
#pragma unroll
                for (int n1 = 0; n1 < Q; n1++)
                    v[g * Q + n1] = __ldg(reinterpret_cast<const float2 *>(p) + 32 * n1 + lane);
            } else if (ok) {
#pragma unroll
                for (int n1 = 0; n1 < Q; n1++) {
                    v[g * Q + n1].x = __ldg(p + 64 * n1 + 2 * lane);
                    v[g * Q + n1].y = __ldg(p + 64 * n1 + 2 * lane + 1);
                }
            } else {
#pragma unroll
                for (int n1 = 0; n1 < Q; n1++) v[g * Q + n1] = make_float2(0.0f, 0.0f);
            }
            if (++f == n_frames) {
                f = 0;
                c++;
            }
        }
#pragma unroll
        for (int n1 = 0; n1 < Q; n1++) {
            const float2 w = s_win[32 * n1 + lane];
#pragma unroll
            for (int g = 0; g < G; g++) v[g * Q + n1] = __fmul2_rn(v[g * Q + n1], w);
        }
        dft_groups<Q>(v);   // v[g*Q + r] = sum_n1 z_g[32 n1 + t] W_Q^(n1 k1), k1 = brevq<Q>(r)
#pragma unroll
        for (int r = 0; r < Q; r++) {
            const int kk1 = brevq<Q>(r);
            const float2 w = s_tw[kk1 * 32 + lane];
#pragma unroll
            for (int g = 0; g < G; g++) {
                const float2 val = (kk1 == 0) ? v[g * Q + r] : cmul(v[g * Q + r], w);
                s_x[(g * Q + kk1) * 33 + lane] = val;
            }
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = s_x[lane * 33 + n2];
        __syncwarp();
        dft32(v);   // v[r] = Z_g[k1 + Q*brev5(r)] of this lane's (g, k1)

        float *o = out + myc * out_stride_c + myf * out_stride_f;
#pragma unroll
        for (int m = 0; m < 16; m++) {
            const float2 z = v[brev5(m)];
            const float2 other = v[brev5(31 - m)];
            float2 zp;
            zp.x = __shfl_sync(0xffffffffu, other.x, src_lane);
            zp.y = __shfl_sync(0xffffffffu, other.y, src_lane);
            if (k1 == 0) zp = (m == 0) ? v[0] : v[brev5(32 - m)];
            const float2 E = make_float2(z.x + zp.x, z.y - zp.y);
            const float2 O = make_float2(z.x - zp.x, z.y + zp.y);
            const int k = k1 + Q * m;
            const float2 T = cmul(O, s_post[k]);
            const float2 X = cadd(E, T);
            const float2 Y = csub(E, T);
            const float p1 = fmaf(X.x, X.x, X.y * X.y);
            const float p2 = fmaf(Y.x, Y.x, Y.y * Y.y);
            if (myok) {
                o[k] = finish<MODE>(p1, scale);
                o[M - k] = finish<MODE>(p2, scale);
            }
        }
        if (k1 == 0 && myok) {
            const float2 z = v[brev5(16)];   // bin M/2: X = conj(Z)
            o[M / 2] = finish<MODE>(4.0f * fmaf(z.x, z.x, z.y * z.y), scale);
        }
    }
}

// ------------------------------------------------------------------------------------------
// N = 4096 / 8192, 8-byte aligned input: the stft_multi_kernel decomposition (R warps per frame, each a
// 1024-point complex FFT of a decimated sequence, then a radix-R combine + split step) rebuilt
// around shared memory: one CTA of 16 warps (16/R frame groups) per SM, every table resident in shared
// memory, the frame fetched by 8-byte cp.async that DE-INTERLEAVES it on the way in (element e of the
// frame lands in warp e % R's tile at e / R, staggered by w*32/R so that the scatter is conflict
// free), named barriers per group instead of block barriers.
// The Hann window is computed on the fly (this kernel is bound by shared-memory wavefronts, not by
// the FMA pipe): w[n] = 0.5 - 0.5 cos(theta n), theta = 2 pi/(N-1), and for warp w, lane t, register
// n1, element e the sample index is n = 64 R n1 + (2 R t + 2 w + e) = A-part + B-part, so
// cos(theta n) = cos A cos B - sin A sin B with 32 (cos A, sin A)/2 pairs passed as a kernel
// argument (constant bank) and the four B factors of a lane held in registers.
constexpr int LARGE_WARPS = 16;
struct HannA {
    float2 v[32];     // (0.5 cos(theta 64 R n1), 0.5 sin(theta 64 R n1))
};
template <int R>
constexpr size_t large_smem() {
    // tw [1024] + comb [R-1][1024] + post [512 R + 1 -> even] + tiles
    return sizeof(float2) * (1024 + 1024 * (R - 1) + 512 * R + 2 + (size_t)LARGE_WARPS * FAST_TILE);
}

__device__ __forceinline__ void group_barrier(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cp_async8(void *dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// split twiddle U[kk] = -j W_N^kk for kk in [0, M] from a table of M/2 + 1 entries:
// U[kk + M/2] = -j U[kk]
__device__ __forceinline__ float2 post_lookup(const float2 *s_post, int kk, int half_m) {
    if (kk <= half_m) return s_post[kk];
    return mul_mj(s_post[kk - half_m]);
}

template <int R, int MODE>
__device__ __forceinline__ void split_store_s(float *__restrict__ o, int kk, float2 z, float2 zp,
                                              const float2 *s_post, float scale) {
    constexpr int M = 1024 * R;
    const float2 E = make_float2(z.x + zp.x, z.y - zp.y);
    const float2 O = make_float2(z.x - zp.x, z.y + zp.y);
    const float2 T = cmul(O, post_lookup(s_post, kk, M / 2));
    const float2 X = cadd(E, T);
    const float2 Y = csub(E, T);
    o[kk] = finish<MODE>(fmaf(X.x, X.x, X.y * X.y), scale);
    o[M - kk] = finish<MODE>(fmaf(Y.x, Y.x, Y.y * Y.y), scale);
}

template <int R>
__device__ __forceinline__ void load_group_s(const float2 *tiles, const float2 *s_comb, int k0,
                                             float2 (&t)[R]) {
    t[0] = tiles[k0];
#pragma unroll
    for (int w = 1; w < R; w++) t[w] = cmul(tiles[w * FAST_TILE + k0], s_comb[(w - 1) * 1024 + k0]);
    combine<R>(t);
}

template <int R, int MODE>
__global__ void __launch_bounds__(LARGE_WARPS * 32, 1)
stft_large_kernel(const float *__restrict__ x, long long x_stride, long long n_frames, int hop,
                  float *__restrict__ out, long long out_stride_c, long long out_stride_f,
                  const float4 *__restrict__ wlane, const float2 *__restrict__ tw32,
                  const float2 *__restrict__ comb, const float2 *__restrict__ post,
                  long long total_items, const __grid_constant__ HannA hann_a) {
    // STAG: 8-byte shared accesses are served per half-warp; the 16/R elements a half-warp sends to
    // each of the R tiles must fall on different banks
    constexpr int M = 1024 * R, N = 2 * M, NG = LARGE_WARPS / R, GT = 32 * R, STAG = 16 / R;
    extern __shared__ float2 smem[];
    float2 *s_tw = smem;                          // [32][32] W_1024^(k1 t)
    float2 *s_comb = s_tw + 1024;                 // [R-1][1024] W_M^(w k0)
    float2 *s_post = s_comb + 1024 * (R - 1);     // [M/2 + 1] U[k]
    float2 *s_tiles_all = s_post + 512 * R + 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp / R, w = warp % R, gtid = tid - grp * GT;
    float2 *s_tiles = s_tiles_all + (size_t)grp * R * FAST_TILE;   // this group's R tiles
    float2 *s_x = s_tiles + w * FAST_TILE;
    for (int i = tid; i < 1024; i += blockDim.x) s_tw[i] = tw32[i];
    for (int i = tid; i < 1024 * (R - 1); i += blockDim.x) s_comb[i] = comb[i];
    for (int i = tid; i <= 512 * R; i += blockDim.x) s_post[i] = post[i];
    __syncthreads();

    const long long groups_total = (long long)gridDim.x * NG;
    const long long gg = (long long)blockIdx.x * NG + grp;
    const long long per = (total_items + groups_total - 1) / groups_total;
    long long item = gg * per;
    long long item_end = item + per;
    if (item_end > total_items) item_end = total_items;
    const float scale = 1.0f / (4.0f * (float)N * (float)N);
    long long c = (item < item_end) ? item / n_frames : 0;
    long long f = item - c * n_frames;
    const float4 wb = wlane[w * 32 + lane];       // cos B (e = 0, 1), sin B (e = 0, 1)
    const float2 wcb = make_float2(wb.x, wb.y), wsb = make_float2(wb.z, wb.w);

    auto fetch = [&](long long cc, long long ff) {
        const float2 *p2 = reinterpret_cast<const float2 *>(x + cc * x_stride + ff * hop);
#pragma unroll 8
        for (int e = gtid; e < M; e += GT) {
            const int ww = e % R, i = e / R;
            cp_async8(s_tiles + ww * FAST_TILE + ww * STAG + i, p2 + e);
        }
    };
    if (item < item_end) fetch(c, f);
    for (; item < item_end; item++) {
        cp_async_commit_wait_all();
        group_barrier(1 + grp, GT);          // the whole frame has landed
        float2 v[32];
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) v[n1] = s_x[w * STAG + 32 * n1 + lane];
        __syncwarp();                        // inputs are in registers: the tile becomes the exchange buffer
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {
            const float ca = hann_a.v[n1].x, sa = hann_a.v[n1].y;
            float2 wv = __ffma2_rn(make_float2(-ca, -ca), wcb, make_float2(0.5f, 0.5f));
            wv = __ffma2_rn(make_float2(sa, sa), wsb, wv);
            v[n1] = __fmul2_rn(v[n1], wv);
        }
        dft32(v);
#pragma unroll
        for (int r = 0; r < 32; r++) {
            const int k1 = brev5(r);
            const float2 val = (k1 == 0) ? v[r] : cmul(v[r], s_tw[k1 * 32 + lane]);
            s_x[k1 * 33 + lane] = val;
        }
        __syncwarp();
#pragma unroll
        for (int n2 = 0; n2 < 32; n2++) v[n2] = s_x[lane * 33 + n2];
        __syncwarp();
        dft32(v);   // v[r] = Z_w[lane + 32*brev5(r)]
#pragma unroll
        for (int r = 0; r < 32; r++) s_x[lane + 32 * brev5(r)] = v[r];   // natural order
        group_barrier(1 + grp, GT);

        float *o = out + c * out_stride_c + f * out_stride_f;
        for (int k0 = 1 + gtid; k0 < 512; k0 += GT) {
            float2 A[R], B[R];
            load_group_s<R>(s_tiles, s_comb, k0, A);
            load_group_s<R>(s_tiles, s_comb, 1024 - k0, B);
#pragma unroll
            for (int q = 0; q < R; q++)
                split_store_s<R, MODE>(o, k0 + 1024 * q, A[q], B[R - 1 - q], s_post, scale);
        }
        if (gtid == 0) {           // k0 = 0: partners inside the group, q <-> R - q
            float2 A[R];
            load_group_s<R>(s_tiles, s_comb, 0, A);
            split_store_s<R, MODE>(o, 0, A[0], A[0], s_post, scale);
#pragma unroll
            for (int q = 1; q <= R / 2; q++) split_store_s<R, MODE>(o, 1024 * q, A[q], A[R - q], s_post, scale);
        }
        if (gtid == 32) {          // k0 = 512: partners inside the group, q <-> R - 1 - q
            float2 A[R];
            load_group_s<R>(s_tiles, s_comb, 512, A);
#pragma unroll
            for (int q = 0; q < R / 2; q++)
                split_store_s<R, MODE>(o, 512 + 1024 * q, A[q], A[R - 1 - q], s_post, scale);
        }
        if (++f == n_frames) {
            f = 0;
            c++;
        }
        group_barrier(1 + grp, GT);          // every Z_w has been read: the tiles are free again
        if (item + 1 < item_end) fetch(c, f);
    }
}

// ------------------------------------------------------------------------------------------
// Generic path: one CTA per frame, Stockham autosort in shared memory.
__device__ __forceinline__ void stockham_radix4(const float2 *__restrict__ in,
                                                float2 *__restrict__ outb,
                                                const float2 *__restrict__ tw, int M, int Ns,
                                                int tid, int nthreads) {
    const int quarter = M >> 2;
    const int tw_step = M / (Ns * 4);
    for (int j = tid; j < quarter; j += nthreads) {
        const int k = j & (Ns - 1);
        float2 a0 = in[j], a1 = in[j + quarter], a2 = in[j + 2 * quarter],
               a3 = in[j + 3 * quarter];
        if (Ns > 1) {
            a1 = cmul(a1, tw[k * tw_step]);
            a2 = cmul(a2, tw[2 * k * tw_step]);
            a3 = cmul(a3, tw[3 * k * tw_step]);
        }
        const float2 s02 = cadd(a0, a2), d02 = csub(a0, a2);
        const float2 s13 = cadd(a1, a3), d13 = mul_mj(csub(a1, a3));
        const int j0 = ((j - k) << 2) + k;
        outb[j0] = cadd(s02, s13);
        outb[j0 + Ns] = cadd(d02, d13);
        outb[j0 + 2 * Ns] = csub(s02, s13);
        outb[j0 + 3 * Ns] = csub(d02, d13);
    }
}

__global__ void stft_generic_kernel(const float *__restrict__ x, long long x_stride,
                                    long long n_frames, int hop, float *__restrict__ out,
                                    long long out_stride_c, long long out_stride_f,
                                    const float *__restrict__ win,
                                    const float2 *__restrict__ tw,
                                    const float2 *__restrict__ post, int mode, int n_fft,
                                    int log2m) {
    extern __shared__ float2 smem[];
    const int M = n_fft >> 1;
    float2 *buf0 = smem;
    float2 *buf1 = smem + M;
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const long long item = blockIdx.x;
    const long long c = item / n_frames;
    const long long f = item - c * n_frames;
    const float *p = x + c * x_stride + f * hop;

    for (int n = tid; n < M; n += nthreads)
        buf0[n] = make_float2(__ldg(p + 2 * n) * __ldg(win + 2 * n),
                              __ldg(p + 2 * n + 1) * __ldg(win + 2 * n + 1));
    __syncthreads();

    float2 *in = buf0, *ob = buf1;
    int Ns = 1;
    if (log2m & 1) {   // one radix-2 pass first (no twiddles at Ns = 1)
        const int halfm = M >> 1;
        for (int j = tid; j < halfm; j += nthreads) {
            const float2 a = in[j], b = in[j + halfm];
            ob[2 * j] = cadd(a, b);
            ob[2 * j + 1] = csub(a, b);
        }
        __syncthreads();
        float2 *t = in; in = ob; ob = t;
        Ns = 2;
    }
    for (; Ns < M; Ns <<= 2) {
        stockham_radix4(in, ob, tw, M, Ns, tid, nthreads);
        __syncthreads();
        float2 *t = in; in = ob; ob = t;
    }
    // `in` now holds Z[0..M)
    const float scale = 1.0f / (4.0f * (float)n_fft * (float)n_fft);
    float *o = out + c * out_stride_c + f * out_stride_f;
    for (int k = tid; k <= M / 2; k += nthreads) {
        const float2 z = in[k];
        const float2 zp = in[(M - k) & (M - 1)];
        const float2 E = make_float2(z.x + zp.x, z.y - zp.y);
        const float2 O = make_float2(z.x - zp.x, z.y + zp.y);
        const float2 T = cmul(O, post[k]);
        const float2 X = cadd(E, T);
        const float2 Y = csub(E, T);
        const float p1 = fmaf(X.x, X.x, X.y * X.y), p2 = fmaf(Y.x, Y.x, Y.y * Y.y);
        if (mode == FRT_STFT_POWER) {
            o[k] = finish<FRT_STFT_POWER>(p1, scale);
            o[M - k] = finish<FRT_STFT_POWER>(p2, scale);
        } else {
            o[k] = finish<FRT_STFT_LOGPOWER>(p1, scale);
            o[M - k] = finish<FRT_STFT_LOGPOWER>(p2, scale);
        }
    }
}

template <int MODE, int VEC>
cudaError_t set_fast_smem_one() {
    return cudaFuncSetAttribute(stft2048_kernel<MODE, VEC>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FAST_SMEM);
}
cudaError_t set_fast_smem() {
    cudaError_t e = set_fast_smem_one<0, 0>();
    if (e == cudaSuccess) e = set_fast_smem_one<0, 1>();
    if (e == cudaSuccess) e = set_fast_smem_one<0, 2>();
    if (e == cudaSuccess) e = set_fast_smem_one<1, 0>();
    if (e == cudaSuccess) e = set_fast_smem_one<1, 1>();
    if (e == cudaSuccess) e = set_fast_smem_one<1, 2>();
    return e;
}

template <int MODE, int VEC>
void launch_fast(unsigned blocks, cudaStream_t st, const float *x, long long x_stride,
                 long long n_frames, int hop, float *out, long long osc, long long osf,
                 const StftPlan &pl, long long total) {
    stft2048_kernel<MODE, VEC><<<blocks, FAST_WARPS * 32, FAST_SMEM, st>>>(
        x, x_stride, n_frames, hop, out, osc, osf, reinterpret_cast<const float2 *>(pl.win_dev),
        pl.tw_dev, pl.post_dev, pl.wlane_dev, total);
}

template <int Q>
cudaError_t set_small_smem() {
    cudaError_t e = cudaFuncSetAttribute(stft_small_kernel<Q, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)small_smem<Q>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(stft_small_kernel<Q, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)small_smem<Q>());
    return e;
}
template <int R>
cudaError_t set_large_smem() {
    cudaError_t e = cudaFuncSetAttribute(stft_large_kernel<R, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)large_smem<R>());
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(stft_large_kernel<R, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)large_smem<R>());
    return e;
}

template <int Q>
void launch_small(int mode, unsigned blocks, cudaStream_t st, const float *x, long long x_stride,
                  long long n_frames, int hop, float *out, long long osc, long long osf,
                  const StftPlan &pl, long long total, int vec_ok) {
    const float2 *w2 = reinterpret_cast<const float2 *>(pl.win_dev);
    if (mode == FRT_STFT_POWER)
        stft_small_kernel<Q, FRT_STFT_POWER><<<blocks, SMALL_WARPS * 32, small_smem<Q>(), st>>>(
            x, x_stride, n_frames, hop, out, osc, osf, w2, pl.tw_dev, pl.post_dev, total, vec_ok);
    else
        stft_small_kernel<Q, FRT_STFT_LOGPOWER><<<blocks, SMALL_WARPS * 32, small_smem<Q>(), st>>>(
            x, x_stride, n_frames, hop, out, osc, osf, w2, pl.tw_dev, pl.post_dev, total, vec_ok);
}
template <int R>
void launch_large(int mode, unsigned blocks, cudaStream_t st, const float *x, long long x_stride,
                  long long n_frames, int hop, float *out, long long osc, long long osf,
                  const StftPlan &pl, long long total) {
    const float4 *wl = reinterpret_cast<const float4 *>(pl.wlane_dev);
    HannA ha;
    const double theta = 2.0 * 3.14159265358979323846 / (double)(2048 * R - 1);
    for (int n1 = 0; n1 < 32; n1++) {
        const double a = theta * 64.0 * R * n1;
        ha.v[n1] = make_float2((float)(0.5 * cos(a)), (float)(0.5 * sin(a)));
    }
    if (mode == FRT_STFT_POWER)
        stft_large_kernel<R, FRT_STFT_POWER><<<blocks, LARGE_WARPS * 32, large_smem<R>(), st>>>(
            x, x_stride, n_frames, hop, out, osc, osf, wl, pl.tw_dev, pl.comb_dev, pl.post_dev, total, ha);
    else
        stft_large_kernel<R, FRT_STFT_LOGPOWER><<<blocks, LARGE_WARPS * 32, large_smem<R>(), st>>>(
            x, x_stride, n_frames, hop, out, osc, osf, wl, pl.tw_dev, pl.comb_dev, pl.post_dev, total, ha);
}

size_t multi_smem(int r) { return sizeof(float2) * (FAST_M + (size_t)r * FAST_TILE); }

int ilog2(int v) {
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

}   // namespace

extern "C" int frt_stft_plan(frt_handle h, int n_fft) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_fft >= 32 && n_fft <= 16384 && (n_fft & (n_fft - 1)) == 0,
                  "n_fft must be a power of two in [32, 16384]");
    if (h->stft.n_fft == n_fft) return FRT_OK;
    // plans are cached per size: the spectrum (8192) and spectrogram (4096) widgets share a handle
    auto cached = h->stft_cache.find(n_fft);
    if (cached != h->stft_cache.end()) {
        h->stft = cached->second;
        return FRT_OK;
    }
    StftPlan pl;
    // a failed cudaMalloc / cudaMemcpy below returns early: free what the half-built plan owns
    struct PlanGuard {
        StftPlan *p;
        ~PlanGuard() {
            if (!p) return;
            if (p->win_dev) cudaFree(p->win_dev);
            if (p->tw_dev) cudaFree(p->tw_dev);
            if (p->post_dev) cudaFree(p->post_dev);
            if (p->wlane_dev) cudaFree(p->wlane_dev);
            if (p->comb_dev) cudaFree(p->comb_dev);
        }
    } guard{&pl};
    const int N = n_fft, M = N / 2;
    const double PI = 3.14159265358979323846;
    // symmetric Hann, friture/audioproc.py:76-81
    pl.win_host.resize(N);
    for (int n = 0; n < N; n++)
        pl.win_host[n] = (float)(0.5 * (1.0 - cos(2.0 * PI * n / (double)(N - 1))));
    const int multi_r = (N == 4096) ? 2 : ((N == 8192) ? 4 : ((N == 16384) ? 8 : 0));
    std::vector<float2> tw(M), post(multi_r ? M + 33 : M / 2 + 1 + 32);
    if (multi_r) {
        tw.resize(FAST_M);
        for (int k1 = 0; k1 < 32; k1++)
            for (int t = 0; t < 32; t++) {
                const double a = -2.0 * PI * (double)((k1 * t) % FAST_M) / (double)FAST_M;
                tw[k1 * 32 + t] = make_float2((float)cos(a), (float)sin(a));
            }
        std::vector<float2> comb((size_t)(multi_r - 1) * 1024);
        for (int w = 1; w < multi_r; w++)
            for (int k0 = 0; k0 < 1024; k0++) {
                const double a = -2.0 * PI * (double)(w * k0) / (double)M;
                comb[(size_t)(w - 1) * 1024 + k0] = make_float2((float)cos(a), (float)sin(a));
            }
        FRT_CUDA(h, cudaMalloc(&pl.comb_dev, sizeof(float2) * comb.size()));
        FRT_CUDA(h, cudaMemcpy(pl.comb_dev, comb.data(), sizeof(float2) * comb.size(),
                               cudaMemcpyHostToDevice));
    } else if (N == FAST_N) {
        for (int k1 = 0; k1 < 32; k1++)
            for (int t = 0; t < 32; t++) {
                const double a = -2.0 * PI * (double)((k1 * t) % M) / (double)M;
                tw[k1 * 32 + t] = make_float2((float)cos(a), (float)sin(a));
            }
    } else if (N >= 64 && N <= 1024) {   // stft_small_kernel: [Q][32] W_M^(k1 t), M = 32 Q
        for (int k1 = 0; k1 < M / 32; k1++)
            for (int t = 0; t < 32; t++) {
                const double a = -2.0 * PI * (double)((k1 * t) % M) / (double)M;
                tw[k1 * 32 + t] = make_float2((float)cos(a), (float)sin(a));
            }
    } else {
        for (int i = 0; i < M; i++) {
            const double a = -2.0 * PI * (double)i / (double)M;
            tw[i] = make_float2((float)cos(a), (float)sin(a));
        }
    }
    for (size_t k = 0; k < post.size(); k++) {   // U[k] = -j * W_N^k
        const double a = 2.0 * PI * (double)k / (double)N;
        post[k] = make_float2((float)(-sin(a)), (float)(-cos(a)));
    }
    if (N == FAST_N) {   // per-lane part of the Hann phase: cos/sin(theta (2t + e)), t = 0..31
        std::vector<float2> wl(64);
        const double theta = 2.0 * PI / (double)(N - 1);
        for (int t = 0; t < 32; t++) {
            wl[t] = make_float2((float)cos(theta * (2 * t)), (float)cos(theta * (2 * t + 1)));
            wl[32 + t] = make_float2((float)sin(theta * (2 * t)), (float)sin(theta * (2 * t + 1)));
        }
        FRT_CUDA(h, cudaMalloc(&pl.wlane_dev, sizeof(float2) * 64));
        FRT_CUDA(h, cudaMemcpy(pl.wlane_dev, wl.data(), sizeof(float2) * 64,
                               cudaMemcpyHostToDevice));
    }
    if (N == 4096 || N == 8192) {   // stft_large_kernel: per-(warp, lane) part of the Hann phase
        const int R = N / 2048;
        std::vector<float> wl((size_t)R * 32 * 4);
        const double theta = 2.0 * PI / (double)(N - 1);
        for (int w = 0; w < R; w++)
            for (int t = 0; t < 32; t++) {
                const double b0 = theta * (2 * R * t + 2 * w), b1 = theta * (2 * R * t + 2 * w + 1);
                float *q = &wl[((size_t)w * 32 + t) * 4];
                q[0] = (float)cos(b0);
                q[1] = (float)cos(b1);
                q[2] = (float)sin(b0);
                q[3] = (float)sin(b1);
            }
        FRT_CUDA(h, cudaMalloc(&pl.wlane_dev, sizeof(float) * wl.size()));
        FRT_CUDA(h, cudaMemcpy(pl.wlane_dev, wl.data(), sizeof(float) * wl.size(), cudaMemcpyHostToDevice));
    }
    FRT_CUDA(h, cudaMalloc(&pl.win_dev, sizeof(float) * N));
    FRT_CUDA(h, cudaMalloc(&pl.tw_dev, sizeof(float2) * tw.size()));
    FRT_CUDA(h, cudaMalloc(&pl.post_dev, sizeof(float2) * post.size()));
    FRT_CUDA(h, cudaMemcpy(pl.win_dev, pl.win_host.data(), sizeof(float) * N,
                           cudaMemcpyHostToDevice));
    FRT_CUDA(h, cudaMemcpy(pl.tw_dev, tw.data(), sizeof(float2) * tw.size(),
                           cudaMemcpyHostToDevice));
    FRT_CUDA(h, cudaMemcpy(pl.post_dev, post.data(), sizeof(float2) * post.size(),
                           cudaMemcpyHostToDevice));
    if (N == FAST_N) {
        FRT_CUDA(h, set_fast_smem());
    } else if (multi_r) {
        FRT_CUDA(h, cudaFuncSetAttribute(stft_multi_kernel<2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)multi_smem(2)));
        FRT_CUDA(h, cudaFuncSetAttribute(stft_multi_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)multi_smem(2)));
        FRT_CUDA(h, cudaFuncSetAttribute(stft_multi_kernel<4, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)multi_smem(4)));
        FRT_CUDA(h, cudaFuncSetAttribute(stft_multi_kernel<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)multi_smem(4)));
        FRT_CUDA(h, cudaFuncSetAttribute(stft_multi_kernel<8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)multi_smem(8)));
        FRT_CUDA(h, cudaFuncSetAttribute(stft_multi_kernel<8, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)multi_smem(8)));
        FRT_CUDA(h, set_large_smem<2>());
        FRT_CUDA(h, set_large_smem<4>());
    } else if (N >= 64 && N <= 1024) {
        FRT_CUDA(h, set_small_smem<1>());
        FRT_CUDA(h, set_small_smem<2>());
        FRT_CUDA(h, set_small_smem<4>());
        FRT_CUDA(h, set_small_smem<8>());
        FRT_CUDA(h, set_small_smem<16>());
    } else {
        FRT_CUDA(h, cudaFuncSetAttribute(stft_generic_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(float2) * 2 * 8192)));
    }
    pl.n_fft = n_fft;
    guard.p = nullptr;          // the cache owns the device tables from here on
    h->stft_cache[n_fft] = pl;
    h->stft = pl;
    return FRT_OK;
}

extern "C" int frt_stft_window(frt_handle h, float *window_host) {
    if (!h) return FRT_EINVAL;
    if (!h->stft.n_fft) return frt_fail(h, FRT_ESTATE, "frt_stft_window: no plan");
    FRT_CHECK_ARG(h, window_host != nullptr, "window_host is NULL");
    memcpy(window_host, h->stft.win_host.data(), sizeof(float) * h->stft.n_fft);
    return FRT_OK;
}

extern "C" int frt_stft_process(frt_handle h, const float *x_dev, int64_t x_stride,
                                int n_channels, int64_t n_frames, int hop, float *out_dev,
                                int64_t out_stride_c, int64_t out_stride_f, int mode,
                                void *stream) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    const StftPlan &pl = h->stft;
    if (!pl.n_fft) return frt_fail(h, FRT_ESTATE, "frt_stft_process: call frt_stft_plan first");
    FRT_CHECK_ARG(h, n_channels >= 0 && n_frames >= 0, "negative shape");
    FRT_CHECK_ARG(h, hop >= 1, "hop must be >= 1");
    FRT_CHECK_ARG(h, mode == FRT_STFT_POWER || mode == FRT_STFT_LOGPOWER, "unknown mode");
    if (n_channels == 0 || n_frames == 0) return FRT_OK;   // empty input: nothing to do
    FRT_CHECK_ARG(h, x_dev != nullptr && out_dev != nullptr, "NULL buffer");
    const int nbins = pl.n_fft / 2 + 1;
    FRT_CHECK_ARG(h, out_stride_f >= nbins, "out_stride_f smaller than n_fft/2+1");
    const long long total = (long long)n_channels * n_frames;
    cudaStream_t st = (cudaStream_t)stream;
    if (pl.n_fft == FAST_N) {
        int vec_ok = (((uintptr_t)x_dev & 7) == 0) && ((x_stride & 1) == 0) && ((hop & 1) == 0);
        // 16-byte aligned frames: let the TMA engine stage them (cp.async.bulk)
        static const int no_tma = getenv("FRT_STFT_NO_TMA") ? 1 : 0;   // tuning knob
        if (!no_tma && (((uintptr_t)x_dev & 15) == 0) && ((x_stride & 3) == 0) && ((hop & 3) == 0))
            vec_ok = 2;
        long long blocks = (total + FAST_WARPS - 1) / FAST_WARPS;
        if (blocks > h->sm_count) blocks = h->sm_count;
#define FRT_LAUNCH_FAST(MODE, VEC)                                                         \
    launch_fast<MODE, VEC>((unsigned)blocks, st, x_dev, x_stride, n_frames, hop, out_dev,   \
                           out_stride_c, out_stride_f, pl, total)
        if (mode == FRT_STFT_POWER) {
            if (vec_ok == 2) FRT_LAUNCH_FAST(FRT_STFT_POWER, 2);
            else if (vec_ok) FRT_LAUNCH_FAST(FRT_STFT_POWER, 1);
            else FRT_LAUNCH_FAST(FRT_STFT_POWER, 0);
        } else {
            if (vec_ok == 2) FRT_LAUNCH_FAST(FRT_STFT_LOGPOWER, 2);
            else if (vec_ok) FRT_LAUNCH_FAST(FRT_STFT_LOGPOWER, 1);
            else FRT_LAUNCH_FAST(FRT_STFT_LOGPOWER, 0);
        }
#undef FRT_LAUNCH_FAST
    } else if (pl.n_fft >= 64 && pl.n_fft <= 1024) {
        const int vec_ok = (((uintptr_t)x_dev & 7) == 0) && ((x_stride & 1) == 0) && ((hop & 1) == 0);
        const int q = pl.n_fft / 64, gframes = 32 / q;
        long long blocks = ((total + gframes - 1) / gframes + SMALL_WARPS - 1) / SMALL_WARPS;
        if (blocks > 2LL * h->sm_count) blocks = 2LL * h->sm_count;
#define FRT_LAUNCH_SMALL(QQ)                                                                    \
    launch_small<QQ>(mode, (unsigned)blocks, st, x_dev, x_stride, n_frames, hop, out_dev,        \
                     out_stride_c, out_stride_f, pl, total, vec_ok)
        switch (q) {
            case 1: FRT_LAUNCH_SMALL(1); break;
            case 2: FRT_LAUNCH_SMALL(2); break;
            case 4: FRT_LAUNCH_SMALL(4); break;
            case 8: FRT_LAUNCH_SMALL(8); break;
            default: FRT_LAUNCH_SMALL(16); break;
        }
#undef FRT_LAUNCH_SMALL
    } else if ((pl.n_fft == 4096 || pl.n_fft == 8192) && (((uintptr_t)x_dev & 7) == 0) &&
               ((x_stride & 1) == 0) && ((hop & 1) == 0) && !getenv("FRT_STFT_NO_LARGE")) {
        const int r = pl.n_fft / 2048, ng = LARGE_WARPS / r;
        long long blocks = (total + ng - 1) / ng;
        if (blocks > h->sm_count) blocks = h->sm_count;
        if (r == 2)
            launch_large<2>(mode, (unsigned)blocks, st, x_dev, x_stride, n_frames, hop, out_dev, out_stride_c,
                            out_stride_f, pl, total);
        else
            launch_large<4>(mode, (unsigned)blocks, st, x_dev, x_stride, n_frames, hop, out_dev, out_stride_c,
                            out_stride_f, pl, total);
    } else if (pl.n_fft == 4096 || pl.n_fft == 8192 || pl.n_fft == 16384) {
        const int r = pl.n_fft / 2048;
        const int vec_ok = (((uintptr_t)x_dev & 7) == 0) && ((x_stride & 1) == 0) && ((hop & 1) == 0);
        long long blocks = (long long)h->sm_count * (r == 8 ? 2 : (r == 4 ? 4 : 8));
        if (blocks > total) blocks = total;
        const float2 *w2 = reinterpret_cast<const float2 *>(pl.win_dev);
#define FRT_LAUNCH_MULTI(RR, MODE)                                                          \
    stft_multi_kernel<RR, MODE><<<(unsigned)blocks, 32 * RR, multi_smem(RR), st>>>(          \
        x_dev, x_stride, n_frames, hop, out_dev, out_stride_c, out_stride_f, w2, pl.tw_dev,  \
        pl.comb_dev, pl.post_dev, total, vec_ok)
        if (r == 2) {
            if (mode == FRT_STFT_POWER) FRT_LAUNCH_MULTI(2, FRT_STFT_POWER);
            else FRT_LAUNCH_MULTI(2, FRT_STFT_LOGPOWER);
        } else if (r == 4) {
            if (mode == FRT_STFT_POWER) FRT_LAUNCH_MULTI(4, FRT_STFT_POWER);
            else FRT_LAUNCH_MULTI(4, FRT_STFT_LOGPOWER);
        } else {
            if (mode == FRT_STFT_POWER) FRT_LAUNCH_MULTI(8, FRT_STFT_POWER);
            else FRT_LAUNCH_MULTI(8, FRT_STFT_LOGPOWER);
        }
#undef FRT_LAUNCH_MULTI
    } else {
        FRT_CHECK_ARG(h, total <= 0x7fffffffLL, "too many frames for one launch");
        const int M = pl.n_fft / 2;
        int threads = M / 4;
        if (threads < 32) threads = 32;
        if (threads > 512) threads = 512;
        stft_generic_kernel<<<(unsigned)total, threads, sizeof(float2) * 2 * M, st>>>(
            x_dev, x_stride, n_frames, hop, out_dev, out_stride_c, out_stride_f, pl.win_dev,
            pl.tw_dev, pl.post_dev, mode, pl.n_fft, ilog2(M));
    }
    h->launches++;
    FRT_CUDA(h, cudaGetLastError());
    return FRT_OK;
}

extern "C" int frt_stft_process_host(frt_handle h, const float *x_host, int64_t x_stride,
                                     int n_channels, int64_t n_samples, int hop,
                                     float *out_host, int mode) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    const StftPlan &pl = h->stft;
    if (!pl.n_fft) return frt_fail(h, FRT_ESTATE, "frt_stft_process_host: no plan");
    FRT_CHECK_ARG(h, n_channels >= 0 && n_samples >= 0 && hop >= 1, "bad shape");
    const int N = pl.n_fft, nbins = N / 2 + 1;
    if (n_channels == 0 || n_samples < N) return FRT_OK;
    FRT_CHECK_ARG(h, x_host != nullptr && out_host != nullptr, "NULL buffer");
    FRT_CHECK_ARG(h, x_stride >= n_samples, "x_stride smaller than n_samples");
    const int64_t frames = (n_samples - N) / hop + 1;
    // chunk = (channel group) x (frame range), sized to ~32 MiB of output per buffer
    const size_t budget = (size_t)32 << 20;
    int64_t fchunk = frames;
    const size_t out_per_frame = sizeof(float) * nbins;
    if ((size_t)fchunk * out_per_frame > budget) fchunk = (int64_t)(budget / out_per_frame);
    if (fchunk < 1) fchunk = 1;
    int cgroup = 1;
    if (fchunk == frames) {
        cgroup = (int)(budget / ((size_t)frames * out_per_frame));
        if (cgroup < 1) cgroup = 1;
        if (cgroup > n_channels) cgroup = n_channels;
    }
    const int64_t chunk_samples = (fchunk - 1) * (int64_t)hop + N;
    const size_t in_bytes = sizeof(float) * (size_t)cgroup * chunk_samples;
    const size_t out_bytes = (size_t)cgroup * fchunk * out_per_frame;
    int rc = frt_pipe_ensure(h, in_bytes, out_bytes);
    if (rc) return rc;
    HostPipe &p = h->pipe;
    int it = 0;
    for (int c0 = 0; c0 < n_channels; c0 += cgroup) {
        const int nc = (c0 + cgroup <= n_channels) ? cgroup : (n_channels - c0);
        for (int64_t f0 = 0; f0 < frames; f0 += fchunk, it++) {
            const int64_t nf = (f0 + fchunk <= frames) ? fchunk : (frames - f0);
            const int64_t ns = (nf - 1) * (int64_t)hop + N;
            const int b = it & 1;
            // the H2D may only overwrite d_in[b] after the kernel that read it has finished
            FRT_CUDA(h, cudaStreamWaitEvent(p.s_in, p.ev_cmp[b], 0));
            FRT_CUDA(h, cudaMemcpy2DAsync(p.d_in[b], sizeof(float) * ns,
                                          x_host + (size_t)c0 * x_stride + f0 * hop,
                                          sizeof(float) * x_stride, sizeof(float) * ns, nc,
                                          cudaMemcpyHostToDevice, p.s_in));
            FRT_CUDA(h, cudaEventRecord(p.ev_in[b], p.s_in));
            FRT_CUDA(h, cudaStreamWaitEvent(p.s_cmp, p.ev_in[b], 0));
            FRT_CUDA(h, cudaStreamWaitEvent(p.s_cmp, p.ev_out[b], 0));   // d_out[b] drained
            rc = frt_stft_process(h, p.d_in[b], ns, nc, nf, hop, p.d_out[b],
                                  nf * (int64_t)nbins, nbins, mode, p.s_cmp);
            if (rc) return rc;
            FRT_CUDA(h, cudaEventRecord(p.ev_cmp[b], p.s_cmp));
            FRT_CUDA(h, cudaStreamWaitEvent(p.s_out, p.ev_cmp[b], 0));
            FRT_CUDA(h, cudaMemcpy2DAsync(out_host + ((size_t)c0 * frames + f0) * nbins,
                                          sizeof(float) * frames * nbins, p.d_out[b],
                                          sizeof(float) * nf * nbins,
                                          sizeof(float) * nf * nbins, nc,
                                          cudaMemcpyDeviceToHost, p.s_out));
            FRT_CUDA(h, cudaEventRecord(p.ev_out[b], p.s_out));
        }
    }
    FRT_CUDA(h, cudaStreamSynchronize(p.s_out));
    FRT_CUDA(h, cudaStreamSynchronize(p.s_cmp));
    FRT_CUDA(h, cudaStreamSynchronize(p.s_in));
    return FRT_OK;
}
