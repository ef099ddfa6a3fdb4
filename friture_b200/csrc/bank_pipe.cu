// Lane-pipelined multi-rate IIR filterbank + exponential RMS for sm_100a (bpo = 1 or 3).
//
// Same numerics contract as bank.cu (the reference's IIR bank octave_filter_bank_decimation,
// friture/filter.py:86-118, recursion friture/signal/lfilter.py:131-139, [::2] of
// friture/signal/decimate.py:39-41, fused with exp_smoothed_value(y**2),
// friture/octavespectrum.py:104 / friture/signal/exp_smoothing.py:11-56), different schedule:
//
// A recursion is serial in time, so instead of splitting time (bank.cu: two passes + a scan per
// section) the SECTIONS are spread over lanes and every lane runs its sections serially over a chunk
// of CH samples per step, handing the chunk to the next lane of its chain through a shared-memory
// ring: the chain is a software pipeline, a lane works on the chunk its predecessor finished one
// step earlier.  Each sample of each section is computed exactly once, with the 4-operation
// normalised biquad (bank_internal.cuh).
//
// A CTA of THREE SPECIALISED WARPS serves NCH = 32 / NR channels (NR = bpo + 3 roles per channel;
// 5 channels for 1/3-octave banks; channel PAIRS when two channels are packed in float2 -> FFMA2):
//   warp 0  stage 0 (rate fs): lane = (channel, role), TWO CHAINED SECTIONS PER LANE (their
//           recurrences are independent, which gives the in-order issue two chains to interleave)
//             role r < bpo   band r: both band-pass sections, reads the stage input
//             role bpo + d   decimator sections 2d, 2d+1 (d = 0, 1, 2); d = 0 reads the stage input
//           state in registers for the whole launch, no control flow in the step.
//   warp 1  the same roles for ALL lower-rate stages, time-multiplexed: of the CH sample slots of a
//           step, slots [CH-2 len_j, CH-len_j) belong to stage j (len_j = CH >> j; their states are
//           named registers) and the last slot to the one stage >= JR = log2(CH)+1 whose turn it
//           is (stage JR + ctz(u+1): a binary-ruler schedule, every stage gets exactly its 2^-j
//           share; state rows in shared memory) -- 31/32 of the slots carry work.
//           The last decimator lane of either warp keeps the even samples (x chain gain) and writes
//           them where the next stage's chain heads read them one step later (slot i -> CH/2 + i/2
//           of the multiplexed input vector; single samples of the ruler stages wait in a
//           double-buffered per-stage mailbox).
//   warp 2  everything that is not a recursion, one step behind the section warps: smoothing of y^2
//           with one accumulator per sample slot (lane = slot: acc <- q^len acc + y^2, three
//           instructions per band and step instead of two per sample), the weighted warp
//           reduction of a stage's accumulators into its band energy at a block end (collapsed
//           back to one value, so the state carried between launches is the plain smoothed energy
//           and a stream gives the same bits whether it arrives in one launch or block by block),
//           staging of the band vector in shared memory and ONE coalesced store per block once the
//           slowest stage has reported, and the cp.async prefetch of the input ring.
//           Decays are applied in complement form (acc - (1-q^n) acc): the float32 rounding of q
//           must not bias long time constants.
//   One __syncthreads() per step separates producers and consumers; all buffers that cross it
//   are double-buffered.
//
// tests/bank_pipeline_model.py is an executable NumPy model of exactly this schedule, checked
// against the oracle on CPU; this file mirrors it phase by phase.
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "bank_internal.cuh"

namespace {

constexpr int DEC_DEPTH = 3;     // decimator chain = 3 lanes: a stage trails its parent by 3 steps

template <int PACK> struct VT;
template <> struct VT<1> { using t = float; };
template <> struct VT<2> { using t = float2; };

__device__ __forceinline__ float v_add(float a, float b) { return a + b; }
__device__ __forceinline__ float2 v_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float v_fma(float c, float x, float z) { return fmaf(c, x, z); }
__device__ __forceinline__ float2 v_fma(float c, float2 x, float2 z) {
    return __ffma2_rn(make_float2(c, c), x, z);
}
__device__ __forceinline__ float v_mul(float c, float x) { return c * x; }
__device__ __forceinline__ float2 v_mul(float c, float2 x) { return __fmul2_rn(make_float2(c, c), x); }
__device__ __forceinline__ float v_sq(float a) { return a * a; }
__device__ __forceinline__ float2 v_sq(float2 a) { return __fmul2_rn(a, a); }
__device__ __forceinline__ void v_set(float &d, float a, float) { d = a; }
__device__ __forceinline__ void v_set(float2 &d, float a, float b) { d = make_float2(a, b); }
__device__ __forceinline__ float v_x(float a) { return a; }
__device__ __forceinline__ float v_x(float2 a) { return a.x; }
__device__ __forceinline__ float v_y(float a) { return a; }
__device__ __forceinline__ float v_y(float2 a) { return a.y; }
__device__ __forceinline__ float v_sel(bool p, float a, float b) { return p ? a : b; }
__device__ __forceinline__ float2 v_sel(bool p, float2 a, float2 b) {
    return make_float2(p ? a.x : b.x, p ? a.y : b.y);
}
__device__ __forceinline__ float v_shfl_xor(float a, int d) { return __shfl_xor_sync(0xffffffffu, a, d); }
__device__ __forceinline__ float2 v_shfl_xor(float2 a, int d) {
    return make_float2(__shfl_xor_sync(0xffffffffu, a.x, d), __shfl_xor_sync(0xffffffffu, a.y, d));
}

__device__ __forceinline__ void ld4(const float *p, float (&v)[4]) {
    const float4 t = *reinterpret_cast<const float4 *>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const float2 *p, float2 (&v)[4]) {
    const float4 a = reinterpret_cast<const float4 *>(p)[0];
    const float4 b = reinterpret_cast<const float4 *>(p)[1];
    v[0] = make_float2(a.x, a.y); v[1] = make_float2(a.z, a.w);
    v[2] = make_float2(b.x, b.y); v[3] = make_float2(b.z, b.w);
}
__device__ __forceinline__ void st4(float *p, const float (&v)[4]) {
    *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(float2 *p, const float2 (&v)[4]) {
    reinterpret_cast<float4 *>(p)[0] = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
    reinterpret_cast<float4 *>(p)[1] = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
}
__device__ __forceinline__ void st2(float *p, float a, float b) {
    *reinterpret_cast<float2 *>(p) = make_float2(a, b);
}
__device__ __forceinline__ void st2(float2 *p, float2 a, float2 b) {
    *reinterpret_cast<float4 *>(p) = make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ float lg2_fast(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// The lane groups of a warp must run ONE instruction stream with per-lane predicates; passing a
// loop-invariant flag through an empty asm keeps the compiler from unswitching on it.
__device__ __forceinline__ bool opq(bool p) {
    int v = p ? 1 : 0;
    asm volatile("" : "+r"(v));
    return v != 0;
}

// normalised biquad step (bank_internal.cuh): y = x + z1; z1 = c x - a1 y + z2; z2 = x - a2 y
template <class T>
__device__ __forceinline__ T biquad(T x, T &z1, T &z2, float cc, float na1, float na2) {
    const T y = v_add(x, z1);
    const T t = v_fma(cc, x, z2);
    z1 = v_fma(na1, y, t);
    z2 = v_fma(na2, y, x);
    return y;
}

// Sample slots of a step are grouped in segments (= stages of the multiplexed lanes): segment
// g < LOGCH holds the CH >> (g+1) slots starting at CH - 2 (CH >> (g+1)), segment LOGCH is the
// ruler slot CH-1.
__host__ __device__ constexpr int ilog2c(int v) {
    return v >= 64 ? 6 : v >= 32 ? 5 : v >= 16 ? 4 : v >= 8 ? 3 : v >= 4 ? 2 : v >= 2 ? 1 : 0;
}
__host__ __device__ constexpr int seg_of_slot(int slot, int ch, int logch) {
    return slot < ch / 2 ? 0 : logch - ilog2c(ch - slot);
}

constexpr int EN_RING = 8;      // blocks of band energies staged in shared memory before the flush

// Shared memory of one channel (one channel pair when PACK == 2), in units of T.  128-bit accesses
// are served a quarter-warp (8 lanes x 16 B) at a time: writer lane (c, r) owns a region whose base
// is 16*(NR*c + r) = 16*lane (mod 128) bytes, so the 8 lanes of a quarter store to 8 different
// 16-byte bank groups.
template <int LOGCH, int PACK, int BPO>
struct PipeLayout {
    static constexpr int CH = 1 << LOGCH;
    static constexpr int NR = BPO + DEC_DEPTH;                // roles per lane group
    static constexpr int NCH = 32 / NR;                       // channels per CTA
    static constexpr int TB = 4 * PACK;                       // bytes per T
    static constexpr int PAD = 16 / TB;                       // 16 bytes
    static constexpr int XSLOT = 2 * CH + 8 * PAD;            // ring slot: 128 B of skew + 2 vectors
    static constexpr int X0 = 5 * PAD;                        // stage-0 chunk within a slot (group 5)
    static constexpr int XM = X0 + CH + PAD;                  // multiplexed vector (group 6)
    static constexpr int X = 0;                               // [RX][XSLOT]
    static constexpr int W = X + PIPE_RX * XSLOT;             // [2 NR writer lanes][2 bufs][CH] + 16 B each
    static constexpr int WSTR = 2 * CH + PAD;
    static constexpr int SR = W + 2 * NR * WSTR;              // [MAX_OCT][NR][4]: ruler-stage section states
    static constexpr int ER = SR + BANK_MAX_OCT * NR * 4;     // [MAX_OCT][4]: ruler-stage energies
    static constexpr int MB = ER + BANK_MAX_OCT * 4;          // [MAX_OCT + 2][2] mailboxes
    static constexpr int EN = MB + (BANK_MAX_OCT + 2) * 2;    // [EN_RING][32] staged band energies
    static constexpr int ACC = EN + EN_RING * 32;             // [2 vectors][BPO][CH] smoothing accumulators
    static constexpr int RAW = ACC + 2 * BPO * CH;
    // channel stride = NR*16 (mod 128) bytes
    static constexpr int RAWB = RAW * TB;
    static constexpr int WANT = (NR * 16) % 128;
    static constexpr int STRB = ((RAWB - WANT + 127) / 128) * 128 + WANT;
    static constexpr int STR = STRB / TB;
};

template <int LOGCH, int PACK, int BPO>
__global__ void __launch_bounds__(96)
bank_pipe_kernel(const __grid_constant__ PipeParams P, const BankArgs a) {
    using T = typename VT<PACK>::t;
    using LY = PipeLayout<LOGCH, PACK, BPO>;
    constexpr int CH = 1 << LOGCH, JR = LOGCH + 1, NR = LY::NR, NSEC = 2 * NR, NCH = LY::NCH;
    constexpr int NSL = CH / 32;              // sample slots per lane of the smoothing warp
    constexpr int NG = CH / 4;                // 4-sample groups per step
    constexpr int GB = (PACK == 1 ? 32 : 16) / 4;   // groups whose inputs are loaded ahead
    constexpr int RX = PIPE_RX, PF = PIPE_PF;
    static_assert(LOGCH == 5 || LOGCH == 6, "steps of 32 or 64 samples");
    static_assert(NG % GB == 0, "whole load batches");
    constexpr unsigned FULL = 0xffffffffu;

    extern __shared__ __align__(128) float4 smem4[];
    T *smem = reinterpret_cast<T *>(smem4);
    int *sT = reinterpret_cast<int *>(smem + NCH * LY::STR);    // [MAX_OCT + 2]
    float *sAl = reinterpret_cast<float *>(sT + 12);            // [MAX_OCT + 2]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n_oct = P.n_oct;
    const int nbands = n_oct * BPO;
    const int n_chunks = (int)(a.t_total >> LOGCH);
    const int lognb = 31 - __clz(a.block) - LOGCH;          // block = CH << lognb
    const int nbmask = (1 << lognb) - 1;
    const int logblock = lognb + LOGCH;
    const int n_steps = a.n_steps;
    const int cta_ch0 = blockIdx.x * NCH;                   // first channel group of this CTA

    auto chan_alive = [&](int c) { return (cta_ch0 + c) * PACK < a.n_channels; };
    auto chan_first = [&](int c) { return chan_alive(c) ? (cta_ch0 + c) * PACK : 0; };
    auto chan_second = [&](int c) {
        const int f = chan_first(c);
        return (PACK == 2 && chan_alive(c) && f + 1 < a.n_channels) ? f + 1 : f;
    };

    // ---- prologue (all warps): tables, ruler-stage state, first input chunks
    if (tid <= BANK_MAX_OCT + 1) {
        sT[tid] = tid <= BANK_MAX_OCT ? P.T[tid] : 0x3fffffff;
        sAl[tid] = tid <= BANK_MAX_OCT ? P.alpha[tid] : 1.f;
    }
    for (int i = tid; i < NCH * BANK_MAX_OCT * NSEC * 2; i += 96) {     // sSR[c][j][r][4] == global z[j][2r..2r+1][2]
        const int c = i / (BANK_MAX_OCT * NSEC * 2), o = i - c * (BANK_MAX_OCT * NSEC * 2);
        T z;
        v_set(z, 0.f, 0.f);
        if (o < n_oct * NSEC * 2)
            v_set(z, a.zstate[(size_t)chan_first(c) * n_oct * NSEC * 2 + o],
                  a.zstate[(size_t)chan_second(c) * n_oct * NSEC * 2 + o]);
        smem[c * LY::STR + LY::SR + o] = z;
    }
    for (int i = tid; i < NCH * BANK_MAX_OCT * 4; i += 96) {
        const int c = i / (BANK_MAX_OCT * 4), o = i - c * (BANK_MAX_OCT * 4), j = o >> 2, b = o & 3;
        T e;
        v_set(e, 0.f, 0.f);
        if (j >= JR && j < n_oct && b < BPO)
            v_set(e, a.ema[(size_t)chan_first(c) * nbands + j * BPO + b],
                  a.ema[(size_t)chan_second(c) * nbands + j * BPO + b]);
        smem[c * LY::STR + LY::ER + o] = e;
    }
    const bool vec16 = (PACK == 1) && a.vec_ok;
    auto prefetch = [&](int cn) {          // smoothing warp: input chunk cn of every channel -> ring
        if (cn < n_chunks) {
            if (vec16) {
                for (int i = lane; i < NCH * (CH / 4); i += 32) {
                    const int c = i / (CH / 4), q = i - c * (CH / 4);
                    T *dst = smem + c * LY::STR + LY::X + (cn & (RX - 1)) * LY::XSLOT + LY::X0;
                    cp_async16(dst + 4 * q, a.x + (size_t)chan_first(c) * a.x_stride + (size_t)cn * CH + 4 * q);
                }
            } else {
                for (int i = lane; i < NCH * CH; i += 32) {
                    const int c = i / CH, p = i - c * CH;
                    float *dd = reinterpret_cast<float *>(smem + c * LY::STR + LY::X +
                                                          (cn & (RX - 1)) * LY::XSLOT + LY::X0 + p);
                    cp_async4(dd, a.x + (size_t)chan_first(c) * a.x_stride + (size_t)cn * CH + p);
                    if (PACK == 2)
                        cp_async4(dd + 1, a.x + (size_t)chan_second(c) * a.x_stride + (size_t)cn * CH + p);
                }
            }
        }
        cp_async_commit();
    };
    if (warp == 2) {
#pragma unroll
        for (int cn = 0; cn < PF; cn++) prefetch(cn);
        cp_async_wait<PF - 1>();
    }
    __syncthreads();

    // ---- section warps: lane = (channel c, role r)
    const int c = lane / NR, r = lane - c * NR;
    const bool worker = c < NCH;
    const bool isband = r < BPO;
    const int d = isband ? 0 : r - BPO;                      // skew = position in the decimator chain
    const bool isdec2 = (r == NR - 1);
    const bool ishead = (d == 0);                            // reads the stage input
    const float cA = P.c[2 * r], n1A = P.na1[2 * r], n2A = P.na2[2 * r];
    const float cB = P.c[2 * r + 1], n1B = P.na1[2 * r + 1], n2B = P.na2[2 * r + 1];
    const float gdec = P.gdec;
    T *sm = smem + (worker ? c : 0) * LY::STR;               // this lane's channel
    T *sX = sm + LY::X, *sW = sm + LY::W, *sSR = sm + LY::SR, *sER = sm + LY::ER, *sMB = sm + LY::MB,
      *sEN = sm + LY::EN;
    const int ch0 = chan_first(worker ? c : 0), ch1 = chan_second(worker ? c : 0);
    const bool alive = worker && chan_alive(c);
    const bool has2 = alive && ch1 != ch0;
    float *gz0 = a.zstate + (size_t)ch0 * n_oct * NSEC * 2;
    float *gz1 = a.zstate + (size_t)ch1 * n_oct * NSEC * 2;

    if (warp == 0) {
        // ============================================================ stage 0 (rate fs)
        const int maxstage = isband ? n_oct - 1 : n_oct - 2;     // the last stage's decimator is unused (filter.py:113)
        T zc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) v_set(zc[i], gz0[r * 4 + i], gz1[r * 4 + i]);
        for (int k = 0; k < n_steps; k++) {
            if (worker) {
                const int u = k - d;
                const bool pv0 = (unsigned)u < (unsigned)n_chunks && 0 <= maxstage;
                const T *inp = ishead ? sX + (k & (RX - 1)) * LY::XSLOT + LY::X0
                                      : sW + (r - 1) * LY::WSTR + ((k - 1) & 1) * CH;
                T *outp = sW + r * LY::WSTR + (k & 1) * CH;
                T *xn = sX + ((k + 1) & (RX - 1)) * LY::XSLOT + LY::XM;
                T cur[4], vn[4];
#pragma unroll
                for (int i = 0; i < 4; i++) cur[i] = zc[i];
                ld4(inp, vn);
                // a rolled loop over the 4-sample groups keeps the hot code in the instruction cache;
                // the next group's input is fetched while this one is computed
#pragma unroll 1
                for (int gq = 0; gq < NG; gq++) {
                    T v[4], y[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) v[i] = vn[i];
                    ld4(inp + 4 * ((gq + 1) & (NG - 1)), vn);
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        const T ya = biquad(v[i], cur[0], cur[1], cA, n1A, n2A);
                        y[i] = biquad(ya, cur[2], cur[3], cB, n1B, n2B);
                    }
                    // the last decimator lane keeps the even samples (friture/signal/decimate.py:41)
                    if (!isdec2) st4(outp + 4 * gq, y);
                    else st2(xn + 2 * gq, v_mul(gdec, y[0]), v_mul(gdec, y[2]));
                }
#pragma unroll
                for (int i = 0; i < 4; i++) zc[i] = v_sel(pv0, cur[i], zc[i]);
            }
            __syncthreads();
        }
        if (alive) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                gz0[r * 4 + i] = v_x(zc[i]);
                if (has2) gz1[r * 4 + i] = v_y(zc[i]);
            }
        }
    } else if (warp == 1) {
        // ============================================================ stages >= 1, time-multiplexed
        const int maxstage = isband ? n_oct - 1 : n_oct - 2;
        const float gb2 = isband ? P.gband[r] * P.gband[r] : 0.f;
        for (int k = 0; k < n_steps; k++) {
            if (worker) {
                const int u = k - d;
                const T *inp = ishead ? sX + (k & (RX - 1)) * LY::XSLOT + LY::XM
                                      : sW + (NR + r - 1) * LY::WSTR + ((k - 1) & 1) * CH;
                T *outp = sW + (NR + r) * LY::WSTR + (k & 1) * CH;
                T *xn = sX + ((k + 1) & (RX - 1)) * LY::XSLOT + LY::XM + CH / 2;
                // ruler slot: the stage >= JR whose sample is due
                const int jr = JR - 1 + __ffs(u + 1);
                const int jrc = jr < BANK_MAX_OCT - 1 ? jr : BANK_MAX_OCT - 1;
                const int Tj = sT[jrc];
                const int m = (u - Tj) >> (jrc - LOGCH);
                const bool pvR = jr <= maxstage && u >= Tj && m < (int)(a.t_total >> jrc);
                T *spr = sSR + (jrc * NR + r) * 4;
                T zr[4];
#pragma unroll
                for (int i = 0; i < 4; i++) zr[i] = spr[i];
                const T mb = sMB[jrc * 2 + (m & 1)];
                // section state: one row of sSR per stage; the next row is fetched a segment ahead
                T cur[4], nxt[4], vn[4];
                T *spc = sSR + (1 * NR + r) * 4;
#pragma unroll
                for (int i = 0; i < 4; i++) cur[i] = spc[i];
                bool pvc = (unsigned)(u - DEC_DEPTH) < (unsigned)n_chunks && 1 <= maxstage;
                ld4(inp, vn);
                int gq = 0;
#pragma unroll 1
                for (int g = 0; g <= LOGCH - 3; g++) {      // segments of 4+ samples: stage g+1, CH >> (g+1) slots
                    T *spn = sSR + ((g + 2) * NR + r) * 4;
#pragma unroll
                    for (int i = 0; i < 4; i++) nxt[i] = spn[i];
                    const int ng = CH >> (g + 3);
#pragma unroll 1
                    for (int q = 0; q < ng; q++, gq++) {
                        T v[4], y[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) v[i] = vn[i];
                        ld4(inp + 4 * (gq + 1), vn);
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            const T ya = biquad(v[i], cur[0], cur[1], cA, n1A, n2A);
                            y[i] = biquad(ya, cur[2], cur[3], cB, n1B, n2B);
                        }
                        if (!isdec2) st4(outp + 4 * gq, y);
                        else st2(xn + 2 * gq, v_mul(gdec, y[0]), v_mul(gdec, y[2]));
                    }
                    if (pvc) {
#pragma unroll
                        for (int i = 0; i < 4; i++) spc[i] = cur[i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; i++) cur[i] = nxt[i];
                    spc = spn;
                    pvc = (unsigned)(u - DEC_DEPTH * (g + 2)) < (unsigned)n_chunks && (g + 2) <= maxstage;
                }
                // last group: slots CH-4, CH-3 (stage LOGCH-1, state in cur), CH-2 (stage LOGCH), CH-1 (ruler)
                T y[4];
                {
                    T *spn = sSR + (LOGCH * NR + r) * 4;
#pragma unroll
                    for (int i = 0; i < 4; i++) nxt[i] = spn[i];
                    vn[3] = v_sel(ishead, mb, vn[3]);     // the ruler stage's sample comes from its mailbox
                    T ya = biquad(vn[0], cur[0], cur[1], cA, n1A, n2A);
                    y[0] = biquad(ya, cur[2], cur[3], cB, n1B, n2B);
                    ya = biquad(vn[1], cur[0], cur[1], cA, n1A, n2A);
                    y[1] = biquad(ya, cur[2], cur[3], cB, n1B, n2B);
                    if (pvc) {
#pragma unroll
                        for (int i = 0; i < 4; i++) spc[i] = cur[i];
                    }
                    const int cidxB = u - DEC_DEPTH * LOGCH;
                    const bool pvB = (unsigned)cidxB < (unsigned)n_chunks && LOGCH <= maxstage;
                    ya = biquad(vn[2], nxt[0], nxt[1], cA, n1A, n2A);
                    y[2] = biquad(ya, nxt[2], nxt[3], cB, n1B, n2B);
                    if (pvB) {
#pragma unroll
                        for (int i = 0; i < 4; i++) spn[i] = nxt[i];
                    }
                    ya = biquad(vn[3], zr[0], zr[1], cA, n1A, n2A);
                    y[3] = biquad(ya, zr[2], zr[3], cB, n1B, n2B);
                    if (pvR) {
#pragma unroll
                        for (int i = 0; i < 4; i++) spr[i] = zr[i];
                    }
                    if (!isdec2) {
                        st4(outp + CH - 4, y);
                    } else {
                        xn[CH / 2 - 2] = v_mul(gdec, y[0]);
                        if (pvB && !(cidxB & 1)) sMB[JR * 2 + ((cidxB >> 1) & 1)] = v_mul(gdec, y[2]);
                        if (pvR && !(m & 1)) sMB[(jrc + 1) * 2 + ((m >> 1) & 1)] = v_mul(gdec, y[3]);
                    }
                }
                if (isband) {
                    // exp_smoothed_value of the low-rate stages: e <- (1-alpha) e + y^2 (e/alpha form, raw units)
                    const float alj = sAl[jrc];
                    T e = sER[jrc * 4 + r];
                    e = v_add(v_fma(-alj, e, e), v_sq(y[3]));
                    if (pvR) {
                        sER[jrc * 4 + r] = e;
                        const int bl = logblock - jrc;           // block >> jrc = 2^bl samples
                        if (((m + 1) & ((1 << bl) - 1)) == 0) {
                            const int blk = ((m + 1) >> bl) - 1;
                            sEN[(blk & (EN_RING - 1)) * 32 + (n_oct - 1 - jrc) * BPO + r] = v_mul(alj * gb2, e);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (alive) {
            for (int j = 1; j < n_oct; j++) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const T z = sSR[(j * NR + r) * 4 + i];
                    gz0[(j * NR + r) * 4 + i] = v_x(z);
                    if (has2) gz1[(j * NR + r) * 4 + i] = v_y(z);
                }
                if (isband && j >= JR) {
                    const T e = sER[j * 4 + r];
                    a.ema[(size_t)ch0 * nbands + j * BPO + r] = v_x(e);
                    if (has2) a.ema[(size_t)ch1 * nbands + j * BPO + r] = v_y(e);
                }
            }
        }
    } else {
        // ============================================================ smoothing, energies, input prefetch
        // lane = sample slot (slot = lane + 32 q); one accumulator per slot, band, channel and vector,
        // kept in shared memory so that the channel loops can stay rolled
        int mst[NSL];
        bool mok[NSL], lastslot[NSL];
        float om0[NSL], aqm[NSL], omm[NSL];
#pragma unroll
        for (int q = 0; q < NSL; q++) {
            const int p = lane + 32 * q;
            const int j = 1 + __clz(~((unsigned)p << (32 - LOGCH)));     // stage owning slot p
            mst[q] = j;
            mok[q] = (j < JR) && (j <= n_oct - 1);
            lastslot[q] = (p == CH - (CH >> j) - 1);
            om0[q] = P.om0[q][lane];
            aqm[q] = P.aqm[q][lane];
            omm[q] = P.omm[q][lane];
#pragma unroll 1
            for (int cc = 0; cc < NCH; cc++) {
                const float *e0 = a.ema + (size_t)chan_first(cc) * nbands;
                const float *e1 = a.ema + (size_t)chan_second(cc) * nbands;
                T *ac = smem + cc * LY::STR + LY::ACC;
#pragma unroll
                for (int b = 0; b < BPO; b++) {
                    T z0, zm;
                    v_set(z0, 0.f, 0.f);
                    v_set(zm, 0.f, 0.f);
                    if (p == CH - 1) v_set(z0, e0[b], e1[b]);
                    if (mok[q] && lastslot[q]) v_set(zm, e0[j * BPO + b], e1[j * BPO + b]);
                    ac[b * CH + p] = z0;
                    ac[(BPO + b) * CH + p] = zm;
                }
            }
        }
        __syncwarp();
        const int fdelta = P.fdelta;
        for (int k = 0; k < n_steps; k++) {
            const int kk = k - 1;              // the section warps' step whose band outputs are smoothed now
            const bool valid0 = (unsigned)kk < (unsigned)n_chunks;
            bool vm[NSL];
#pragma unroll
            for (int q = 0; q < NSL; q++) vm[q] = mok[q] && (unsigned)(kk - DEC_DEPTH * mst[q]) < (unsigned)n_chunks;
            const int buf = (kk & 1) * CH;
#pragma unroll 1
            for (int cc = 0; cc < NCH; cc++) {
                const T *wb = smem + cc * LY::STR + LY::W + buf;
                T *ac = smem + cc * LY::STR + LY::ACC;
#pragma unroll
                for (int b = 0; b < BPO; b++) {
#pragma unroll
                    for (int q = 0; q < NSL; q++) {
                        const int p = lane + 32 * q;
                        const T y0 = wb[b * LY::WSTR + p];
                        const T ym = wb[(NR + b) * LY::WSTR + p];
                        const T a0 = ac[b * CH + p];
                        const T am = ac[(BPO + b) * CH + p];
                        if (valid0) ac[b * CH + p] = v_add(v_fma(-P.aq0, a0, a0), v_sq(y0));
                        if (vm[q]) ac[(BPO + b) * CH + p] = v_add(v_fma(-aqm[q], am, am), v_sq(ym));
                    }
                }
            }
            // ---- block ends (warp-uniform conditions: they depend on the step only)
            if (valid0 && (((kk + 1) & nbmask) == 0)) {     // stage 0: weighted sum of its CH accumulators
                const int blk = ((kk + 1) >> lognb) - 1;
#pragma unroll 1
                for (int cc = 0; cc < NCH; cc++) {
                    T *ac = smem + cc * LY::STR + LY::ACC;
                    T *en = smem + cc * LY::STR + LY::EN + (blk & (EN_RING - 1)) * 32 + (n_oct - 1) * BPO;
#pragma unroll
                    for (int b = 0; b < BPO; b++) {
                        T val;
                        v_set(val, 0.f, 0.f);
#pragma unroll
                        for (int q = 0; q < NSL; q++) {
                            const T a0 = ac[b * CH + lane + 32 * q];
                            val = v_add(val, v_fma(-om0[q], a0, a0));
                        }
#pragma unroll
                        for (int dlt = 16; dlt >= 1; dlt >>= 1) val = v_add(val, v_shfl_xor(val, dlt));
#pragma unroll
                        for (int q = 0; q < NSL; q++) {
                            T z;
                            v_set(z, 0.f, 0.f);
                            ac[b * CH + lane + 32 * q] = (lane + 32 * q == CH - 1) ? val : z;
                        }
                        if (lane == 0) en[b] = v_mul(P.alpha[0] * P.gband[b] * P.gband[b], val);
                    }
                }
            }
#pragma unroll
            for (int j = 1; j <= LOGCH; j++) {             // chunked lower-rate stages
                const int cm = kk - DEC_DEPTH * j;
                const bool ev = j <= n_oct - 1 && (unsigned)cm < (unsigned)n_chunks && (((cm + 1) & nbmask) == 0);
                if (!ev) continue;
                const int blk = ((cm + 1) >> lognb) - 1;
                const int len = CH >> j, lo = CH - 2 * len;          // slots [lo, lo + len)
                const int kb0 = (n_oct - 1 - j) * BPO;
                const float alj = sAl[j];
                const int ql = lo >> 5;                              // register (slot / 32) that holds the stage
                const int l0 = lo & 31;
                const bool mine = lane >= l0 && lane < l0 + len;
#pragma unroll 1
                for (int cc = 0; cc < NCH; cc++) {
                    T *ac = smem + cc * LY::STR + LY::ACC + BPO * CH + 32 * ql + lane;
                    T *en = smem + cc * LY::STR + LY::EN + (blk & (EN_RING - 1)) * 32 + kb0;
#pragma unroll
                    for (int b = 0; b < BPO; b++) {
                        T val;
                        v_set(val, 0.f, 0.f);
                        if (mine) {
                            const T am = ac[b * CH];
                            val = v_fma(-omm[ql], am, am);
                        }
#pragma unroll
                        for (int dlt = 1; dlt < 32; dlt <<= 1) {
                            if (dlt < len) val = v_add(val, v_shfl_xor(val, dlt));
                        }
                        if (mine) {
                            T z;
                            v_set(z, 0.f, 0.f);
                            ac[b * CH] = (lane == l0 + len - 1) ? val : z;
                            if (lane == l0) en[b] = v_mul(alj * P.gband[b] * P.gband[b], val);
                        }
                    }
                }
            }
            // ---- flush the band vector of the block every stage has reported (coalesced store)
            if (a.energies && k >= fdelta + nbmask + 1 && (((k - fdelta) & nbmask) == 0)) {
                const int blk = ((k - fdelta) >> lognb) - 1;
                __syncwarp();
                if (lane < nbands) {
                    const float w = (a.db && a.weight) ? __ldg(a.weight + lane) : 0.f;
#pragma unroll 1
                    for (int cc = 0; cc < NCH; cc++) {
                        if (chan_alive(cc)) {
                            const T v = smem[cc * LY::STR + LY::EN + (blk & (EN_RING - 1)) * 32 + lane];
                            float *o = a.energies + ((size_t)chan_first(cc) * a.n_blocks + blk) * nbands + lane;
                            float e0 = v_x(v), e1 = v_y(v);
                            if (a.db) {     // friture/octavespectrum.py:119-121: 10*log10(sp + 1e-30) + w
                                e0 = fmaf(3.01029995663981195f, lg2_fast(e0 + 1e-30f), w);
                                e1 = fmaf(3.01029995663981195f, lg2_fast(e1 + 1e-30f), w);
                            }
                            o[0] = e0;
                            if (PACK == 2 && chan_second(cc) != chan_first(cc)) o[(size_t)a.n_blocks * nbands] = e1;
                        }
                    }
                }
            }
            prefetch(k + PF);
            cp_async_wait<PF - 1>();           // chunk k+1 has landed
            __syncthreads();
        }
        // ---- epilogue: every chunked stage ended on a block boundary, its energy sits in one slot
        __syncwarp();
#pragma unroll
        for (int q = 0; q < NSL; q++) {
            const int p = lane + 32 * q;
#pragma unroll 1
            for (int cc = 0; cc < NCH; cc++) {
                if (!chan_alive(cc)) continue;
                float *e0 = a.ema + (size_t)chan_first(cc) * nbands;
                float *e1 = a.ema + (size_t)chan_second(cc) * nbands;
                const bool two = PACK == 2 && chan_second(cc) != chan_first(cc);
                const T *ac = smem + cc * LY::STR + LY::ACC;
#pragma unroll
                for (int b = 0; b < BPO; b++) {
                    if (p == CH - 1) {
                        const T v = ac[b * CH + p];
                        e0[b] = v_x(v);
                        if (two) e1[b] = v_y(v);
                    }
                    if (mok[q] && lastslot[q]) {
                        const T v = ac[(BPO + b) * CH + p];
                        e0[mst[q] * BPO + b] = v_x(v);
                        if (two) e1[mst[q] * BPO + b] = v_y(v);
                    }
                }
            }
        }
    }
}

template <int LOGCH, int PACK, int BPO>
cudaError_t launch_pipe(const PipeParams &P, const BankArgs &a, cudaStream_t st) {
    using LY = PipeLayout<LOGCH, PACK, BPO>;
    const size_t smem = (size_t)LY::TB * LY::NCH * LY::STR + 2 * 12 * 4 + 128;
    auto kern = bank_pipe_kernel<LOGCH, PACK, BPO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int per_cta = LY::NCH * PACK;
    const unsigned blocks = (unsigned)((a.n_channels + per_cta - 1) / per_cta);
    kern<<<blocks, 96, smem, st>>>(P, a);
    return cudaGetLastError();
}

template <int LOGCH, int PACK>
cudaError_t launch_pipe_bpo(const PipeParams &P, const BankArgs &a, cudaStream_t st) {
    if (P.bpo == 3) return launch_pipe<LOGCH, PACK, 3>(P, a, st);
    return launch_pipe<LOGCH, PACK, 1>(P, a, st);
}

}   // namespace

// Step at which the chain heads of stage j start (T[j]) and the number of steps that drains the
// pipeline for t_total samples per channel.  Stages j <= logch move a chunk of 2^logch >> j samples
// per step, DEC_DEPTH steps (the decimator chain) behind the previous stage; stage j > logch has one
// sample every P_j = 2^(j-logch) steps, at steps u = T_j (mod P_j) with T_j = P_j/2 - 1 (mod P_j),
// which makes the stages' turns disjoint (the ruler sequence JR + ctz(u+1)).
void frt_pipe_schedule(int n_oct, int logch, long long t_total, int *T, int *n_steps) {
    T[0] = 0;
    for (int j = 1; j <= BANK_MAX_OCT; j++) {
        int t = T[j - 1] + DEC_DEPTH;
        if (j > logch) {
            const int P = 1 << (j - logch), arem = P / 2 - 1;
            while (t % P != arem) t++;
        }
        T[j] = t;
    }
    if (!n_steps) return;
    const long long n_chunks = t_total >> logch;
    long long last = 0;
    for (int j = 0; j < n_oct; j++) {
        const int dmax = (j == n_oct - 1) ? 0 : DEC_DEPTH - 1;
        long long l;
        if (j <= logch) l = n_chunks - 1 + T[j] + dmax;
        else l = T[j] + (((t_total >> j) - 1) << (j - logch)) + dmax;
        if (l > last) last = l;
    }
    // the smoothing warp runs one step behind and flushes block b at step (b+1)*NB + delta
    const long long flush_last = n_chunks + frt_pipe_flush_delta(n_oct, logch, T);
    last = flush_last > last + 1 ? flush_last : last + 1;
    *n_steps = (int)(last + 1);
}

// Steps after the end of a block's last stage-0 chunk at which every stage has written the block's
// band energies to the staging ring: chunked stage j reports T_j steps late (plus the smoothing
// warp's lag of one step, which the flush shares), ruler stage j at step T_j - P_j of the block's
// last chunk (visible one step later).
int frt_pipe_flush_delta(int n_oct, int logch, const int *T) {
    int delta = 0;
    for (int j = 1; j < n_oct; j++) {
        const int dj = j <= logch ? T[j] : T[j] - (1 << (j - logch)) + 1;
        if (dj > delta) delta = dj;
    }
    return delta;
}

void frt_pipe_prepare(BankPlan *pl) {
    const BankParams &B = pl->params;
    pl->pipe_ok = (B.bpo == 1 || B.bpo == 3);
    if (!pl->pipe_ok) return;
    for (int v = 0; v < 2; v++) {
        const int logch = 5 + v, CH = 1 << logch;
        PipeParams &P = pl->pipe[v];
        memset(&P, 0, sizeof(P));
        P.n_oct = B.n_oct;
        P.bpo = B.bpo;
        for (int r = 0; r < B.nsec; r++) {
            P.c[r] = B.coef[r][1];
            P.na1[r] = -B.coef[r][3];
            P.na2[r] = -B.coef[r][4];
        }
        for (int b = 0; b < B.bpo; b++) P.gband[b] = B.gband[b];
        P.gdec = B.gdec;
        for (int j = 0; j <= BANK_MAX_OCT; j++) P.alpha[j] = j < B.n_oct ? (float)pl->alphas[j] : 1.f;
        frt_pipe_schedule(B.n_oct, logch, 0, P.T, nullptr);
        P.fdelta = frt_pipe_flush_delta(B.n_oct, logch, P.T);
        const double q0 = 1.0 - pl->alphas[0];
        P.aq0 = (float)(1.0 - pow(q0, CH));
        for (int p = 0; p < CH; p++) {
            P.om0[p >> 5][p & 31] = (float)(1.0 - pow(q0, CH - 1 - p));
            int j = 1;
            while (j <= logch && p >= CH - (CH >> j)) j++;
            if (j <= logch && j < B.n_oct) {
                const int len = CH >> j, pos = p - (CH - 2 * len);
                const double qj = 1.0 - pl->alphas[j];
                P.aqm[p >> 5][p & 31] = (float)(1.0 - pow(qj, len));
                P.omm[p >> 5][p & 31] = (float)(1.0 - pow(qj, len - 1 - pos));
            }
        }
    }
}

cudaError_t frt_pipe_launch(const BankPlan *pl, BankArgs a, int logch, int pack, cudaStream_t st) {
    const PipeParams &P = pl->pipe[logch - 5];
    int T[BANK_MAX_OCT + 1];
    frt_pipe_schedule(P.n_oct, logch, a.t_total, T, &a.n_steps);
    if (logch == 5) return pack == 2 ? launch_pipe_bpo<5, 2>(P, a, st) : launch_pipe_bpo<5, 1>(P, a, st);
    return pack == 2 ? launch_pipe_bpo<6, 2>(P, a, st) : launch_pipe_bpo<6, 1>(P, a, st);
}
