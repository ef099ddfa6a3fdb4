// Multi-rate fractional-octave IIR filterbank + exponential RMS for sm_100a.
//
// Stands behind Octave_Filters.filter (friture/octavefilters.py:49-58) with the numerics of the
// reference's IIR bank octave_filter_bank_decimation (friture/filter.py:86-118: per stage j,
// bands bpo-1..0 filter x_j, then x_{j+1} = lowpass(x_j)[::2], friture/signal/decimate.py:39-41),
// and behind the widget's per-band smoothing exp_smoothed_value(y**2)
// (friture/octavespectrum.py:104, friture/signal/exp_smoothing.py:11-56).  The (b, a) filters of
// friture/generated_filters.py are run as float32 second-order sections (direct form II
// transposed, the recursion of friture/signal/lfilter.py:131-139 per section).
//
// ONE WARP PER CHANNEL.  An IIR recursion is serial in time, so the time axis of a tile of
// 32*L0 samples is split over the 32 lanes (lane l owns samples [l*L, (l+1)*L) of the stage's
// signal, in registers) and every biquad section is run as
//   pass 1  zero-state recursion over the lane's chunk -> local final state f_l
//   scan    s_l = A^L s_{l-1} + f_l over the lanes (5 shuffle steps with A^(L*2^k), 2x2)
//   pass 2  the recursion again from the true incoming state -> outputs, in place.
// The decimated output stays in the same lane (L halves each stage).  Once a stage has only
// 32 samples left per tile (one per lane) the remaining low-rate stages switch to
// "lane = filter chain": lanes 0..bpo-1 run the band chains and lane bpo the decimator chain
// serially over the <=32 samples, which are broadcast with shuffles.
// Filter and smoothing state lives in shared memory while the kernel runs (global in between).
#include <cmath>
#include <cstdlib>

#include "bank_internal.cuh"

namespace {

constexpr int MAX_BPO = BANK_MAX_BPO;
constexpr int MAX_OCT = BANK_MAX_OCT;
constexpr int NQ = BANK_NQ;

__device__ __forceinline__ float lg2_fast(float v) {
    float r;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ float energy_out(float e, int db, const float *weight, int kband) {
    // friture/octavespectrum.py:119-121: 10*log10(sp + 1e-30) + w
    if (!db) return e;
    float v = 3.01029995663981195f * lg2_fast(e + 1e-30f);
    if (weight) v += __ldg(weight + kband);
    return v;
}

// offset of band k in the ragged per-channel y layout (bands concatenated, k = 0 lowest)
__device__ __forceinline__ long long y_offset(int k, int bpo, int n_oct, long long t_total) {
    // bands of stage j (dec 2^j) are k = (n_oct-1-j)*bpo + i; lower k = higher j = shorter
    const int jk = n_oct - 1 - k / bpo;        // stage of band k
    const int i = k - (n_oct - 1 - jk) * bpo;
    // sum over stages j' > jk of bpo * (T >> j')  +  i * (T >> jk)
    long long off = 0;
    for (int j = n_oct - 1; j > jk; j--) off += (long long)bpo * (t_total >> j);
    return off + (long long)i * (t_total >> jk);
}

template <bool WANT_Y>
struct WarpCtxT {
    static constexpr bool kWantY = WANT_Y;   // compile the ragged-y stores in or out
    const BankParams *P;
    float *s_z;       // [n_oct][nsec][2]   this warp's filter state (shared memory)
    float *s_e;       // [n_oct][bpo]       this warp's smoothing state (e / alpha form)
    const float *s_coef;   // [nsec][8] CTA copy of the coefficients for lane-varying access
    float *energies;  // this channel, this block: [nbands] or NULL when not a block end
    float *y;         // this channel's ragged y base or NULL
    long long t_total;
    long long t_off;  // sample offset of this tile within the launch
    int lane;
    int db;
    const float *weight;   // dB offsets per band (db mode) or NULL
};

// ---------------------------------------------------------------- scan-mode section
template <int L> struct Log2 { static constexpr int v = 1 + Log2<L / 2>::v; };
template <> struct Log2<1> { static constexpr int v = 0; };

template <int L, class W>
__device__ __forceinline__ void section_scan(const float (&in)[L], float (&out)[L],
                                             const W &w, int j, int sec) {
    const BankParams &P = *w.P;
    const float b0 = P.coef[sec][0], b1 = P.coef[sec][1], b2 = P.coef[sec][2];
    const float a1 = P.coef[sec][3], a2 = P.coef[sec][4];
    const float B1 = P.coef[sec][5], B2 = P.coef[sec][6];
    float *zp = w.s_z + (j * P.nsec + sec) * 2;
    const float c1 = zp[0], c2 = zp[1];
    // pass 1: state-only recursion from zero state
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < L; k++) {
        const float x = in[k];
        const float n1 = fmaf(-a1, s1, fmaf(B1, x, s2));
        s2 = fmaf(-a2, s1, B2 * x);
        s1 = n1;
    }
    constexpr int q0 = Log2<L>::v;
    if (w.lane == 0) {   // fold the carried state into lane 0's segment
        const float *A = P.apow[sec][q0];
        s1 += fmaf(A[0], c1, A[1] * c2);
        s2 += fmaf(A[2], c1, A[3] * c2);
    }
#pragma unroll
    for (int k = 0; k < 5; k++) {
        const float *A = P.apow[sec][q0 + k];
        const float t1 = __shfl_up_sync(0xffffffffu, s1, 1 << k);
        const float t2 = __shfl_up_sync(0xffffffffu, s2, 1 << k);
        if (w.lane >= (1 << k)) {
            s1 = fmaf(A[0], t1, fmaf(A[1], t2, s1));
            s2 = fmaf(A[2], t1, fmaf(A[3], t2, s2));
        }
    }
    float i1 = __shfl_up_sync(0xffffffffu, s1, 1);
    float i2 = __shfl_up_sync(0xffffffffu, s2, 1);
    if (w.lane == 0) {
        i1 = c1;
        i2 = c2;
    }
    // pass 2: direct form II transposed (friture/signal/lfilter.py:131-139, order 2)
#pragma unroll
    for (int k = 0; k < L; k++) {
        const float x = in[k];
        const float y = fmaf(b0, x, i1);
        i1 = fmaf(-a1, y, fmaf(b1, x, i2));
        i2 = fmaf(-a2, y, b2 * x);
        out[k] = y;
    }
    __syncwarp();
    if (w.lane == 31) {   // state after the tile's last sample
        zp[0] = i1;
        zp[1] = i2;
    }
}

template <class W> __device__ __forceinline__ void serial_stages(float xin, int n, const W &w, int j);

// Band chains of one scan-mode stage (bands bpo-1 .. 0, friture/filter.py:105) with the smoothing
// of y^2: L samples per lane, 32 lanes active.
template <int L, class W>
__device__ __forceinline__ void bands_of_stage(const float (&xc)[L], const W &w, int j) {
    const BankParams &P = *w.P;
    const int bpo = P.bpo;
    float wk[L];
    for (int i = bpo - 1; i >= 0; i--) {
        section_scan<L>(xc, wk, w, j, 2 * i);
        section_scan<L>(wk, wk, w, j, 2 * i + 1);
        const int kband = (P.n_oct - 1 - j) * bpo + i;
        const float gb = P.gband[i];     // chain gain of the normalised sections
        if constexpr (W::kWantY) {
            float *yp = w.y + y_offset(kband, bpo, P.n_oct, w.t_total) + (w.t_off >> j) +
                        w.lane * L;
#pragma unroll
            for (int k = 0; k < L; k++) yp[k] = wk[k] * gb;
        }
        // exponential smoothing of y^2 (friture/signal/exp_smoothing.py:11-56), e/alpha form, in the
        // units of the normalised sections (the squared chain gain is applied to the output);
        // decays in complement form e - (1-q^n) e: the rounding of q must not bias long time constants
        const float om = P.omq[j][0];
        float e = 0.f;
#pragma unroll
        for (int k = 0; k < L; k++) e = fmaf(-om, e, e) + wk[k] * wk[k];
        float *ep = w.s_e + j * bpo + i;
        constexpr int q0 = Log2<L>::v;
        if (w.lane == 0) e += fmaf(-P.omq[j][q0], ep[0], ep[0]);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            const float t = __shfl_up_sync(0xffffffffu, e, 1 << k);
            if (w.lane >= (1 << k)) e += fmaf(-P.omq[j][q0 + k], t, t);
        }
        __syncwarp();
        if (w.lane == 31) {
            ep[0] = e;
            if (w.energies) w.energies[kband] = energy_out(P.alpha[j] * gb * gb * e, w.db, w.weight, kband);
        }
    }
}

// Decimation low-pass of one scan-mode stage: 6 sections, full-rate output in wk
// (the caller keeps the even samples, friture/signal/decimate.py:39-41).
template <int L, class W>
__device__ __forceinline__ void dec_of_stage(const float (&xc)[L], float (&wk)[L], const W &w,
                                             int j) {
    const int bpo = w.P->bpo;
    section_scan<L>(xc, wk, w, j, 2 * bpo);
#pragma unroll 1
    for (int s = 1; s < 6; s++) section_scan<L>(wk, wk, w, j, 2 * bpo + s);
}

// One scan-mode stage, single-warp variant: bands, then the decimator, then the next stage.
template <int L, class W>
__device__ void scan_stage(const float (&xc)[L], const W &w, int j) {
    const BankParams &P = *w.P;
    bands_of_stage<L>(xc, w, j);
    if (j + 1 >= P.n_oct) return;   // the last decimator's output is discarded (filter.py:113)
    float wk[L];
    dec_of_stage<L>(xc, wk, w, j);
    if constexpr (L >= 4) {
        float xn[L / 2];
#pragma unroll
        for (int k = 0; k < L / 2; k++) xn[k] = wk[2 * k] * P.gdec;
        scan_stage<L / 2>(xn, w, j + 1);
    } else {
        // L == 2: one sample per lane is left -> switch to lane = chain
        serial_stages<W>(wk[0] * P.gdec, 32, w, j + 1);
    }
}

// ---- three-warp variant (one CTA of 96 threads per channel) --------------------------------
// warp 0 "D": decimator chain of the scan-mode stages (the critical path to the next stage),
// warp 1 "B": band chains of the scan-mode stages, one barrier behind D per stage,
// warp 2 "S": the low-rate serial stages of the PREVIOUS tile.
// D hands the decimated chunks to B through shared memory; named barrier 1 (64 threads) paces
// D and B stage by stage, __syncthreads() closes a tile.
__device__ __forceinline__ void bar_db() { asm volatile("bar.sync 1, 64;" ::: "memory"); }

template <int L, class W>
__device__ void stage_D(const float (&xc)[L], const W &w, int j, float *s_x, float *s_x32) {
    float wk[L];
    dec_of_stage<L>(xc, wk, w, j);
    if constexpr (L >= 4) {
        float xn[L / 2];
        float *dst = s_x + w.lane * (L / 2);     // stage j+1 chunk of this lane
#pragma unroll
        for (int k = 0; k < L / 2; k++) {
            xn[k] = wk[2 * k] * w.P->gdec;
            dst[k] = xn[k];
        }
        bar_db();
        stage_D<L / 2>(xn, w, j + 1, s_x + 32 * (L / 2), s_x32);
    } else {
        s_x32[w.lane] = wk[0] * w.P->gdec;       // 32 samples for the serial stages (warp S)
        bar_db();
    }
}

template <int L, class W>
__device__ void stage_B(const float (&xc)[L], const W &w, int j, const float *s_x) {
    bands_of_stage<L>(xc, w, j);
    bar_db();
    if constexpr (L >= 4) {
        float xn[L / 2];
        const float *src = s_x + w.lane * (L / 2);
#pragma unroll
        for (int k = 0; k < L / 2; k++) xn[k] = src[k];
        stage_B<L / 2>(xn, w, j + 1, s_x + 32 * (L / 2));
    }
}

// Serial-mode stages: the n samples of stage j sit one per lane (sample m in lane m).  Lane
// r < bpo runs band chain r (2 sections), lane bpo runs the decimator chain (6 sections); samples
// are broadcast one by one.  Continues through the remaining stages (n halves every stage).  One
// rolled copy of this code serves all low-rate stages: instruction-cache footprint matters here.
template <class W>
__device__ __forceinline__ void serial_stages(float xin, int n, const W &w, int j) {
    const BankParams &P = *w.P;
    const int bpo = P.bpo, lane = w.lane;
    const bool is_band = lane < bpo;
#pragma unroll 1
    for (; j < P.n_oct && n >= 1; j++, n >>= 1) {
        const bool last = (j + 1 >= P.n_oct);
        const bool is_dec = (lane == bpo) && !last;
        const int nchain = is_band ? 2 : (is_dec ? 6 : 0);
        const int sec0 = is_band ? 2 * lane : 2 * bpo;
        float z1[6], z2[6], cb0[6], cb1[6], cb2[6], ca1[6], ca2[6];
#pragma unroll
        for (int s = 0; s < 6; s++) {
            const bool on = s < nchain;
            const float *cf = w.s_coef + (sec0 + (on ? s : 0)) * 8;
            cb0[s] = cf[0]; cb1[s] = cf[1]; cb2[s] = cf[2]; ca1[s] = cf[3]; ca2[s] = cf[4];
            const float *zp = w.s_z + (j * P.nsec + sec0 + (on ? s : 0)) * 2;
            z1[s] = zp[0];
            z2[s] = zp[1];
        }
        const float om = P.omq[j][0];
        const float gout = is_band ? P.gband[is_band ? lane : 0] : P.gdec;   // chain gain
        float e = is_band ? w.s_e[j * bpo + lane] : 0.f;
        const int kband = (P.n_oct - 1 - j) * bpo + (is_band ? lane : 0);
        float *yp = nullptr;
        if constexpr (W::kWantY) {
            if (is_band) yp = w.y + y_offset(kband, bpo, P.n_oct, w.t_total) + (w.t_off >> j);
        }
        float xnext = 0.f;
#pragma unroll 1
        for (int m = 0; m < n; m++) {
            float v = __shfl_sync(0xffffffffu, xin, m);
#pragma unroll
            for (int s = 0; s < 6; s++) {
                if (s < nchain) {
                    const float y = fmaf(cb0[s], v, z1[s]);
                    z1[s] = fmaf(-ca1[s], y, fmaf(cb1[s], v, z2[s]));
                    z2[s] = fmaf(-ca2[s], y, cb2[s] * v);
                    v = y;
                }
            }
            if (is_band) {
                e = fmaf(-om, e, e) + v * v;      // raw units; gain applied to the outputs
                if constexpr (W::kWantY) {
                    if (yp) yp[m] = v * gout;
                }
            } else {
                v *= gout;
            }
            if ((m & 1) == 0) {   // keep even samples (friture/signal/decimate.py:41)
                const float o = __shfl_sync(0xffffffffu, v, bpo);
                if (lane == (m >> 1)) xnext = o;
            }
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < 6; s++) {
            if (s < nchain) {
                float *zp = w.s_z + (j * P.nsec + sec0 + s) * 2;
                zp[0] = z1[s];
                zp[1] = z2[s];
            }
        }
        if (is_band) {
            w.s_e[j * bpo + lane] = e;
            if (w.energies) w.energies[kband] = energy_out(P.alpha[j] * gout * gout * e, w.db, w.weight, kband);
        }
        __syncwarp();
        xin = xnext;
    }
}

template <int L0, bool WANT_Y>
__global__ void __launch_bounds__(64)
bank_kernel(const __grid_constant__ BankParams P, const BankArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int warps = blockDim.x >> 5;
    const int nz = P.n_oct * P.nsec * 2, ne = P.n_oct * P.bpo;
    float *s_coef = smem;                        // [nsec][8]
    float *s_z = s_coef + P.nsec * 8 + wid * (nz + ne);
    float *s_e = s_z + nz;
    for (int i = threadIdx.x; i < P.nsec * 8; i += blockDim.x)
        s_coef[i] = P.coef[i >> 3][i & 7];
    const int c = blockIdx.x * warps + wid;
    const bool active = c < a.n_channels;
    if (active) {
        const float *gz = a.zstate + (size_t)c * nz;
        const float *ge = a.ema + (size_t)c * ne;
        for (int i = lane; i < nz; i += 32) s_z[i] = gz[i];
        for (int i = lane; i < ne; i += 32) s_e[i] = ge[i];   // e/alpha form: e = q e + y^2
    }
    __syncthreads();
    if (!active) return;

    constexpr int TILE = 32 * L0;
    const int nbands = P.n_oct * P.bpo;
    WarpCtxT<WANT_Y> w;
    w.P = &P;
    w.s_z = s_z;
    w.s_e = s_e;
    w.s_coef = s_coef;
    w.lane = lane;
    w.db = a.db;
    w.weight = a.weight;
    w.t_total = a.t_total;
    w.y = a.y ? a.y + (size_t)c * a.y_stride : nullptr;
    const float *xch = a.x + (size_t)c * a.x_stride;
    const int n_blocks = a.n_tiles / a.tiles_per_block;
    for (int t = 0; t < a.n_tiles; t++) {
        const float *xp = xch + (size_t)t * TILE + lane * L0;
        float xc[L0];
        if (a.vec_ok) {
#pragma unroll
            for (int k = 0; k < L0; k += 4) {
                const float4 v = __ldg(reinterpret_cast<const float4 *>(xp + k));
                xc[k] = v.x; xc[k + 1] = v.y; xc[k + 2] = v.z; xc[k + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < L0; k++) xc[k] = __ldg(xp + k);
        }
        const bool block_end = ((t + 1) % a.tiles_per_block) == 0;
        w.energies = (a.energies && block_end)
                         ? a.energies + (size_t)c * a.e_stride + (size_t)(t / a.tiles_per_block) * nbands
                         : nullptr;
        w.t_off = (long long)t * TILE;
        scan_stage<L0>(xc, w, 0);
        __syncwarp();
    }
    // write the state back
    float *gz = a.zstate + (size_t)c * nz;
    float *ge = a.ema + (size_t)c * ne;
    for (int i = lane; i < nz; i += 32) gz[i] = s_z[i];
    for (int i = lane; i < ne; i += 32) ge[i] = s_e[i];
}

template <int L0, bool WANT_Y>
__global__ void __launch_bounds__(96)
bank_kernel3(const __grid_constant__ BankParams P, const BankArgs a) {
    extern __shared__ float smem[];
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    const int nz = P.n_oct * P.nsec * 2, ne = P.n_oct * P.bpo;
    constexpr int TILE = 32 * L0;
    float *s_coef = smem;                        // [nsec][8]
    float *s_z = s_coef + P.nsec * 8;            // this channel's filter state
    float *s_e = s_z + nz;                       // this channel's smoothing state
    float *s_x = s_e + ne;                       // chunks of stages 1.. (TILE/2 + TILE/4 + ...)
    float *s_x32 = s_x + TILE;                   // [2][32] double-buffered input of the serial stages
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < P.nsec * 8; i += blockDim.x) s_coef[i] = P.coef[i >> 3][i & 7];
    {
        const float *gz = a.zstate + (size_t)c * nz;
        const float *ge = a.ema + (size_t)c * ne;
        for (int i = threadIdx.x; i < nz; i += blockDim.x) s_z[i] = gz[i];
        for (int i = threadIdx.x; i < ne; i += blockDim.x) s_e[i] = ge[i];
    }
    __syncthreads();
    const int nbands = P.n_oct * P.bpo;
    WarpCtxT<WANT_Y> w;
    w.P = &P;
    w.s_z = s_z;
    w.s_e = s_e;
    w.s_coef = s_coef;
    w.lane = lane;
    w.db = a.db;
    w.weight = a.weight;
    w.t_total = a.t_total;
    w.y = a.y ? a.y + (size_t)c * a.y_stride : nullptr;
    const float *xch = a.x + (size_t)c * a.x_stride;
    const int n_blocks = a.n_tiles / a.tiles_per_block;
    constexpr int NSCAN = Log2<L0>::v;           // scan-mode stages 0 .. NSCAN-1 (L = L0 .. 2)
    for (int t = 0; t <= a.n_tiles; t++) {
        if (role < 2 && t < a.n_tiles) {
            const float *xp = xch + (size_t)t * TILE + lane * L0;
            float xc[L0];
            if (a.vec_ok) {
#pragma unroll
                for (int k = 0; k < L0; k += 4) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(xp + k));
                    xc[k] = v.x; xc[k + 1] = v.y; xc[k + 2] = v.z; xc[k + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int k = 0; k < L0; k++) xc[k] = __ldg(xp + k);
            }
            const bool block_end = ((t + 1) % a.tiles_per_block) == 0;
            w.energies = (a.energies && block_end)
                             ? a.energies + (size_t)c * a.e_stride + (size_t)(t / a.tiles_per_block) * nbands
                             : nullptr;
            w.t_off = (long long)t * TILE;
            if (role == 0) stage_D<L0>(xc, w, 0, s_x, s_x32 + (t & 1) * 32);
            else stage_B<L0>(xc, w, 0, s_x);
        } else if (role == 2 && t >= 1) {
            const int tp = t - 1;
            const bool block_end = ((tp + 1) % a.tiles_per_block) == 0;
            w.energies = (a.energies && block_end)
                             ? a.energies + (size_t)c * a.e_stride + (size_t)(tp / a.tiles_per_block) * nbands
                             : nullptr;
            w.t_off = (long long)tp * TILE;
            if (NSCAN < P.n_oct) serial_stages<WarpCtxT<WANT_Y>>(s_x32[(tp & 1) * 32 + lane], 32, w, NSCAN);
        }
        __syncthreads();
    }
    float *gz = a.zstate + (size_t)c * nz;
    float *ge = a.ema + (size_t)c * ne;
    for (int i = threadIdx.x; i < nz; i += blockDim.x) gz[i] = s_z[i];
    for (int i = threadIdx.x; i < ne; i += blockDim.x) ge[i] = s_e[i];
}

// ---- stand-alone N-fold decimation (decimate_multiple, friture/signal/decimate.py:45-71) ------
// The same 6-section low-pass as the bank's decimator, run NDEC times with [::2] in between, one
// warp per channel, tiles of 256 samples (L = 8, 4, ... per lane).  Used by the delay estimator
// (48 kHz -> 12 kHz with Ndec = 2, friture/delay_estimator.py:53-59,97-98).
struct DecArgs {
    const float *x;
    long long x_stride;
    float *out;
    long long out_stride;
    int n_channels, n_tiles, n_stages;
    float *zstate;   // [C][n_stages][6][2]
};

template <int L, class W>
__device__ void dec_chain(const float (&xc)[L], const W &w, int j, int n_stages, float *dst) {
    float wk[L];
    dec_of_stage<L>(xc, wk, w, j);
    if (j + 1 == n_stages) {
#pragma unroll
        for (int k = 0; k < L / 2; k++) dst[w.lane * (L / 2) + k] = wk[2 * k] * w.P->gdec;
        return;
    }
    if constexpr (L >= 4) {
        float xn[L / 2];
#pragma unroll
        for (int k = 0; k < L / 2; k++) xn[k] = wk[2 * k] * w.P->gdec;
        dec_chain<L / 2>(xn, w, j + 1, n_stages, dst);
    }
}

__global__ void __launch_bounds__(64)
decimate_kernel(const __grid_constant__ BankParams P, const DecArgs a) {
    extern __shared__ float smem[];
    constexpr int L0 = 8, TILE = 256;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, warps = blockDim.x >> 5;
    const int nz = a.n_stages * P.nsec * 2;
    float *s_z = smem + wid * nz;
    const int c = blockIdx.x * warps + wid;
    if (c >= a.n_channels) return;
    float *gz = a.zstate + (size_t)c * nz;
    for (int i = lane; i < nz; i += 32) s_z[i] = gz[i];
    __syncwarp();
    WarpCtxT<false> w;
    w.P = &P;
    w.s_z = s_z;
    w.s_e = nullptr;
    w.s_coef = nullptr;
    w.energies = nullptr;
    w.y = nullptr;
    w.lane = lane;
    w.db = 0;
    w.weight = nullptr;
    w.t_total = 0;
    w.t_off = 0;
    const int out_tile = TILE >> a.n_stages;
    for (int t = 0; t < a.n_tiles; t++) {
        const float *xp = a.x + (size_t)c * a.x_stride + (size_t)t * TILE + lane * L0;
        float xc[L0];
#pragma unroll
        for (int k = 0; k < L0; k++) xc[k] = __ldg(xp + k);
        dec_chain<L0>(xc, w, 0, a.n_stages, a.out + (size_t)c * a.out_stride + (size_t)t * out_tile);
        __syncwarp();
    }
    for (int i = lane; i < nz; i += 32) gz[i] = s_z[i];
}

}   // namespace

void frt_bank_release(frt_ctx *h) {
    if (!h->bank) return;
    if (h->bank->zstate) cudaFree(h->bank->zstate);
    if (h->bank->ema) cudaFree(h->bank->ema);
    if (h->bank->weight) cudaFree(h->bank->weight);
    delete h->bank;
    h->bank = nullptr;
}

static void mat2_mul(const double a[4], const double b[4], double o[4]) {
    o[0] = a[0] * b[0] + a[1] * b[2];
    o[1] = a[0] * b[1] + a[1] * b[3];
    o[2] = a[2] * b[0] + a[3] * b[2];
    o[3] = a[2] * b[1] + a[3] * b[3];
}

extern "C" int frt_bank_plan(frt_handle h, int n_channels, int bands_per_octave, int n_octaves,
                             const double *sos_band, const double *sos_dec,
                             const double *alphas) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_channels >= 1, "n_channels must be >= 1");
    FRT_CHECK_ARG(h, bands_per_octave >= 1 && bands_per_octave <= MAX_BPO,
                  "bands_per_octave must be in [1, 24]");
    FRT_CHECK_ARG(h, n_octaves >= 1 && n_octaves <= MAX_OCT, "n_octaves must be in [1, 10]");
    FRT_CHECK_ARG(h, sos_band && sos_dec && alphas, "NULL coefficient table");
    frt_bank_release(h);
    BankPlan *pl = new (std::nothrow) BankPlan();
    if (!pl) return frt_fail(h, FRT_ENOMEM, "out of host memory");
    BankParams &P = pl->params;
    memset(&P, 0, sizeof(P));
    P.bpo = bands_per_octave;
    P.n_oct = n_octaves;
    P.nsec = 2 * bands_per_octave + 6;
    double gchain[BANK_MAX_BPO + 1];
    for (int i = 0; i <= bands_per_octave; i++) gchain[i] = 1.0;
    for (int sec = 0; sec < P.nsec; sec++) {
        const double *s = sec < 2 * bands_per_octave ? sos_band + (size_t)sec * 6
                                                     : sos_dec + (size_t)(sec - 2 * bands_per_octave) * 6;
        if (fabs(s[3] - 1.0) > 1e-12) {
            delete pl;
            return frt_fail(h, FRT_EINVAL, "SOS section %d has a0 != 1", sec);
        }
        // elliptic sections have their zeros on the unit circle: b = g (1, c, 1)
        if (!(fabs(s[0]) > 0.0) || fabs(s[2] / s[0] - 1.0) > 1e-7) {
            delete pl;
            return frt_fail(h, FRT_EINVAL, "SOS section %d: b2 != b0 (zeros off the unit circle)", sec);
        }
        gchain[sec < 2 * bands_per_octave ? sec / 2 : bands_per_octave] *= s[0];
        const double cmid = s[1] / s[0], a1 = s[4], a2 = s[5];
        P.coef[sec][0] = 1.f; P.coef[sec][1] = (float)cmid; P.coef[sec][2] = 1.f;
        P.coef[sec][3] = (float)a1; P.coef[sec][4] = (float)a2;
        // the scan must propagate the state of the float32 filter that pass 2 runs
        const double fb1 = P.coef[sec][1], fa1 = P.coef[sec][3], fa2 = P.coef[sec][4];
        P.coef[sec][5] = (float)(fb1 - fa1);
        P.coef[sec][6] = (float)(1.0 - fa2);
        double A[4] = {-fa1, 1.0, -fa2, 0.0};
        for (int q = 0; q < NQ; q++) {
            for (int i = 0; i < 4; i++) P.apow[sec][q][i] = (float)A[i];
            double A2[4];
            mat2_mul(A, A, A2);
            memcpy(A, A2, sizeof(A));
        }
    }
    for (int i = 0; i < bands_per_octave; i++) P.gband[i] = (float)gchain[i];
    P.gdec = (float)gchain[bands_per_octave];
    for (int j = 0; j < n_octaves; j++) {
        if (!(alphas[j] > 0.0 && alphas[j] <= 1.0)) {
            delete pl;
            return frt_fail(h, FRT_EINVAL, "alpha[%d] must be in (0, 1]", j);
        }
        P.alpha[j] = (float)alphas[j];
        pl->alphas[j] = alphas[j];
        double q = 1.0 - alphas[j];
        for (int k = 0; k < NQ; k++) {
            P.omq[j][k] = (float)(1.0 - q);
            q *= q;
        }
    }
    frt_pipe_prepare(pl);
    pl->n_channels = n_channels;
    pl->nz = (size_t)n_octaves * P.nsec * 2;
    pl->ne = (size_t)n_octaves * bands_per_octave;
    cudaError_t e = cudaMalloc(&pl->zstate, sizeof(float) * pl->nz * n_channels);
    if (e == cudaSuccess) e = cudaMalloc(&pl->ema, sizeof(float) * pl->ne * n_channels);
    if (e == cudaSuccess) e = cudaMemset(pl->zstate, 0, sizeof(float) * pl->nz * n_channels);
    if (e == cudaSuccess) e = cudaMemset(pl->ema, 0, sizeof(float) * pl->ne * n_channels);
    // the memsets run on the legacy stream; callers process on their own (possibly non-blocking) streams
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (pl->zstate) cudaFree(pl->zstate);
        if (pl->ema) cudaFree(pl->ema);
        delete pl;
        return frt_fail(h, FRT_ECUDA, "frt_bank_plan: %s", cudaGetErrorString(e));
    }
    h->bank = pl;
    return FRT_OK;
}

extern "C" int frt_bank_reset(frt_handle h) {
    if (!h) return FRT_EINVAL;
    if (!h->bank) return frt_fail(h, FRT_ESTATE, "frt_bank_reset: no plan");
    DeviceGuard g(h->device);
    BankPlan *pl = h->bank;
    FRT_CUDA(h, cudaMemset(pl->zstate, 0, sizeof(float) * pl->nz * pl->n_channels));
    FRT_CUDA(h, cudaMemset(pl->ema, 0, sizeof(float) * pl->ne * pl->n_channels));
    FRT_CUDA(h, cudaDeviceSynchronize());
    return FRT_OK;
}

extern "C" int frt_bank_set_weighting(frt_handle h, const float *weight_db_host) {
    if (!h) return FRT_EINVAL;
    if (!h->bank) return frt_fail(h, FRT_ESTATE, "frt_bank_set_weighting: no plan");
    DeviceGuard g(h->device);
    BankPlan *pl = h->bank;
    const size_t nb = (size_t)pl->params.n_oct * pl->params.bpo;
    FRT_CUDA(h, cudaDeviceSynchronize());
    if (!weight_db_host) {
        if (pl->weight) cudaFree(pl->weight);
        pl->weight = nullptr;
        return FRT_OK;
    }
    if (!pl->weight) FRT_CUDA(h, cudaMalloc(&pl->weight, sizeof(float) * nb));
    FRT_CUDA(h, cudaMemcpy(pl->weight, weight_db_host, sizeof(float) * nb, cudaMemcpyHostToDevice));
    return FRT_OK;
}

extern "C" int frt_bank_schedule2(int n_octaves, int log2_chunk, int sections_per_lane, int64_t n_samples,
                                  int *stage_start, int *n_steps) {
    if (n_octaves < 1 || n_octaves > MAX_OCT || (log2_chunk != 5 && log2_chunk != 6) || !stage_start ||
        n_samples < 0 || (sections_per_lane != 1 && sections_per_lane != 2))
        return FRT_EINVAL;
    int T[BANK_MAX_OCT + 1];
    frt_pipe_schedule(n_octaves, log2_chunk, sections_per_lane, n_samples, T, n_steps);
    for (int j = 0; j < MAX_OCT; j++) stage_start[j] = T[j];
    return FRT_OK;
}

extern "C" int frt_bank_schedule(int n_octaves, int log2_chunk, int64_t n_samples, int *stage_start,
                                 int *n_steps) {
    return frt_bank_schedule2(n_octaves, log2_chunk, 2, n_samples, stage_start, n_steps);
}

extern "C" int frt_bank_state_size(frt_handle h, int64_t *z_floats, int64_t *ema_floats) {
    if (!h) return FRT_EINVAL;
    if (!h->bank) return frt_fail(h, FRT_ESTATE, "frt_bank_state_size: no plan");
    if (z_floats) *z_floats = (int64_t)(h->bank->nz * h->bank->n_channels);
    if (ema_floats) *ema_floats = (int64_t)(h->bank->ne * h->bank->n_channels);
    return FRT_OK;
}

extern "C" int frt_bank_get_state(frt_handle h, float *z_host, float *ema_host) {
    if (!h) return FRT_EINVAL;
    if (!h->bank) return frt_fail(h, FRT_ESTATE, "frt_bank_get_state: no plan");
    DeviceGuard g(h->device);
    BankPlan *pl = h->bank;
    FRT_CUDA(h, cudaDeviceSynchronize());
    if (z_host)
        FRT_CUDA(h, cudaMemcpy(z_host, pl->zstate, sizeof(float) * pl->nz * pl->n_channels,
                               cudaMemcpyDeviceToHost));
    if (ema_host)
        FRT_CUDA(h, cudaMemcpy(ema_host, pl->ema, sizeof(float) * pl->ne * pl->n_channels,
                               cudaMemcpyDeviceToHost));
    return FRT_OK;
}

extern "C" int frt_bank_set_state(frt_handle h, const float *z_host, const float *ema_host) {
    if (!h) return FRT_EINVAL;
    if (!h->bank) return frt_fail(h, FRT_ESTATE, "frt_bank_set_state: no plan");
    DeviceGuard g(h->device);
    BankPlan *pl = h->bank;
    FRT_CUDA(h, cudaDeviceSynchronize());
    if (z_host)
        FRT_CUDA(h, cudaMemcpy(pl->zstate, z_host, sizeof(float) * pl->nz * pl->n_channels,
                               cudaMemcpyHostToDevice));
    if (ema_host)
        FRT_CUDA(h, cudaMemcpy(pl->ema, ema_host, sizeof(float) * pl->ne * pl->n_channels,
                               cudaMemcpyHostToDevice));
    return FRT_OK;
}

template <int L0, bool WANT_Y>
static cudaError_t launch_bank3(const BankPlan *pl, const BankArgs &a, cudaStream_t st) {
    const BankParams &P = pl->params;
    const size_t smem = sizeof(float) * ((size_t)P.nsec * 8 + pl->nz + pl->ne + 32 * L0 + 64);
    cudaError_t e = cudaFuncSetAttribute(bank_kernel3<L0, WANT_Y>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    bank_kernel3<L0, WANT_Y><<<(unsigned)a.n_channels, 96, smem, st>>>(P, a);
    return cudaGetLastError();
}

template <int L0, bool WANT_Y>
static cudaError_t launch_bank_y(const BankPlan *pl, const BankArgs &a, cudaStream_t st) {
    // few channels: three warps per channel (latency-bound regime); many channels: one warp
    // per channel already saturates the schedulers with less synchronisation.  Measured on B200
    // (blocks/s, 512-sample blocks, one / three warps): 1024 ch 4.1e7 / 5.3e7, 2048 ch 5.6e7 / 5.0e7,
    // 4096 ch 6.7e7 / 5.2e7, 8192 ch 6.8e7 / 5.5e7
    static const char *force = getenv("FRT_BANK_WARPS");
    // the three-warp kernel runs all log2(L0) scan-mode stages unconditionally: it needs more
    // octaves than that (always true for the 9- and 10-octave banks of the reference)
    const bool can_three = pl->params.n_oct > Log2<L0>::v;
    const bool three = can_three && (force ? (force[0] == '3') : (a.n_channels < 1536 && L0 <= 16));
    if (three) return launch_bank3<L0, WANT_Y>(pl, a, st);
    const BankParams &P = pl->params;
    const int warps = 2;
    const size_t smem = sizeof(float) * ((size_t)P.nsec * 8 + (size_t)warps * (pl->nz + pl->ne));
    cudaError_t e = cudaFuncSetAttribute(bank_kernel<L0, WANT_Y>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned blocks = (unsigned)((a.n_channels + warps - 1) / warps);
    bank_kernel<L0, WANT_Y><<<blocks, warps * 32, smem, st>>>(P, a);
    return cudaGetLastError();
}

template <int L0>
static cudaError_t launch_bank(const BankPlan *pl, const BankArgs &a, cudaStream_t st) {
    return a.y ? launch_bank_y<L0, true>(pl, a, st) : launch_bank_y<L0, false>(pl, a, st);
}

static int bank_process_impl(frt_handle h, const float *x_dev, int64_t x_stride, int block,
                             int n_blocks, float *energies_dev, int64_t e_stride, float *y_dev,
                             int64_t y_stride, int db, void *stream);

extern "C" int frt_bank_process(frt_handle h, const float *x_dev, int64_t x_stride, int block,
                                int n_blocks, float *energies_dev, float *y_dev,
                                int64_t y_stride, int db, void *stream) {
    if (!h) return FRT_EINVAL;
    const int64_t nb = h->bank ? (int64_t)h->bank->params.n_oct * h->bank->params.bpo : 0;
    return bank_process_impl(h, x_dev, x_stride, block, n_blocks, energies_dev, (int64_t)n_blocks * nb,
                             y_dev, y_stride, db, stream);
}

extern "C" int frt_bank_process_strided(frt_handle h, const float *x_dev, int64_t x_stride, int block,
                                        int n_blocks, float *energies_dev, int64_t e_stride_c, int db,
                                        void *stream) {
    if (!h) return FRT_EINVAL;
    if (h->bank && energies_dev) {
        const int64_t nb = (int64_t)h->bank->params.n_oct * h->bank->params.bpo;
        if (e_stride_c < (int64_t)n_blocks * nb)
            return frt_fail(h, FRT_EINVAL, "e_stride_c smaller than n_blocks * nbands");
    }
    return bank_process_impl(h, x_dev, x_stride, block, n_blocks, energies_dev, e_stride_c, nullptr, 0, db,
                             stream);
}

static int bank_process_impl(frt_handle h, const float *x_dev, int64_t x_stride, int block,
                             int n_blocks, float *energies_dev, int64_t e_stride, float *y_dev,
                             int64_t y_stride, int db, void *stream) {
    if (!h) return FRT_EINVAL;
    if (!h->bank) return frt_fail(h, FRT_ESTATE, "frt_bank_process: call frt_bank_plan first");
    DeviceGuard g(h->device);
    BankPlan *pl = h->bank;
    const BankParams &P = pl->params;
    FRT_CHECK_ARG(h, n_blocks >= 0 && block >= 0, "negative shape");
    if (n_blocks == 0 || block == 0) return FRT_OK;   // empty chunk: octavespectrum.py:94-95
    // blocks must stay even-length at every stage (decimate.py:41 restarts the [::2] phase
    // on every call), i.e. block % 2^(n_oct-1) == 0; the kernel tiles are 256/512/1024
    FRT_CHECK_ARG(h, block % 256 == 0, "block must be a multiple of 256 samples");
    FRT_CHECK_ARG(h, x_dev != nullptr, "x_dev is NULL");
    FRT_CHECK_ARG(h, x_stride >= (int64_t)block * n_blocks, "x_stride too small");
    int tile = (block % 1024 == 0) ? 1024 : ((block % 512 == 0) ? 512 : 256);
    FRT_CHECK_ARG(h, (tile >> (P.n_oct - 1)) >= 1, "block too short for this many octaves");
    const long long t_total = (long long)block * n_blocks;
    if (y_dev) {
        long long need = 0;
        for (int j = 0; j < P.n_oct; j++) need += (long long)P.bpo * (t_total >> j);
        FRT_CHECK_ARG(h, y_stride >= need, "y_stride smaller than the ragged output size");
    }
    BankArgs a;
    a.x = x_dev;
    a.x_stride = x_stride;
    a.n_channels = pl->n_channels;
    a.n_tiles = (int)(t_total / tile);
    a.tiles_per_block = block / tile;
    a.zstate = pl->zstate;
    a.ema = pl->ema;
    a.energies = energies_dev;
    a.e_stride = e_stride;
    a.y = y_dev;
    a.y_stride = y_stride;
    a.t_total = t_total;
    a.db = db;
    a.vec_ok = (((uintptr_t)x_dev & 15) == 0) && ((x_stride & 3) == 0);
    a.weight = db ? pl->weight : nullptr;
    a.block = block;
    a.n_blocks = n_blocks;
    a.n_steps = 0;
    cudaError_t e;
    cudaStream_t st = (cudaStream_t)stream;
    // Kernel choice.  The lane-pipelined kernel (bank_pipe.cu) computes every section sample once
    // and is the fast path for the fused energies of 1- and 1/3-octave banks; the chunk-scan
    // kernels below serve the ragged band outputs (`.filter()`), 6/12/24 bands per octave and
    // block lengths that are not a power of two.  FRT_BANK_KERNEL=scan|pipe, FRT_BANK_PACK=1|2 and
    // FRT_BANK_LOGCH=5|6 override the choice (tuning knobs).
    const char *force_k = getenv("FRT_BANK_KERNEL");
    const char *force_p = getenv("FRT_BANK_PACK");
    const char *force_c = getenv("FRT_BANK_LOGCH");
    const bool pow2 = (block & (block - 1)) == 0;
    bool use_pipe = pl->pipe_ok && !y_dev && pow2 && (block >> (P.n_oct - 1)) >= 1;
    if (force_k && force_k[0] == 's') use_pipe = false;
    if (use_pipe) {
        // measured on B200 (blocks/s of 512 samples, 27 bands; steps of 32 / 64 samples, one / two
        // channels per lane): 1024 channels 6.1e7 (32) / 8.7e7 (64); 2048: 9.9e7 / 1.35e8; 4096: 1.04e8 /
        // 1.33e8 / 1.44e8 (32, two per lane); 8192: 1.29e8 / 1.34e8 / 1.44e8.  128-sample steps brought
        // nothing over 64, the one-section-per-lane layout (FRT_BANK_SPL=1) nothing at any channel count
        // (profiles/r2_bank_variants.txt).  The choice depends on the plan and the block length only, so
        // a stream gives the same bits however it is cut into launches.
        int pack = pl->n_channels > 3072 ? 2 : 1;
        int logch = (block >= 512 && pl->n_channels <= 3072) ? 6 : 5;
        int spl = 2;
        const char *force_s = getenv("FRT_BANK_SPL");
        if (force_p) pack = force_p[0] == '2' ? 2 : 1;
        if (force_c) logch = force_c[0] == '6' ? 6 : 5;
        if (force_s) spl = force_s[0] == '1' ? 1 : 2;
        // fall back to shorter steps, then to the half-warp layout, when the block is too short for the
        // variant's pipeline depth (frt_pipe_supported)
        while (logch > 5 && !frt_pipe_supported(pl, block, logch, spl)) logch--;
        if (!frt_pipe_supported(pl, block, logch, spl)) spl = 2;
        while (logch > 5 && !frt_pipe_supported(pl, block, logch, spl)) logch--;
        e = frt_pipe_launch(pl, a, logch, pack, spl, st);
    } else if (tile == 1024) e = launch_bank<32>(pl, a, st);
    else if (tile == 512) e = launch_bank<16>(pl, a, st);
    else e = launch_bank<8>(pl, a, st);
    h->launches++;
    if (e != cudaSuccess)
        return frt_fail(h, FRT_ECUDA, "bank kernel launch: %s", cudaGetErrorString(e));
    return FRT_OK;
}

// ---------------------------------------------------------------------------------------------
struct DecPlan {
    BankParams params;
    int n_channels = 0, n_stages = 0;
    float *zstate = nullptr;
};

void frt_dec_release(frt_ctx *h) {
    DecPlan *pl = reinterpret_cast<DecPlan *>(h->dec);
    if (!pl) return;
    if (pl->zstate) cudaFree(pl->zstate);
    delete pl;
    h->dec = nullptr;
}

extern "C" int frt_decimate_plan(frt_handle h, int n_channels, int n_stages, const double *sos_dec) {
    if (!h) return FRT_EINVAL;
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_channels >= 1, "n_channels must be >= 1");
    FRT_CHECK_ARG(h, n_stages >= 1 && n_stages <= 2, "n_stages must be 1 or 2");
    FRT_CHECK_ARG(h, sos_dec != nullptr, "NULL coefficient table");
    frt_dec_release(h);
    DecPlan *pl = new (std::nothrow) DecPlan();
    if (!pl) return frt_fail(h, FRT_ENOMEM, "out of host memory");
    BankParams &P = pl->params;
    memset(&P, 0, sizeof(P));
    P.bpo = 0;          // section index 2*bpo + s = s: the six decimator sections
    P.n_oct = n_stages;
    P.nsec = 6;
    double gdec = 1.0;
    for (int sec = 0; sec < 6; sec++) {
        const double *s = sos_dec + (size_t)sec * 6;
        if (!(fabs(s[0]) > 0.0) || fabs(s[2] / s[0] - 1.0) > 1e-7 || fabs(s[3] - 1.0) > 1e-12) {
            delete pl;
            return frt_fail(h, FRT_EINVAL, "decimator section %d is not of the form g(1,c,1)/(1,a1,a2)", sec);
        }
        gdec *= s[0];
        P.coef[sec][0] = 1.f; P.coef[sec][1] = (float)(s[1] / s[0]); P.coef[sec][2] = 1.f;
        P.coef[sec][3] = (float)s[4]; P.coef[sec][4] = (float)s[5];
        const double fb1 = P.coef[sec][1], fa1 = P.coef[sec][3], fa2 = P.coef[sec][4];
        P.coef[sec][5] = (float)(fb1 - fa1);
        P.coef[sec][6] = (float)(1.0 - fa2);
        double A[4] = {-fa1, 1.0, -fa2, 0.0};
        for (int q = 0; q < NQ; q++) {
            for (int i = 0; i < 4; i++) P.apow[sec][q][i] = (float)A[i];
            double A2[4];
            mat2_mul(A, A, A2);
            memcpy(A, A2, sizeof(A));
        }
    }
    P.gdec = (float)gdec;
    pl->n_channels = n_channels;
    pl->n_stages = n_stages;
    const size_t nz = (size_t)n_stages * 6 * 2 * n_channels;
    cudaError_t e = cudaMalloc(&pl->zstate, sizeof(float) * nz);
    if (e == cudaSuccess) e = cudaMemset(pl->zstate, 0, sizeof(float) * nz);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        if (pl->zstate) cudaFree(pl->zstate);
        delete pl;
        return frt_fail(h, FRT_ECUDA, "frt_decimate_plan: %s", cudaGetErrorString(e));
    }
    h->dec = pl;
    return FRT_OK;
}

extern "C" int frt_decimate_process(frt_handle h, const float *x_dev, int64_t x_stride,
                                    int64_t n_samples, float *out_dev, int64_t out_stride,
                                    void *stream) {
    if (!h) return FRT_EINVAL;
    DecPlan *pl = reinterpret_cast<DecPlan *>(h->dec);
    if (!pl) return frt_fail(h, FRT_ESTATE, "frt_decimate_process: call frt_decimate_plan first");
    DeviceGuard g(h->device);
    FRT_CHECK_ARG(h, n_samples >= 0, "negative length");
    if (n_samples == 0) return FRT_OK;   // decimate_multiple returns empty input unchanged (:56-57)
    FRT_CHECK_ARG(h, n_samples % 256 == 0, "n_samples must be a multiple of 256");
    FRT_CHECK_ARG(h, x_dev && out_dev, "NULL buffer");
    FRT_CHECK_ARG(h, x_stride >= n_samples && out_stride >= (n_samples >> pl->n_stages),
                  "stride too small");
    DecArgs a;
    a.x = x_dev;
    a.x_stride = x_stride;
    a.out = out_dev;
    a.out_stride = out_stride;
    a.n_channels = pl->n_channels;
    a.n_tiles = (int)(n_samples / 256);
    a.n_stages = pl->n_stages;
    a.zstate = pl->zstate;
    const int warps = 2;
    const size_t smem = sizeof(float) * warps * pl->n_stages * 12;
    decimate_kernel<<<(pl->n_channels + warps - 1) / warps, warps * 32, smem,
                      (cudaStream_t)stream>>>(pl->params, a);
    h->launches++;
    FRT_CUDA(h, cudaGetLastError());
    return FRT_OK;
}
