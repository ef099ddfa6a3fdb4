// Internal declarations shared by the translation units of libfrt_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/frt.h"

struct StftPlan {
    int n_fft = 0;
    float *win_dev = nullptr;      // [n_fft] symmetric Hann (audioproc.py:76-81)
    float2 *tw_dev = nullptr;      // fast path: [32][32] W_M^(k1*t); generic: [M] W_M^k
    float2 *post_dev = nullptr;    // [M/2+1] U[k] = -j*W_N^k (real-FFT split twiddles)
    float2 *wlane_dev = nullptr;   // fast path: per-lane Hann phase factors
    float2 *comb_dev = nullptr;    // N = 4096/8192: [(R-1)][1024] W_M^(w k0)
    std::vector<float> win_host;
};

struct HostPipe {                 // staging for the *_host entry points
    cudaStream_t s_in = nullptr, s_cmp = nullptr, s_out = nullptr;
    float *d_in[2] = {nullptr, nullptr};
    float *d_out[2] = {nullptr, nullptr};
    size_t in_bytes = 0, out_bytes = 0;
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_cmp[2] = {nullptr, nullptr},
                ev_out[2] = {nullptr, nullptr};
};

struct BankPlan;   // bank.cu
struct GccPlan;    // gcc_phat.cu

struct frt_ctx {
    int device = 0;
    int sm_count = 0;
    int64_t launches = 0;
    std::string err;
    StftPlan stft;                       // current plan (a copy of one cache entry)
    std::map<int, StftPlan> stft_cache;  // every size planned so far owns its device tables
    HostPipe pipe;
    BankPlan *bank = nullptr;
    GccPlan *gcc = nullptr;
    void *dec = nullptr;          // DecPlan (bank.cu)
    void *comb = nullptr;         // CombPipe (combined.cu)
    void *fir = nullptr;          // FirPlan (bank_fir.cu)
};

int frt_fail(frt_ctx *h, int code, const char *fmt, ...);

#define FRT_CUDA(h, call)                                                                  \
    do {                                                                                   \
        cudaError_t e__ = (call);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return frt_fail((h), FRT_ECUDA, "%s failed: %s (%s:%d)", #call,                \
                            cudaGetErrorString(e__), __FILE__, __LINE__);                  \
    } while (0)

#define FRT_CHECK_ARG(h, cond, msg)                                                        \
    do {                                                                                   \
        if (!(cond)) return frt_fail((h), FRT_EINVAL, "%s (%s)", msg, #cond);              \
    } while (0)

// Each entry point runs on the handle's device.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// bank.cu / gcc_phat.cu clean-up hooks
void frt_bank_release(frt_ctx *h);
void frt_gcc_release(frt_ctx *h);
void frt_dec_release(frt_ctx *h);
void frt_comb_release(frt_ctx *h);
void frt_fir_release(frt_ctx *h);
int frt_pipe_ensure(frt_ctx *h, size_t in_bytes, size_t out_bytes);
