// Shared helpers for the libwgs_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace wgs {

// ---- error plumbing (C ABI returns int codes, message via wgs_last_error) ---------------------
void set_error(const std::string& msg);
int  fail(const char* file, int line, const std::string& msg);

#define WGS_REQUIRE(cond, msg)                                                   \
    do { if (!(cond)) return ::wgs::fail(__FILE__, __LINE__, std::string(msg)); } while (0)

#define WGS_CUDA(expr)                                                           \
    do { cudaError_t e__ = (expr);                                               \
         if (e__ != cudaSuccess)                                                 \
             return ::wgs::fail(__FILE__, __LINE__, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

#define WGS_LAUNCH_CHECK() WGS_CUDA(cudaGetLastError())

// counts kernels launched through the library (bench.py reports it as gpu_launches)
extern unsigned long long g_launches;
inline void count_launch(int n = 1) { g_launches += (unsigned long long)n; }

inline int num_sms() {
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    return sms;
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// ---- device helpers ------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; `red` is >= 32 floats of shared memory. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
    if (warp == 0) {
        t = warp_sum(t);
        if (lane == 0) red[0] = t;
    }
    __syncthreads();
    t = red[0];
    return t;
}

// bf16 hi/lo split of an fp32 value: x ~= hi + lo with |x - hi - lo| <= 2^-17 |x|
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
    hi = __float2bfloat16_rn(x);
    lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}

// two values at once: one packed cvt.rn.bf16x2.f32 per pair for hi and for lo
__device__ __forceinline__ void split_bf16x2(float a, float b, __nv_bfloat162& hi, __nv_bfloat162& lo) {
    hi = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(hi);
    lo = __floats2bfloat162_rn(a - hf.x, b - hf.y);
}

}  // namespace wgs
