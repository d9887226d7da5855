// Fused Adam over one flat fp32 buffer (torch.optim.Adam defaults, lib/trainer.py:153,156,253-254):
// both optimisers' parameters live in flat storage, so a step is one launch and the gradient
// all-reduce is one NCCL call over the same flat gradient buffer.
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt, float grad_scale,
            const int* __restrict__ step_ptr) {
    if (step_ptr) {                      // step counter lives on the device (CUDA-graph replay): derive the corrections here
        const double st = (double)(*step_ptr);
        bc1 = (float)(1.0 - pow((double)b1, st));
        bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, st));
    }
    const long long n4 = n >> 2;
    const float step = lr / bc1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 p4 = reinterpret_cast<float4*>(p)[i];
        const float4 g4 = reinterpret_cast<const float4*>(g)[i];
        float4 m4 = reinterpret_cast<float4*>(m)[i];
        float4 v4 = reinterpret_cast<float4*>(v)[i];
        float pp[4] = {p4.x, p4.y, p4.z, p4.w}, gg[4] = {g4.x, g4.y, g4.z, g4.w};
        float mm[4] = {m4.x, m4.y, m4.z, m4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = gg[k] * grad_scale;
            mm[k] = b1 * mm[k] + (1.f - b1) * gk;
            vv[k] = b2 * vv[k] + (1.f - b2) * gk * gk;
            pp[k] -= step * mm[k] / (sqrtf(vv[k]) / bc2_sqrt + eps);
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pp[0], pp[1], pp[2], pp[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mm[0], mm[1], mm[2], mm[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    // tail
    if (blockIdx.x == 0) {
        for (long long i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            const float gk = g[i] * grad_scale;
            const float mk = b1 * m[i] + (1.f - b1) * gk;
            const float vk = b2 * v[i] + (1.f - b2) * gk * gk;
            m[i] = mk; v[i] = vk;
            p[i] -= step * mk / (sqrtf(vk) / bc2_sqrt + eps);
        }
    }
}

__global__ void step_inc_kernel(int* step) { *step += 1; }

}  // namespace wgs
using namespace wgs;

extern "C" int wgs_step_increment(int* step_dev, void* stream) {
    step_inc_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2,
                             float eps, int step, float grad_scale, const int* step_dev, void* stream) {
    WGS_REQUIRE(n >= 0 && (step >= 1 || step_dev), "adam_step: bad arguments");
    WGS_REQUIRE(((uintptr_t)p & 15) == 0 && ((uintptr_t)g & 15) == 0 && ((uintptr_t)m & 15) == 0 && ((uintptr_t)v & 15) == 0,
                "adam_step: buffers must be 16-byte aligned");
    if (n == 0) return 0;
    const double bc1 = 1.0 - pow((double)b1, std::max(1, step)), bc2 = 1.0 - pow((double)b2, std::max(1, step));
    const int blocks = (int)std::min<long long>((n / 4 + 255) / 256 + 1, (long long)num_sms() * 8);
    adam_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, b1, b2, eps, (float)bc1, (float)sqrt(bc2),
                                                         grad_scale, step_dev);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
