// Op-level boundary of the reference's two native extensions: drop-ins for the pybind
// functions `fused.fused_bias_act` (models/StyleGAN2/op/fused_bias_act.cpp:11-20, kernel fused_bias_act_kernel.cu:18-99)
// and `upfirdn2d_op.upfirdn2d` (op/upfirdn2d.cpp:12-22, kernel upfirdn2d_kernel.cu:52-272) with the same argument
// meaning.  The generator's hot path does NOT go through these - there the bias / activation lives in the conv epilogue
// and the FIR in fir4_act / torgb (sg2.cu) - they exist so that code written against the reference's op API (its
// Discriminator, projector scripts, third-party StyleGAN2 code) finds the same operators here.
//
// Both are HBM-bound element-wise / stencil kernels: 128-bit loads where the layout allows, FIR taps in shared memory,
// grid = a multiple of the SM count with a grid-stride loop.
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

// y = act(x + b[(i / step_b) % size_b]) * scale, the reference's `act * 10 + grad` switch:
//   act 1 linear, act 3 leaky ReLU; grad 0 forward, grad 1 first derivative w.r.t. x selected by the sign of `ref`
//   (the forward OUTPUT), grad 2 second derivative (= 0).
__device__ __forceinline__ float bias_act_one(float x, float ref, int act, int grad, float alpha, float scale) {
    float y;
    if (act == 1) y = grad == 0 ? x : (grad == 1 ? x : 0.f);
    else y = grad == 0 ? (x > 0.f ? x : x * alpha) : (grad == 1 ? (ref > 0.f ? x : x * alpha) : 0.f);
    return y * scale;
}

__global__ void __launch_bounds__(256)
fused_bias_act_kernel(const float* __restrict__ x, const float* __restrict__ b, const float* __restrict__ ref,
                      float* __restrict__ out, long long n, long long step_b, int size_b, int act, int grad, float alpha,
                      float scale, int vec) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (vec) {                                       // step_b % 4 == 0 and 16-byte aligned: one bias per float4
        const long long n4 = n >> 2;
        for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
            const float bb = b ? __ldg(b + ((i * 4) / step_b) % size_b) : 0.f;
            float4 r = ref ? __ldg(reinterpret_cast<const float4*>(ref) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            v.x = bias_act_one(v.x + bb, r.x, act, grad, alpha, scale);
            v.y = bias_act_one(v.y + bb, r.y, act, grad, alpha, scale);
            v.z = bias_act_one(v.z + bb, r.z, act, grad, alpha, scale);
            v.w = bias_act_one(v.w + bb, r.w, act, grad, alpha, scale);
            reinterpret_cast<float4*>(out)[i] = v;
        }
        return;
    }
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
        const float bb = b ? __ldg(b + (i / step_b) % size_b) : 0.f;
        out[i] = bias_act_one(__ldg(x + i) + bb, ref ? __ldg(ref + i) : 0.f, act, grad, alpha, scale);
    }
}

// out[m, oy, ox, c] = sum_{ky,kx} kernel[kh-1-ky, kw-1-kx] * up_pad(in)[m, oy*down_y + ky, ox*down_x + kx, c]
// where up_pad zero-inserts (up_x, up_y), then pads (negative pad = crop).  One thread per output element, the FIR in
// shared memory; x is the fastest output index for minor == 1 (the only layout the reference's Python wrapper produces,
// op/upfirdn2d.py:98), so loads and stores of a warp are contiguous.
__global__ void __launch_bounds__(256)
upfirdn2d_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out, int major,
                 int in_h, int in_w, int minor, int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0,
                 int pad_y0, int out_h, int out_w) {
    extern __shared__ float k_s[];
    for (int i = threadIdx.x; i < kh * kw; i += blockDim.x) k_s[i] = kernel[(kh - 1 - i / kw) * kw + (kw - 1 - i % kw)];
    __syncthreads();
    const long long total = (long long)major * out_h * out_w * minor;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % minor);
        long long r = i / minor;
        const int ox = (int)(r % out_w); r /= out_w;
        const int oy = (int)(r % out_h);
        const long long m = r / out_h;
        const float* src = in + m * in_h * (long long)in_w * minor + c;
        float acc = 0.f;
        for (int ky = 0; ky < kh; ++ky) {
            const int py = oy * down_y + ky - pad_y0;
            if (py < 0 || py % up_y) continue;
            const int iy = py / up_y;
            if (iy >= in_h) continue;
            for (int kx = 0; kx < kw; ++kx) {
                const int px = ox * down_x + kx - pad_x0;
                if (px < 0 || px % up_x) continue;
                const int ix = px / up_x;
                if (ix >= in_w) continue;
                acc += k_s[ky * kw + kx] * __ldg(src + ((long long)iy * in_w + ix) * minor);
            }
        }
        out[i] = acc;
    }
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_fused_bias_act(const float* x, const float* b, const float* ref, float* out, long long n,
                                  long long step_b, int size_b, int act, int grad, float alpha, float scale, void* stream) {
    WGS_REQUIRE(n >= 0 && step_b >= 1 && size_b >= 1, "fused_bias_act: bad sizes");
    WGS_REQUIRE(act == 1 || act == 3, "fused_bias_act: act must be 1 (linear) or 3 (leaky ReLU)");
    WGS_REQUIRE(grad >= 0 && grad <= 2, "fused_bias_act: grad must be 0, 1 or 2");
    WGS_REQUIRE(grad == 0 || act == 1 || ref != nullptr, "fused_bias_act: the leaky-ReLU derivative needs the forward output as ref");
    if (n == 0) return 0;
    const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(ref)) & 15) == 0;
    const int vec = (aligned && n % 4 == 0 && (b == nullptr || step_b % 4 == 0)) ? 1 : 0;
    const long long work = vec ? n / 4 : n;
    const int blocks = (int)std::min<long long>((work + 255) / 256, (long long)num_sms() * 16);
    fused_bias_act_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, b, ref, out, n, step_b, size_b, act, grad, alpha, scale, vec);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_upfirdn2d(const float* in, const float* kernel, float* out, int major, int in_h, int in_w, int minor,
                             int kh, int kw, int up_x, int up_y, int down_x, int down_y, int pad_x0, int pad_x1, int pad_y0,
                             int pad_y1, void* stream) {
    WGS_REQUIRE(major >= 0 && in_h >= 1 && in_w >= 1 && minor >= 1, "upfirdn2d: bad input shape");
    WGS_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= 4096, "upfirdn2d: bad FIR shape");
    WGS_REQUIRE(up_x >= 1 && up_y >= 1 && down_x >= 1 && down_y >= 1, "upfirdn2d: up / down factors must be >= 1");
    const int out_h = (in_h * up_y + pad_y0 + pad_y1 - kh) / down_y + 1;
    const int out_w = (in_w * up_x + pad_x0 + pad_x1 - kw) / down_x + 1;
    WGS_REQUIRE(out_h >= 1 && out_w >= 1 && in_h * up_y + pad_y0 + pad_y1 >= kh && in_w * up_x + pad_x0 + pad_x1 >= kw,
                "upfirdn2d: empty output");
    if (major == 0) return 0;
    const long long total = (long long)major * out_h * out_w * minor;
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
    upfirdn2d_kernel<<<blocks, 256, (size_t)kh * kw * sizeof(float), (cudaStream_t)stream>>>(
        in, kernel, out, major, in_h, in_w, minor, kh, kw, up_x, up_y, down_x, down_y, pad_x0, pad_y0, out_h, out_w);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
