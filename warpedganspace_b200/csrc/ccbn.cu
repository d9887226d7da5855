// Class-conditional BatchNorm (eval mode) + ReLU fused with the split32 operand pack of the convolution that follows it,
// and its backward - the glue of BigGAN's GBlock (models/BigGAN/layers.py:303-322 ccbn, :395-405 GBlock.forward).
//
// The generator is frozen and in eval mode, so `F.batch_norm(x, stored_mean, stored_var) * gain(y) + bias(y)` is a
// per-sample, per-channel affine map  o[n, p, c] = A[n, c] * x[n, p, c] + B[n, c]  with
// A = gain / sqrt(var + eps), B = bias - mean * A  (both [N, C], formed by the host from the two small ccbn linears, so
// autograd carries their gradient on to z).  Reference: batch_norm + mul + add + relu + interpolate = five full passes
// per ccbn; here ONE pass reads x and writes the next conv's operand (the nearest x2 up-sample never materialises: the conv
// runs as four output-phase convs over this low-resolution operand, generators.py).
// Backward: ONE pass reads the incoming gradient and x, writes the operand of the next data-gradient conv (or fp32) and
// reduces dA[n, c] = sum_p dpre * x, dB[n, c] = sum_p dpre per sample (registers -> shared -> one atomic per channel
// per block).  HBM-bound element-wise kernels; C % 4 == 0 (BigGAN: 1536 ... 96).
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

__global__ void __launch_bounds__(256)
affine_act_pack_kernel(const float* __restrict__ x, const float* __restrict__ A, const float* __restrict__ B, long long R, int C,
                       long long rows_per_group, int relu, __nv_bfloat16* __restrict__ out_split, float* __restrict__ out_f32) {
    const int C4 = C >> 2, chunks = (C + 31) >> 5;
    const long long total = R * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        const long long r = i / C4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C + c));
        float o[4] = {v.x, v.y, v.z, v.w};
        if (A) {
            const long long g = (r / rows_per_group) * C + c;
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(A + g)), b4 = __ldg(reinterpret_cast<const float4*>(B + g));
            o[0] = fmaf(o[0], a4.x, b4.x); o[1] = fmaf(o[1], a4.y, b4.y); o[2] = fmaf(o[2], a4.z, b4.z); o[3] = fmaf(o[3], a4.w, b4.w);
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] > 0.f ? o[k] : 0.f;
        }
        if (out_f32) *reinterpret_cast<float4*>(out_f32 + r * C + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (out_split) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
            __nv_bfloat16* sp = out_split + r * (long long)(chunks * 64) + (c >> 5) * 64 + (c & 31);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// grid = (blocks per sample, samples); block = C4 * PL threads: thread (q, pl) owns channel quad q for rows pl, pl+PL, ...
__global__ void __launch_bounds__(1024)
affine_act_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ x, const float* __restrict__ A,
                      const float* __restrict__ B, long long rows_per_group, int C, int relu,
                      __nv_bfloat16* __restrict__ dx_split, float* __restrict__ dx_f32, float* __restrict__ dA,
                      float* __restrict__ dB, int PL) {
    extern __shared__ float sm[];
    const int C4 = C >> 2, chunks = (C + 31) >> 5;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4, c = q * 4;
    const long long n = blockIdx.y;
    const long long per = (rows_per_group + gridDim.x - 1) / gridDim.x;
    const long long r0 = n * rows_per_group + blockIdx.x * per, r1 = min((n + 1) * rows_per_group, r0 + per);
    float4 a4 = make_float4(1.f, 1.f, 1.f, 1.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (A) {
        a4 = __ldg(reinterpret_cast<const float4*>(A + n * C + c));
        b4 = __ldg(reinterpret_cast<const float4*>(B + n * C + c));
    }
    float sa[4] = {0.f, 0.f, 0.f, 0.f}, sb[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = r0 + pl; r < r1; r += PL) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(dz + r * C + c));
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * C + c));
        float gv[4] = {g.x, g.y, g.z, g.w};
        const float xv[4] = {v.x, v.y, v.z, v.w}, av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (relu && !(fmaf(xv[k], av[k], bv[k]) > 0.f)) gv[k] = 0.f;
            sa[k] += gv[k] * xv[k];
            sb[k] += gv[k];
            o[k] = gv[k] * av[k];
        }
        if (dx_f32) *reinterpret_cast<float4*>(dx_f32 + r * C + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (dx_split) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
            __nv_bfloat16* sp = dx_split + r * (long long)(chunks * 64) + (c >> 5) * 64 + (c & 31);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
    if (!dA) return;                                  // (uniform: every thread of the grid takes the same branch)
    float* s0 = sm;                                   // [PL][C] partial dA, then [PL][C] partial dB
    float* s1 = sm + (size_t)PL * C;
    *reinterpret_cast<float4*>(s0 + (size_t)pl * C + c) = make_float4(sa[0], sa[1], sa[2], sa[3]);
    *reinterpret_cast<float4*>(s1 + (size_t)pl * C + c) = make_float4(sb[0], sb[1], sb[2], sb[3]);
    __syncthreads();
    for (int cc = threadIdx.x; cc < C; cc += blockDim.x) {
        float ta = 0.f, tb = 0.f;
        for (int l = 0; l < PL; ++l) { ta += s0[(size_t)l * C + cc]; tb += s1[(size_t)l * C + cc]; }
        atomicAdd(dA + n * C + cc, ta);
        atomicAdd(dB + n * C + cc, tb);
    }
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_affine_act_pack(const float* x, const float* A, const float* B, long long R, int C, long long rows_per_group,
                                   int relu, void* out_split, float* out_f32, void* stream) {
    WGS_REQUIRE(R >= 0 && C >= 4 && C % 4 == 0, "affine_act_pack: C must be a multiple of 4");
    WGS_REQUIRE((A == nullptr) == (B == nullptr), "affine_act_pack: A and B come together");
    WGS_REQUIRE(!A || (rows_per_group >= 1 && R % rows_per_group == 0), "affine_act_pack: R must be groups * rows_per_group");
    WGS_REQUIRE(out_split != nullptr || out_f32 != nullptr, "affine_act_pack: no output requested");
    if (R == 0) return 0;
    const long long total = R * (C / 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16);
    affine_act_pack_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, A, B, R, C, rows_per_group ? rows_per_group : R, relu,
                                                                     (__nv_bfloat16*)out_split, out_f32);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_affine_act_bwd(const float* dz, const float* x, const float* A, const float* B, int N, long long rows_per_group,
                                  int C, int relu, void* dx_split, float* dx_f32, float* dA, float* dB, void* stream) {
    WGS_REQUIRE(N >= 0 && rows_per_group >= 1 && C >= 4 && C % 4 == 0 && C <= 4096, "affine_act_bwd: bad sizes (C % 4 == 0, C <= 4096)");
    WGS_REQUIRE((A == nullptr) == (B == nullptr) && (dA == nullptr) == (dB == nullptr), "affine_act_bwd: A/B and dA/dB come in pairs");
    WGS_REQUIRE(dx_split != nullptr || dx_f32 != nullptr || dA != nullptr, "affine_act_bwd: no output requested");
    if (N == 0) return 0;
    const int C4 = C / 4;
    const int PL = std::max(1, 256 / C4);
    const int threads = C4 * PL;
    WGS_REQUIRE(threads <= 1024, "affine_act_bwd: C too large for one row per block");
    const size_t smem = dA ? 2 * (size_t)PL * C * sizeof(float) : 0;
    // enough blocks per sample for ~8 waves over the GPU, each with at least 4 rows per row-lane
    const long long want = std::max<long long>(1, (long long)num_sms() * 8 / std::max(1, N));
    const int bps = (int)std::max<long long>(1, std::min<long long>(want, rows_per_group / (4LL * PL) + 1));
    static bool attr = false;
    if (!attr) {
        WGS_CUDA(cudaFuncSetAttribute(affine_act_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        attr = true;
    }
    WGS_REQUIRE(smem <= 64 * 1024, "affine_act_bwd: reduction buffer too large");
    affine_act_bwd_kernel<<<dim3(bps, N), threads, smem, (cudaStream_t)stream>>>(dz, x, A, B, rows_per_group, C, relu,
                                                                                  (__nv_bfloat16*)dx_split, dx_f32, dA, dB, PL);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
