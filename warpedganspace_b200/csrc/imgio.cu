// Traversal / sampling output stage: generator image (fp32, NHWC) -> uint8 pixels on the device.
//
// Reference: tensor2image (traverse_latent_space.py:26-41, sample_gan.py:10-25) runs on the HOST per image:
//   adaptive:  t = (x - min(x)) / (max(x) - min(x));  u8 = uint8(255 * t)
//   otherwise: t = (x + 1) / 2  (the clamp result is discarded);  u8 = uint8(255 * t)
// on a CHW tensor that ToPILImage then transposes to HWC.  Here the min/max reduction and the conversion run on the
// device over the NHWC image the generator already produced, and only 1 byte per value crosses PCIe (4x less than the
// reference's fp32 .cpu()).  Same fp32 operations in the same order -> bit-identical pixels.
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

constexpr int IMG_THREADS = 256;

// partial[n][b] = (min, max) over block b's slice of image n
__global__ void __launch_bounds__(IMG_THREADS)
image_minmax_kernel(const float* __restrict__ x, long long count, float2* __restrict__ partial) {
    __shared__ float smn[IMG_THREADS / 32], smx[IMG_THREADS / 32];
    const float* img = x + (size_t)blockIdx.y * count;
    float mn = INFINITY, mx = -INFINITY;
    for (long long i = blockIdx.x * (long long)IMG_THREADS + threadIdx.x; i < count; i += (long long)gridDim.x * IMG_THREADS) {
        const float v = __ldg(img + i);
        mn = fminf(mn, v);
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { smn[threadIdx.x >> 5] = mn; smx[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < IMG_THREADS / 32; ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
        partial[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = make_float2(mn, mx);
    }
}

__global__ void __launch_bounds__(IMG_THREADS)
image_to_u8_kernel(const float* __restrict__ x, unsigned char* __restrict__ out, long long count,
                   const float2* __restrict__ partial, int n_partial, int adaptive) {
    __shared__ float s_mn, s_mx;
    if (adaptive) {
        if (threadIdx.x == 0) {
            float mn = INFINITY, mx = -INFINITY;
            for (int b = 0; b < n_partial; ++b) {
                const float2 p = partial[(size_t)blockIdx.y * n_partial + b];
                mn = fminf(mn, p.x);
                mx = fmaxf(mx, p.y);
            }
            s_mn = mn; s_mx = mx;
        }
        __syncthreads();
    }
    const float mn = adaptive ? s_mn : 0.f;
    const float range = adaptive ? __fsub_rn(s_mx, s_mn) : 0.f;
    const float* img = x + (size_t)blockIdx.y * count;
    unsigned char* dst = out + (size_t)blockIdx.y * count;
    for (long long i = blockIdx.x * (long long)IMG_THREADS + threadIdx.x; i < count; i += (long long)gridDim.x * IMG_THREADS) {
        const float v = __ldg(img + i);
        const float t = adaptive ? __fdiv_rn(__fsub_rn(v, mn), range) : __fdiv_rn(__fadd_rn(v, 1.f), 2.f);
        dst[i] = (unsigned char)(int)__fmul_rn(255.f, t);       // truncation, as Tensor.to(torch.uint8)
    }
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_image_to_u8(const float* images, int n, long long count, int adaptive, float* workspace,
                               int workspace_pairs, unsigned char* out, void* stream) {
    WGS_REQUIRE(n >= 0 && count > 0, "image_to_u8: bad sizes");
    WGS_REQUIRE(!adaptive || (workspace != nullptr && workspace_pairs >= 1), "image_to_u8: adaptive mode needs a workspace");
    if (n == 0) return 0;
    const cudaStream_t st = (cudaStream_t)stream;
    const int want = (int)std::min<long long>((count + IMG_THREADS * 8 - 1) / (IMG_THREADS * 8), 4 * num_sms());
    int blocks = std::max(1, want);
    if (adaptive) {
        blocks = std::max(1, std::min(blocks, workspace_pairs / std::max(1, n)));
        WGS_REQUIRE((long long)blocks * n <= workspace_pairs, "image_to_u8: workspace too small (need >= n pairs)");
        image_minmax_kernel<<<dim3(blocks, n), IMG_THREADS, 0, st>>>(images, count, reinterpret_cast<float2*>(workspace));
        count_launch();
    }
    image_to_u8_kernel<<<dim3(std::max(1, want), n), IMG_THREADS, 0, st>>>(images, out, count,
                                                                         reinterpret_cast<const float2*>(workspace), blocks, adaptive);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
