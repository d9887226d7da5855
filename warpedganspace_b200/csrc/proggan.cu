// ProgGAN glue kernels: PixelNormLayer fused with the split32 operand pack of the following convolution (forward), and its
// backward fused with the LeakyReLU backward of the preceding block and the pack of the data-gradient conv's operand.
//
// Reference block (models/ProgGAN/model.py:42-62):  x -> x / sqrt(mean_c x^2 + 1e-8) -> [nearest x2] -> conv -> *scale + b
// -> leaky_relu(0.2): three element-wise kernels for the norm, one for the up-sample, two for WScale, one for the
// activation, each a full read + write of a tensor of up to 16 x 1024^2 floats per image.  Here:
//   * WScale's scale is folded into the (frozen) conv weights, bias + LeakyReLU run in the conv epilogue (act = 2);
//   * nearest x2 + 3x3 conv is run as four output-phase 2x2 convs over the LOW-resolution operand with pre-summed taps
//     (2.25x fewer MACs, 4x fewer operand bytes; host side, generators.py);
//   * what is left per block is ONE pass: fp32 activation -> pixel-norm -> split32 operand (this file).
// HBM-bound: 4 B read + 4 B written per element; one warp serves 32 / (C/4) pixels so that every load is a coalesced
// 128-bit load and the channel reduction is a sub-warp shuffle.
#include <stdlib.h>
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

constexpr int PN_THREADS = 256;
constexpr int PN_MAX_V = 4;                      // float4 slots per lane: C <= 32 * 4 * PN_MAX_V = 512

__device__ __forceinline__ float group_sum(float v, int lanes) {
    for (int o = lanes >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// mode 0: out = split32(a * r)                                   (forward)
// mode 1: backward.  g = r * (dxn - xn * mean_c(dxn * xn)),  xn = a * r;  if slope >= 0: g *= (a > 0 ? 1 : slope)
//         (LeakyReLU backward of the block that produced a; sign(a) = sign of its pre-activation);
//         writes split32(g) to out_split and / or fp32 g to out_f32.
// NV float4 slots per lane and pixel, U pixels per lane and trip (NV * U = 4): every load of a trip is issued before the first
// use, so a lane keeps 4 (forward) / 8 (backward) 128-bit loads in flight instead of 1 / 2 - the <= 128-channel layers, where
// one slot covers the pixel, were latency-bound at 3.5 - 3.8 TB/s.
template <int MODE, int NV, int U>
__global__ void __launch_bounds__(PN_THREADS)
pixelnorm_kernel(const float* __restrict__ a, const float* __restrict__ dxn, long long R, int C, float eps, float slope,
                 __nv_bfloat16* __restrict__ out_split, float* __restrict__ out_f32, int lanes) {
    const int lane = threadIdx.x & 31;
    const int sub = lane % lanes, grp = lane / lanes, per_warp = 32 / lanes;
    const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
    const int chunks = (C + 31) >> 5;
    for (long long base = warp_global * per_warp * U; base < R; base += n_warps * per_warp * U) {
        float4 v[U][NV], d[U][NV];
        bool ok[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = base + (long long)u * per_warp + grp;
            ok[u] = r < R;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = (i * lanes + sub) * 4;
                v[u][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                d[u][i] = v[u][i];
                if (ok[u] && c < C) {
                    v[u][i] = __ldg(reinterpret_cast<const float4*>(a + r * C + c));
                    if (MODE == 1) d[u][i] = __ldg(reinterpret_cast<const float4*>(dxn + r * C + c));
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long r = base + (long long)u * per_warp + grp;
            float ss = 0.f, dot = 0.f;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                ss += v[u][i].x * v[u][i].x + v[u][i].y * v[u][i].y + v[u][i].z * v[u][i].z + v[u][i].w * v[u][i].w;
                if (MODE == 1) dot += v[u][i].x * d[u][i].x + v[u][i].y * d[u][i].y + v[u][i].z * d[u][i].z + v[u][i].w * d[u][i].w;
            }
            ss = group_sum(ss, lanes);                                         // (all lanes take part: no early exit above)
            const float rn = rsqrtf(ss / (float)C + eps);
            float m = 0.f;
            if (MODE == 1) m = group_sum(dot, lanes) * rn / (float)C;          // mean_c(dxn * xn)
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int c = (i * lanes + sub) * 4;
                if (!(ok[u] && c < C)) continue;
                float o[4];
                if (MODE == 0) {
                    o[0] = v[u][i].x * rn; o[1] = v[u][i].y * rn; o[2] = v[u][i].z * rn; o[3] = v[u][i].w * rn;
                } else {
                    const float av[4] = {v[u][i].x, v[u][i].y, v[u][i].z, v[u][i].w};
                    const float dv[4] = {d[u][i].x, d[u][i].y, d[u][i].z, d[u][i].w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float g = rn * (dv[k] - av[k] * rn * m);
                        if (slope >= 0.f) g *= (av[k] > 0.f ? 1.f : slope);
                        o[k] = g;
                    }
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
            // channels C .. chunks*32 of a partially filled last chunk (C = 16): zeros, so the MMA contracts nothing there
            if (out_split && ok[u] && (C & 31)) {
                for (int c = C + sub * 4; c < chunks * 32; c += lanes * 4) {
                    __nv_bfloat16* sp = out_split + r * (long long)(chunks * 64) + (c >> 5) * 64 + (c & 31);
                    *reinterpret_cast<uint2*>(sp) = make_uint2(0u, 0u);
                    *reinterpret_cast<uint2*>(sp + 32) = make_uint2(0u, 0u);
                }
            }
        }
    }
}

// WGS_PN_UNROLL=0: one pixel per lane and trip for every channel count (the former schedule; A/B switch)
static bool pn_unroll() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("WGS_PN_UNROLL");
        on = (e && e[0] == '0') ? 0 : 1;
    }
    return on == 1;
}

template <int MODE>
static void pn_launch(int blocks, cudaStream_t st, int nv, const float* a, const float* dxn, long long R, int C, float eps,
                      float slope, __nv_bfloat16* out_split, float* out_f32, int lanes) {
    if (!pn_unroll()) nv = 4;
    if (nv == 1) pixelnorm_kernel<MODE, 1, 4><<<blocks, PN_THREADS, 0, st>>>(a, dxn, R, C, eps, slope, out_split, out_f32, lanes);
    else if (nv == 2) pixelnorm_kernel<MODE, 2, 2><<<blocks, PN_THREADS, 0, st>>>(a, dxn, R, C, eps, slope, out_split, out_f32, lanes);
    else pixelnorm_kernel<MODE, 4, 1><<<blocks, PN_THREADS, 0, st>>>(a, dxn, R, C, eps, slope, out_split, out_f32, lanes);
}

static int pn_geometry(int C, int* lanes, int* nv) {
    if (C % 4 != 0 || C < 4 || C > 32 * 4 * PN_MAX_V) return -1;
    int l = 1;
    while (l < 32 && l * 4 < C) l <<= 1;         // lanes per pixel: power of two, l * 4 * nv >= C
    *lanes = l;
    *nv = (C / 4 + l - 1) / l;
    if (*nv == 3) *nv = 4;                       // slots come as 1, 2 or 4 (the c < C guard skips the empty one)
    return (*nv <= PN_MAX_V) ? 0 : -1;
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_pixelnorm_pack(const float* a, long long R, int C, float eps, void* out_split, float* out_f32, void* stream) {
    int lanes = 0, nv = 0;
    WGS_REQUIRE(R >= 0 && pn_geometry(C, &lanes, &nv) == 0, "pixelnorm_pack: C must be a multiple of 4, 4 <= C <= 512");
    WGS_REQUIRE(out_split != nullptr || out_f32 != nullptr, "pixelnorm_pack: no output requested");
    if (R == 0) return 0;
    if (!pn_unroll()) nv = 4;
    const long long per_trip = (long long)(32 / lanes) * (4 / nv);             // pixels per warp and trip
    const long long warps = (R + per_trip - 1) / per_trip;
    const int blocks = (int)std::min<long long>((warps * 32 + PN_THREADS - 1) / PN_THREADS, (long long)num_sms() * 8);
    pn_launch<0>(blocks, (cudaStream_t)stream, nv, a, nullptr, R, C, eps, -1.f, (__nv_bfloat16*)out_split, out_f32, lanes);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_pixelnorm_bwd_pack(const float* dxn, const float* a, long long R, int C, float eps, float lrelu_slope,
                                      void* out_split, float* out_f32, void* stream) {
    int lanes = 0, nv = 0;
    WGS_REQUIRE(R >= 0 && pn_geometry(C, &lanes, &nv) == 0, "pixelnorm_bwd_pack: C must be a multiple of 4, 4 <= C <= 512");
    WGS_REQUIRE(out_split != nullptr || out_f32 != nullptr, "pixelnorm_bwd_pack: no output requested");
    if (R == 0) return 0;
    if (!pn_unroll()) nv = 4;
    const long long per_trip = (long long)(32 / lanes) * (4 / nv);
    const long long warps = (R + per_trip - 1) / per_trip;
    const int blocks = (int)std::min<long long>((warps * 32 + PN_THREADS - 1) / PN_THREADS, (long long)num_sms() * 8);
    pn_launch<1>(blocks, (cudaStream_t)stream, nv, a, dxn, R, C, eps, lrelu_slope, (__nv_bfloat16*)out_split, out_f32, lanes);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
