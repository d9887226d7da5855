// Data-gradient glue of the StyleGAN2 synthesis network (the frozen generator gets no weight
// gradients; only d(loss)/d(styles) is needed to reach the latent shift — SURVEY.md §7 step 4).
// For one styled layer   a = sqrt2*lrelu( d[n,c]*conv(W, s[n,i]*a_prev) + nw*noise + b )   the backward is
//   dpre  = da * sqrt2 * (a > 0 ? 1 : 0.2)                         (sg2_act_bwd)
//   dd    = sum_p dpre * yraw,   yraw = (pre - nw*noise - b)/d     (sg2_act_bwd, pre recovered from a)
//   dx~   = conv^T(W, d * dpre)                                    (tensor-core conv, csrc/conv.cu)
//   da_prev = s * dx~ ;  ds = sum_p dx~ * a_prev                   (sg2_mod_bwd)
// replacing autograd through F.conv2d(groups=B) / F.conv_transpose2d and the per-sample weight tensors
// (models/StyleGAN2/model.py:187-228), whose weight-gradient GEMMs are never formed here.
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

constexpr int RED_THREADS = 256;
constexpr float SQRT2 = 1.41421356237309515f;

// Shared tail: reduce per-thread float4 partials over the pixel lanes of the block, then one atomicAdd
// per channel per block.  Thread t owns channel quad t % C4 and pixel lane t / C4.
__device__ __forceinline__ void reduce_quads_to_global(const float acc[4], float* sm, int C, int C4, int PL,
                                                       float* dst /* [C] of image n */) {
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    *reinterpret_cast<float4*>(sm + (size_t)pl * C + q * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += RED_THREADS) {
        float t = 0.f;
        for (int l = 0; l < PL; ++l) t += sm[(size_t)l * C + c];
        atomicAdd(dst + c, t);
    }
}

// da, a: [N, P, C];  dpre out: [N, P, C] (may alias da);  dd: [N, C] accumulated.
__global__ void __launch_bounds__(RED_THREADS)
sg2_act_bwd_kernel(const float* __restrict__ da, const float* __restrict__ a, const float* __restrict__ demod,
                   const float* __restrict__ bias, const float* __restrict__ noise, float noise_w,
                   float* __restrict__ dpre, float* __restrict__ dd, __nv_bfloat16* __restrict__ g_split, long long P,
                   int C) {
    extern __shared__ float sm[];
    const int n = blockIdx.y;
    const int C4 = C >> 2, PL = RED_THREADS / C4;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    const long long per = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    const float4 d4 = __ldg(reinterpret_cast<const float4*>(demod + (size_t)n * C + q * 4));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + q * 4));
    const float dv[4] = {d4.x, d4.y, d4.z, d4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long p = p0 + pl; p < p1; p += PL) {
        const size_t off = ((size_t)n * P + p) * C + q * 4;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(da + off));
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(a + off));
        const float nz = noise ? noise_w * __ldg(noise + p) : 0.f;
        const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
        float o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool pos = av[k] > 0.f;
            const float pre = pos ? av[k] * (1.f / SQRT2) : av[k] * (1.f / (0.2f * SQRT2));
            o[k] = gv[k] * (pos ? SQRT2 : 0.2f * SQRT2);
            acc[k] += o[k] * (pre - nz - bv[k]) / dv[k];
        }
        if (dpre) *reinterpret_cast<float4*>(dpre + off) = make_float4(o[0], o[1], o[2], o[3]);
        if (g_split) {                                   // operand of the data-gradient conv: d[n,c] * dpre, split32
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(o[k] * dv[k], hi[k], lo[k]);
            __nv_bfloat16* sp = g_split + ((size_t)n * P + p) * (size_t)(C * 2) + (size_t)(q >> 3) * 64 + ((q & 7) << 2);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
    reduce_quads_to_global(acc, sm, C, C4, PL, dd + (size_t)n * C);
}

// dx: [N,P,C] gradient wrt the modulated input; a_prev: [N,P,C] (or [P,C] broadcast when a_bcast);
// da_prev = s * dx (+ existing when accumulate; skipped when NULL); ds[n,c] += sum_p dx * a_prev.
__global__ void __launch_bounds__(RED_THREADS)
sg2_mod_bwd_kernel(const float* __restrict__ dx, const float* __restrict__ a_prev, int a_bcast,
                   const float* __restrict__ s, long long s_ld, float* __restrict__ da_prev, int accumulate,
                   float* __restrict__ ds, long long ds_ld, long long P, int C) {
    extern __shared__ float sm[];
    const int n = blockIdx.y;
    const int C4 = C >> 2, PL = RED_THREADS / C4;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    const long long per = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(s + (size_t)n * s_ld + q * 4));
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long p = p0 + pl; p < p1; p += PL) {
        const size_t off = ((size_t)n * P + p) * C + q * 4;
        const size_t aoff = a_bcast ? (size_t)p * C + q * 4 : off;
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(dx + off));
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(a_prev + aoff));
        acc[0] += g4.x * a4.x; acc[1] += g4.y * a4.y; acc[2] += g4.z * a4.z; acc[3] += g4.w * a4.w;
        if (da_prev) {
            float4 o = make_float4(g4.x * s4.x, g4.y * s4.y, g4.z * s4.z, g4.w * s4.w);
            if (accumulate) {
                const float4 e = *reinterpret_cast<const float4*>(da_prev + off);
                o.x += e.x; o.y += e.y; o.z += e.z; o.w += e.w;
            }
            *reinterpret_cast<float4*>(da_prev + off) = o;
        }
    }
    reduce_quads_to_global(acc, sm, C, C4, PL, ds + (size_t)n * ds_ld);
}

// ToRGB backward: da[n,p,c] (+)= sum_o drgb[n,p,o] * wscale * W[o,c] * s[n,c];
//                 ds[n,c] += sum_p a[n,p,c] * sum_o drgb[n,p,o] * wscale * W[o,c]
__global__ void __launch_bounds__(RED_THREADS)
sg2_torgb_bwd_kernel(const float* __restrict__ drgb, const float* __restrict__ a, const float* __restrict__ s,
                     long long s_ld, const float* __restrict__ W, float wscale, float* __restrict__ da,
                     int accumulate, float* __restrict__ ds, long long ds_ld, long long P, int C) {
    extern __shared__ float sm[];
    const int n = blockIdx.y;
    const int C4 = C >> 2, PL = RED_THREADS / C4;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    const long long per = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    float w[3][4], sv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        sv[k] = __ldg(s + (size_t)n * s_ld + q * 4 + k);
#pragma unroll
        for (int o = 0; o < 3; ++o) w[o][k] = wscale * __ldg(W + (size_t)o * C + q * 4 + k);
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long p = p0 + pl; p < p1; p += PL) {
        const float* g = drgb + ((size_t)n * P + p) * 3;
        const float g0 = __ldg(g), g1 = __ldg(g + 1), g2 = __ldg(g + 2);
        const size_t off = ((size_t)n * P + p) * C + q * 4;
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(a + off));
        const float av[4] = {a4.x, a4.y, a4.z, a4.w};
        float o4[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float t = g0 * w[0][k] + g1 * w[1][k] + g2 * w[2][k];
            acc[k] += t * av[k];
            o4[k] = t * sv[k];
        }
        if (accumulate) {
            const float4 e = *reinterpret_cast<const float4*>(da + off);
            o4[0] += e.x; o4[1] += e.y; o4[2] += e.z; o4[3] += e.w;
        }
        *reinterpret_cast<float4*>(da + off) = make_float4(o4[0], o4[1], o4[2], o4[3]);
    }
    reduce_quads_to_global(acc, sm, C, C4, PL, ds + (size_t)n * ds_ld);
}

// One pass over a layer boundary of the data-gradient chain, fusing sg2_mod_bwd (of the layer above), sg2_torgb_bwd
// and sg2_act_bwd (of this layer):
//   da    = s_up[n,c] * dx_up            (gradient arriving through the next conv's modulated input; optional)
//         + s_rgb[n,c] * sum_o drgb[o] * wscale * Wrgb[o,c]                         (ToRGB branch; optional)
//   ds_up[n,c]  += sum_p dx_up * a ;   ds_rgb[n,c] += sum_p a * sum_o drgb[o] * wscale * Wrgb[o,c]
//   dpre  = da * sqrt2 * (a > 0 ? 1 : 0.2) ;   dd[n,c] += sum_p dpre * (pre - nw*noise - b) / d
//   out:  split32(d * dpre)  (operand of this layer's data-gradient conv)  and / or  dpre in fp32
// Every tensor is read once: 3 reads + 1 write instead of the 9 passes of the three separate kernels.
__global__ void __launch_bounds__(RED_THREADS)
sg2_layer_bwd_kernel(const float* __restrict__ dx_up, const float* __restrict__ s_up, long long s_up_ld,
                     float* __restrict__ ds_up, long long ds_up_ld, const float* __restrict__ drgb,
                     const float* __restrict__ s_rgb, long long s_rgb_ld, const float* __restrict__ Wrgb, float wscale,
                     float* __restrict__ ds_rgb, long long ds_rgb_ld, const float* __restrict__ a,
                     const float* __restrict__ demod, const float* __restrict__ bias, const float* __restrict__ noise,
                     float noise_w, float* __restrict__ dd, float* __restrict__ dpre, __nv_bfloat16* __restrict__ g_split,
                     long long P, int C) {
    extern __shared__ float sm[];
    const int n = blockIdx.y;
    const int C4 = C >> 2, PL = RED_THREADS / C4;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    const long long per = (P + gridDim.x - 1) / gridDim.x;
    const long long p0 = blockIdx.x * per, p1 = min(P, p0 + per);
    float dv[4], bv[4], su[4] = {0.f, 0.f, 0.f, 0.f}, sr[4] = {0.f, 0.f, 0.f, 0.f}, w[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int c = q * 4 + k;
        dv[k] = __ldg(demod + (size_t)n * C + c);
        bv[k] = __ldg(bias + c);
        if (dx_up) su[k] = __ldg(s_up + (size_t)n * s_up_ld + c);
        if (drgb) {
            sr[k] = __ldg(s_rgb + (size_t)n * s_rgb_ld + c);
#pragma unroll
            for (int o = 0; o < 3; ++o) w[o][k] = wscale * __ldg(Wrgb + (size_t)o * C + c);
        } else {
#pragma unroll
            for (int o = 0; o < 3; ++o) w[o][k] = 0.f;
        }
    }
    float acc_u[4] = {0.f, 0.f, 0.f, 0.f}, acc_r[4] = {0.f, 0.f, 0.f, 0.f}, acc_d[4] = {0.f, 0.f, 0.f, 0.f};
    // LBW_U pixels per trip with every load issued before the first use: one pixel per trip leaves a single batch of
    // loads in flight per thread and 66 % of the stall samples on the two first uses (profiles/r01_misc_ncu.md)
    constexpr int LBW_U = 4;
    for (long long pb = p0 + pl; pb < p1; pb += (long long)PL * LBW_U) {
        float4 a4[LBW_U], g4[LBW_U];
        float r0[LBW_U], r1[LBW_U], r2[LBW_U], nzv[LBW_U];
#pragma unroll
        for (int u = 0; u < LBW_U; ++u) {
            const long long pu = pb + (long long)u * PL;
            const long long p = pu < p1 ? pu : pb;                       // clamp: the tail re-reads a valid pixel
            const size_t off = ((size_t)n * P + p) * C + q * 4;
            a4[u] = __ldg(reinterpret_cast<const float4*>(a + off));
            if (dx_up) g4[u] = __ldg(reinterpret_cast<const float4*>(dx_up + off));
            if (drgb) {
                const float* g = drgb + ((size_t)n * P + p) * 3;
                r0[u] = __ldg(g); r1[u] = __ldg(g + 1); r2[u] = __ldg(g + 2);
            }
            nzv[u] = noise ? noise_w * __ldg(noise + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < LBW_U; ++u) {
            const long long p = pb + (long long)u * PL;
            if (p >= p1) break;
            const size_t off = ((size_t)n * P + p) * C + q * 4;
            const float av[4] = {a4[u].x, a4[u].y, a4[u].z, a4[u].w};
            float da[4] = {0.f, 0.f, 0.f, 0.f};
            if (dx_up) {
                const float gv[4] = {g4[u].x, g4[u].y, g4[u].z, g4[u].w};
#pragma unroll
                for (int k = 0; k < 4; ++k) { acc_u[k] += gv[k] * av[k]; da[k] = gv[k] * su[k]; }
            }
            if (drgb) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float t = r0[u] * w[0][k] + r1[u] * w[1][k] + r2[u] * w[2][k];
                    acc_r[k] += t * av[k];
                    da[k] += t * sr[k];
                }
            }
            const float nz = nzv[u];
            float o[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const bool pos = av[k] > 0.f;
                const float pre = pos ? av[k] * (1.f / SQRT2) : av[k] * (1.f / (0.2f * SQRT2));
                o[k] = da[k] * (pos ? SQRT2 : 0.2f * SQRT2);
                acc_d[k] += o[k] * (pre - nz - bv[k]) / dv[k];
            }
            if (dpre) *reinterpret_cast<float4*>(dpre + off) = make_float4(o[0], o[1], o[2], o[3]);
            if (g_split) {
                __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) split_bf16(o[k] * dv[k], hi[k], lo[k]);
                __nv_bfloat16* sp = g_split + ((size_t)n * P + p) * (size_t)(C * 2) + (size_t)(q >> 3) * 64 + ((q & 7) << 2);
                *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
                *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
            }
        }
    }
    reduce_quads_to_global(acc_d, sm, C, C4, PL, dd + (size_t)n * C);
    if (dx_up) {
        __syncthreads();
        reduce_quads_to_global(acc_u, sm, C, C4, PL, ds_up + (size_t)n * ds_up_ld);
    }
    if (drgb) {
        __syncthreads();
        reduce_quads_to_global(acc_r, sm, C, C4, PL, ds_rgb + (size_t)n * ds_rgb_ld);
    }
}

// Transpose of the FIR skip upsample (Upsample, model.py:29-45): drgb [N,H,W,3] -> dprev [N,H/2,W/2,3]
//   dprev[i] = k0*d[2i-1] + k1*d[2i] + k2*d[2i+1] + k3*d[2i+2]   per axis
__global__ void rgb_up_bwd_kernel(const float* __restrict__ drgb, float* __restrict__ dprev, int N, int H, int W,
                                  float k0, float k1, float k2, float k3) {
    const int h2 = H >> 1, w2 = W >> 1;
    const long long total = (long long)N * h2 * w2;
    const float kk[4] = {k0, k1, k2, k3};
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w2);
        const int y = (int)((i / w2) % h2);
        const int n = (int)(i / ((long long)w2 * h2));
        float o[3] = {0.f, 0.f, 0.f};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int yy = 2 * y - 1 + a;
            if (yy < 0 || yy >= H) continue;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int xx = 2 * x - 1 + b;
                if (xx < 0 || xx >= W) continue;
                const float wgt = kk[a] * kk[b];
                const float* g = drgb + (((size_t)n * H + yy) * W + xx) * 3;
                o[0] += wgt * __ldg(g); o[1] += wgt * __ldg(g + 1); o[2] += wgt * __ldg(g + 2);
            }
        }
        float* d = dprev + i * 3;
        d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
    }
}

static int red_blocks(long long P, int N) {
    const long long want = std::max<long long>(1, ((long long)num_sms() * 16) / std::max(1, N));
    return (int)std::max<long long>(1, std::min<long long>(want, (P + 31) / 32));
}

static bool red_ok(int C) { return C >= 4 && C % 4 == 0 && (C / 4) <= RED_THREADS && RED_THREADS % (C / 4) == 0; }

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_sg2_act_bwd(const float* da, const float* a, const float* demod, const float* bias,
                               const float* noise, float noise_w, float* dpre, float* dd, void* g_split, int N,
                               long long P, int C, void* stream) {
    WGS_REQUIRE(!g_split || C % 32 == 0, "sg2_act_bwd: split32 output needs C % 32 == 0");
    WGS_REQUIRE(red_ok(C), "sg2_act_bwd: channel count must be a power of two in [4, 1024]");
    const size_t smem = (size_t)(RED_THREADS / (C / 4)) * C * sizeof(float);
    sg2_act_bwd_kernel<<<dim3(red_blocks(P, N), N), RED_THREADS, smem, (cudaStream_t)stream>>>(
        da, a, demod, bias, noise, noise_w, dpre, dd, (__nv_bfloat16*)g_split, P, C);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_mod_bwd(const float* dx, const float* a_prev, int a_bcast, const float* s, long long s_ld,
                               float* da_prev, int accumulate, float* ds, long long ds_ld, int N, long long P, int C,
                               void* stream) {
    WGS_REQUIRE(red_ok(C), "sg2_mod_bwd: channel count must be a power of two in [4, 1024]");
    const size_t smem = (size_t)(RED_THREADS / (C / 4)) * C * sizeof(float);
    sg2_mod_bwd_kernel<<<dim3(red_blocks(P, N), N), RED_THREADS, smem, (cudaStream_t)stream>>>(
        dx, a_prev, a_bcast, s, s_ld, da_prev, accumulate, ds, ds_ld, P, C);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_torgb_bwd(const float* drgb, const float* a, const float* s, long long s_ld, const float* W,
                                 float wscale, float* da, int accumulate, float* ds, long long ds_ld, int N,
                                 long long P, int C, void* stream) {
    WGS_REQUIRE(red_ok(C), "sg2_torgb_bwd: channel count must be a power of two in [4, 1024]");
    const size_t smem = (size_t)(RED_THREADS / (C / 4)) * C * sizeof(float);
    sg2_torgb_bwd_kernel<<<dim3(red_blocks(P, N), N), RED_THREADS, smem, (cudaStream_t)stream>>>(
        drgb, a, s, s_ld, W, wscale, da, accumulate, ds, ds_ld, P, C);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_rgb_up_bwd(const float* drgb, float* dprev, int N, int H, int W, const float* h_taps4,
                                  void* stream) {
    WGS_REQUIRE(N > 0 && H % 2 == 0 && W % 2 == 0, "rgb_up_bwd: even sizes required");
    const float k0 = h_taps4 ? h_taps4[0] : 0.25f, k1 = h_taps4 ? h_taps4[1] : 0.75f,
                k2 = h_taps4 ? h_taps4[2] : 0.75f, k3 = h_taps4 ? h_taps4[3] : 0.25f;
    const long long total = (long long)N * (H / 2) * (W / 2);
    rgb_up_bwd_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), 256, 0,
                        (cudaStream_t)stream>>>(drgb, dprev, N, H, W, k0, k1, k2, k3);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_layer_bwd(const float* dx_up, const float* s_up, long long s_up_ld, float* ds_up, long long ds_up_ld,
                                 const float* drgb, const float* s_rgb, long long s_rgb_ld, const float* Wrgb, float wscale,
                                 float* ds_rgb, long long ds_rgb_ld, const float* a, const float* demod, const float* bias,
                                 const float* noise, float noise_w, float* dd, float* dpre, void* g_split, int N,
                                 long long P, int C, void* stream) {
    WGS_REQUIRE(red_ok(C), "sg2_layer_bwd: channel count must be a power of two in [4, 1024]");
    WGS_REQUIRE(!g_split || C % 32 == 0, "sg2_layer_bwd: split32 output needs C % 32 == 0");
    WGS_REQUIRE(dx_up || drgb, "sg2_layer_bwd: no incoming gradient");
    const size_t smem = (size_t)(RED_THREADS / (C / 4)) * C * sizeof(float);
    sg2_layer_bwd_kernel<<<dim3(red_blocks(P, N), N), RED_THREADS, smem, (cudaStream_t)stream>>>(
        dx_up, s_up, s_up_ld, ds_up, ds_up_ld, drgb, s_rgb, s_rgb_ld, Wrgb, wscale, ds_rgb, ds_rgb_ld, a, demod, bias, noise,
        noise_w, dd, dpre, (__nv_bfloat16*)g_split, P, C);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
