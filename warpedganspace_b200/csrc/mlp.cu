// A chain of equal-width small-batch linears in ONE launch on a thread-block cluster: StyleGAN2's mapping network
// (PixelNorm + 8 x EqualLinear 512 -> 512 with fused leaky-ReLU, models/StyleGAN2/model.py:110-131,291-295) and its backward
// pass.  As eight dependent launches each layer cost ~16 us of pure latency (a 1 MB weight matrix, 8 rows of input):
// 0.13 ms forward + 0.13 ms backward per step, 1.5 % of it.  Here a cluster of 8 CTAs keeps the activations of every batch row
// in shared memory; per layer each CTA computes 1/8 of the output features (one warp per feature, the weight row streamed
// once with 128-bit loads and reused for every batch row), writes its slice into the NEXT-layer buffer of all eight CTAs
// through distributed shared memory (st.shared::cluster) and the cluster barrier is the only synchronisation between
// layers - no grid-wide sync, no round trip through L2 for the activations.
//
//   layer l:  out_l[b, o] = epi( wscale * sum_i f(in[b, i], aux_l[b, i]) * W_l[o, i] + bscale * bias_l[o] ),  in = out_{l-1}
//   f = identity (forward) or in * lrelu'(aux) with aux = the layer's forward output (backward of the fused leaky-ReLU);
//   epi = sqrt(2) * lrelu_0.2 (forward) or identity (backward).
#include "common.cuh"
#include "ptx.cuh"
#include "wgs_b200.h"

namespace wgs {

constexpr int MLP_CLUSTER = 8;
constexpr int MLP_THREADS = 512;                  // 16 warps x MLP_F output features each per pass (d = 512: one pass per layer)
constexpr int MLP_F = 4;                          // features per warp per pass: every activation load is used for four rows of W
constexpr int MLP_BT = 8;

struct MlpChain {
    wgs_mlp_layer layer[WGS_MLP_MAX_LAYERS];
    const float* x;
    long long x_ld;
    int L, B, d, in_mode, epi;
    float wscale, bscale;
};

__device__ __forceinline__ void st_dsmem_f32(uint32_t cluster_addr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(cluster_addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(MLP_THREADS)
mlp_chain_kernel(const __grid_constant__ MlpChain g) {
    extern __shared__ float act[];                                  // [2][B][d] ping-pong activation buffers
    const int B = g.B, d = g.d;
    const uint32_t rank = ptx::cluster_ctarank();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* cur = act;
    float* nxt = act + (size_t)B * d;
    // Every weight row this CTA will ever read (L layers x d/8 rows x d floats = 1 MB at d = 512, L = 8) and the backward's
    // aux activations are requested into L2 up front: the layers are strictly dependent, so without this each one pays a
    // full DRAM round trip for its rows after the cluster barrier (measured 11 us per layer for 128 KB per CTA).
    {
        const int per0 = d / MLP_CLUSTER;
        const int lines_per_layer = per0 * d / 32;                    // 128-byte lines of this CTA's row slice
        for (int l = 0; l < g.L; ++l) {
            const char* base = reinterpret_cast<const char*>(g.layer[l].W + (size_t)rank * per0 * d);
            for (int i = threadIdx.x; i < lines_per_layer; i += MLP_THREADS)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * 128));
            if (g.in_mode == 2) {
                const char* ab = reinterpret_cast<const char*>(g.layer[l].aux);
                for (int i = threadIdx.x; i < B * d / 32; i += MLP_THREADS)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(ab + (size_t)i * 128));
            }
        }
    }
    for (int i = threadIdx.x; i < B * d; i += MLP_THREADS) cur[i] = __ldg(g.x + (size_t)(i / d) * g.x_ld + i % d);
    // cluster-wide, not CTA-wide: a peer's shared memory may only be written once that CTA is known to have started
    // (compute-sanitizer racecheck: "block that might not have entered yet")
    ptx::cluster_sync();
    const int per = d / MLP_CLUSTER;
    for (int l = 0; l < g.L; ++l) {
        const wgs_mlp_layer& q = g.layer[l];
        if (g.in_mode == 2) {                                       // input * lrelu'(forward output of this layer), in place
            const float up = 1.41421356237309515f, dn = 0.2f * 1.41421356237309515f;
            for (int i = threadIdx.x; i < B * d; i += MLP_THREADS) cur[i] *= (__ldg(q.aux + i) > 0.f ? up : dn);
            __syncthreads();
        }
        const uint32_t nxt_s = ptx::smem_u32(nxt);
        // A warp computes MLP_F output features at once: the weight rows (MLP_F x 4 float4 per lane for d = 512) are all in
        // flight together and every activation quad read from shared memory feeds MLP_F dot products - one feature per warp
        // re-read the whole [B, d] activation block per feature, 1 MB of shared-memory traffic per CTA and layer.
        for (int o0 = rank * per + warp * MLP_F; o0 < (int)(rank + 1) * per; o0 += (MLP_THREADS / 32) * MLP_F) {
            for (int b0 = 0; b0 < B; b0 += MLP_BT) {
                float acc[MLP_F][MLP_BT];
#pragma unroll
                for (int f = 0; f < MLP_F; ++f)
#pragma unroll
                    for (int t = 0; t < MLP_BT; ++t) acc[f][t] = 0.f;
                for (int i0 = lane * 4; i0 < d; i0 += 512) {
                    float4 w4[MLP_F][4];
#pragma unroll
                    for (int f = 0; f < MLP_F; ++f) {
                        const bool f_ok = o0 + f < (int)(rank + 1) * per;
                        const float* wrow = q.W + (size_t)(f_ok ? o0 + f : o0) * d;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = i0 + j * 128;
                            w4[f][j] = (f_ok && i < d) ? __ldg(reinterpret_cast<const float4*>(wrow + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
#pragma unroll
                    for (int t = 0; t < MLP_BT; ++t) {
                        if (b0 + t < B) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int i = i0 + j * 128;
                                if (i < d) {
                                    const float4 x4 = *reinterpret_cast<const float4*>(cur + (size_t)(b0 + t) * d + i);
#pragma unroll
                                    for (int f = 0; f < MLP_F; ++f)
                                        acc[f][t] += w4[f][j].x * x4.x + w4[f][j].y * x4.y + w4[f][j].z * x4.z + w4[f][j].w * x4.w;
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int f = 0; f < MLP_F; ++f) {
#pragma unroll
                    for (int t = 0; t < MLP_BT; ++t) acc[f][t] = warp_sum(acc[f][t]);
                }
                // lane (f * MLP_BT + t) finishes feature o0 + f of batch row b0 + t
                const int f_mine = lane / MLP_BT, t_mine = lane % MLP_BT;
                float v = 0.f;
#pragma unroll
                for (int f = 0; f < MLP_F; ++f)
#pragma unroll
                    for (int t = 0; t < MLP_BT; ++t) if (f == f_mine && t == t_mine) v = acc[f][t];
                const int o = o0 + f_mine, b = b0 + t_mine;
                if (f_mine < MLP_F && o < (int)(rank + 1) * per && b < B) {
                    v *= g.wscale;
                    if (q.bias) v += g.bscale * __ldg(q.bias + o);
                    if (g.epi == 1) v = 1.41421356237309515f * (v > 0.f ? v : 0.2f * v);
                    if (q.out) q.out[(size_t)b * d + o] = v;
                    const uint32_t off = nxt_s + (uint32_t)((size_t)b * d + o) * 4u;
#pragma unroll
                    for (int r = 0; r < MLP_CLUSTER; ++r) st_dsmem_f32(ptx::map_to_cta(off, (uint32_t)r), v);
                }
            }
        }
        ptx::cluster_sync();                                         // every CTA's next-layer buffer is complete
        float* t = cur; cur = nxt; nxt = t;
    }
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_mlp_layer_size(void) { return (int)sizeof(wgs_mlp_layer); }

extern "C" int wgs_mlp_chain(const float* x, long long x_ld, const wgs_mlp_layer* h_layers, int L, int B, int d, float wscale,
                             float bscale, int in_mode, int epi, void* stream) {
    WGS_REQUIRE(x && h_layers && L >= 1 && L <= WGS_MLP_MAX_LAYERS, "mlp_chain: 1..WGS_MLP_MAX_LAYERS layers");
    WGS_REQUIRE(B >= 1 && d >= 128 && d % 128 == 0 && d % MLP_CLUSTER == 0, "mlp_chain: d must be a multiple of 128");
    WGS_REQUIRE(in_mode == 0 || in_mode == 2, "mlp_chain: in_mode 0 (identity) or 2 (times lrelu'(aux))");
    WGS_REQUIRE(epi == 0 || epi == 1, "mlp_chain: epi 0 (linear) or 1 (sqrt2 * lrelu)");
    const size_t smem = 2 * (size_t)B * d * sizeof(float);
    WGS_REQUIRE(smem <= 200 * 1024, "mlp_chain: batch too large for the shared-memory activation buffers");
    MlpChain g;
    memset(&g, 0, sizeof(g));
    for (int l = 0; l < L; ++l) {
        WGS_REQUIRE(h_layers[l].W != nullptr && (in_mode != 2 || h_layers[l].aux != nullptr), "mlp_chain: missing weight / aux pointer");
        g.layer[l] = h_layers[l];
    }
    g.x = x; g.x_ld = x_ld; g.L = L; g.B = B; g.d = d; g.in_mode = in_mode; g.epi = epi; g.wscale = wscale; g.bscale = bscale;
    static bool attr = false;
    if (!attr) {
        WGS_CUDA(cudaFuncSetAttribute(mlp_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr = true;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(MLP_CLUSTER, 1, 1);
    cfg.blockDim = dim3(MLP_THREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = MLP_CLUSTER;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    WGS_CUDA(cudaLaunchKernelEx(&cfg, mlp_chain_kernel, g));
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
