// Tap-list implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// One kernel family serves every dense contraction on the hot path: StyleGAN2's modulated 3x3 convs
// (restated as  d[b,o] * conv(W, s[b,i]*x) — no per-sample weights, no grouped conv; reference
// models/StyleGAN2/model.py:187-228), its stride-2 transposed convs (four output phases, each a short
// tap list; :201-212), ProgGAN / BigGAN / SNGAN 3x3 and 1x1 convs, the Reconstructor's ResNet-18 convs
// (7x7/2, 3x3/1, 3x3/2, 1x1/2) and all of their data-gradients (same kernel, flipped/transposed taps).
//
//   out[n, oy*ystep+y0, ox*xstep+x0, co] = act( alpha[n,co] * sum_{t,ci} W[t][co][ci] *
//                                           in[n, oy*stride+dy_t, ox*stride+dx_t, ci] + beta[co] )
//
// Operands are "split32" bf16: every 32 fp32 channels are stored as one 128-byte row
// [hi(32) | lo(32)] with x = hi + lo to 2^-17; the product is formed as hi*hi + hi*lo + lo*hi with fp32
// accumulation in TMEM (3 MMAs per logical MAC, error ~1e-5 instead of bf16's ~1e-2).
//
// Structure (one 128-pixel x BN-channel output tile per CTA):
//   warp 0    TMA producer: per (tap, 32-channel chunk) one 5-D box load of the shifted input patch
//             (out-of-bounds = zero fill = conv padding; elementStrides = conv stride) and one 4-D box
//             of the weight slice, both 128B-swizzled, into an N-stage mbarrier ring;
//   warp 1    TMEM allocation + single-thread tcgen05.mma issue (6 MMAs of K=16 per stage);
//   warps 2-5 epilogue: tcgen05.ld of the fp32 accumulators, alpha/beta/activation, NHWC stores.
#include "common.cuh"
#include "ptx.cuh"
#include "wgs_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace wgs {

constexpr int CONV_THREADS = 192;
constexpr int A_STAGE_BYTES = 128 * 128;

struct ConvKernelParams {
    int out_n, grid_h, grid_w;
    int bh, bw, bn;
    int tiles_x, tiles_y, tiles_n, n_tiles_co;
    int in_stride;
    int c_chunks, cout, BN, stages, tmem_cols;
    int num_taps;
    float* out;
    long long out_sn, out_sy, out_sx;
    int out_y0, out_x0, out_ystep, out_xstep;
    const float* alpha;
    const float* beta;
    const float* noise;      // [grid plane in output coords: out_h x out_w] or null
    float noise_w;
    int noise_ld;
    int act, accumulate;
    // fused consumers of the activation (all optional)
    void* out_split;             // split32 [out_n][grid_h][grid_w][cout/32][64] of act * split_scale[n, co]
    const float* split_scale;    // [out_n, cout] (row stride split_scale_ld) or null (= 1)
    long long split_scale_ld;
    int out_from_n;              // fp32 `out` is written only for images n >= out_from_n
    const float* rgb_w;          // [out_n][3][cout] modulated ToRGB weights; rgb_out += act . rgb_w
    float* rgb_out;              // [out_n][grid_h][grid_w][3], pre-initialised with bias + upsampled skip
    // halo variant: the (bh + wy - 1) x (bw + wx - 1) input patch of a chunk is loaded once and every tap
    // addresses it through its UMMA descriptor
    int dy0, dx0, halo_h, halo_w, a_stage_bytes;
    int group_size, group_w, out_h, out_w;   // phase-packed output (0 = off)
    int ksplit;                              // cluster split-K: CTAs per output tile (1 = off)
    int coalesce;                            // stage epilogue stores through shared memory (coalesced 64-byte rows)
    int tiles_per_cta, x_groups;             // multi-tile halo kernel: consecutive x tiles handled by one CTA
    int w_cout;                              // rows of the weight tensor (stacked layout addressing, SIMT twin)
    float pixnorm_eps;                       // > 0: out_split holds pixel_norm(act) = act * rsqrt(mean_c act^2 + eps) (ProgGAN)
    int split_hw;                            // out_split addressed by OUTPUT pixel (strided output mapping), dims out_h x out_w
    float* stat_sum;                         // BatchNorm statistics of the output (plain epilogue): shifted sum / sum of squares
    float* stat_sumsq;
    const float* stat_shift;
    signed char tap_dy[WGS_MAX_TAPS], tap_dx[WGS_MAX_TAPS];
    unsigned char tap_w[WGS_MAX_TAPS];
};

__device__ __forceinline__ float apply_act(float v, int act) {
    if (act == 1) return v > 0.f ? v : 0.f;
    if (act == 2) return v > 0.f ? v : 0.2f * v;
    if (act == 3) return 1.41421356237309515f * (v > 0.f ? v : 0.2f * v);      // FusedLeakyReLU
    if (act == 4) return tanhf(v);                                              // BigGAN / SNGAN output layer
    return v;
}

// Warp-collective coalesced store of one 64-byte row per lane.  In the TMEM 32x32b layout a thread owns a pixel, so
// a direct 128-bit store instruction touches 32 different cache lines, 16 bytes each; with two or three output tensors
// per tile the LSU/L2 request rate, not bandwidth, bounds the epilogue (fused 32 -> 32 @1024^2: 1.04 ms vs 0.58 ms with
// the stores removed, profiles/r01_conv_ncu_step.md).  The rows are transposed through a 2 KB per-warp staging
// buffer (XOR-swizzled 16-byte chunks, conflict-free both ways) so that every store instruction writes 8 rows x 64
// contiguous bytes.  Chunk c of a row goes to byte offset (c & 1) * 16 + (c >> 1) * hi_stride of that lane's pointer
// (fp32 block: hi_stride 32 -> 64 contiguous bytes; split32 block: hi_stride 64 -> hi half, lo half).
__device__ __forceinline__ void warp_store_rows64(uint32_t stg_s, int lane, const uint4 (&q)[4], const void* row_ptr,
                                                  uint32_t hi_stride) {
    const uint32_t sw = ((uint32_t)lane >> 1) & 3u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg_s + (uint32_t)lane * 64u + (((uint32_t)k ^ sw) << 4)),
                     "r"(q[k].x), "r"(q[k].y), "r"(q[k].z), "r"(q[k].w) : "memory");
    __syncwarp();
    const unsigned long long mine = reinterpret_cast<unsigned long long>(row_ptr);
    const uint32_t c = (uint32_t)lane & 3u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = j * 8 + (lane >> 2);
        uint4 t;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w)
                     : "r"(stg_s + (uint32_t)r * 64u + ((c ^ (((uint32_t)r >> 1) & 3u)) << 4)));
        const unsigned long long pr = __shfl_sync(0xffffffffu, mine, r);
        if (pr) *reinterpret_cast<uint4*>(pr + (c & 1u) * 16u + (c >> 1) * hi_stride) = t;
    }
    __syncwarp();
}


// Phase-packed output whose groups are narrower than a 16-channel block on a DENSE few-channel NHWC tensor (out_sx == GSZ,
// two phases per row): the stem's data gradient, [N, H, W, 3] or [N, H, W, 6].  Column cc of the accumulator row belongs
// to output row gy = cc / (2*GSZ) at float offset cc % (2*GSZ) from the row's first pixel - constant divisors, staged
// constants, no per-element 64-bit address arithmetic (the generic element-wise path below made this launch
// epilogue-bound: 0.46 ms with the tensor pipe 24 % active).
template <int GSZ>
__device__ __forceinline__ void small_group_store(const float (&v)[16], int c, int co, int cout, float* dst, long long out_sy,
                                                  bool lane_ok, int py0, int out_h, uint32_t ep_s, int BN, float nz, int act) {
    constexpr int RUN = 2 * GSZ;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int cc = co + i;
        if (cc < cout) {
            const int gy = cc / RUN, rem = cc - gy * RUN;
            float r = fmaf(v[i], ptx::lds32(ep_s + (uint32_t)(c + i) * 4u), nz) + ptx::lds32(ep_s + (uint32_t)(BN + c + i) * 4u);
            r = apply_act(r, act);
            if (lane_ok && py0 + gy < out_h) dst[(long long)gy * out_sy + rem] = r;
        }
    }
}

// Column sums of a 32 (lanes = pixels) x 16 (registers = channels) block by recursive halving: 16 shuffles instead of the 80 of
// sixteen butterfly reductions.  Afterwards lane l holds the sum of channel ((l >> 1) & 15 with its bits reversed as below)
// over all 32 lanes: channel = 8*bit4 + 4*bit3 + 2*bit2 + bit1 of the lane index (both lanes of a pair hold the same value).
__device__ __forceinline__ float warp_colsum16(const float (&v)[16], int lane) {
    float a[8], b[4], c[2], d;
    const bool u4 = lane & 16, u3 = lane & 8, u2 = lane & 4, u1 = lane & 2;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (u4 ? v[i + 8] : v[i]) + __shfl_xor_sync(0xffffffffu, u4 ? v[i] : v[i + 8], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (u3 ? a[i + 4] : a[i]) + __shfl_xor_sync(0xffffffffu, u3 ? a[i] : a[i + 4], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (u2 ? b[i + 2] : b[i]) + __shfl_xor_sync(0xffffffffu, u2 ? b[i] : b[i + 2], 4);
    d = (u1 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, u1 ? c[0] : c[1], 2);
    return d + __shfl_xor_sync(0xffffffffu, d, 1);
}

// Epilogue shared by the tensor-core conv kernels (executed by warps 2..5, threads 64..191).
template <bool FUSED, bool STACK>
__device__ __forceinline__ void conv_epilogue(const ConvKernelParams& p, uint32_t tmem_base, float* ep, uint64_t* acc_bar,
                                              int warp, int lane, int n0, int oy0, int ox0, int co0,
                                              uint8_t* stg = nullptr, bool stage_consts = true, uint32_t parity = 0,
                                              uint32_t peer_s = 0, int n_ranks = 0, int my_rank = 0, bool flush_stats = true) {
    // peer_s / n_ranks / my_rank: cluster split-K - every CTA of the cluster parks its partial accumulator at shared offset
    // peer_s in [column/4][row] float4 order; each rank then finishes BN / n_ranks of the columns (a reduce-scatter over
    // distributed shared memory: one rank summing everything serialised 0.9 MB of remote reads per tile and was slower
    // than not splitting at all)
    // stg: >= 8 KB of shared memory that is free while the epilogue runs (4 warps x 2 KB staging for coalesced stores),
    // or nullptr for direct per-thread stores
    // epilogue: warp (2..5) may only touch TMEM lanes 32*(warp%4) .. +31
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int xl = row % p.bw, yl = (row / p.bw) % p.bh, nl = row / (p.bw * p.bh);
    const int n = n0 + nl, oy = oy0 + yl, ox = ox0 + xl;
    const bool valid = (n < p.out_n) && (oy < p.grid_h) && (ox < p.grid_w);
    float* dst = p.out + (long long)n * p.out_sn + (long long)(oy * p.out_ystep + p.out_y0) * p.out_sy +
                 (long long)(ox * p.out_xstep + p.out_x0) * p.out_sx;
    // lanes of a multi-image tile beyond the batch stay in the warp-collective store path: their per-image reads use a
    // clamped image index (the values are never stored)
    const int n_rd = n < p.out_n ? n : p.out_n - 1;
    const float* alpha = p.alpha ? p.alpha + (size_t)n_rd * p.cout : nullptr;
    const float nz = (p.noise && valid)
        ? p.noise_w * __ldg(p.noise + (size_t)(oy * p.out_ystep + p.out_y0) * p.noise_ld + (ox * p.out_xstep + p.out_x0))
        : 0.f;
    // One image per tile (the common case): stage alpha / beta / next-layer style / ToRGB weights of this
    // CTA's channel slice in shared memory once instead of re-loading them per row from global memory.
    const bool cs = (p.bn == 1);
    const int BN = p.BN;
    if (cs && stage_consts) {
        const bool n_ok = n0 < p.out_n;
        for (int i = threadIdx.x - 64; i < BN; i += 128) {
            const int cc = co0 + i;
            const bool ok = n_ok && cc < p.cout;
            ep[i] = (ok && p.alpha) ? __ldg(p.alpha + (size_t)n0 * p.cout + cc) : 1.f;
            ep[BN + i] = (ok && p.beta) ? __ldg(p.beta + cc) : 0.f;
            if constexpr (!FUSED) {
                if (p.stat_sum) {        // per-CTA accumulators of the BatchNorm statistics and the staged shift
                    ep[2 * BN + i] = 0.f;
                    ep[3 * BN + i] = 0.f;
                    ep[4 * BN + i] = (cc < p.cout && p.stat_shift) ? __ldg(p.stat_shift + cc) : 0.f;
                }
            }
            if constexpr (FUSED) {
                ep[2 * BN + i] = (ok && p.split_scale) ? __ldg(p.split_scale + (size_t)n0 * p.split_scale_ld + cc) : 1.f;
#pragma unroll
                for (int o = 0; o < 3; ++o)
                    ep[(3 + o) * BN + i] = (ok && p.rgb_w) ? __ldg(p.rgb_w + ((size_t)n0 * 3 + o) * p.cout + cc) : 0.f;
            }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    ptx::mbar_wait(acc_bar, parity);
    ptx::tc_fence_after();
    const bool write_f32 = FUSED ? (p.out != nullptr && n >= p.out_from_n) : true;
    const int gsz = FUSED ? 0 : p.group_size;                // phase-packed output: plain epilogue only
    const bool g_uniform = (gsz == 0) || (gsz % 16 == 0);    // a 16-channel block never straddles two groups
    const int py0 = oy * p.out_ystep + p.out_y0, px0 = ox * p.out_xstep + p.out_x0;
    const size_t pix = ((size_t)n * p.grid_h + oy) * p.grid_w + ox;
    const uint32_t ep_s = ptx::smem_u32(ep);                 // explicit ld.shared: the generic pointer costs LD.E + a stall per use
    const uint32_t stg_s = stg ? ptx::smem_u32(stg) + (uint32_t)q * 2048u : 0u;
    float rgb0 = 0.f, rgb1 = 0.f, rgb2 = 0.f;
    // split32 output addressed by grid pixel (identity mapping) or, for the output-phase launches of an up-sampling conv,
    // by OUTPUT pixel
    const size_t pix_split = (FUSED && p.split_hw) ? ((size_t)n * p.out_h + py0) * p.out_w + px0 : pix;
    // ProgGAN's pixel norm of the activation (models/ProgGAN/model.py:17-18), fused: the tile holds every channel of its
    // pixels (one channel tile, host-checked), a thread owns a pixel, so mean_c act^2 is a private sum over a first pass
    // through the accumulator columns (TMEM is re-read below: cheaper than 256 live registers)
    float pn = 1.f;
    if constexpr (FUSED) {
        if (p.pixnorm_eps > 0.f) {
            float ss = 0.f;
            for (int c = 0; c < BN; c += 16) {
                float t[16];
                if constexpr (STACK)
                    ptx::tmem_ld16_sum(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c,
                                       tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c), t);
                else
                    ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, t);
                const uint32_t ea = ep_s + (uint32_t)c * 4u, eb = ea + (uint32_t)BN * 4u;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 al = ptx::lds128(ea + i * 4), be = ptx::lds128(eb + i * 4);
                    const float a0 = apply_act(fmaf(t[i + 0], al.x, nz) + be.x, p.act), a1 = apply_act(fmaf(t[i + 1], al.y, nz) + be.y, p.act),
                                a2 = apply_act(fmaf(t[i + 2], al.z, nz) + be.z, p.act), a3 = apply_act(fmaf(t[i + 3], al.w, nz) + be.w, p.act);
                    if (co0 + c + i + 0 < p.cout) ss += a0 * a0;
                    if (co0 + c + i + 1 < p.cout) ss += a1 * a1;
                    if (co0 + c + i + 2 < p.cout) ss += a2 * a2;
                    if (co0 + c + i + 3 < p.cout) ss += a3 * a3;
                }
            }
            pn = rsqrtf(ss / (float)p.cout + p.pixnorm_eps);
        }
    }
    const int c_lo = n_ranks ? my_rank * (BN / n_ranks) : 0, c_hi = n_ranks ? c_lo + BN / n_ranks : BN;
    for (int c = c_lo; c < c_hi; c += 16) {
        float v[16];
        if constexpr (STACK)      // columns [0, BN) hold hi*hi + lo*hi, columns [BN, 2BN) hold hi*lo
            ptx::tmem_ld16_sum(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c,
                               tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c), v);
        else
            ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
        if (n_ranks) {
            const uint32_t mine = peer_s + (uint32_t)(((c >> 2) * 128 + row) * 16);
            for (int pr = 0; pr < n_ranks; ++pr) {
                if (pr == my_rank) continue;
                const uint32_t ra = ptx::map_to_cta(mine, (uint32_t)pr);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 t4 = ptx::ld_dsmem128(ra + (uint32_t)k * 2048u);
                    v[4 * k] += t4.x; v[4 * k + 1] += t4.y; v[4 * k + 2] += t4.z; v[4 * k + 3] += t4.w;
                }
            }
        }
        const int co = co0 + c;
        if (co >= p.cout) continue;                          // warp-uniform
        // destination of this 16-channel block: `dptr + cof` (identity unless the output is phase-packed)
        bool lane_ok = valid;
        float* dptr = dst;
        int cof = co;
        if (gsz && g_uniform) {
            const int g = co / gsz, gy = g / p.group_w, gx = g - gy * p.group_w;
            if (py0 + gy >= p.out_h || px0 + gx >= p.out_w) lane_ok = false;
            dptr = dst + (long long)gy * p.out_sy + (long long)gx * p.out_sx;
            cof = co - g * gsz;
        }
        if (!stg_s && !lane_ok) continue;                    // direct stores: nothing below is warp-collective
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(dptr + cof) & 15) == 0);
        const bool fast = cs && g_uniform && (co + 16 <= p.cout);   // whole 16-channel block valid, constants staged in smem
        if (fast) {
            const uint32_t ea = ep_s + (uint32_t)c * 4u, eb = ea + (uint32_t)BN * 4u;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const float4 al = ptx::lds128(ea + i * 4), be = ptx::lds128(eb + i * 4);
                v[i + 0] = fmaf(v[i + 0], al.x, nz) + be.x;
                v[i + 1] = fmaf(v[i + 1], al.y, nz) + be.y;
                v[i + 2] = fmaf(v[i + 2], al.z, nz) + be.z;
                v[i + 3] = fmaf(v[i + 3], al.w, nz) + be.w;
            }
            if (p.accumulate && lane_ok) {
                if (vec_ok) {
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 o = *reinterpret_cast<const float4*>(dptr + cof + i);
                        v[i] += o.x; v[i + 1] += o.y; v[i + 2] += o.z; v[i + 3] += o.w;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += dptr[cof + i];
                }
            }
            if (p.act == 1) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            } else if (p.act == 2) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
            } else if (p.act == 3) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = 1.41421356237309515f * (v[i] > 0.f ? v[i] : 0.2f * v[i]);
            } else if (p.act == 4) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = tanhf(v[i]);
            }
            if constexpr (!FUSED) {
                if (p.stat_sum) {
                    // shifted first / second moments of this 32-pixel x 16-channel block -> the CTA's shared accumulators
                    float dv[16], dq[16];
                    const uint32_t es = ep_s + (uint32_t)(4 * BN + c) * 4u;
#pragma unroll
                    for (int i = 0; i < 16; i += 4) {
                        const float4 sh = ptx::lds128(es + i * 4);
                        dv[i + 0] = lane_ok ? v[i + 0] - sh.x : 0.f; dv[i + 1] = lane_ok ? v[i + 1] - sh.y : 0.f;
                        dv[i + 2] = lane_ok ? v[i + 2] - sh.z : 0.f; dv[i + 3] = lane_ok ? v[i + 3] - sh.w : 0.f;
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) dq[i] = dv[i] * dv[i];
                    const float s1 = warp_colsum16(dv, lane), s2 = warp_colsum16(dq, lane);
                    if ((lane & 1) == 0) {
                        const int ch = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
                        atomicAdd(ep + 2 * BN + c + ch, s1);
                        atomicAdd(ep + 3 * BN + c + ch, s2);
                    }
                }
            }
        } else if (!g_uniform && cs && !p.accumulate && p.group_w == 2 && p.out_sx == gsz && (gsz == 3 || gsz == 6) &&
                   px0 + 2 <= p.out_w) {
            if (gsz == 3) small_group_store<3>(v, c, co, p.cout, dst, p.out_sy, lane_ok, py0, p.out_h, ep_s, BN, nz, p.act);
            else small_group_store<6>(v, c, co, p.cout, dst, p.out_sy, lane_ok, py0, p.out_h, ep_s, BN, nz, p.act);
            continue;
        } else if (!g_uniform) {
            // phase-packed output whose groups are narrower than a 16-channel block (e.g. the 6-channel stem data
            // gradient): element-wise destination, written here
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int cc = co + i;
                if (cc >= p.cout || !lane_ok) break;
                const int g = cc / gsz, gy = g / p.group_w, gx = g - gy * p.group_w;
                if (py0 + gy >= p.out_h || px0 + gx >= p.out_w) continue;
                float r = v[i];
                if (cs) r = r * ep[c + i] + nz + ep[BN + c + i];
                else {
                    if (alpha) r *= __ldg(alpha + cc);
                    r += nz;
                    if (p.beta) r += __ldg(p.beta + cc);
                }
                float* d1 = dst + (long long)gy * p.out_sy + (long long)gx * p.out_sx + (cc - g * gsz);
                if (p.accumulate) r += *d1;
                *d1 = apply_act(r, p.act);
            }
            continue;
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int cc = co + i;
                if (cc < p.cout) {
                    float r = v[i];
                    if (cs) r = r * ep[c + i] + nz + ep[BN + c + i];
                    else {
                        if (alpha) r *= __ldg(alpha + cc);
                        r += nz;
                        if (p.beta) r += __ldg(p.beta + cc);
                    }
                    if (p.accumulate && lane_ok) r += dptr[cof + i];
                    v[i] = apply_act(r, p.act);
                } else {
                    v[i] = 0.f;
                }
            }
        }
        if (FUSED && p.rgb_w) {
            if (fast) {
                const uint32_t er = ep_s + (uint32_t)(3 * BN + c) * 4u;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float4 w0 = ptx::lds128(er + i * 4), w1 = ptx::lds128(er + (BN + i) * 4),
                                 w2 = ptx::lds128(er + (2 * BN + i) * 4);
                    rgb0 += v[i] * w0.x + v[i + 1] * w0.y + v[i + 2] * w0.z + v[i + 3] * w0.w;
                    rgb1 += v[i] * w1.x + v[i + 1] * w1.y + v[i + 2] * w1.z + v[i + 3] * w1.w;
                    rgb2 += v[i] * w2.x + v[i + 1] * w2.y + v[i + 2] * w2.z + v[i + 3] * w2.w;
                }
            } else if (cs) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    rgb0 += v[i] * ep[3 * BN + c + i];
                    rgb1 += v[i] * ep[4 * BN + c + i];
                    rgb2 += v[i] * ep[5 * BN + c + i];
                }
            } else {
                const float* wm = p.rgb_w + (size_t)n_rd * 3 * p.cout + co;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    if (co + i < p.cout) {
                        rgb0 += v[i] * __ldg(wm + i);
                        rgb1 += v[i] * __ldg(wm + p.cout + i);
                        rgb2 += v[i] * __ldg(wm + 2 * p.cout + i);
                    }
                }
            }
        }
        if (FUSED && p.out_split) {
            __align__(16) __nv_bfloat16 hi[16], lo[16];
            if (fast) {
                const uint32_t es = ep_s + (uint32_t)(2 * BN + c) * 4u;
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    float4 sc4 = ptx::lds128(es + i * 4);
                    sc4.x *= pn; sc4.y *= pn; sc4.z *= pn; sc4.w *= pn;
                    split_bf16x2(v[i + 0] * sc4.x, v[i + 1] * sc4.y, reinterpret_cast<__nv_bfloat162*>(hi)[i >> 1],
                                 reinterpret_cast<__nv_bfloat162*>(lo)[i >> 1]);
                    split_bf16x2(v[i + 2] * sc4.z, v[i + 3] * sc4.w, reinterpret_cast<__nv_bfloat162*>(hi)[(i >> 1) + 1],
                                 reinterpret_cast<__nv_bfloat162*>(lo)[(i >> 1) + 1]);
                }
            } else {
                const float* sc = (!cs && p.split_scale) ? p.split_scale + (size_t)n_rd * p.split_scale_ld + co : nullptr;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    split_bf16(pn * (cs ? v[i] * ep[2 * BN + c + i] : (sc ? v[i] * __ldg(sc + i) : v[i])), hi[i], lo[i]);
            }
            __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(p.out_split) + pix_split * (size_t)(((p.cout + 31) >> 5) * 64) +
                                (size_t)(co >> 5) * 64 + (co & 16);
            // cout = 16 (mod 32): the last chunk is half full - its upper 16 channels are written as zeros (warp-uniform)
            const bool pad_half = ((p.cout & 31) == 16) && (co + 16 == p.cout);
            if (stg_s) {
                const uint4 q4[4] = {reinterpret_cast<const uint4*>(hi)[0], reinterpret_cast<const uint4*>(hi)[1],
                                     reinterpret_cast<const uint4*>(lo)[0], reinterpret_cast<const uint4*>(lo)[1]};
                warp_store_rows64(stg_s, lane, q4, lane_ok ? sp : nullptr, 64u);
                if (pad_half) {
                    const uint4 z4[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
                    warp_store_rows64(stg_s, lane, z4, lane_ok ? sp + 16 : nullptr, 64u);
                }
            } else if (lane_ok) {
                reinterpret_cast<uint4*>(sp)[0] = reinterpret_cast<const uint4*>(hi)[0];
                reinterpret_cast<uint4*>(sp)[1] = reinterpret_cast<const uint4*>(hi)[1];
                reinterpret_cast<uint4*>(sp + 32)[0] = reinterpret_cast<const uint4*>(lo)[0];
                reinterpret_cast<uint4*>(sp + 32)[1] = reinterpret_cast<const uint4*>(lo)[1];
                if (pad_half) {
                    reinterpret_cast<uint4*>(sp + 16)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(sp + 16)[1] = make_uint4(0, 0, 0, 0);
                    reinterpret_cast<uint4*>(sp + 48)[0] = make_uint4(0, 0, 0, 0); reinterpret_cast<uint4*>(sp + 48)[1] = make_uint4(0, 0, 0, 0);
                }
            }
        }
        // (write_f32 is per image, hence per lane when a tile spans images; the collective path needs every lane)
        const bool st_ok = write_f32 && lane_ok;
        if (stg_s && co + 16 <= p.cout && __all_sync(0xffffffffu, vec_ok || !st_ok)) {
            if (__any_sync(0xffffffffu, st_ok)) {
                uint4 q4[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    q4[k] = make_uint4(__float_as_uint(v[4 * k]), __float_as_uint(v[4 * k + 1]), __float_as_uint(v[4 * k + 2]),
                                       __float_as_uint(v[4 * k + 3]));
                warp_store_rows64(stg_s, lane, q4, st_ok ? dptr + cof : nullptr, 32u);
            }
        } else if (st_ok) {
            if (vec_ok && co + 16 <= p.cout) {
#pragma unroll
                for (int i = 0; i < 16; i += 4)
                    *reinterpret_cast<float4*>(dptr + cof + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
                for (int i = 0; i < 16 && co + i < p.cout; ++i) dptr[cof + i] = v[i];
            }
        }
    }
    if (FUSED && p.rgb_out && valid) {
        float* ro = p.rgb_out + pix * 3;
        atomicAdd(ro, rgb0); atomicAdd(ro + 1, rgb1); atomicAdd(ro + 2, rgb2);
    }
    if constexpr (!FUSED) {
        if (p.stat_sum && flush_stats) {         // one global atomic per channel per CTA (after its last tile)
            asm volatile("bar.sync 1, 128;" ::: "memory");
            for (int i = threadIdx.x - 64 + c_lo; i < c_hi; i += 128) {
                const int cc = co0 + i;
                if (cc < p.cout) {
                    atomicAdd(p.stat_sum + cc, ep[2 * BN + i]);
                    atomicAdd(p.stat_sumsq + cc, ep[3 * BN + i]);
                }
            }
        }
    }
}

// FUSED = false: plain epilogue (demod / noise / bias / activation / accumulate -> fp32), lean register budget.
// FUSED = true: additionally emits the next layer's split32 operand, ToRGB partial sums, selective fp32 stores.
// STACK = true (stacked weight layout, BN <= 64): per K slice one N = 2*BN MMA a_hi x [b_hi; b_lo] and one N = BN MMA
// a_lo x b_hi instead of three N = BN MMAs - a narrow MMA costs ~(32 + N/4) cycles of operand fetch, not N/2 of math.
template <bool FUSED, bool STACK>
__global__ void __launch_bounds__(CONV_THREADS, 4)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ ConvKernelParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_stage_bytes = p.BN * 128;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * A_STAGE_BYTES;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* acc_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    float* ep = reinterpret_cast<float*>(tmem_slot + 4);
    ep += ((16u - (ptx::smem_u32(ep) & 15u)) & 15u) >> 2;   // [6][BN] per-channel epilogue constants

    // tile coordinates: co tile fastest so CTAs sharing an input patch are co-scheduled (L2 reuse).
    // Cluster split-K (ksplit > 1, launched as clusters of ksplit CTAs along x): the ksplit CTAs of a cluster share one
    // output tile and each contracts a contiguous range of the (tap, chunk) blocks; ranks 1.. park their accumulators in
    // their own shared memory and rank 0 adds them through distributed shared memory inside its normal epilogue.  Used
    // where M is tiny (512-channel layers at 4^2 .. 32^2: 8 CTAs x 144 blocks ran 63 us on 5 % of the SMs).
    int t = blockIdx.x;
    const int ks = t % p.ksplit; t /= p.ksplit;
    const int co_tile = t % p.n_tiles_co; t /= p.n_tiles_co;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int tn = t;
    const int ox0 = tx * p.bw, oy0 = ty * p.bh, n0 = tn * p.bn, co0 = co_tile * p.BN;
    const int k_total = p.num_taps * p.c_chunks;
    const int k_per = (k_total + p.ksplit - 1) / p.ksplit;
    const int kb_lo = ks * k_per, kb_hi = min(k_total, kb_lo + k_per);
    const int k_blocks = kb_hi - kb_lo;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < p.stages; ++s) {
            ptx::mbar_init(full_bar + s, 1);
            ptx::mbar_init(empty_bar + s, 1);
        }
        ptx::mbar_init(acc_bar, 1);
        ptx::fence_mbar_init();
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int tap = kb_lo / p.c_chunks, ch = kb_lo % p.c_chunks;
            for (int kb = kb_lo; kb < kb_hi; ++kb) {
                const int ix = ox0 * p.in_stride + p.tap_dx[tap];
                const int iy = oy0 * p.in_stride + p.tap_dy[tap];
                const int wt = p.tap_w[tap];
                ptx::mbar_wait(empty_bar + stage, phase ^ 1);
                ptx::mbar_expect_tx(full_bar + stage, (uint32_t)(A_STAGE_BYTES + b_stage_bytes));
                ptx::tma_load_5d(smem_a + (size_t)stage * A_STAGE_BYTES, &tmap_a, full_bar + stage, 0, ch, ix, iy, n0);
                if constexpr (STACK)
                    ptx::tma_load_5d(smem_b + (size_t)stage * b_stage_bytes, &tmap_b, full_bar + stage, 0, co0, 0, ch, wt);
                else
                    ptx::tma_load_4d(smem_b + (size_t)stage * b_stage_bytes, &tmap_b, full_bar + stage, 0, ch, co0, wt);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                if (++ch == p.c_chunks) { ch = 0; ++tap; }
            }
        }
    } else if (warp == 1) {
        // whole warp with uniform control flow (descriptor arithmetic stays in uniform registers); one elected lane
        // issues the MMAs and commits
        const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)p.BN);
        const uint32_t idesc2 = ptx::umma_idesc_bf16(128, (uint32_t)(2 * p.BN));
        const uint32_t dhi = (uint32_t)(ptx::umma_desc_sw128(0) >> 32);
        const uint32_t bhi64 = (uint32_t)(ptx::umma_desc_sw64(0) >> 32);
        const uint32_t a_lo0 = (ptx::smem_u32(smem_a) & 0x3FFFFu) >> 4, b_lo0 = (ptx::smem_u32(smem_b) & 0x3FFFFu) >> 4;
        const uint32_t a_step = A_STAGE_BYTES >> 4, b_step = (uint32_t)b_stage_bytes >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < k_blocks; ++kb) {
            ptx::mbar_wait(full_bar + stage, phase);
            ptx::tc_fence_after();
            const uint32_t da = a_lo0 + (uint32_t)stage * a_step, db = b_lo0 + (uint32_t)stage * b_step;
            if constexpr (STACK) {
                if (ptx::elect_one()) {
                    // B: 64-byte rows [k0 | k1], BN hi rows then BN lo rows
                    ptx::mma_f16_lh(tmem_base, da + 0, dhi, db + 0, bhi64, idesc2, kb > 0 ? 1u : 0u);   // a_hi x [b_hi; b_lo]
                    ptx::mma_f16_lh(tmem_base, da + 2, dhi, db + 2, bhi64, idesc2, 1u);
                    ptx::mma_f16_lh(tmem_base, da + 4, dhi, db + 0, bhi64, idesc, 1u);                   // a_lo x b_hi
                    ptx::mma_f16_lh(tmem_base, da + 6, dhi, db + 2, bhi64, idesc, 1u);
                    ptx::mma_commit(empty_bar + stage);
                }
            } else if (ptx::elect_one()) {
                // 128-byte row = [hi k0 | hi k1 | lo k0 | lo k1], 32 B each -> descriptor address +2 per slot
                ptx::mma_f16_lh(tmem_base, da + 0, dhi, db + 0, dhi, idesc, kb > 0 ? 1u : 0u);   // hi*hi
                ptx::mma_f16_lh(tmem_base, da + 2, dhi, db + 2, dhi, idesc, 1u);
                ptx::mma_f16_lh(tmem_base, da + 0, dhi, db + 4, dhi, idesc, 1u);                 // hi*lo
                ptx::mma_f16_lh(tmem_base, da + 2, dhi, db + 6, dhi, idesc, 1u);
                ptx::mma_f16_lh(tmem_base, da + 4, dhi, db + 0, dhi, idesc, 1u);                 // lo*hi
                ptx::mma_f16_lh(tmem_base, da + 6, dhi, db + 2, dhi, idesc, 1u);
                ptx::mma_commit(empty_bar + stage);          // frees the smem slot when these MMAs retire
            }
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::mma_commit(acc_bar);       // accumulator complete
        __syncwarp();
    } else {
        // every MMA has retired when the accumulator barrier fires: the operand ring is free and stages the stores
        if (p.ksplit == 1) {
            conv_epilogue<FUSED, STACK>(p, tmem_base, ep, acc_bar, warp, lane, n0, oy0, ox0, co0, p.coalesce ? smem_a : nullptr);
        } else {
            // park this CTA's partial accumulator in its (now idle) operand ring: [column / 4][row] float4
            // (its own slice of the columns stays in TMEM)
            const int q = warp & 3, row = q * 32 + lane;
            ptx::mbar_wait(acc_bar, 0);
            ptx::tc_fence_after();
            const uint32_t base_s = ptx::smem_u32(smem_a) + (uint32_t)row * 16u;
            const int own_lo = ks * (p.BN / p.ksplit), own_hi = own_lo + p.BN / p.ksplit;
            for (int c = 0; c < p.BN; c += 16) {
                if (c >= own_lo && c < own_hi) continue;
                float v[16];
                ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(base_s + (uint32_t)((c >> 2) + k) * 2048u),
                                 "f"(v[4 * k]), "f"(v[4 * k + 1]), "f"(v[4 * k + 2]), "f"(v[4 * k + 3]) : "memory");
            }
        }
    }
    if (p.ksplit > 1) {
        ptx::tc_fence_before();
        ptx::cluster_sync();                                  // every rank's partial accumulator is in place
        if (warp >= 2)
            conv_epilogue<FUSED, STACK>(p, tmem_base, ep, acc_bar, warp, lane, n0, oy0, ox0, co0, nullptr, true, 0,
                                        ptx::smem_u32(smem_a), p.ksplit, ks);
        ptx::cluster_sync();                                  // nobody leaves while a peer still reads its shared memory
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}


// ---------------------------------------------------------------------------------------------------
// Halo variant for stride-1 tap lists on <= 64-channel layers (L2-fabric bound in conv_tc_kernel, where every tap
// re-fetches its shifted 128-pixel patch).  Tile = 16 rows x 8 columns of output pixels.  Per 32-channel chunk ONE
// stage holds the (16 + wy - 1) x (8 + wx - 1) input patch (one TMA box) plus the weight slices of all taps; the A
// operand of tap (ty, tx) is the patch viewed through a descriptor whose start address is advanced by
// (ty * halo_w + tx) 128-byte rows and whose stride-byte-offset is the patch row pitch: the eight pixels of one output
// row are eight consecutive rows (one swizzle group), successive output rows are halo_w rows apart.  TMA and UMMA both
// apply the 128B swizzle as a function of the absolute shared-memory address, so shifted views stay consistent
// (descriptor base_offset = 0; measured in round 1, profiles/r01_conv_halo_ncu.md).  One tile per CTA, 3 CTAs per SM.
template <bool FUSED, bool STACK>
__global__ void __launch_bounds__(CONV_THREADS, 3)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ ConvKernelParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_slice = p.BN * 128;
    const int stage_bytes = p.a_stage_bytes + p.num_taps * b_slice;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* acc_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);
    uint32_t* tap_off = tmem_slot + 4;                                    // [WGS_MAX_TAPS] descriptor offsets (>>4)
    float* ep = reinterpret_cast<float*>(tap_off + WGS_MAX_TAPS);
    ep += ((16u - (ptx::smem_u32(ep) & 15u)) & 15u) >> 2;

    int t = blockIdx.x;
    const int co_tile = t % p.n_tiles_co; t /= p.n_tiles_co;
    const int tx = t % p.tiles_x; t /= p.tiles_x;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int n0 = t;
    const int ox0 = tx * p.bw, oy0 = ty * p.bh, co0 = co_tile * p.BN;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < p.stages; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
        ptx::mbar_init(acc_bar, 1);
        ptx::fence_mbar_init();
    }
    if (threadIdx.x >= 64 && threadIdx.x - 64 < p.num_taps) {
        const int i = threadIdx.x - 64;
        tap_off[i] = (uint32_t)(((p.tap_dy[i] - p.dy0) * p.halo_w + (p.tap_dx[i] - p.dx0)) * 128) >> 4;
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t bytes = (uint32_t)(p.halo_h * p.halo_w * 128 + p.num_taps * b_slice);
            for (int ch = 0; ch < p.c_chunks; ++ch) {
                ptx::mbar_wait(empty_bar + stage, phase ^ 1);
                ptx::mbar_expect_tx(full_bar + stage, bytes);
                uint8_t* sa = smem + (size_t)stage * stage_bytes;
                ptx::tma_load_5d(sa, &tmap_a, full_bar + stage, 0, ch, ox0 + p.dx0, oy0 + p.dy0, n0);
                for (int tap = 0; tap < p.num_taps; ++tap) {
                    if constexpr (STACK)
                        ptx::tma_load_5d(sa + p.a_stage_bytes + (size_t)tap * b_slice, &tmap_b, full_bar + stage, 0, co0, 0, ch,
                                         (int)p.tap_w[tap]);
                    else
                        ptx::tma_load_4d(sa + p.a_stage_bytes + (size_t)tap * b_slice, &tmap_b, full_bar + stage, 0, ch, co0,
                                         (int)p.tap_w[tap]);
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)p.BN);
        const uint32_t idesc2 = ptx::umma_idesc_bf16(128, (uint32_t)(2 * p.BN));
        const uint32_t b_hi = (uint32_t)((STACK ? ptx::umma_desc_sw64(0) : ptx::umma_desc_sw128(0)) >> 32);
        // A: K-major SW128, SBO = patch row pitch, version 1, base_offset 0
        const uint32_t a_hi = (uint32_t)((((uint64_t)((uint32_t)p.halo_w * 128u >> 4) << 32) | ((uint64_t)1 << 46) |
                                          ((uint64_t)2 << 61)) >> 32);
        const uint32_t lo0 = (ptx::smem_u32(smem) & 0x3FFFFu) >> 4;
        const uint32_t st_step = (uint32_t)stage_bytes >> 4, a_sz = (uint32_t)p.a_stage_bytes >> 4, b_step = (uint32_t)b_slice >> 4;
        int stage = 0;
        uint32_t phase = 0, accum = 0;
        for (int ch = 0; ch < p.c_chunks; ++ch) {
            ptx::mbar_wait(full_bar + stage, phase);
            ptx::tc_fence_after();
            const uint32_t a_lo = lo0 + (uint32_t)stage * st_step;
            const uint32_t b_lo0 = a_lo + a_sz;
            for (int tap = 0; tap < p.num_taps; ++tap) {
                const uint32_t da = a_lo + tap_off[tap];
                const uint32_t db = b_lo0 + (uint32_t)tap * b_step;
                if constexpr (STACK) {
                    if (ptx::elect_one()) {
                        ptx::mma_f16_lh(tmem_base, da + 0, a_hi, db + 0, b_hi, idesc2, accum);  // a_hi x [b_hi; b_lo]
                        ptx::mma_f16_lh(tmem_base, da + 2, a_hi, db + 2, b_hi, idesc2, 1u);
                        ptx::mma_f16_lh(tmem_base, da + 4, a_hi, db + 0, b_hi, idesc, 1u);      // a_lo x b_hi
                        ptx::mma_f16_lh(tmem_base, da + 6, a_hi, db + 2, b_hi, idesc, 1u);
                    }
                } else if (ptx::elect_one()) {
                    ptx::mma_f16_lh(tmem_base, da + 0, a_hi, db + 0, b_hi, idesc, accum);   // hi*hi
                    ptx::mma_f16_lh(tmem_base, da + 2, a_hi, db + 2, b_hi, idesc, 1u);
                    ptx::mma_f16_lh(tmem_base, da + 0, a_hi, db + 4, b_hi, idesc, 1u);      // hi*lo
                    ptx::mma_f16_lh(tmem_base, da + 2, a_hi, db + 6, b_hi, idesc, 1u);
                    ptx::mma_f16_lh(tmem_base, da + 4, a_hi, db + 0, b_hi, idesc, 1u);      // lo*hi
                    ptx::mma_f16_lh(tmem_base, da + 6, a_hi, db + 2, b_hi, idesc, 1u);
                }
                __syncwarp();
                accum = 1u;
            }
            if (ptx::elect_one()) ptx::mma_commit(empty_bar + stage);
            __syncwarp();
            if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (ptx::elect_one()) ptx::mma_commit(acc_bar);
        __syncwarp();
    } else {
        conv_epilogue<FUSED, STACK>(p, tmem_base, ep, acc_bar, warp, lane, n0, oy0, ox0, co0, p.coalesce ? smem : nullptr);
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ---------------------------------------------------------------------------------------------------
// Multi-tile halo kernel for single-chunk layers (<= 32 input channels: the 32 -> 32 convs at 1024^2 and their
// data-gradients).  The one-tile-per-CTA kernel above is bound by its own serial chain (set-up, TMA round trip, MMAs,
// epilogue: ~6.5 us per CTA at 3 CTAs / SM, tensor pipe 22 % active, profiles/r01_conv_ncu_step.md).  Here a CTA
// walks `tiles_per_cta` x-adjacent tiles of one image row band with
//   * the tap weights resident in shared memory (loaded once: they were 60 % of the per-tile L2->SM bytes),
//   * a ring of input patches filled ahead by the TMA warp,
//   * two TMEM accumulators, so the MMAs of tile i+1 overlap the epilogue of tile i.
// Stacked weight layout only (two MMAs per K slice).  2 CTAs / SM.
// Also serves TWO-chunk layers (33..64 input channels: the 64 -> 64 convs at 512^2 / 256^2 and their data-gradients): the
// taps of both chunks stay resident (147 KB for 9 taps x 64 channels), the patch ring then holds (tile, chunk) units and one
// CTA owns the SM - the per-tap kernel moved 432 KB of operands through L2 per 128-pixel tile, this one 46 KB.
constexpr int MT_STAGES = 2;                                       // patch ring depth, single-chunk layers (runtime: p.stages)
constexpr int MT_MAX_STAGES = 4;
constexpr int MT_STG_BYTES = 8192;                                 // 4 epilogue warps x 2 KB store staging

template <bool FUSED>
__global__ void __launch_bounds__(CONV_THREADS, 2)
conv_halo_mt_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ ConvKernelParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b_slice = p.BN * 128;
    const int n_ch = p.c_chunks, n_stages = p.stages;
    const int w_bytes = n_ch * p.num_taps * b_slice;                       // multiple of 1024 (BN >= 16 -> 2 KB slices)
    uint8_t* smem_w = smem;
    uint8_t* smem_p = smem + w_bytes;
    uint8_t* stg = smem_p + (size_t)n_stages * p.a_stage_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + MT_STG_BYTES);
    uint64_t* empty_bar = full_bar + n_stages;
    uint64_t* tfull_bar = empty_bar + n_stages;                            // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;                                  // [2] accumulator drained
    uint64_t* w_bar = tempty_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
    uint32_t* tap_off = tmem_slot + 4;
    float* ep = reinterpret_cast<float*>(tap_off + WGS_MAX_TAPS);
    ep += ((16u - (ptx::smem_u32(ep) & 15u)) & 15u) >> 2;

    int t = blockIdx.x;
    const int co_tile = t % p.n_tiles_co; t /= p.n_tiles_co;
    const int xg = t % p.x_groups; t /= p.x_groups;
    const int ty = t % p.tiles_y; t /= p.tiles_y;
    const int n0 = t;
    const int tx0 = xg * p.tiles_per_cta;
    const int n_tiles = min(p.tiles_per_cta, p.tiles_x - tx0);
    const int oy0 = ty * p.bh, co0 = co_tile * p.BN;
    const uint32_t acc_cols = (uint32_t)(2 * p.BN);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < n_stages; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(tfull_bar + a, 1); ptx::mbar_init(tempty_bar + a, 4); }
        ptx::mbar_init(w_bar, 1);
        ptx::fence_mbar_init();
    }
    if (threadIdx.x >= 64 && threadIdx.x - 64 < p.num_taps) {
        const int i = threadIdx.x - 64;
        tap_off[i] = (uint32_t)(((p.tap_dy[i] - p.dy0) * p.halo_w + (p.tap_dx[i] - p.dx0)) * 128) >> 4;
    }
    if (warp == 1) ptx::tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            ptx::mbar_expect_tx(w_bar, (uint32_t)w_bytes);
            for (int ch = 0; ch < n_ch; ++ch)
                for (int tap = 0; tap < p.num_taps; ++tap)
                    ptx::tma_load_5d(smem_w + (size_t)(ch * p.num_taps + tap) * b_slice, &tmap_b, w_bar, 0, co0, 0, ch,
                                     (int)p.tap_w[tap]);
            const uint32_t bytes = (uint32_t)(p.halo_h * p.halo_w * 128);
            int stage = 0;
            uint32_t phase = 0;
            for (int i = 0; i < n_tiles; ++i) {
                for (int ch = 0; ch < n_ch; ++ch) {
                    ptx::mbar_wait(empty_bar + stage, phase ^ 1);
                    ptx::mbar_expect_tx(full_bar + stage, bytes);
                    ptx::tma_load_5d(smem_p + (size_t)stage * p.a_stage_bytes, &tmap_a, full_bar + stage, 0, ch,
                                     (tx0 + i) * p.bw + p.dx0, oy0 + p.dy0, n0);
                    if (++stage == n_stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)p.BN);
        const uint32_t idesc2 = ptx::umma_idesc_bf16(128, (uint32_t)(2 * p.BN));
        const uint32_t b_hi = (uint32_t)(ptx::umma_desc_sw64(0) >> 32);
        const uint32_t a_hi = (uint32_t)((((uint64_t)((uint32_t)p.halo_w * 128u >> 4) << 32) | ((uint64_t)1 << 46) |
                                          ((uint64_t)2 << 61)) >> 32);
        const uint32_t w_lo0 = (ptx::smem_u32(smem_w) & 0x3FFFFu) >> 4;
        const uint32_t p_lo0 = (ptx::smem_u32(smem_p) & 0x3FFFFu) >> 4;
        const uint32_t st_step = (uint32_t)p.a_stage_bytes >> 4, b_step = (uint32_t)b_slice >> 4;
        ptx::mbar_wait(w_bar, 0);
        int stage = 0;
        uint32_t phase = 0;
        for (int i = 0; i < n_tiles; ++i) {
            const int acc = i & 1;
            const uint32_t use = (uint32_t)(i >> 1) & 1u;
            ptx::mbar_wait(tempty_bar + acc, use ^ 1);                     // the epilogue has drained this accumulator
            const uint32_t d_tmem = tmem_base + (uint32_t)acc * acc_cols;
            uint32_t accum = 0;
            for (int ch = 0; ch < n_ch; ++ch) {
                ptx::mbar_wait(full_bar + stage, phase);
                ptx::tc_fence_after();
                const uint32_t a_lo = p_lo0 + (uint32_t)stage * st_step;
                const uint32_t w_ch = w_lo0 + (uint32_t)(ch * p.num_taps) * b_step;
                for (int tap = 0; tap < p.num_taps; ++tap) {
                    const uint32_t da = a_lo + tap_off[tap];
                    const uint32_t db = w_ch + (uint32_t)tap * b_step;
                    if (ptx::elect_one()) {
                        ptx::mma_f16_lh(d_tmem, da + 0, a_hi, db + 0, b_hi, idesc2, accum);   // a_hi x [b_hi; b_lo]
                        ptx::mma_f16_lh(d_tmem, da + 2, a_hi, db + 2, b_hi, idesc2, 1u);
                        ptx::mma_f16_lh(d_tmem, da + 4, a_hi, db + 0, b_hi, idesc, 1u);       // a_lo x b_hi
                        ptx::mma_f16_lh(d_tmem, da + 6, a_hi, db + 2, b_hi, idesc, 1u);
                    }
                    __syncwarp();
                    accum = 1u;
                }
                if (ptx::elect_one()) {
                    ptx::mma_commit(empty_bar + stage);                    // patch slot free once these MMAs retire
                    if (ch == n_ch - 1) ptx::mma_commit(tfull_bar + acc);  // accumulator complete
                }
                __syncwarp();
                if (++stage == n_stages) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        for (int i = 0; i < n_tiles; ++i) {
            const int acc = i & 1;
            conv_epilogue<FUSED, true>(p, tmem_base + (uint32_t)acc * acc_cols, ep, tfull_bar + acc, warp, lane, n0, oy0,
                                       (tx0 + i) * p.bw, co0, p.coalesce ? stg : nullptr, i == 0, (uint32_t)(i >> 1) & 1u,
                                       0, 0, 0, i == n_tiles - 1);
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar + acc);
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// Same contract on CUDA cores (debug / cross-check path; selected with WGS_CONV_IMPL=simt).
__global__ void conv_simt_kernel(const __nv_bfloat16* __restrict__ in, const __nv_bfloat16* __restrict__ w,
                                 int in_n, int in_h, int in_w, int w_cout, int w_layout, const ConvKernelParams p) {
    const long long total = (long long)p.out_n * p.grid_h * p.grid_w * p.cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % p.cout);
        long long r = i / p.cout;
        const int ox = (int)(r % p.grid_w); r /= p.grid_w;
        const int oy = (int)(r % p.grid_h);
        const int n = (int)(r / p.grid_h);
        float acc = 0.f;
        for (int tp = 0; tp < p.num_taps; ++tp) {
            const int iy = oy * p.in_stride + p.tap_dy[tp], ix = ox * p.in_stride + p.tap_dx[tp];
            if (iy < 0 || iy >= in_h || ix < 0 || ix >= in_w) continue;
            const __nv_bfloat16* a = in + (((size_t)n * in_h + iy) * in_w + ix) * p.c_chunks * 64;
            const __nv_bfloat16* b = w + ((size_t)p.tap_w[tp] * w_cout + co) * p.c_chunks * 64;
            for (int ch = 0; ch < p.c_chunks; ++ch)
                for (int c = 0; c < 32; ++c) {
                    const float ah = __bfloat162float(a[ch * 64 + c]), al = __bfloat162float(a[ch * 64 + 32 + c]);
                    float bh, bl;
                    if (w_layout == 1) {        // stacked [tap][chunk][hi | lo][w_cout][32]
                        const __nv_bfloat16* bs = w + ((((size_t)p.tap_w[tp] * p.c_chunks + ch) * 2) * w_cout + co) * 32 + c;
                        bh = __bfloat162float(bs[0]); bl = __bfloat162float(bs[(size_t)w_cout * 32]);
                    } else {
                        bh = __bfloat162float(b[ch * 64 + c]); bl = __bfloat162float(b[ch * 64 + 32 + c]);
                    }
                    acc += ah * bh + ah * bl + al * bh;
                }
        }
        int gy = 0, gx = 0, cw = co;
        if (p.group_size) {
            const int g = co / p.group_size;
            gy = g / p.group_w; gx = g - gy * p.group_w; cw = co - g * p.group_size;
            if (oy * p.out_ystep + p.out_y0 + gy >= p.out_h || ox * p.out_xstep + p.out_x0 + gx >= p.out_w) continue;
        }
        float* dst = p.out + (long long)n * p.out_sn + (long long)(oy * p.out_ystep + p.out_y0 + gy) * p.out_sy +
                     (long long)(ox * p.out_xstep + p.out_x0 + gx) * p.out_sx + cw;
        if (p.alpha) acc *= p.alpha[(size_t)n * p.cout + co];
        if (p.noise)
            acc += p.noise_w * p.noise[(size_t)(oy * p.out_ystep + p.out_y0) * p.noise_ld + (ox * p.out_xstep + p.out_x0)];
        if (p.beta) acc += p.beta[co];
        if (p.accumulate) acc += *dst;
        acc = apply_act(acc, p.act);
        if (p.out && n >= p.out_from_n) *dst = acc;
        const size_t pix = ((size_t)n * p.grid_h + oy) * p.grid_w + ox;
        if (p.rgb_w)
            for (int o = 0; o < 3; ++o) atomicAdd(p.rgb_out + pix * 3 + o, acc * p.rgb_w[((size_t)n * 3 + o) * p.cout + co]);
        if (p.out_split) {
            __nv_bfloat16 hi, lo;
            split_bf16(p.split_scale ? acc * p.split_scale[(size_t)n * p.split_scale_ld + co] : acc, hi, lo);
            __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(p.out_split) + pix * (size_t)(p.cout * 2) + (size_t)(co >> 5) * 64 + (co & 31);
            sp[0] = hi; sp[32] = lo;
        }
    }
}

// ---- host side ------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

static int next_pow2(int v) { int r = 1; while (r < v) r <<= 1; return r; }
static int k_blocks_total(const wgs_conv_desc* d) { return d->num_taps * d->c_chunks; }

static int conv_impl_is_simt() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("WGS_CONV_IMPL");
        v = (e && std::string(e) == "simt") ? 1 : 0;
    }
    return v;
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_conv_split32(const wgs_conv_desc* d, void* stream) {
    WGS_REQUIRE(d != nullptr, "conv: null descriptor");
    WGS_REQUIRE(d->num_taps >= 1 && d->num_taps <= WGS_MAX_TAPS, "conv: bad tap count");
    WGS_REQUIRE(d->c_chunks >= 1 && d->cout >= 1, "conv: bad channel counts");
    WGS_REQUIRE(d->in_stride >= 1 && d->in_stride <= 8, "conv: bad input stride");
    WGS_REQUIRE(d->out_n >= 1 && d->grid_h >= 1 && d->grid_w >= 1, "conv: empty output grid");
    WGS_REQUIRE((reinterpret_cast<uintptr_t>(d->in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0,
                "conv: operands must be 16-byte aligned");
    ConvKernelParams p;
    memset(&p, 0, sizeof(p));
    p.out_n = d->out_n; p.grid_h = d->grid_h; p.grid_w = d->grid_w;
    p.bw = std::min(16, next_pow2(d->grid_w));
    p.bh = std::min(128 / p.bw, next_pow2(d->grid_h));
    p.bn = 128 / (p.bw * p.bh);
    p.tiles_x = ceil_div(d->grid_w, p.bw);
    p.tiles_y = ceil_div(d->grid_h, p.bh);
    p.tiles_n = ceil_div(d->out_n, p.bn);
    p.in_stride = d->in_stride;
    p.c_chunks = d->c_chunks; p.cout = d->cout; p.num_taps = d->num_taps;
    p.out = d->out; p.out_sn = d->out_sn; p.out_sy = d->out_sy; p.out_sx = d->out_sx;
    p.out_y0 = d->out_y0; p.out_x0 = d->out_x0; p.out_ystep = d->out_ystep; p.out_xstep = d->out_xstep;
    p.alpha = d->alpha; p.beta = d->beta; p.act = d->act; p.accumulate = d->accumulate;
    p.noise = d->noise; p.noise_w = d->noise_w; p.noise_ld = d->noise_ld;
    p.out_split = d->out_split; p.split_scale = d->split_scale; p.split_scale_ld = d->split_scale_ld;
    p.out_from_n = d->out_from_n; p.rgb_w = d->rgb_w; p.rgb_out = d->rgb_out;
    p.group_size = d->group_size; p.group_w = d->group_w; p.out_h = d->out_h; p.out_w = d->out_w;
    p.stat_sum = d->stat_sum; p.stat_sumsq = d->stat_sumsq; p.stat_shift = d->stat_shift;
    p.pixnorm_eps = d->pixnorm_eps;
    WGS_REQUIRE(d->pixnorm_eps >= 0.f, "conv: bad pixnorm_eps");
    if (d->pixnorm_eps > 0.f) {
        WGS_REQUIRE(d->out_split != nullptr && d->cout <= 256,
                    "conv: the fused pixel norm needs a split32 output and at most 256 output channels (one channel tile)");
        WGS_REQUIRE(p.bn == 1, "conv: the fused pixel norm needs output maps of at least 128 pixels (one image per tile)");
        WGS_REQUIRE(!conv_impl_is_simt(), "conv: the fused pixel norm is not available on the SIMT cross-check path");
    }
    WGS_REQUIRE((d->stat_sum == nullptr) == (d->stat_sumsq == nullptr), "conv: stat_sum and stat_sumsq go together");
    if (d->stat_sum) {
        WGS_REQUIRE(d->out != nullptr && !d->out_split && !d->rgb_out && d->out_from_n == 0 && d->group_size == 0 &&
                    !d->accumulate && d->cout % 16 == 0,
                    "conv: output statistics need the plain fp32 epilogue and cout % 16 == 0");
        WGS_REQUIRE(p.bn == 1, "conv: output statistics need output maps of at least 128 pixels (one image per tile)");
        WGS_REQUIRE(!conv_impl_is_simt(), "conv: output statistics are not available on the SIMT cross-check path");
    }
    WGS_REQUIRE(d->group_size >= 0, "conv: bad group_size");
    if (d->group_size > 0) {
        WGS_REQUIRE(d->cout % d->group_size == 0 && d->group_w >= 1 && (d->cout / d->group_size) % d->group_w == 0,
                    "conv: phase-packed output needs cout = groups * group_size with groups % group_w == 0");
        WGS_REQUIRE(d->out != nullptr && !d->out_split && !d->rgb_out && !d->noise && d->out_from_n == 0,
                    "conv: phase-packed output is fp32 only");
        WGS_REQUIRE(d->out_h > 0 && d->out_w > 0, "conv: phase-packed output needs out_h / out_w");
    }
    WGS_REQUIRE(d->out != nullptr || d->out_split != nullptr || d->rgb_out != nullptr, "conv: no output requested");
    const bool identity_map = d->out_ystep == 1 && d->out_xstep == 1 && d->out_y0 == 0 && d->out_x0 == 0;
    WGS_REQUIRE(!d->rgb_out || identity_map, "conv: the fused ToRGB output needs the identity output mapping");
    WGS_REQUIRE(!d->out_split || identity_map || (d->group_size == 0 && d->out_h > 0 && d->out_w > 0),
                "conv: a split32 output under a strided output mapping needs out_h / out_w (the dims of that tensor)");
    p.split_hw = (d->out_split && !identity_map) ? 1 : 0;
    WGS_REQUIRE(!d->out_split || d->cout % 16 == 0, "conv: split32 output needs cout % 16 == 0");
    WGS_REQUIRE(!d->accumulate || d->out != nullptr, "conv: accumulate needs an fp32 output");
    for (int i = 0; i < d->num_taps; ++i) {
        WGS_REQUIRE(d->tap_dy[i] >= -64 && d->tap_dy[i] <= 64 && d->tap_dx[i] >= -64 && d->tap_dx[i] <= 64,
                    "conv: tap offset out of range");
        WGS_REQUIRE(d->tap_w[i] >= 0 && d->tap_w[i] < d->w_taps && d->tap_w[i] < 256, "conv: bad weight tap index");
        p.tap_dy[i] = (signed char)d->tap_dy[i]; p.tap_dx[i] = (signed char)d->tap_dx[i];
        p.tap_w[i] = (unsigned char)d->tap_w[i];
    }
    // N tile: as wide as possible (fewer re-reads of the input patch) but keep >= ~1 wave of CTAs
    const int m_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    int BN = std::min(256, (d->cout + 15) / 16 * 16);
    static int bn_fill = -1;
    if (bn_fill < 0) {
        // percent of the SMs a launch must fill before its N tile stops halving.  80: a launch that reaches 118 .. 147 CTAs with the
        // widest tile keeps it (256 -> 256 @64^2 x 4 images: 128 CTAs of N = 256 instead of 256 CTAs of N = 128 - these
        // launches are L2 -> SM bound, and the wide tile re-loads the pixel operand half as often; same call 16.22 vs 16.37 ms)
        const char* e = getenv("WGS_BN_FILL");
        bn_fill = e ? atoi(e) : 80;
    }
    while (BN > 64 && m_tiles * ceil_div(d->cout, BN) * 100 < num_sms() * bn_fill && BN % 32 == 0) BN /= 2;
    if (d->pixnorm_eps > 0.f) BN = (d->cout + 15) / 16 * 16;      // every channel of a pixel in one tile
    if (d->force_bn > 0) BN = d->force_bn;
    WGS_REQUIRE(BN % 16 == 0 && BN >= 16 && BN <= 256, "conv: bad N tile");
    static int coalesce_mode = -1;
    if (coalesce_mode < 0) {
        const char* e = getenv("WGS_CONV_COALESCE");              // 0 = direct per-thread stores (A/B switch)
        coalesce_mode = (e && e[0] == '0') ? 0 : 1;
    }
    p.coalesce = coalesce_mode;
    static int splitk_mode = -1;
    if (splitk_mode < 0) {
        // -1 = per launch (wgs_conv_desc.split_k), 1 / 0 = forced on / off.  The split factor depends on the batch size, so
        // G(z) inside a batch of 8 and alone would differ by fp32 re-association (5.7e-6) instead of being bit-identical
        // (tests/test_step_gpu.py::test_full_size_properties): the generator never asks for it, the Reconstructor does
        const char* e = getenv("WGS_CONV_SPLITK");
        splitk_mode = e ? ((e[0] == '1') ? 1 : 0) : -1;
    }
    // (the fused pixel norm needs every channel of a pixel in ONE accumulator: never split)
    const bool splitk_on = (splitk_mode < 0 ? (d->split_k != 0) : (splitk_mode == 1)) && d->pixnorm_eps <= 0.f;
    const bool stack = d->w_layout == 1;
    // Cluster split-K for tiny-M, deep-K launches: widest N tile (math-bound MMAs, one patch load per tile), then as many
    // K splits as fill about two CTAs per SM (<= 8: portable cluster size)
    p.ksplit = 1;
    if (splitk_on && d->split_k == 2) {
        // geometry-only decision (frozen generators): the split depends on the feature-map size, channel and tap counts
        // alone, never on the batch, so an image computed alone or inside any batch goes through the same sums in the same
        // order.  Maps up to 16 x 16 (StyleGAN2 / ProgGAN / BigGAN layers at 4^2 .. 16^2: 8 - 64 CTAs of 144 dependent
        // stages each, ~60 us per launch whatever the batch)
        if (!stack && d->force_bn == 0 && (d->group_size == 0 || d->group_size % 16 == 0) && k_blocks_total(d) >= 16 &&
            d->grid_h * d->grid_w <= 256) {
            const int wide = std::min(256, (d->cout + 15) / 16 * 16);
            const int tiles1 = p.tiles_x * p.tiles_y * ceil_div(d->cout, wide);
            int ks = 1;
            while (ks < 8 && tiles1 * ks * 2 <= 64 && k_blocks_total(d) / (ks * 2) >= 4) ks *= 2;
            while (ks > 1 && (wide % (16 * ks)) != 0) ks /= 2;
            if (ks > 1) { BN = wide; p.ksplit = ks; }
        }
    } else if (splitk_on && !stack && d->force_bn == 0 && d->group_size == 0 && k_blocks_total(d) >= 16 &&
        (BN <= 64 || m_tiles * ceil_div(d->cout, BN) * 2 <= num_sms())) {   // narrow-N tiles or under half a wave
        const int wide = std::min(256, (d->cout + 15) / 16 * 16);
        const int tiles = m_tiles * ceil_div(d->cout, wide);
        int ks = 1;
        while (ks < 8 && tiles * ks * 2 <= 2 * num_sms() && k_blocks_total(d) / (ks * 2) >= 4) ks *= 2;
        while (ks > 1 && (wide % (16 * ks)) != 0) ks /= 2;     // every rank finishes a multiple of 16 columns
        if (ks > 1) { BN = wide; p.ksplit = ks; }
    }
    WGS_REQUIRE(d->w_layout == 0 || d->w_layout == 1, "conv: bad w_layout");
    WGS_REQUIRE(!stack || (d->w_cout <= 64 && BN <= 64), "conv: the stacked weight layout is for w_cout <= 64");
    p.BN = BN;
    p.w_cout = d->w_cout;
    p.n_tiles_co = ceil_div(d->cout, BN);
    p.tmem_cols = std::max(32, next_pow2(stack ? 2 * BN : BN));
    const int stage_bytes = A_STAGE_BYTES + BN * 128;
    // Short contractions (few taps x chunks) are latency-bound per tile: keep the ring shallow so that several
    // CTAs fit on one SM (smem and TMEM columns permitting) and overlap each other's prologue / epilogue.
    const int k_blocks = ceil_div(d->num_taps * d->c_chunks, p.ksplit);      // per CTA
    int ctas_per_sm = 1;
    if (k_blocks <= 12) ctas_per_sm = 4;
    else if (k_blocks <= 24) ctas_per_sm = 3;
    else if (k_blocks <= 48) ctas_per_sm = 2;
    ctas_per_sm = std::max(1, std::min(ctas_per_sm, 512 / p.tmem_cols));
    const int smem_budget = (220 * 1024) / ctas_per_sm - 2048 - 6 * BN * 4;
    p.stages = std::max(2, std::min(std::min(8, k_blocks), smem_budget / stage_bytes));
    if (p.ksplit > 1)                                            // the ring also parks a 128 x BN fp32 partial accumulator
        p.stages = std::max(p.stages, ceil_div(128 * BN * 4, stage_bytes));
    const cudaStream_t st = (cudaStream_t)stream;

    if (conv_impl_is_simt()) {
        const long long total = (long long)p.out_n * p.grid_h * p.grid_w * p.cout;
        const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
        conv_simt_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)d->in, (const __nv_bfloat16*)d->w, d->in_n,
                                                 d->in_h, d->in_w, d->w_cout, d->w_layout, p);
        count_launch();
        WGS_LAUNCH_CHECK();
        return 0;
    }

    auto encode = get_encode();
    WGS_REQUIRE(encode != nullptr, "conv: cuTensorMapEncodeTiled entry point not available");
    // weights: rows layout -> 4-D {64, chunks, cout, taps} SWIZZLE_128B box {64, 1, bn, 1};
    //          stacked     -> 5-D {32, cout, 2, chunks, taps} SWIZZLE_64B box {32, bn, 2, 1, 1} (bn hi rows then bn lo rows)
    auto encode_weights = [&](CUtensorMap* tm, int bn) -> CUresult {
        if (stack) {
            const cuuint64_t dims[5] = {32, (cuuint64_t)d->w_cout, 2, (cuuint64_t)d->c_chunks, (cuuint64_t)d->w_taps};
            const cuuint64_t s1 = 64, s2 = s1 * d->w_cout, s3 = s2 * 2, s4 = s3 * d->c_chunks;
            const cuuint64_t strides[4] = {s1, s2, s3, s4};
            const cuuint32_t box[5] = {32, (cuuint32_t)bn, 2, 1, 1};
            const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
            return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->w), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        }
        const cuuint64_t dims[4] = {64, (cuuint64_t)d->c_chunks, (cuuint64_t)d->w_cout, (cuuint64_t)d->w_taps};
        const cuuint64_t s1 = 128, s2 = s1 * d->c_chunks, s3 = s2 * d->w_cout;
        const cuuint64_t strides[3] = {s1, s2, s3};
        const cuuint32_t box[4] = {64, 1, (cuuint32_t)bn, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(d->w), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    };
    // ---- halo variant ---------------------------------------------------------------------------------
    {
        int dy0 = 127, dy1 = -127, dx0 = 127, dx1 = -127;
        for (int i = 0; i < d->num_taps; ++i) {
            dy0 = std::min(dy0, d->tap_dy[i]); dy1 = std::max(dy1, d->tap_dy[i]);
            dx0 = std::min(dx0, d->tap_dx[i]); dx1 = std::max(dx1, d->tap_dx[i]);
        }
        const int wy = dy1 - dy0 + 1, wx = dx1 - dx0 + 1;
        static int halo_mode = -1;
        if (halo_mode < 0) {
            const char* e = getenv("WGS_CONV_HALO");           // 0 = off, 1 = on (default)
            halo_mode = (e && e[0] == '0') ? 0 : 1;
        }
        int hBN = std::min(64, (d->cout + 15) / 16 * 16);
        const int halo_h = 16 + wy - 1, halo_w = 8 + wx - 1;
        const int a_bytes = (halo_h * halo_w * 128 + 1023) / 1024 * 1024;
        // 3 CTAs / SM up to 72 KB per stage; long tap lists (the 16 shifts of the merged 7x7/2 stem data-gradient) may
        // take up to 108 KB (2 CTAs / SM): still far less L2->SM traffic than re-loading the patch per tap
        static int halo_wide = -1;
        if (halo_wide < 0) {
            const char* e = getenv("WGS_HALO_WIDE");            // 1 = also 64 -> 64 two-chunk layers, 108 KB stages (experiment)
            halo_wide = (e && e[0] == '1') ? 1 : 0;
        }
        const int h_limit = ((d->num_taps >= 12 || halo_wide) ? 108 : 72) * 1024;
        // long tap lists on a single chunk (the 4x4-tap space-to-depth stem, 16 x 8 KB of weights): the multi-tile kernel
        // with ALL taps resident and the full N tile, one CTA per SM (the two TMEM accumulators still overlap MMAs and
        // epilogue) beats halving N, which re-loads every patch twice and the weights once per tile
        const bool big_mt = stack && d->c_chunks == 1 && d->force_bn == 0 && d->num_taps >= 12 &&
                            ceil_div(d->grid_w, 8) >= 32 &&
                            d->num_taps * hBN * 128 + MT_STAGES * a_bytes + MT_STG_BYTES <= 200 * 1024;
        // two-chunk layers (64 -> 64 3x3 and data-gradients; the 16-shift stem data-gradient): both chunks' taps resident,
        // ring of three (tile, chunk) patches, one CTA per SM
        static int mt_env = -1;
        if (mt_env < 0) {
            const char* e = getenv("WGS_HALO_MT");                // 0 = multi-tile kernels off (A/B switch)
            mt_env = (e && atoi(e) == 0) ? 0 : 1;
        }
        const bool mt2 = mt_env && halo_mode && stack && d->c_chunks == 2 && d->force_bn == 0 && d->in_stride == 1 &&
                         d->num_taps >= 4 && wy <= 7 && wx <= 7 && d->grid_h >= 16 && ceil_div(d->grid_w, 8) >= 16 &&
                         2 * d->num_taps * hBN * 128 + 3 * a_bytes + MT_STG_BYTES + 4096 <= 227 * 1024;
        while (!big_mt && !mt2 && hBN > 32 && hBN % 32 == 0 && a_bytes + d->num_taps * hBN * 128 > h_limit) hBN /= 2;
        const int h_stage = a_bytes + d->num_taps * hBN * 128;
        // (single-chunk 1 x 1 convs on wide maps too - ProgGAN's 16 -> 3 to-RGB layer at 1024^2 and its data gradient: on the
        // one-tile-per-CTA kernel that is 131072 CTAs and 1.14 ms for a bandwidth-bound 1 GB read; the multi-tile kernel
        // keeps its single tap resident and streams the pixel rows)
        const bool narrow_1x1 = stack && d->c_chunks == 1 && d->num_taps < 3 && ceil_div(d->grid_w, 8) >= 32 && d->grid_h >= 64;
        const bool eligible = halo_mode && p.ksplit == 1 && d->in_stride == 1 && (d->num_taps >= 3 || narrow_1x1) && wy <= 7 && wx <= 7 &&
                              (d->pixnorm_eps <= 0.f || hBN >= (d->cout + 15) / 16 * 16) &&
                              d->grid_h >= 16 && d->grid_w >= 8 && d->force_bn == 0 && (h_stage <= h_limit || big_mt || mt2) &&
                              (d->c_chunks == 1 || mt2 || (d->c_chunks == 2 && d->cout <= 32 && d->num_taps >= 4) ||
                               (halo_wide && d->c_chunks == 2 && d->cout <= 64 && d->num_taps >= 4));
        // (64 -> 64 3x3, two chunks x two channel tiles, measured faster on the per-tap kernel: 0.73 vs 0.89 ms)
        if (eligible) {
            p.bw = 8; p.bh = 16; p.bn = 1;
            p.tiles_x = ceil_div(d->grid_w, 8); p.tiles_y = ceil_div(d->grid_h, 16); p.tiles_n = d->out_n;
            p.dy0 = dy0; p.dx0 = dx0; p.halo_h = halo_h; p.halo_w = halo_w; p.a_stage_bytes = a_bytes;
            p.BN = hBN;
            p.n_tiles_co = ceil_div(d->cout, hBN);
            p.tmem_cols = std::max(32, next_pow2(stack ? 2 * hBN : hBN));
            p.stages = 1;
            alignas(64) CUtensorMap ta, tb;
            {
                const cuuint64_t dims[5] = {64, (cuuint64_t)d->c_chunks, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h,
                                            (cuuint64_t)d->in_n};
                const cuuint64_t s1 = 128, s2 = s1 * d->c_chunks, s3 = s2 * d->in_w, s4 = s3 * d->in_h;
                const cuuint64_t strides[4] = {s1, s2, s3, s4};
                const cuuint32_t box[5] = {64, 1, (cuuint32_t)halo_w, (cuuint32_t)halo_h, 1};
                const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
                CUresult r = encode(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->in), dims, strides, box,
                                    estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
                WGS_REQUIRE(r == CUDA_SUCCESS, "conv(halo): cuTensorMapEncodeTiled(input) failed with code " + std::to_string((int)r));
            }
            {
                CUresult r = encode_weights(&tb, hBN);
                WGS_REQUIRE(r == CUDA_SUCCESS, "conv(halo): cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r));
            }
            static int mt_mode = -1;
            if (mt_mode < 0) {
                const char* e = getenv("WGS_HALO_MT");                // 0 = off, N = tiles per CTA, default: by row width
                mt_mode = e ? atoi(e) : -2;
            }
            // measured on 32 -> 32 @1024^2 x 8 images: 16 tiles per CTA 0.641 ms, 8: 0.665, one-tile kernel: 0.925
            const int mt_tiles = mt_mode == -2 ? (p.tiles_x >= 64 ? 16 : 8) : mt_mode;
            const int hgrid_tiles = p.tiles_y * p.tiles_n * p.n_tiles_co;
            const int mt_use = mt2 ? std::min(mt_tiles > 1 ? mt_tiles : 8, std::max(2, p.tiles_x / 2)) : mt_tiles;
            if (mt_use > 1 && stack && p.tiles_x >= 2 * mt_use &&
                (mt2 || (d->c_chunks == 1 &&
                         ((hBN <= 32 && d->num_taps * hBN * 128 + MT_STAGES * a_bytes + MT_STG_BYTES <= 108 * 1024) || big_mt)))) {
                p.tiles_per_cta = mt_use;
                p.x_groups = ceil_div(p.tiles_x, mt_use);
                p.stages = mt2 ? 3 : MT_STAGES;
                p.tmem_cols = std::max(32, next_pow2(4 * hBN));          // two stacked accumulators
                const size_t msmem = (size_t)d->c_chunks * d->num_taps * hBN * 128 + (size_t)p.stages * a_bytes + MT_STG_BYTES +
                                     (2 * MT_MAX_STAGES + 5) * 8 + 32 + WGS_MAX_TAPS * 4 + 6 * hBN * 4 + 1024;
                WGS_REQUIRE(msmem <= 227 * 1024, "conv(halo, multi-tile): shared memory budget exceeded");
                static bool mattr = false;
                if (!mattr) {
                    WGS_CUDA(cudaFuncSetAttribute(conv_halo_mt_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
                    WGS_CUDA(cudaFuncSetAttribute(conv_halo_mt_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
                    mattr = true;
                }
                const bool mfused = d->out_split != nullptr || d->rgb_out != nullptr || d->out_from_n > 0 || d->out == nullptr;
                const int mgrid = hgrid_tiles * p.x_groups;
                if (mfused) conv_halo_mt_kernel<true><<<mgrid, CONV_THREADS, msmem, st>>>(ta, tb, p);
                else conv_halo_mt_kernel<false><<<mgrid, CONV_THREADS, msmem, st>>>(ta, tb, p);
                count_launch();
                WGS_LAUNCH_CHECK();
                return 0;
            }
            const size_t hsmem = (size_t)p.stages * h_stage + (2 * p.stages + 1) * 8 + 32 + WGS_MAX_TAPS * 4 + 6 * hBN * 4 + 1024;
            static bool hattr = false;
            if (!hattr) {
                WGS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
                WGS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
                WGS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
                WGS_CUDA(cudaFuncSetAttribute(conv_halo_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
                hattr = true;
            }
            const int hgrid = p.tiles_x * p.tiles_y * p.tiles_n * p.n_tiles_co;
            const bool hfused = d->out_split != nullptr || d->rgb_out != nullptr || d->out_from_n > 0 || d->out == nullptr;
            if (stack) {
                if (hfused) conv_halo_kernel<true, true><<<hgrid, CONV_THREADS, hsmem, st>>>(ta, tb, p);
                else conv_halo_kernel<false, true><<<hgrid, CONV_THREADS, hsmem, st>>>(ta, tb, p);
            } else {
                if (hfused) conv_halo_kernel<true, false><<<hgrid, CONV_THREADS, hsmem, st>>>(ta, tb, p);
                else conv_halo_kernel<false, false><<<hgrid, CONV_THREADS, hsmem, st>>>(ta, tb, p);
            }
            count_launch();
            WGS_LAUNCH_CHECK();
            return 0;
        }
    }

    alignas(64) CUtensorMap tmap_a, tmap_b;
    {
        const cuuint64_t dims[5] = {64, (cuuint64_t)d->c_chunks, (cuuint64_t)d->in_w, (cuuint64_t)d->in_h,
                                    (cuuint64_t)d->in_n};
        const cuuint64_t s1 = 128, s2 = s1 * d->c_chunks, s3 = s2 * d->in_w, s4 = s3 * d->in_h;
        const cuuint64_t strides[4] = {s1, s2, s3, s4};
        const cuuint32_t box[5] = {64, 1, (cuuint32_t)(p.bw * d->in_stride), (cuuint32_t)(p.bh * d->in_stride),
                                   (cuuint32_t)p.bn};
        const cuuint32_t estr[5] = {1, 1, (cuuint32_t)d->in_stride, (cuuint32_t)d->in_stride, 1};
        CUresult r = encode(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(d->in), dims, strides, box,
                            estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        WGS_REQUIRE(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(input) failed with code " + std::to_string((int)r));
    }
    {
        CUresult r = encode_weights(&tmap_b, BN);
        WGS_REQUIRE(r == CUDA_SUCCESS, "conv: cuTensorMapEncodeTiled(weights) failed with code " + std::to_string((int)r));
    }
    const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * 8 + 32 + 6 * BN * 4 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        WGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        WGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        WGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        WGS_CUDA(cudaFuncSetAttribute(conv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        attr_set = true;
    }
    const int grid = m_tiles * p.n_tiles_co * p.ksplit;
    const bool fused = d->out_split != nullptr || d->rgb_out != nullptr || d->out_from_n > 0 || d->out == nullptr;
    if (p.ksplit > 1) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)grid, 1, 1);
        cfg.blockDim = dim3(CONV_THREADS, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = (unsigned)p.ksplit;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (fused) WGS_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<true, false>, tmap_a, tmap_b, p));
        else WGS_CUDA(cudaLaunchKernelEx(&cfg, conv_tc_kernel<false, false>, tmap_a, tmap_b, p));
    } else if (stack) {
        if (fused) conv_tc_kernel<true, true><<<grid, CONV_THREADS, smem, st>>>(tmap_a, tmap_b, p);
        else conv_tc_kernel<false, true><<<grid, CONV_THREADS, smem, st>>>(tmap_a, tmap_b, p);
    } else {
        if (fused) conv_tc_kernel<true, false><<<grid, CONV_THREADS, smem, st>>>(tmap_a, tmap_b, p);
        else conv_tc_kernel<false, false><<<grid, CONV_THREADS, smem, st>>>(tmap_a, tmap_b, p);
    }
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_conv_desc_size(void) { return (int)sizeof(wgs_conv_desc); }
