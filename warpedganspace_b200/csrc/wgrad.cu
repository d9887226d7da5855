// Weight gradient of a convolution on tcgen05 tensor cores (sm_100a).
//
//   dW[t][co][ci] = sum_{n,oy,ox} dy[n,oy,ox,co] * x[n, oy*s - p + ky_t, ox*s - p + kx_t, ci]
//
// Per tap this is a GEMM whose contraction runs over PIXELS, so with NHWC data both operands are
// MN-major (channels contiguous).  A split32 row (128 B = hi32|lo32 of one 32-channel chunk) is
// exactly one 64-element MN atom of the 128B-swizzled MN-major UMMA layout, and a TMA box of 64 pixels
// x 1 chunk lands as the canonical [8 K-rows x 128 B] x 8 block.  M = 128 covers two co chunks, N = 64*NB
// covers NB ci chunks; the accumulator holds the four products hi*hi, hi*lo, lo*hi, lo*lo in separate
// (lane, column) blocks and the epilogue folds them while reducing into dW with atomics (split-K over
// pixel ranges).  Replaces cuDNN's wgrad behind loss.backward() for the Reconstructor
// (lib/trainer.py:250; torchvision resnet18 convs, lib/reconstructor.py:54-61).
#include "common.cuh"
#include "ptx.cuh"
#include "wgs_b200.h"
#include <stdlib.h>
#include <cuda.h>
#include <cudaTypedefs.h>

namespace wgs {

constexpr int WG_THREADS = 192;
constexpr int WG_BK = 64;                        // pixels per pipeline stage
constexpr int WG_BLOCK_BYTES = WG_BK * 128;      // one chunk x 64 pixels

struct WgradParams {
    int n, h, w, ci_chunks, oh, ow, co_chunks, kh, kw, stride, pad;
    int bw, bh, bn, tiles_x, tiles_y, tiles_n;
    int NB, n_blocks, m_blocks, ksplit, stages, tmem_cols;
    int layout, out_co, out_ci;      // layout 0: dw[T][co_pad][ci_pad]; 1: torch dw[out_co][out_ci][T]
    int row_mode, xw, b_block;       // row mode: one CTA = one kernel ROW (kw taps) on an x-halo input patch of xw pixels
    float* dw;
};

// MN-major, 128B swizzle: LBO = stride between 64-element MN atoms, SBO = stride between 8-row K groups
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__global__ void __launch_bounds__(WG_THREADS, 2)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_dy, const __grid_constant__ CUtensorMap tmap_x,
                const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);   // keeps the shared address space (LDS/STS, not generic LD/ST)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_bytes = 2 * WG_BLOCK_BYTES, b_bytes = p.NB * p.b_block;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + (size_t)p.stages * a_bytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_bytes);
    uint64_t* empty_bar = full_bar + p.stages;
    uint64_t* acc_bar = empty_bar + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_bar + 1);

    int t = blockIdx.x;
    const int ks = t % p.ksplit; t /= p.ksplit;
    const int nb = t % p.n_blocks; t /= p.n_blocks;
    const int mb = t % p.m_blocks; t /= p.m_blocks;
    const int tap = t;                                            // row mode: t = ky
    const int ky = p.row_mode ? t : tap / p.kw, kx = p.row_mode ? 0 : tap % p.kw;
    const int n_acc = p.row_mode ? p.kw : 1;
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    const int per = (total_tiles + p.ksplit - 1) / p.ksplit;
    const int tile_lo = ks * per, tile_hi = min(total_tiles, tile_lo + per);
    const int k_steps = max(0, tile_hi - tile_lo);

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_dy);
        ptx::prefetch_tmap(&tmap_x);
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

    if (k_steps > 0) {
        if (warp == 0) {
            if (lane == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int tile = tile_lo; tile < tile_hi; ++tile) {
                    int r = tile;
                    const int tx = r % p.tiles_x; r /= p.tiles_x;
                    const int ty = r % p.tiles_y; r /= p.tiles_y;
                    const int ox0 = tx * p.bw, oy0 = ty * p.bh, n0 = r * p.bn;
                    ptx::mbar_wait(empty_bar + stage, phase ^ 1);
                    ptx::mbar_expect_tx(full_bar + stage, (uint32_t)(a_bytes + b_bytes));
                    uint8_t* sa = smem_a + (size_t)stage * a_bytes;
                    uint8_t* sb = smem_b + (size_t)stage * b_bytes;
                    for (int c = 0; c < 2; ++c)
                        ptx::tma_load_5d(sa + c * WG_BLOCK_BYTES, &tmap_dy, full_bar + stage, 0, mb * 2 + c, ox0, oy0, n0);
                    for (int c = 0; c < p.NB; ++c)
                        ptx::tma_load_5d(sb + c * p.b_block, &tmap_x, full_bar + stage, 0, nb * p.NB + c,
                                         ox0 * p.stride - p.pad + kx, oy0 * p.stride - p.pad + ky, n0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (warp == 1) {
            // whole warp, uniform control flow; one elected lane issues.  Both operands MN-major: a_major [15] = 1,
            // b_major [16] = 1
            const uint32_t idesc = ptx::umma_idesc_bf16(128, (uint32_t)(64 * p.NB)) | (1u << 15) | (1u << 16);
            const uint32_t dhi = (uint32_t)(umma_desc_mn_sw128(0, WG_BLOCK_BYTES, 1024) >> 32);
            const uint32_t dlo_extra = (uint32_t)(umma_desc_mn_sw128(0, WG_BLOCK_BYTES, 1024) & 0xFFFFFFFFu);   // LBO bits
            // B: K groups (8 pixels of one image row) are one patch row apart, chunks one patch apart
            const uint32_t b_sbo = p.row_mode ? (uint32_t)p.xw * 128u : 1024u;
            const uint32_t dhi_b = (uint32_t)(umma_desc_mn_sw128(0, (uint32_t)p.b_block, b_sbo) >> 32);
            const uint32_t dlo_extra_b = (uint32_t)(umma_desc_mn_sw128(0, (uint32_t)p.b_block, b_sbo) & 0xFFFFFFFFu);
            const uint32_t a_lo0 = ((ptx::smem_u32(smem_a) & 0x3FFFFu) >> 4) | dlo_extra;
            const uint32_t b_lo0 = ((ptx::smem_u32(smem_b) & 0x3FFFFu) >> 4) | dlo_extra_b;
            const uint32_t a_step = (uint32_t)a_bytes >> 4, b_step = (uint32_t)b_bytes >> 4;
            const uint32_t b_kstep = (2u * b_sbo) >> 4;               // one K = 16 slice = two K groups
            const uint32_t acc_cols = (uint32_t)(64 * p.NB);
            int stage = 0;
            uint32_t phase = 0;
            for (int k = 0; k < k_steps; ++k) {
                ptx::mbar_wait(full_bar + stage, phase);
                ptx::tc_fence_after();
                const uint32_t a0 = a_lo0 + (uint32_t)stage * a_step, b0 = b_lo0 + (uint32_t)stage * b_step;
                if (ptx::elect_one()) {
                    for (int ax = 0; ax < n_acc; ++ax) {
#pragma unroll
                        for (int kk = 0; kk < WG_BK / 16; ++kk)
                            ptx::mma_f16_lh(tmem_base + (uint32_t)ax * acc_cols, a0 + kk * (2048 >> 4), dhi,
                                            b0 + (uint32_t)ax * (128u >> 4) + (uint32_t)kk * b_kstep, dhi_b, idesc,
                                            (k > 0 || kk > 0) ? 1u : 0u);
                    }
                    ptx::mma_commit(empty_bar + stage);
                }
                __syncwarp();
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (ptx::elect_one()) ptx::mma_commit(acc_bar);
            __syncwarp();
        } else {
            const int q = warp & 3;
            const int m = q * 32 + lane;                         // accumulator row
            const int co = (mb * 2 + m / 64) * 32 + (m % 32);    // hi and lo rows of a chunk fold into one co
            const int ci_pad = p.ci_chunks * 32;
            const bool row_ok = (mb * 2 + m / 64) < p.co_chunks && (p.layout == 0 || co < p.out_co);
            const int T = p.kh * p.kw;
            ptx::mbar_wait(acc_bar, 0);
            ptx::tc_fence_after();
            // Rows m and m + 32 of a 64-row block are the hi and lo halves of the SAME output channels: the lo warp hands its
            // 32 column sums to the hi warp through shared memory (the operand ring is idle once the accumulator barrier has
            // fired), so every weight gets one atomic per CTA instead of two - split-K atomics, not MMAs, bound this kernel.
            float* xch = reinterpret_cast<float*>(smem) + (size_t)(q >> 1) * (32 * 33);
            const int pair_bar = 1 + (q >> 1);
            const bool is_lo = (q & 1) != 0;
            for (int ax = 0; ax < n_acc; ++ax) {
                const int tap_o = p.row_mode ? ky * p.kw + ax : tap;
                float* dst_row = p.dw + ((size_t)tap_o * p.co_chunks * 32 + co) * ci_pad;
                for (int c = 0; c < p.NB; ++c) {
                    float v[64];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        ptx::tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((ax * p.NB + c) * 64 + j * 16), v + j * 16);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += v[j + 32];                 // x hi + x lo columns
                    if (is_lo) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) xch[lane * 33 + j] = v[j];
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
                    if (!is_lo) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += xch[lane * 33 + j];
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");    // xch is reused by the next block
                    const int chunk = nb * p.NB + c;
                    if (is_lo || !row_ok || chunk >= p.ci_chunks) continue;
                    if (p.layout == 0) {
                        float* dst = dst_row + chunk * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) atomicAdd(dst + j, v[j]);
                    } else {
                        float* dst = p.dw + ((size_t)co * p.out_ci + chunk * 32) * T + tap_o;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (chunk * 32 + j < p.out_ci) atomicAdd(dst + (size_t)j * T, v[j]);
                    }
                }
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        ptx::tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// CUDA-core twin (debug / cross-check): one thread per (tap, co, ci)
__global__ void wgrad_simt_kernel(const __nv_bfloat16* __restrict__ xs, const __nv_bfloat16* __restrict__ dys,
                                  const WgradParams p) {
    const int ci_pad = p.ci_chunks * 32, co_pad = p.co_chunks * 32;
    const long long total = (long long)p.kh * p.kw * co_pad * ci_pad;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int ci = (int)(i % ci_pad);
        const int co = (int)((i / ci_pad) % co_pad);
        const int tap = (int)(i / ((long long)ci_pad * co_pad));
        const int ky = tap / p.kw, kx = tap % p.kw;
        float acc = 0.f;
        for (int n = 0; n < p.n; ++n)
            for (int oy = 0; oy < p.oh; ++oy) {
                const int iy = oy * p.stride - p.pad + ky;
                if (iy < 0 || iy >= p.h) continue;
                for (int ox = 0; ox < p.ow; ++ox) {
                    const int ix = ox * p.stride - p.pad + kx;
                    if (ix < 0 || ix >= p.w) continue;
                    const __nv_bfloat16* a = dys + ((((size_t)n * p.oh + oy) * p.ow + ox) * p.co_chunks + co / 32) * 64 + co % 32;
                    const __nv_bfloat16* b = xs + ((((size_t)n * p.h + iy) * p.w + ix) * p.ci_chunks + ci / 32) * 64 + ci % 32;
                    const float ah = __bfloat162float(a[0]), al = __bfloat162float(a[32]);
                    const float bh = __bfloat162float(b[0]), bl = __bfloat162float(b[32]);
                    acc += ah * bh + ah * bl + al * bh + al * bl;
                }
            }
        if (p.layout == 0) p.dw[i] += acc;
        else if (co < p.out_co && ci < p.out_ci) p.dw[((size_t)co * p.out_ci + ci) * (p.kh * p.kw) + tap] += acc;
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 get_encode_w() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* q = nullptr;
        cudaDriverEntryPointQueryResult r;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
            r == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(q);
    }
    return fn;
}

static int pow2ceil(int v) { int r = 1; while (r < v) r <<= 1; return r; }

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_conv_wgrad_split32(const void* xs, int n, int h, int w, int ci_chunks, const void* dys, int oh,
                                      int ow, int co_chunks, int kh, int kw, int stride, int pad, float* dw,
                                      int layout, int out_co, int out_ci, void* stream) {
    WGS_REQUIRE(layout == 0 || (layout == 1 && out_co > 0 && out_ci > 0), "wgrad: bad output layout");
    WGS_REQUIRE(n > 0 && h > 0 && w > 0 && oh > 0 && ow > 0 && ci_chunks > 0 && co_chunks > 0, "wgrad: bad sizes");
    WGS_REQUIRE(kh >= 1 && kw >= 1 && kh * kw <= 64 && stride >= 1 && stride <= 8 && pad >= 0, "wgrad: bad kernel geometry");
    WgradParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.h = h; p.w = w; p.ci_chunks = ci_chunks; p.oh = oh; p.ow = ow; p.co_chunks = co_chunks;
    p.kh = kh; p.kw = kw; p.stride = stride; p.pad = pad; p.dw = dw;
    p.layout = layout; p.out_co = out_co; p.out_ci = out_ci;
    p.bw = std::min(8, pow2ceil(ow));
    p.bh = std::min(WG_BK / p.bw, pow2ceil(oh));
    p.bn = WG_BK / (p.bw * p.bh);
    p.tiles_x = ceil_div(ow, p.bw); p.tiles_y = ceil_div(oh, p.bh); p.tiles_n = ceil_div(n, p.bn);
    static int row_env = -1;
    if (row_env < 0) {
        const char* e = getenv("WGS_WGRAD_ROW");                  // 1 = row mode (default: one CTA per tap)
        row_env = (e && e[0] == '1') ? 1 : 0;                     // measured: no gain (the kernel is bound by its atomics)
    }
    // row mode keeps kw accumulators of 64*NB columns in the 512 TMEM columns
    // on by default for long tap lists (the 4x4-tap space-to-depth stem: one CTA per tap re-read x and dy 16 times and was
    // L2-bound, 0.45 ms; one CTA per kernel ROW loads dy once for its 4 taps: -0.29 ms on the training step); measured
    // neutral for 3x3 layers, where it stays opt-in
    p.row_mode = ((row_env || kh * kw >= 12) && stride == 1 && kw >= 2 && kw * 64 <= 512 && p.bw == 8 && p.bh == 8) ? 1 : 0;
    p.NB = std::min(p.row_mode ? std::max(1, 512 / (64 * kw)) : 4, std::min(4, ci_chunks));
    p.n_blocks = ceil_div(ci_chunks, p.NB);
    p.m_blocks = ceil_div(co_chunks, 2);
    const int total_tiles = p.tiles_x * p.tiles_y * p.tiles_n;
    p.xw = p.row_mode ? p.bw + kw - 1 : p.bw;
    p.b_block = p.row_mode ? (p.bh * p.xw * 128 + 1023) / 1024 * 1024 : WG_BLOCK_BYTES;
    const int base = (p.row_mode ? kh : kh * kw) * p.n_blocks * p.m_blocks;
    static int wg_ctas = -1;
    if (wg_ctas < 0) {
        const char* e = getenv("WGS_WGRAD_CTAS");                 // CTAs per SM: 1 = deep ring, 2 = two shallower rings
        wg_ctas = (e && atoi(e) == 1) ? 1 : 2;
    }
    // two CTAs per SM overlap one CTA's prologue / atomic epilogue with the other's main loop; row mode with three
    // 128-column accumulators (384 -> 512 TMEM columns) cannot share an SM
    const int tmem_need = 64 * p.NB * (p.row_mode ? kw : 1);
    const int ctas = (wg_ctas == 2 && tmem_need <= 256) ? 2 : 1;
    // every extra K split adds one atomic per weight: aim for one full wave of CTAs, not several
    static int wg_waves = -1;
    if (wg_waves < 0) {
        const char* e = getenv("WGS_WGRAD_WAVES");                // CTA waves targeted by the K split (x ctas per SM)
        wg_waves = e ? std::max(1, atoi(e)) : 1;
    }
    int ksplit = std::max(1, (wg_waves * ctas * num_sms()) / base);
    ksplit = std::min(ksplit, std::max(1, total_tiles / 8));
    p.ksplit = std::max(1, std::min(ksplit, total_tiles));
    const int stage_bytes = 2 * WG_BLOCK_BYTES + p.NB * p.b_block;
    p.stages = std::max(2, std::min(8, ((ctas == 2 ? 104 : 200) * 1024) / stage_bytes));
    p.tmem_cols = std::max(32, pow2ceil(64 * p.NB * (p.row_mode ? kw : 1)));
    const cudaStream_t st = (cudaStream_t)stream;

    const char* impl = getenv("WGS_CONV_IMPL");
    if (impl && std::string(impl) == "simt") {
        const long long total = (long long)kh * kw * co_chunks * 32 * ci_chunks * 32;
        wgrad_simt_kernel<<<(int)std::min<long long>((total + 127) / 128, 148 * 32), 128, 0, st>>>(
            (const __nv_bfloat16*)xs, (const __nv_bfloat16*)dys, p);
        count_launch();
        WGS_LAUNCH_CHECK();
        return 0;
    }
    auto encode = get_encode_w();
    WGS_REQUIRE(encode != nullptr, "wgrad: cuTensorMapEncodeTiled entry point not available");
    alignas(64) CUtensorMap tmap_dy, tmap_x;
    {
        const cuuint64_t dims[5] = {64, (cuuint64_t)co_chunks, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)n};
        const cuuint64_t s1 = 128, s2 = s1 * co_chunks, s3 = s2 * ow, s4 = s3 * oh;
        const cuuint64_t strides[4] = {s1, s2, s3, s4};
        const cuuint32_t box[5] = {64, 1, (cuuint32_t)p.bw, (cuuint32_t)p.bh, (cuuint32_t)p.bn};
        const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
        CUresult r = encode(&tmap_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dys), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        WGS_REQUIRE(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled(dy) failed with code " + std::to_string((int)r));
    }
    {
        const cuuint64_t dims[5] = {64, (cuuint64_t)ci_chunks, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
        const cuuint64_t s1 = 128, s2 = s1 * ci_chunks, s3 = s2 * w, s4 = s3 * h;
        const cuuint64_t strides[4] = {s1, s2, s3, s4};
        const cuuint32_t box[5] = {64, 1, (cuuint32_t)(p.row_mode ? p.xw : p.bw * stride), (cuuint32_t)(p.bh * stride), (cuuint32_t)p.bn};
        const cuuint32_t estr[5] = {1, 1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
        CUresult r = encode(&tmap_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(xs), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        WGS_REQUIRE(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled(x) failed with code " + std::to_string((int)r));
    }
    const size_t smem = (size_t)p.stages * stage_bytes + (2 * p.stages + 1) * 8 + 16 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        WGS_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        attr_set = true;
    }
    const int grid = base * p.ksplit;
    wgrad_tc_kernel<<<grid, WG_THREADS, smem, st>>>(tmap_dy, tmap_x, p);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
