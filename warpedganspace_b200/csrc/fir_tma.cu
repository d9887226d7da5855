// TMA-fed variant of fir4_act (sg2.cu) for the large up-sampling layers (C % 32 == 0, >= 128 pixels wide).
//
// The register-strip kernel in sg2.cu executes ~1400 instructions per thread (address arithmetic, bounds predicates and 44
// global loads per 8-row strip) and sits at 44 % of HBM peak, latency- and issue-bound (profiles/r02_top_ncu.md).  Here a
// producer warp streams (16+3) x (16+3) x 32-channel fp32 input tiles into a two-stage shared-memory ring with one
// cp.async.bulk.tensor per tile - out-of-range rows / columns (the FIR's padding) are zero-filled by the TMA unit, so the
// consumers have no bounds checks at all - and 256 consumer threads run the same separable 4-tap strip algorithm out of
// shared memory with conflict-free 128-bit loads (a warp reads 4 pixels x 128 B = 512 contiguous bytes).  Persistent CTAs
// (2 per SM) walk the tile list; the stores are the same 128-bit fp32 / 64-bit split32 stores as before.
//
// Same contract as wgs_fir4_act: out = act(alpha[n,c] * fir(y)[Y,X,c] + noise_w * noise[Y,X] + beta[c]), fp32 for images
// n >= out_from_n and / or split32(out * split_scale[n,c]).  Replaces upfirdn2d (op/upfirdn2d_kernel.cu:52-137, blur mode) +
// NoiseInjection + FusedLeakyReLU (models/StyleGAN2/model.py:231-241, op/fused_bias_act_kernel.cu:18-49).
#include "common.cuh"
#include "ptx.cuh"
#include "wgs_b200.h"
#include <cuda.h>
#include <cudaTypedefs.h>

namespace wgs {

constexpr int FT_TH = 16, FT_TW = 16, FT_CB = 32;                 // outputs per tile: 16 x 16 pixels x 32 channels
constexpr int FT_IH = FT_TH + 3, FT_IW = FT_TW + 3;
constexpr int FT_STAGE_BYTES = FT_IH * FT_IW * FT_CB * 4;         // 46208
constexpr int FT_STAGES = 2;
constexpr int FT_CONSUMERS = 256;
constexpr int FT_THREADS = FT_CONSUMERS + 32;
constexpr int FT_ROWS = FT_TH / 2;                                 // rows per consumer thread (two thread groups per tile)

struct FirTmaParams {
    float* out;
    __nv_bfloat16* out_split;
    const float* alpha;
    const float* beta;
    const float* noise;
    const float* split_scale;
    long long split_scale_ld;
    float noise_w, k0, k1, k2, k3;
    int N, Hout, Wout, C, pad0, act, out_from_n;
    int tiles_x, tiles_y, c_blocks, total_tiles;
};

__global__ void __launch_bounds__(FT_THREADS, 2)
fir4_act_tma_kernel(const __grid_constant__ CUtensorMap tmap_y, const __grid_constant__ FirTmaParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (ptx::smem_u32(smem_raw) & 127u)) & 127u);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + FT_STAGES * FT_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + FT_STAGES;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        ptx::prefetch_tmap(&tmap_y);
        for (int s = 0; s < FT_STAGES; ++s) { ptx::mbar_init(full_bar + s, 1); ptx::mbar_init(empty_bar + s, FT_CONSUMERS / 32); }
        ptx::fence_mbar_init();
    }
    __syncthreads();
    // tile index -> (c block fastest, then x, y, n): neighbouring CTAs share halo rows / columns in L2
    if (warp == FT_CONSUMERS / 32) {
        if ((threadIdx.x & 31) == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                int r = t;
                const int cb = r % p.c_blocks; r /= p.c_blocks;
                const int tx = r % p.tiles_x; r /= p.tiles_x;
                const int ty = r % p.tiles_y;
                const int n = r / p.tiles_y;
                ptx::mbar_wait(empty_bar + stage, phase ^ 1);
                ptx::mbar_expect_tx(full_bar + stage, (uint32_t)FT_STAGE_BYTES);
                ptx::tma_load_4d(smem + (size_t)stage * FT_STAGE_BYTES, &tmap_y, full_bar + stage, cb * FT_CB,
                                 tx * FT_TW - p.pad0, ty * FT_TH - p.pad0, n);
                if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
        return;
    }
    // consumers: thread -> (channel quad q, column x, row half)
    const int q = threadIdx.x & 7, xl = (threadIdx.x >> 3) & 15, half = threadIdx.x >> 7;
    const float kf[4] = {p.k3, p.k2, p.k1, p.k0};                     // correlation with the flipped kernel
    const int chunk_stride = ((p.C + 31) >> 5) * 64;
    int stage = 0;
    uint32_t phase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        int r = t;
        const int cb = r % p.c_blocks; r /= p.c_blocks;
        const int tx = r % p.tiles_x; r /= p.tiles_x;
        const int ty = r % p.tiles_y;
        const int n = r / p.tiles_y;
        const int c = cb * FT_CB + q * 4;
        const int X = tx * FT_TW + xl, Y0 = ty * FT_TH + half * FT_ROWS;
        // per-tile constants and the noise of this thread's rows: issued before the wait so their latency overlaps it
        float al[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {1.f, 1.f, 1.f, 1.f};
        if (p.alpha) { const float4 v = __ldg(reinterpret_cast<const float4*>(p.alpha + (size_t)n * p.C + c)); al[0] = v.x; al[1] = v.y; al[2] = v.z; al[3] = v.w; }
        if (p.beta) { const float4 v = __ldg(reinterpret_cast<const float4*>(p.beta + c)); be[0] = v.x; be[1] = v.y; be[2] = v.z; be[3] = v.w; }
        if (p.split_scale) {
#pragma unroll
            for (int k = 0; k < 4; ++k) sc[k] = __ldg(p.split_scale + (size_t)n * p.split_scale_ld + c + k);
        }
        float nzv[FT_ROWS];
#pragma unroll
        for (int oy = 0; oy < FT_ROWS; ++oy)
            nzv[oy] = (p.noise && Y0 + oy < p.Hout && X < p.Wout) ? p.noise_w * __ldg(p.noise + (size_t)(Y0 + oy) * p.Wout + X) : 0.f;
        ptx::mbar_wait(full_bar + stage, phase);
        const uint32_t base = ptx::smem_u32(smem) + (uint32_t)stage * FT_STAGE_BYTES +
                              (uint32_t)(((half * FT_ROWS) * FT_IW + xl) * FT_CB + q * 4) * 4u;
        float4 h[FT_ROWS + 3];
#pragma unroll
        for (int a = 0; a < FT_ROWS + 3; ++a) {
            const uint32_t rowp = base + (uint32_t)(a * FT_IW * FT_CB) * 4u;
            const float4 v0 = ptx::lds128(rowp), v1 = ptx::lds128(rowp + FT_CB * 4), v2 = ptx::lds128(rowp + 2 * FT_CB * 4),
                         v3 = ptx::lds128(rowp + 3 * FT_CB * 4);
            h[a].x = kf[0] * v0.x + kf[1] * v1.x + kf[2] * v2.x + kf[3] * v3.x;
            h[a].y = kf[0] * v0.y + kf[1] * v1.y + kf[2] * v2.y + kf[3] * v3.y;
            h[a].z = kf[0] * v0.z + kf[1] * v1.z + kf[2] * v2.z + kf[3] * v3.z;
            h[a].w = kf[0] * v0.w + kf[1] * v1.w + kf[2] * v2.w + kf[3] * v3.w;
        }
        // the stage is consumed (everything is in registers): release it before the stores
        __syncwarp();
        if ((threadIdx.x & 31) == 0) ptx::mbar_arrive(empty_bar + stage);
        if (++stage == FT_STAGES) { stage = 0; phase ^= 1; }
        if (X >= p.Wout) continue;
        const bool f32 = p.out && n >= p.out_from_n;
#pragma unroll
        for (int oy = 0; oy < FT_ROWS; ++oy) {
            const int Y = Y0 + oy;
            if (Y >= p.Hout) break;
            float v4[4];
            v4[0] = kf[0] * h[oy].x + kf[1] * h[oy + 1].x + kf[2] * h[oy + 2].x + kf[3] * h[oy + 3].x;
            v4[1] = kf[0] * h[oy].y + kf[1] * h[oy + 1].y + kf[2] * h[oy + 2].y + kf[3] * h[oy + 3].y;
            v4[2] = kf[0] * h[oy].z + kf[1] * h[oy + 1].z + kf[2] * h[oy + 2].z + kf[3] * h[oy + 3].z;
            v4[3] = kf[0] * h[oy].w + kf[1] * h[oy + 1].w + kf[2] * h[oy + 2].w + kf[3] * h[oy + 3].w;
            const float nz = nzv[oy];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float tv = v4[k] * al[k] + nz + be[k];
                if (p.act == 3) tv = 1.41421356237309515f * (tv > 0.f ? tv : 0.2f * tv);
                else if (p.act == 2) tv = tv > 0.f ? tv : 0.2f * tv;
                else if (p.act == 1) tv = tv > 0.f ? tv : 0.f;
                v4[k] = tv;
            }
            const size_t pix = ((size_t)n * p.Hout + Y) * p.Wout + X;
            if (f32) *reinterpret_cast<float4*>(p.out + pix * p.C + c) = make_float4(v4[0], v4[1], v4[2], v4[3]);
            if (p.out_split) {
                __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) split_bf16(v4[k] * sc[k], hi[k], lo[k]);
                __nv_bfloat16* sp = p.out_split + pix * (size_t)chunk_stride + (size_t)(c >> 5) * 64 + (c & 31);
                *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
                *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
            }
        }
    }
}

static PFN_cuTensorMapEncodeTiled_v12000 fir_get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// returns 1 when this variant took the launch, 0 when the caller should use the register-strip kernel, < 0 on error
int fir4_act_tma_launch(const float* y, float* out, int N, int Hin, int Win, int Hout, int Wout, int C, int pad0,
                        const float* taps4, const float* alpha, const float* beta, const float* noise, float noise_w, int act,
                        void* out_split, const float* split_scale, long long split_scale_ld, int out_from_n, void* stream) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("WGS_FIR_TMA");                     // 0 = register-strip kernel everywhere (A/B switch)
        mode = (e && e[0] == '0') ? 0 : 1;
    }
    if (!mode || C % FT_CB != 0 || Wout < 128 || Hout < 64) return 0;
    if ((reinterpret_cast<uintptr_t>(y) & 15) != 0 || (alpha && (reinterpret_cast<uintptr_t>(alpha) & 15)) ||
        (beta && (reinterpret_cast<uintptr_t>(beta) & 15)))
        return 0;
    auto encode = fir_get_encode();
    if (!encode) return 0;
    alignas(64) CUtensorMap tm;
    const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)Win, (cuuint64_t)Hin, (cuuint64_t)N};
    const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)Win * C * 4, (cuuint64_t)Hin * Win * C * 4};
    const cuuint32_t box[4] = {FT_CB, FT_IW, FT_IH, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    if (encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(y), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 0;
    FirTmaParams p;
    memset(&p, 0, sizeof(p));
    p.out = out; p.out_split = (__nv_bfloat16*)out_split; p.alpha = alpha; p.beta = beta; p.noise = noise;
    p.split_scale = split_scale; p.split_scale_ld = split_scale_ld; p.noise_w = noise_w;
    p.k0 = taps4[0]; p.k1 = taps4[1]; p.k2 = taps4[2]; p.k3 = taps4[3];
    p.N = N; p.Hout = Hout; p.Wout = Wout; p.C = C; p.pad0 = pad0; p.act = act; p.out_from_n = out_from_n;
    p.tiles_x = ceil_div(Wout, FT_TW); p.tiles_y = ceil_div(Hout, FT_TH); p.c_blocks = C / FT_CB;
    p.total_tiles = N * p.tiles_y * p.tiles_x * p.c_blocks;
    const size_t smem = (size_t)FT_STAGES * FT_STAGE_BYTES + 2 * FT_STAGES * 8 + 128;
    static bool attr = false;
    if (!attr) {
        if (cudaFuncSetAttribute(fir4_act_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        attr = true;
    }
    const int grid = std::min(p.total_tiles, 2 * num_sms());
    fir4_act_tma_kernel<<<grid, FT_THREADS, smem, (cudaStream_t)stream>>>(tm, p);
    count_launch();
    if (cudaGetLastError() != cudaSuccess) return fail(__FILE__, __LINE__, "fir4_act (TMA variant): launch failed");
    return 1;
}

}  // namespace wgs
