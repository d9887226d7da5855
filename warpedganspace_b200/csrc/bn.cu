// Train-mode BatchNorm (+ residual + ReLU) for the Reconstructor, fused with the operand packing of the next
// convolution.  Replaces, per conv block, cuDNN batch-norm forward/backward, ATen relu / threshold_backward / add and a
// separate split32 pack (lib/reconstructor.py:54-69 -> torchvision BasicBlock; train mode set at lib/trainer.py:150):
//   forward : bn_stats (one read of y)  ->  bn_act_fwd (y [, residual] -> z fp32 and split32(z))
//   backward: bn_act_bwd_reduce (dz, z, y -> sum dzr, sum dzr*xhat)  ->  bn_act_bwd_apply (-> split32(dy) [, dzr])
// All tensors are NHWC [R, C] (R = N*H*W rows); C is a multiple of 4 with (C/4) dividing 256.
#include "common.cuh"
#include "wgs_b200.h"

namespace wgs {

constexpr int BN_THREADS = 256;

__device__ __forceinline__ void bn_reduce2_to_global(const float a[4], const float b[4], float* sm, int C, int C4, int PL,
                                                     float* dst_a, float* dst_b) {
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    float* sa = sm;
    float* sb = sm + (size_t)PL * C;
    *reinterpret_cast<float4*>(sa + (size_t)pl * C + q * 4) = make_float4(a[0], a[1], a[2], a[3]);
    *reinterpret_cast<float4*>(sb + (size_t)pl * C + q * 4) = make_float4(b[0], b[1], b[2], b[3]);
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        float ta = 0.f, tb = 0.f;
        for (int l = 0; l < PL; ++l) { ta += sa[(size_t)l * C + c]; tb += sb[(size_t)l * C + c]; }
        atomicAdd(dst_a + c, ta);
        atomicAdd(dst_b + c, tb);
    }
}

// SHIFTED single-pass statistics: sum[c] += sum_r (y[r,c] - s_c), sumsq[c] += sum_r (y[r,c] - s_c)^2 with the shift
// s_c = y[0, c] (the channel's first sample, so |mean - s_c| ~ std): var = E[(y-s)^2] - E[y-s]^2 then loses O(1) digits to
// cancellation instead of O(mean^2 / var) - a large-mean channel (|mean| >> std, e.g. after a bias) would otherwise lose
// all of its fp32 variance digits (cuDNN / ATen use Welford for the same reason).
__global__ void __launch_bounds__(BN_THREADS)
bn_stats_kernel(const float* __restrict__ y, long long R, int C, float* __restrict__ sum, float* __restrict__ sumsq) {
    extern __shared__ float sm[];
    const int C4 = C >> 2, PL = BN_THREADS / C4;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    const long long per = (R + gridDim.x - 1) / gridDim.x;
    const long long r0 = blockIdx.x * per, r1 = min(R, r0 + per);
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    const float4 sh = __ldg(reinterpret_cast<const float4*>(y + q * 4));            // row 0
    for (long long r = r0 + pl; r < r1; r += PL) {
        float4 v = __ldg(reinterpret_cast<const float4*>(y + r * C + q * 4));
        v.x -= sh.x; v.y -= sh.y; v.z -= sh.z; v.w -= sh.w;
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
        b[0] += v.x * v.x; b[1] += v.y * v.y; b[2] += v.z * v.z; b[3] += v.w * v.w;
    }
    bn_reduce2_to_global(a, b, sm, C, C4, PL, sum, sumsq);
}

// mean / rstd from the sums, running-stat update (momentum, unbiased variance), one thread per channel
__device__ __forceinline__ void bn_moments(float sum, float sumsq, float shift, long long R, float eps, float& mean, float& rstd,
                                           double& var_out) {
    const double ms = (double)sum / (double)R;                    // mean of (y - shift)
    double var = (double)sumsq / (double)R - ms * ms;
    if (var < 0.0) var = 0.0;
    mean = (float)(ms + (double)shift);
    rstd = (float)(1.0 / sqrt(var + (double)eps));
    var_out = var;
}

__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sumsq, const float* __restrict__ shift,
                                   long long R, int C, float eps, float momentum, float* __restrict__ mean,
                                   float* __restrict__ rstd, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double var;
    float mf, rf;
    bn_moments(sum[c], sumsq[c], shift ? shift[c] : 0.f, R, eps, mf, rf, var);
    const double m = (double)mf;
    mean[c] = mf;
    rstd[c] = rf;
    if (running_mean) {
        const double unbiased = R > 1 ? var * (double)R / (double)(R - 1) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

// z = [relu]( gamma * (y - mean) * rstd + beta [+ residual] );  writes z (fp32, optional) and split32(z) (optional)
// FINALIZE = true: mean / rstd are derived here from the shifted sums of bn_stats (every thread owns one channel quad for
// its whole grid-stride loop: C/4 divides the stride), block 0 publishes them for the backward pass and updates the
// running statistics - the separate one-thread-per-channel finalize launch between bn_stats and this kernel is gone.
template <bool FINALIZE>
__global__ void __launch_bounds__(BN_THREADS)
bn_act_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ residual,
                  int relu, float* __restrict__ z, __nv_bfloat16* __restrict__ zs, long long R, int C,
                  const float* __restrict__ sum, const float* __restrict__ sumsq, float eps, float momentum,
                  float* __restrict__ mean_out, float* __restrict__ rstd_out, float* __restrict__ running_mean,
                  float* __restrict__ running_var, const float* __restrict__ shift) {
    const int C4 = C >> 2;
    const long long total = R * C4;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int q = (int)(i0 % C4), c = q * 4;                       // fixed per thread (host guarantees stride % C4 == 0)
    float4 m4, s4;
    if (FINALIZE) {
        // fp32 is enough here: the sums are SHIFTED (bn_stats), so E[(y-s)^2] - E[y-s]^2 does not cancel catastrophically;
        // only block 0 redoes the moments in double for the published mean / rstd / running statistics
        const float4 su = __ldg(reinterpret_cast<const float4*>(sum + c)), sq = __ldg(reinterpret_cast<const float4*>(sumsq + c));
        // the shift the sums were formed with: row 0 of y (wgs_bn_stats) or the caller's vector (statistics formed in the
        // conv epilogue, wgs_conv_desc.stat_shift)
        const float4 sh = __ldg(reinterpret_cast<const float4*>(shift ? shift + c : y + c));
        if (blockIdx.x == 0 && threadIdx.x < C4) {
            double v0, v1, v2, v3;
            bn_moments(su.x, sq.x, sh.x, R, eps, m4.x, s4.x, v0);
            bn_moments(su.y, sq.y, sh.y, R, eps, m4.y, s4.y, v1);
            bn_moments(su.z, sq.z, sh.z, R, eps, m4.z, s4.z, v2);
            bn_moments(su.w, sq.w, sh.w, R, eps, m4.w, s4.w, v3);
            *reinterpret_cast<float4*>(mean_out + c) = m4;
            *reinterpret_cast<float4*>(rstd_out + c) = s4;
            if (running_mean) {
                const double ub = R > 1 ? (double)R / (double)(R - 1) : 1.0;
                const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
                const double vv[4] = {v0, v1, v2, v3};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    running_mean[c + k] = (1.f - momentum) * running_mean[c + k] + momentum * mm[k];
                    running_var[c + k] = (1.f - momentum) * running_var[c + k] + momentum * (float)(vv[k] * ub);
                }
            }
        }
        const float inv_r = 1.f / (float)R;
        const float a0 = su.x * inv_r, a1 = su.y * inv_r, a2 = su.z * inv_r, a3 = su.w * inv_r;
        m4 = make_float4(a0 + sh.x, a1 + sh.y, a2 + sh.z, a3 + sh.w);
        s4 = make_float4(rsqrtf(fmaxf(sq.x * inv_r - a0 * a0, 0.f) + eps), rsqrtf(fmaxf(sq.y * inv_r - a1 * a1, 0.f) + eps),
                         rsqrtf(fmaxf(sq.z * inv_r - a2 * a2, 0.f) + eps), rsqrtf(fmaxf(sq.w * inv_r - a3 * a3, 0.f) + eps));
    } else {
        m4 = __ldg(reinterpret_cast<const float4*>(mean + c));
        s4 = __ldg(reinterpret_cast<const float4*>(rstd + c));
    }
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 k4 = make_float4(s4.x * g4.x, s4.y * g4.y, s4.z * g4.z, s4.w * g4.w);
    for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / C4;
        const float4 v = __ldg(reinterpret_cast<const float4*>(y + r * C + c));
        float o[4] = {(v.x - m4.x) * k4.x + b4.x, (v.y - m4.y) * k4.y + b4.y,
                      (v.z - m4.z) * k4.z + b4.z, (v.w - m4.w) * k4.w + b4.w};
        if (residual) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(residual + r * C + c));
            o[0] += e.x; o[1] += e.y; o[2] += e.z; o[3] += e.w;
        }
        if (relu) {
#pragma unroll
            for (int k = 0; k < 4; ++k) o[k] = o[k] > 0.f ? o[k] : 0.f;
        }
        if (z) *reinterpret_cast<float4*>(z + r * C + c) = make_float4(o[0], o[1], o[2], o[3]);
        if (zs) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
            __nv_bfloat16* sp = zs + r * (long long)(((C + 31) >> 5) * 64) + (c >> 5) * 64 + (c & 31);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// dzr = dz * (relu ? z > 0 : 1);  sum_dz[c] += sum_r dzr;  sum_dzx[c] += sum_r dzr * (y - mean) * rstd
__global__ void __launch_bounds__(BN_THREADS)
bn_act_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y,
                         const float* __restrict__ mean, const float* __restrict__ rstd, int relu, long long R, int C,
                         float* __restrict__ sum_dz, float* __restrict__ sum_dzx, const float* __restrict__ gamma,
                         const float* __restrict__ beta) {
    extern __shared__ float sm[];
    const int C4 = C >> 2, PL = BN_THREADS / C4;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4;
    const long long per = (R + gridDim.x - 1) / gridDim.x;
    const long long r0 = blockIdx.x * per, r1 = min(R, r0 + per);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + q * 4));
    const float4 s4 = __ldg(reinterpret_cast<const float4*>(rstd + q * 4));
    // z == nullptr (no residual): the ReLU mask is re-derived from y - the normalised activation was never stored
    float4 k4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = k4;
    if (relu && !z) {
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + q * 4));
        b4 = __ldg(reinterpret_cast<const float4*>(beta + q * 4));
        k4 = make_float4(s4.x * g4.x, s4.y * g4.y, s4.z * g4.z, s4.w * g4.w);
    }
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    for (long long r = r0 + pl; r < r1; r += PL) {
        const long long off = r * C + q * 4;
        float4 g = __ldg(reinterpret_cast<const float4*>(dz + off));
        const float4 v = __ldg(reinterpret_cast<const float4*>(y + off));
        if (relu && !z) {
            g.x = (v.x - m4.x) * k4.x + b4.x > 0.f ? g.x : 0.f; g.y = (v.y - m4.y) * k4.y + b4.y > 0.f ? g.y : 0.f;
            g.z = (v.z - m4.z) * k4.z + b4.z > 0.f ? g.z : 0.f; g.w = (v.w - m4.w) * k4.w + b4.w > 0.f ? g.w : 0.f;
        } else if (relu) {
            const float4 zz = __ldg(reinterpret_cast<const float4*>(z + off));
            g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
            g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
        }
        a[0] += g.x; a[1] += g.y; a[2] += g.z; a[3] += g.w;
        b[0] += g.x * (v.x - m4.x) * s4.x; b[1] += g.y * (v.y - m4.y) * s4.y;
        b[2] += g.z * (v.z - m4.z) * s4.z; b[3] += g.w * (v.w - m4.w) * s4.w;
    }
    bn_reduce2_to_global(a, b, sm, C, C4, PL, sum_dz, sum_dzx);
}

// dy = gamma * rstd * (dzr - sum_dz/R - xhat * sum_dzx/R)  -> split32 (operand of the dgrad / wgrad convs) and/or fp32;
// dres = dzr (gradient of the residual branch, optional)
__global__ void __launch_bounds__(BN_THREADS)
bn_act_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ z, const float* __restrict__ y,
                        const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                        const float* __restrict__ sum_dz, const float* __restrict__ sum_dzx, int relu, long long R, int C,
                        __nv_bfloat16* __restrict__ dys, float* __restrict__ dy, float* __restrict__ dres,
                        const float* __restrict__ beta) {
    const int C4 = C >> 2;
    const long long total = R * C4;
    const float inv_r = 1.f / (float)R;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int q = (int)(i % C4);
        const long long r = i / C4;
        const int c = q * 4;
        const long long off = r * C + c;
        float4 g = __ldg(reinterpret_cast<const float4*>(dz + off));
        const float4 v = __ldg(reinterpret_cast<const float4*>(y + off));
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + c));
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(rstd + c));
        const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
        if (relu && !z) {                           // mask re-derived from y (same expression as the reduce pass)
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
            g.x = (v.x - m4.x) * (s4.x * g4.x) + b4.x > 0.f ? g.x : 0.f; g.y = (v.y - m4.y) * (s4.y * g4.y) + b4.y > 0.f ? g.y : 0.f;
            g.z = (v.z - m4.z) * (s4.z * g4.z) + b4.z > 0.f ? g.z : 0.f; g.w = (v.w - m4.w) * (s4.w * g4.w) + b4.w > 0.f ? g.w : 0.f;
        } else if (relu) {
            const float4 zz = __ldg(reinterpret_cast<const float4*>(z + off));
            g.x = zz.x > 0.f ? g.x : 0.f; g.y = zz.y > 0.f ? g.y : 0.f;
            g.z = zz.z > 0.f ? g.z : 0.f; g.w = zz.w > 0.f ? g.w : 0.f;
        }
        if (dres) *reinterpret_cast<float4*>(dres + off) = g;
        const float4 a4 = __ldg(reinterpret_cast<const float4*>(sum_dz + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(sum_dzx + c));
        float o[4];
        o[0] = g4.x * s4.x * (g.x - a4.x * inv_r - (v.x - m4.x) * s4.x * b4.x * inv_r);
        o[1] = g4.y * s4.y * (g.y - a4.y * inv_r - (v.y - m4.y) * s4.y * b4.y * inv_r);
        o[2] = g4.z * s4.z * (g.z - a4.z * inv_r - (v.z - m4.z) * s4.z * b4.z * inv_r);
        o[3] = g4.w * s4.w * (g.w - a4.w * inv_r - (v.w - m4.w) * s4.w * b4.w * inv_r);
        if (dy) *reinterpret_cast<float4*>(dy + off) = make_float4(o[0], o[1], o[2], o[3]);
        if (dys) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
            __nv_bfloat16* sp = dys + r * (long long)(((C + 31) >> 5) * 64) + (c >> 5) * 64 + (c & 31);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

static bool bn_ok(int C) { return C >= 4 && C % 4 == 0 && (C / 4) <= BN_THREADS && BN_THREADS % (C / 4) == 0; }
// reduction grids: ~8 float4 loads per thread (the deep layers have few rows but wide channels: 64 rows per block left
// layer 4 at 64 blocks of 32 dependent iterations each - 19 us for 17 MB)
static int bn_red_blocks(long long R, int C) {
    const long long quads = R * (C / 4);
    return (int)std::max<long long>(1, std::min<long long>((long long)num_sms() * 8, (quads + 2047) / 2048));
}
static int bn_ew_blocks(long long total) { return (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)num_sms() * 32)); }

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_bn_stats(const float* y, long long R, int C, float* sum, float* sumsq, void* stream) {
    WGS_REQUIRE(bn_ok(C) && R > 0, "bn_stats: C must be a multiple of 4 with C/4 dividing 256");
    const size_t smem = 2 * (size_t)(BN_THREADS / (C / 4)) * C * sizeof(float);
    bn_stats_kernel<<<bn_red_blocks(R, C), BN_THREADS, smem, (cudaStream_t)stream>>>(y, R, C, sum, sumsq);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_finalize(const float* sum, const float* sumsq, const float* shift, long long R, int C, float eps,
                               float momentum, float* mean, float* rstd, float* running_mean, float* running_var, void* stream) {
    WGS_REQUIRE(C > 0 && R > 0, "bn_finalize: bad sizes");
    bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, (cudaStream_t)stream>>>(sum, sumsq, shift, R, C, eps, momentum, mean, rstd,
                                                                         running_mean, running_var);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_act_fwd(const float* y, const float* mean, const float* rstd, const float* gamma, const float* beta,
                              const float* residual, int relu, float* z, void* zs, long long R, int C, void* stream) {
    WGS_REQUIRE(bn_ok(C) && R > 0, "bn_act_fwd: C must be a multiple of 4 with C/4 dividing 256");
    bn_act_fwd_kernel<false><<<bn_ew_blocks(R * (C / 4)), BN_THREADS, 0, (cudaStream_t)stream>>>(
        y, mean, rstd, gamma, beta, residual, relu, z, (__nv_bfloat16*)zs, R, C, nullptr, nullptr, 0.f, 0.f, nullptr, nullptr,
        nullptr, nullptr, nullptr);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_fwd_fused(const float* y, const float* sum, const float* sumsq, const float* shift, long long R, int C,
                                float eps, float momentum, const float* gamma, const float* beta, const float* residual, int relu,
                                float* z, void* zs, float* mean, float* rstd, float* running_mean, float* running_var,
                                void* stream) {
    WGS_REQUIRE(bn_ok(C) && R > 0, "bn_fwd_fused: C must be a multiple of 4 with C/4 dividing 256");
    WGS_REQUIRE(sum && sumsq && mean && rstd, "bn_fwd_fused: sums in, mean / rstd out are required");
    bn_act_fwd_kernel<true><<<bn_ew_blocks(R * (C / 4)), BN_THREADS, 0, (cudaStream_t)stream>>>(
        y, nullptr, nullptr, gamma, beta, residual, relu, z, (__nv_bfloat16*)zs, R, C, sum, sumsq, eps, momentum, mean, rstd,
        running_mean, running_var, shift);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_act_bwd_reduce(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, int relu, long long R, int C, float* sum_dz,
                                     float* sum_dzx, void* stream) {
    WGS_REQUIRE(bn_ok(C) && R > 0, "bn_act_bwd_reduce: C must be a multiple of 4 with C/4 dividing 256");
    WGS_REQUIRE(!relu || z || (gamma && beta), "bn_act_bwd_reduce: without z the ReLU mask needs gamma and beta");
    const size_t smem = 2 * (size_t)(BN_THREADS / (C / 4)) * C * sizeof(float);
    bn_act_bwd_reduce_kernel<<<bn_red_blocks(R, C), BN_THREADS, smem, (cudaStream_t)stream>>>(dz, z, y, mean, rstd, relu, R, C,
                                                                                         sum_dz, sum_dzx, gamma, beta);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_act_bwd_apply(const float* dz, const float* z, const float* y, const float* mean, const float* rstd,
                                    const float* gamma, const float* beta, const float* sum_dz, const float* sum_dzx, int relu,
                                    long long R, int C, void* dys, float* dy, float* dres, void* stream) {
    WGS_REQUIRE(C % 4 == 0 && R > 0, "bn_act_bwd_apply: C must be a multiple of 4");
    WGS_REQUIRE(!relu || z || beta, "bn_act_bwd_apply: without z the ReLU mask needs beta");
    bn_act_bwd_apply_kernel<<<bn_ew_blocks(R * (C / 4)), BN_THREADS, 0, (cudaStream_t)stream>>>(
        dz, z, y, mean, rstd, gamma, sum_dz, sum_dzx, relu, R, C, (__nv_bfloat16*)dys, dy, dres, beta);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

// ---------------------------------------------------------------------------------------------------
// 3x3 / stride 2 / pad 1 max-pool of the ResNet stem (torchvision resnet18.maxpool), NHWC, fused with the split32
// packing of the first BasicBlock's operand; the backward is a gather (each input pixel looks at the <= 4 windows that
// contain it), so no memset / scatter-atomics are needed.  idx holds the argmax tap (0..8) per element.
namespace wgs {

__global__ void __launch_bounds__(256)
maxpool3s2_fwd_kernel(const float* __restrict__ z, int N, int H, int W, int C, float* __restrict__ out,
                      unsigned char* __restrict__ idx, __nv_bfloat16* __restrict__ outs) {
    const int OH = (H + 1) / 2, OW = (W + 1) / 2, C4 = C >> 2;
    const long long total = (long long)N * OH * OW * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long r = i / C4;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        unsigned char arg[4] = {0, 0, 0, 0};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int y = 2 * oy - 1 + t / 3, x = 2 * ox - 1 + t % 3;
            if (y < 0 || y >= H || x < 0 || x >= W) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(z + (((size_t)n * H + y) * W + x) * C + c));
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (vv[k] > best[k]) { best[k] = vv[k]; arg[k] = (unsigned char)t; }
        }
        const size_t pix = ((size_t)n * OH + oy) * OW + ox;
        *reinterpret_cast<float4*>(out + pix * C + c) = make_float4(best[0], best[1], best[2], best[3]);
        *reinterpret_cast<uchar4*>(idx + pix * C + c) = make_uchar4(arg[0], arg[1], arg[2], arg[3]);
        if (outs) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(best[k], hi[k], lo[k]);
            __nv_bfloat16* sp = outs + pix * (size_t)(((C + 31) >> 5) * 64) + (size_t)(c >> 5) * 64 + (c & 31);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ idx, int N, int H, int W, int C,
                      float* __restrict__ dz) {
    const int OH = (H + 1) / 2, OW = (W + 1) / 2, C4 = C >> 2;
    const long long total = (long long)N * H * W * C4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C4) * 4;
        long long r = i / C4;
        const int x = (int)(r % W); r /= W;
        const int y = (int)(r % H);
        const int n = (int)(r / H);
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        // windows containing (y, x): oy in {floor(y/2), floor((y+1)/2)} (deduplicated), same for x
        const int oy_a = y >> 1, oy_b = (y + 1) >> 1, ox_a = x >> 1, ox_b = (x + 1) >> 1;
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            const int oy = a ? oy_b : oy_a;
            if ((a && oy_b == oy_a) || oy >= OH) continue;
            const int ty = y - (2 * oy - 1);
            if (ty < 0 || ty > 2) continue;
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                const int ox = b ? ox_b : ox_a;
                if ((b && ox_b == ox_a) || ox >= OW) continue;
                const int tx = x - (2 * ox - 1);
                if (tx < 0 || tx > 2) continue;
                const unsigned char t = (unsigned char)(ty * 3 + tx);
                const size_t pix = ((size_t)n * OH + oy) * OW + ox;
                const uchar4 id = *reinterpret_cast<const uchar4*>(idx + pix * C + c);
                const float4 g = __ldg(reinterpret_cast<const float4*>(dout + pix * C + c));
                if (id.x == t) acc[0] += g.x;
                if (id.y == t) acc[1] += g.y;
                if (id.z == t) acc[2] += g.z;
                if (id.w == t) acc[3] += g.w;
            }
        }
        *reinterpret_cast<float4*>(dz + (((size_t)n * H + y) * W + x) * C + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

}  // namespace wgs

// ---------------------------------------------------------------------------------------------------
// The ResNet stem's BatchNorm + ReLU + 3x3/2 max-pool as ONE forward kernel and two backward kernels (torchvision resnet18
// bn1 / relu / maxpool, lib/reconstructor.py:54-69).  The normalised activation z = relu(bn(y)) (268 MB at 4 x 512^2 x 64) is
// never stored: the forward pools it on the fly out of the conv output y, the backward re-derives the ReLU mask from y and
// gathers the pooled gradient through the arg-max table - instead of bn_act_fwd (write z) -> maxpool fwd (read z) and
// maxpool bwd (write dz) -> bn reduce (read dz, z, y) -> bn apply (read dz, z, y).
namespace wgs {

__device__ __forceinline__ float bn_affine(float v, float m, float k, float b) { return (v - m) * k + b; }

__global__ void __launch_bounds__(256)
bn_relu_pool_fwd_kernel(const float* __restrict__ y, const float* __restrict__ sum, const float* __restrict__ sumsq,
                        int N, int H, int W, int C, float eps, float momentum, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ out, unsigned char* __restrict__ idx,
                        __nv_bfloat16* __restrict__ outs, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                        float* __restrict__ running_mean, float* __restrict__ running_var, const float* __restrict__ shift) {
    const int OH = (H + 1) / 2, OW = (W + 1) / 2, C4 = C >> 2;
    const long long R = (long long)N * H * W;
    const long long total = (long long)N * OH * OW * C4;
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = (int)(i0 % C4) * 4;                                 // fixed per thread (host: stride % C4 == 0)
    const float4 su = __ldg(reinterpret_cast<const float4*>(sum + c)), sq = __ldg(reinterpret_cast<const float4*>(sumsq + c));
    const float4 sh = __ldg(reinterpret_cast<const float4*>(shift ? shift + c : y + c));
    float4 m4, s4;
    if (blockIdx.x == 0 && threadIdx.x < C4) {
        double v0, v1, v2, v3;
        bn_moments(su.x, sq.x, sh.x, R, eps, m4.x, s4.x, v0);
        bn_moments(su.y, sq.y, sh.y, R, eps, m4.y, s4.y, v1);
        bn_moments(su.z, sq.z, sh.z, R, eps, m4.z, s4.z, v2);
        bn_moments(su.w, sq.w, sh.w, R, eps, m4.w, s4.w, v3);
        *reinterpret_cast<float4*>(mean_out + c) = m4;
        *reinterpret_cast<float4*>(rstd_out + c) = s4;
        if (running_mean) {
            const double ub = R > 1 ? (double)R / (double)(R - 1) : 1.0;
            const float mm[4] = {m4.x, m4.y, m4.z, m4.w};
            const double vv[4] = {v0, v1, v2, v3};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                running_mean[c + k] = (1.f - momentum) * running_mean[c + k] + momentum * mm[k];
                running_var[c + k] = (1.f - momentum) * running_var[c + k] + momentum * (float)(vv[k] * ub);
            }
        }
    }
    // every thread uses the SAME fp32 moments the backward kernels read back (mean_out / rstd_out are the double-rounded
    // values of block 0): recompute them identically here
    {
        double v;
        bn_moments(su.x, sq.x, sh.x, R, eps, m4.x, s4.x, v);
        bn_moments(su.y, sq.y, sh.y, R, eps, m4.y, s4.y, v);
        bn_moments(su.z, sq.z, sh.z, R, eps, m4.z, s4.z, v);
        bn_moments(su.w, sq.w, sh.w, R, eps, m4.w, s4.w, v);
    }
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 k4 = make_float4(s4.x * g4.x, s4.y * g4.y, s4.z * g4.z, s4.w * g4.w);
    for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long r = i / C4;
        const int ox = (int)(r % OW); r /= OW;
        const int oy = (int)(r % OH);
        const int n = (int)(r / OH);
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        unsigned char arg[4] = {0, 0, 0, 0};
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int yy = 2 * oy - 1 + t / 3, xx = 2 * ox - 1 + t % 3;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const float4 v = __ldg(reinterpret_cast<const float4*>(y + (((size_t)n * H + yy) * W + xx) * C + c));
            const float vv[4] = {fmaxf(bn_affine(v.x, m4.x, k4.x, b4.x), 0.f), fmaxf(bn_affine(v.y, m4.y, k4.y, b4.y), 0.f),
                                 fmaxf(bn_affine(v.z, m4.z, k4.z, b4.z), 0.f), fmaxf(bn_affine(v.w, m4.w, k4.w, b4.w), 0.f)};
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (vv[k] > best[k]) { best[k] = vv[k]; arg[k] = (unsigned char)t; }
        }
        const size_t pix = ((size_t)n * OH + oy) * OW + ox;
        *reinterpret_cast<float4*>(out + pix * C + c) = make_float4(best[0], best[1], best[2], best[3]);
        *reinterpret_cast<uchar4*>(idx + pix * C + c) = make_uchar4(arg[0], arg[1], arg[2], arg[3]);
        if (outs) {
            __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) split_bf16(best[k], hi[k], lo[k]);
            __nv_bfloat16* sp = outs + pix * (size_t)(((C + 31) >> 5) * 64) + (size_t)(c >> 5) * 64 + (c & 31);
            *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
            *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
        }
    }
}

// Backward, reduction pass, from the POOLED side: every pooled element routes its gradient to exactly one input position
// (its arg-max tap), so  sum_dz[c] = sum_pooled dout * [z(argmax) > 0]  and  sum_dzx[c] = sum_pooled dout * [..] * xhat(argmax)
// - a quarter of the elements of the input-side formulation, one gathered y value each.
__global__ void __launch_bounds__(BN_THREADS)
bn_pool_bwd_reduce_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ idx, const float* __restrict__ y,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                          const float* __restrict__ beta, int N, int H, int W, int C, float* __restrict__ sum_dz,
                          float* __restrict__ sum_dzx) {
    extern __shared__ float sm[];
    const int C4 = C >> 2, PL = BN_THREADS / C4, OH = (H + 1) / 2, OW = (W + 1) / 2;
    const int q = threadIdx.x % C4, pl = threadIdx.x / C4, c = q * 4;
    const int RP = N * OH * OW;                                        // pooled rows
    const int per = (RP + gridDim.x - 1) / gridDim.x;
    const int r0 = blockIdx.x * per, r1 = min(RP, r0 + per);
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + c)), s4 = __ldg(reinterpret_cast<const float4*>(rstd + c));
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c)), b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
    const float kk[4] = {s4.x * g4.x, s4.y * g4.y, s4.z * g4.z, s4.w * g4.w}, bb[4] = {b4.x, b4.y, b4.z, b4.w};
    float a[4] = {0.f, 0.f, 0.f, 0.f}, b[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = r0 + pl; r < r1; r += PL) {
        const int ox = r % OW, t1 = r / OW, oy = t1 % OH, n = t1 / OH;
        const float4 g = __ldg(reinterpret_cast<const float4*>(dout + (size_t)r * C + c));
        const uchar4 id = *reinterpret_cast<const uchar4*>(idx + (size_t)r * C + c);
        const float gg[4] = {g.x, g.y, g.z, g.w};
        const unsigned char tt[4] = {id.x, id.y, id.z, id.w};
        const long long yb = (((long long)n * H + (2 * oy - 1)) * W + (2 * ox - 1)) * C + c;  // tap (0, 0) of the window (may lie outside)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int ty = tt[k] / 3, tx = tt[k] - ty * 3;
            const float v = __ldg(y + (yb + ((long long)ty * W + tx) * C + k));       // the arg-max tap itself is always inside
            if (bn_affine(v, mm[k], kk[k], bb[k]) > 0.f) {
                a[k] += gg[k];
                b[k] += gg[k] * (v - mm[k]) * ss[k];
            }
        }
    }
    bn_reduce2_to_global(a, b, sm, C, C4, PL, sum_dz, sum_dzx);
}

// Backward, apply pass: dy = gamma * rstd * (dzr - sum_dz/R - xhat * sum_dzx/R) -> split32 (operand of the stem's dgrad /
// wgrad).  One thread per 2x2 block of input pixels and channel quad: the block's pixels look at the four pooling windows
// (Y, X), (Y, X+1), (Y+1, X), (Y+1, X+1) only, whose gradient / arg-max entries are loaded once for all four pixels
// (pixel (2Y+dy, 2X+dx) is tap (1+dy-2a, 1+dx-2b) of window (Y+a, X+b)).
__global__ void __launch_bounds__(BN_THREADS)
bn_pool_bwd_apply_kernel(const float* __restrict__ dout, const unsigned char* __restrict__ idx, const float* __restrict__ y,
                         const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                         const float* __restrict__ beta, const float* __restrict__ sum_dz, const float* __restrict__ sum_dzx,
                         int N, int H, int W, int C, __nv_bfloat16* __restrict__ dys) {
    const int C4 = C >> 2, OH = (H + 1) / 2, OW = (W + 1) / 2;
    const long long total = (long long)N * OH * OW * C4;
    const float inv_r = 1.f / (float)((long long)N * H * W);
    const long long i0 = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int c = (int)(i0 % C4) * 4;                                 // fixed per thread (host: stride % C4 == 0)
    const float4 m4 = __ldg(reinterpret_cast<const float4*>(mean + c)), s4 = __ldg(reinterpret_cast<const float4*>(rstd + c));
    const float4 g4 = __ldg(reinterpret_cast<const float4*>(gamma + c)), b4 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 k4 = make_float4(s4.x * g4.x, s4.y * g4.y, s4.z * g4.z, s4.w * g4.w);
    const float4 a4 = __ldg(reinterpret_cast<const float4*>(sum_dz + c)), q4 = __ldg(reinterpret_cast<const float4*>(sum_dzx + c));
    const int chunks64 = ((C + 31) >> 5) * 64;
    for (long long i = i0; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / C4);
        const int X = r % OW, t1 = r / OW, Y = t1 % OH, n = t1 / OH;
        float4 wg[2][2];
        uchar4 wi[2][2];
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) {
                wg[a][b] = make_float4(0.f, 0.f, 0.f, 0.f);
                wi[a][b] = make_uchar4(255, 255, 255, 255);
                if (Y + a < OH && X + b < OW) {
                    const size_t pix = ((size_t)n * OH + Y + a) * OW + X + b;
                    wg[a][b] = __ldg(reinterpret_cast<const float4*>(dout + pix * C + c));
                    wi[a][b] = *reinterpret_cast<const uchar4*>(idx + pix * C + c);
                }
            }
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int yy = 2 * Y + dy;
            if (yy >= H) continue;
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const int xx = 2 * X + dx;
                if (xx >= W) continue;
                float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int a = 0; a <= dy; ++a)
#pragma unroll
                    for (int b = 0; b <= dx; ++b) {
                        const unsigned char t = (unsigned char)((1 + dy - 2 * a) * 3 + (1 + dx - 2 * b));
                        if (wi[a][b].x == t) g.x += wg[a][b].x;
                        if (wi[a][b].y == t) g.y += wg[a][b].y;
                        if (wi[a][b].z == t) g.z += wg[a][b].z;
                        if (wi[a][b].w == t) g.w += wg[a][b].w;
                    }
                const size_t row = ((size_t)n * H + yy) * W + xx;
                const float4 v = __ldg(reinterpret_cast<const float4*>(y + row * C + c));
                g.x = bn_affine(v.x, m4.x, k4.x, b4.x) > 0.f ? g.x : 0.f;
                g.y = bn_affine(v.y, m4.y, k4.y, b4.y) > 0.f ? g.y : 0.f;
                g.z = bn_affine(v.z, m4.z, k4.z, b4.z) > 0.f ? g.z : 0.f;
                g.w = bn_affine(v.w, m4.w, k4.w, b4.w) > 0.f ? g.w : 0.f;
                float o[4];
                o[0] = g4.x * s4.x * (g.x - a4.x * inv_r - (v.x - m4.x) * s4.x * q4.x * inv_r);
                o[1] = g4.y * s4.y * (g.y - a4.y * inv_r - (v.y - m4.y) * s4.y * q4.y * inv_r);
                o[2] = g4.z * s4.z * (g.z - a4.z * inv_r - (v.z - m4.z) * s4.z * q4.z * inv_r);
                o[3] = g4.w * s4.w * (g.w - a4.w * inv_r - (v.w - m4.w) * s4.w * q4.w * inv_r);
                __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) split_bf16(o[k], hi[k], lo[k]);
                __nv_bfloat16* sp = dys + row * chunks64 + (c >> 5) * 64 + (c & 31);
                *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
                *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
            }
        }
    }
}

// grid of element-wise kernels whose threads keep one channel quad: a multiple of C/4 threads in total
static int bn_quad_blocks(long long total, int C4, int threads) {
    long long b = std::max<long long>(1, std::min<long long>((total + threads - 1) / threads, (long long)num_sms() * 32));
    while ((b * threads) % C4 != 0) ++b;
    return (int)b;
}

}  // namespace wgs

extern "C" int wgs_bn_pool_fwd(const float* y, const float* sum, const float* sumsq, const float* shift, int N, int H, int W,
                               int C, float eps, float momentum, const float* gamma, const float* beta, float* out, void* idx,
                               void* outs, float* mean, float* rstd, float* running_mean, float* running_var, void* stream) {
    WGS_REQUIRE(wgs::bn_ok(C) && N > 0 && H > 0 && W > 0, "bn_pool_fwd: C must be a multiple of 4 with C/4 dividing 256");
    WGS_REQUIRE(y && sum && sumsq && gamma && beta && out && idx && mean && rstd, "bn_pool_fwd: missing tensor");
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    wgs::bn_relu_pool_fwd_kernel<<<wgs::bn_quad_blocks(total, C / 4, 256), 256, 0, (cudaStream_t)stream>>>(
        y, sum, sumsq, N, H, W, C, eps, momentum, gamma, beta, out, (unsigned char*)idx, (__nv_bfloat16*)outs, mean, rstd,
        running_mean, running_var, shift);
    wgs::count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_pool_bwd_reduce(const float* dout, const void* idx, const float* y, const float* mean, const float* rstd,
                                      const float* gamma, const float* beta, int N, int H, int W, int C, float* sum_dz,
                                      float* sum_dzx, void* stream) {
    WGS_REQUIRE(wgs::bn_ok(C) && N > 0 && H > 0 && W > 0, "bn_pool_bwd_reduce: C must be a multiple of 4 with C/4 dividing 256");
    const size_t smem = 2 * (size_t)(wgs::BN_THREADS / (C / 4)) * C * sizeof(float);
    WGS_REQUIRE((long long)N * H * W < (1ll << 31), "bn_pool_bwd_reduce: too many rows");
    wgs::bn_pool_bwd_reduce_kernel<<<wgs::bn_red_blocks((long long)N * ((H + 1) / 2) * ((W + 1) / 2), C), wgs::BN_THREADS, smem, (cudaStream_t)stream>>>(
        dout, (const unsigned char*)idx, y, mean, rstd, gamma, beta, N, H, W, C, sum_dz, sum_dzx);
    wgs::count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_bn_pool_bwd_apply(const float* dout, const void* idx, const float* y, const float* mean, const float* rstd,
                                     const float* gamma, const float* beta, const float* sum_dz, const float* sum_dzx, int N,
                                     int H, int W, int C, void* dys, void* stream) {
    WGS_REQUIRE(wgs::bn_ok(C) && N > 0 && H > 0 && W > 0 && dys, "bn_pool_bwd_apply: C must be a multiple of 4 with C/4 dividing 256");
    WGS_REQUIRE((long long)N * H * W < (1ll << 31), "bn_pool_bwd_apply: too many rows");
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    wgs::bn_pool_bwd_apply_kernel<<<wgs::bn_quad_blocks(total, C / 4, wgs::BN_THREADS), wgs::BN_THREADS, 0, (cudaStream_t)stream>>>(
        dout, (const unsigned char*)idx, y, mean, rstd, gamma, beta, sum_dz, sum_dzx, N, H, W, C, (__nv_bfloat16*)dys);
    wgs::count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_maxpool3s2_fwd(const float* z, int N, int H, int W, int C, float* out, void* idx, void* outs,
                                  void* stream) {
    WGS_REQUIRE(N > 0 && H > 0 && W > 0 && C % 4 == 0, "maxpool3s2_fwd: C must be a multiple of 4");
    const long long total = (long long)N * ((H + 1) / 2) * ((W + 1) / 2) * (C / 4);
    wgs::maxpool3s2_fwd_kernel<<<wgs::bn_ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(
        z, N, H, W, C, out, (unsigned char*)idx, (__nv_bfloat16*)outs);
    wgs::count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_maxpool3s2_bwd(const float* dout, const void* idx, int N, int H, int W, int C, float* dz, void* stream) {
    WGS_REQUIRE(N > 0 && H > 0 && W > 0 && C % 4 == 0, "maxpool3s2_bwd: C must be a multiple of 4");
    const long long total = (long long)N * H * W * (C / 4);
    wgs::maxpool3s2_bwd_kernel<<<wgs::bn_ew_blocks(total), 256, 0, (cudaStream_t)stream>>>(
        dout, (const unsigned char*)idx, N, H, W, C, dz);
    wgs::count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
