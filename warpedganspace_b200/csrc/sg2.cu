// StyleGAN2 glue kernels around the tensor-core convs: small-batch linears (mapping network,
// per-layer modulation, demodulation coefficients), FIR blur + noise + bias + activation, ToRGB with
// the FIR-upsampled skip.  All HBM/latency-bound CUDA-core work.
#include "common.cuh"
#include "wgs_b200.h"
#include <string.h>
#include <stdlib.h>

namespace wgs {

int fir4_act_tma_launch(const float* y, float* out, int N, int Hin, int Win, int Hout, int Wout, int C, int pad0,
                        const float* taps4, const float* alpha, const float* beta, const float* noise, float noise_w, int act,
                        void* out_split, const float* split_scale, long long split_scale_ld, int out_from_n, void* stream);

// ------------------------------------------------------------------------------------------------
// out[b, o] (+)= mul[b, o] * epi( wscale * sum_i f(x[b, i], x2[b, i]) * W[o, i]  (+ bscale * bias[o]) )     B small, I % 4 == 0
//   f (in_mode): 0 x;  1 x^2;  2 x * dlrelu(x2) with dlrelu = sqrt2 (x2 > 0) or 0.2 sqrt2 (backward of 'fused_lrelu',
//               x2 = the layer's forward OUTPUT);  3 x * x2^3 (demodulation backward: dd * d^3)
//   epi: 0 linear, 1 sqrt(2)*lrelu(0.2), 2 rsqrt(. + eps)
// One warp per output feature: the weight row is streamed once with float4 loads (four in flight per lane) and reused
// for every batch row (EqualLinear, models/StyleGAN2/model.py:110-131; demod :194-195 restated as a linear on s^2).
// A launch covers `count` independent problems of the same batch size (blockIdx.x -> problem through block_start).
constexpr int LIN_BT = 8;
constexpr int LIN_WARPS = 4;

struct LinearGroup {
    wgs_linear_problem prob[WGS_MAX_LINEAR_GROUP];
    int block_start[WGS_MAX_LINEAR_GROUP + 1];
    int count, B;
};

__device__ __forceinline__ float4 lin_in(const wgs_linear_problem& q, int b, int i) {
    float4 x4 = __ldg(reinterpret_cast<const float4*>(q.x + (size_t)b * q.x_ld + i));
    if (q.in_mode == 1) {
        x4.x *= x4.x; x4.y *= x4.y; x4.z *= x4.z; x4.w *= x4.w;
    } else if (q.in_mode == 2) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(q.x2 + (size_t)b * q.x2_ld + i));
        const float up = 1.41421356237309515f, dn = 0.2f * 1.41421356237309515f;
        x4.x *= h.x > 0.f ? up : dn; x4.y *= h.y > 0.f ? up : dn; x4.z *= h.z > 0.f ? up : dn; x4.w *= h.w > 0.f ? up : dn;
    } else if (q.in_mode == 3) {
        const float4 h = __ldg(reinterpret_cast<const float4*>(q.x2 + (size_t)b * q.x2_ld + i));
        x4.x *= h.x * h.x * h.x; x4.y *= h.y * h.y * h.y; x4.z *= h.z * h.z * h.z; x4.w *= h.w * h.w * h.w;
    }
    return x4;
}

__global__ void __launch_bounds__(LIN_WARPS * 32)
linear_group_kernel(const __grid_constant__ LinearGroup g) {
    int pi = 0;
    while (pi + 1 < g.count && (int)blockIdx.x >= g.block_start[pi + 1]) ++pi;
    const wgs_linear_problem& q = g.prob[pi];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int o = ((int)blockIdx.x - g.block_start[pi]) * LIN_WARPS + warp;
    if (o >= q.O) return;
    const float* wrow = q.W + (size_t)o * q.w_ld;
    const int B = g.B, I = q.I;
    for (int b0 = 0; b0 < B; b0 += LIN_BT) {
        float acc[LIN_BT];
#pragma unroll
        for (int t = 0; t < LIN_BT; ++t) acc[t] = 0.f;
        for (int i0 = lane * 4; i0 < I; i0 += 512) {
            float4 w4[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j * 128;
                w4[j] = i < I ? __ldg(reinterpret_cast<const float4*>(wrow + i)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int t = 0; t < LIN_BT; ++t) {
                if (b0 + t < B) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int i = i0 + j * 128;
                        if (i < I) {
                            const float4 x4 = lin_in(q, b0 + t, i);
                            acc[t] += w4[j].x * x4.x + w4[j].y * x4.y + w4[j].z * x4.z + w4[j].w * x4.w;
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int t = 0; t < LIN_BT; ++t) acc[t] = warp_sum(acc[t]);
        if (lane < LIN_BT && b0 + lane < B) {
            float v = 0.f;
#pragma unroll
            for (int t = 0; t < LIN_BT; ++t) if (t == lane) v = acc[t];
            v *= q.wscale;
            if (q.epi == 2) {
                v = rsqrtf(v + q.eps);
            } else {
                if (q.bias) v += q.bscale * __ldg(q.bias + o);
                if (q.epi == 1) v = 1.41421356237309515f * (v > 0.f ? v : 0.2f * v);
            }
            if (q.mul) v *= __ldg(q.mul + (size_t)(b0 + lane) * q.mul_ld + o);
            float* dst = q.out + (size_t)(b0 + lane) * q.out_ld + o;
            if (q.accumulate == 2) atomicAdd(dst, v);          // several problems of one launch share this output (split-I)
            else *dst = q.accumulate ? (*dst + v) : v;
        }
    }
}

// x * rsqrt(mean(x^2) + 1e-8) per row (PixelNorm on latents, model.py:9-15)
__global__ void pixelnorm_rows_kernel(const float* __restrict__ x, float* __restrict__ out, int d) {
    __shared__ float red[32];
    const float* r = x + (size_t)blockIdx.x * d;
    float s = 0.f;
    for (int i = threadIdx.x; i < d; i += blockDim.x) s += r[i] * r[i];
    s = block_sum(s, red);
    const float k = rsqrtf(s / d + 1e-8f);
    for (int i = threadIdx.x; i < d; i += blockDim.x) out[(size_t)blockIdx.x * d + i] = r[i] * k;
}

// ------------------------------------------------------------------------------------------------
// Separable 4-tap FIR (upfirdn2d with up = down = 1, op/upfirdn2d_kernel.cu:52-137) fused with the
// StyledConv tail: out = act(alpha[n,c] * fir(y)[Y,X,c] + noise_w * noise[Y,X] + beta[c]).
// y: [N, Hin, Win, C] NHWC, out: [N, Hout, Wout, C]; fir(y)[Y,X] = sum_ij kf[i] kf[j] y[Y+i-pad0, X+j-pad0].
// FIR_ROWS = output rows per thread: 8 = 5.5 loads per output at 126 registers (2 blocks / SM); 4 = 7 loads per output
// at <= 64 registers (4 blocks / SM) - the kernel is latency-bound on its loads (55 % long-scoreboard stalls at 3.9
// warps per scheduler, profiles/r01_step_ncu.md), so occupancy buys more than the extra L1 traffic costs.
template <int FIR_ROWS, int MIN_BLOCKS>
__global__ void __launch_bounds__(256, MIN_BLOCKS)
fir4_act_kernel(const float* __restrict__ y, float* __restrict__ out, int N, int Hin, int Win, int Hout, int Wout,
                int C, int pad0, float k0, float k1, float k2, float k3, const float* __restrict__ alpha,
                const float* __restrict__ beta, const float* __restrict__ noise, float noise_w, int act,
                __nv_bfloat16* __restrict__ out_split, const float* __restrict__ split_scale, long long split_scale_ld,
                int out_from_n) {
    // each thread produces an 8-row strip of one channel quad: (8 + 3) x 4 float4 loads for 8 outputs (5.5 loads
    // per output instead of 16); horizontal pass first, then the vertical combination in registers.  Interior
    // strips take a path without per-load bounds checks (the kernel is instruction-bound, not DRAM-bound:
    // profiles/r01_misc_ncu.md).
    const int c4n = C >> 2;
    const int yb_n = (Hout + FIR_ROWS - 1) / FIR_ROWS;
    const long long total = (long long)N * yb_n * Wout * c4n;
    const float kf[4] = {k3, k2, k1, k0};                       // correlation with the flipped kernel
    const long long row_stride = (long long)Win * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % c4n) * 4;
        long long r = i / c4n;
        const int X = (int)(r % Wout); r /= Wout;
        const int Y0 = (int)(r % yb_n) * FIR_ROWS;
        const int n = (int)(r / yb_n);
        const int yy0 = Y0 - pad0, xx0 = X - pad0;
        const float* base = y + ((long long)n * Hin + yy0) * row_stride + (long long)xx0 * C + c;
        // noise of the 8 output rows up front: loading it inside the row loop serialises 8 global-load latencies
        // behind the stores (47 % of this kernel's stall samples, profiles/r01_misc_ncu.md)
        float nzv[FIR_ROWS];
#pragma unroll
        for (int oy = 0; oy < FIR_ROWS; ++oy)
            nzv[oy] = (noise && Y0 + oy < Hout) ? noise_w * __ldg(noise + (size_t)(Y0 + oy) * Wout + X) : 0.f;
        float4 h[FIR_ROWS + 3];
        if (yy0 >= 0 && yy0 + FIR_ROWS + 2 < Hin && xx0 >= 0 && xx0 + 3 < Win) {
#pragma unroll
            for (int a = 0; a < FIR_ROWS + 3; ++a) {
                const float* rowp = base + a * row_stride;
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(rowp));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(rowp + C));
                const float4 v2 = __ldg(reinterpret_cast<const float4*>(rowp + 2 * C));
                const float4 v3 = __ldg(reinterpret_cast<const float4*>(rowp + 3 * C));
                h[a].x = kf[0] * v0.x + kf[1] * v1.x + kf[2] * v2.x + kf[3] * v3.x;
                h[a].y = kf[0] * v0.y + kf[1] * v1.y + kf[2] * v2.y + kf[3] * v3.y;
                h[a].z = kf[0] * v0.z + kf[1] * v1.z + kf[2] * v2.z + kf[3] * v3.z;
                h[a].w = kf[0] * v0.w + kf[1] * v1.w + kf[2] * v2.w + kf[3] * v3.w;
            }
        } else {
#pragma unroll
            for (int a = 0; a < FIR_ROWS + 3; ++a) {
                const int yy = yy0 + a;
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (yy >= 0 && yy < Hin) {
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int xx = xx0 + b;
                        if (xx >= 0 && xx < Win) {
                            const float4 v = __ldg(reinterpret_cast<const float4*>(base + a * row_stride + (long long)b * C));
                            t.x += kf[b] * v.x; t.y += kf[b] * v.y; t.z += kf[b] * v.z; t.w += kf[b] * v.w;
                        }
                    }
                }
                h[a] = t;
            }
        }
        float al[4] = {1.f, 1.f, 1.f, 1.f}, be[4] = {0.f, 0.f, 0.f, 0.f}, sc[4] = {1.f, 1.f, 1.f, 1.f};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (alpha) al[k] = __ldg(alpha + (size_t)n * C + c + k);
            if (beta) be[k] = __ldg(beta + c + k);
            if (split_scale) sc[k] = __ldg(split_scale + (size_t)n * split_scale_ld + c + k);
        }
        const bool f32 = out && n >= out_from_n;
        const int chunk_stride = ((C + 31) >> 5) * 64;
#pragma unroll
        for (int oy = 0; oy < FIR_ROWS; ++oy) {
            const int Y = Y0 + oy;
            if (Y >= Hout) break;
            float v4[4];
            v4[0] = kf[0] * h[oy].x + kf[1] * h[oy + 1].x + kf[2] * h[oy + 2].x + kf[3] * h[oy + 3].x;
            v4[1] = kf[0] * h[oy].y + kf[1] * h[oy + 1].y + kf[2] * h[oy + 2].y + kf[3] * h[oy + 3].y;
            v4[2] = kf[0] * h[oy].z + kf[1] * h[oy + 1].z + kf[2] * h[oy + 2].z + kf[3] * h[oy + 3].z;
            v4[3] = kf[0] * h[oy].w + kf[1] * h[oy + 1].w + kf[2] * h[oy + 2].w + kf[3] * h[oy + 3].w;
            const float nz = nzv[oy];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t = v4[k] * al[k] + nz + be[k];
                if (act == 3) t = 1.41421356237309515f * (t > 0.f ? t : 0.2f * t);
                else if (act == 2) t = t > 0.f ? t : 0.2f * t;
                else if (act == 1) t = t > 0.f ? t : 0.f;
                v4[k] = t;
            }
            const size_t pix = ((size_t)n * Hout + Y) * Wout + X;
            if (f32) *reinterpret_cast<float4*>(out + pix * C + c) = make_float4(v4[0], v4[1], v4[2], v4[3]);
            if (out_split) {
                __align__(8) __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) split_bf16(v4[k] * sc[k], hi[k], lo[k]);
                __nv_bfloat16* sp = out_split + pix * (size_t)chunk_stride + (size_t)(c >> 5) * 64 + (c & 31);
                *reinterpret_cast<uint2*>(sp) = *reinterpret_cast<const uint2*>(hi);
                *reinterpret_cast<uint2*>(sp + 32) = *reinterpret_cast<const uint2*>(lo);
            }
        }
    }
}

// (A shared-memory tiled variant - window loaded once per block, horizontal pass out of shared memory - was measured
// and dropped: the LDS + STS + barrier instructions it adds outweigh the L1 hits it removes, 202 vs 210 pairs/s.)

// rgb[n,y,x,:] = bias + up2(prev)[n,y,x,:]  — initialises the ToRGB accumulator that the conv epilogue adds into.
// One thread per FOUR consecutive output pixels of a row (W % 4 == 0): the 2 x 4 window of the half-resolution skip image
// is read once (24 loads instead of 48) and the 12 outputs leave as three 128-bit stores.
__global__ void rgb_init_kernel(const float* __restrict__ bias, const float* __restrict__ prev, float* __restrict__ rgb,
                                int N, int H, int W, float k0, float k1, float k2, float k3) {
    const int W4 = W >> 2;
    const long long total = (long long)N * H * W4;
    const int h2 = H >> 1, w2 = W >> 1;
    const float b0 = __ldg(bias), b1 = __ldg(bias + 1), b2 = __ldg(bias + 2);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int x0 = (int)(i % W4) * 4;
        const int yy = (int)((i / W4) % H);
        const int n = (int)(i / ((long long)W4 * H));
        float o[12] = {b0, b1, b2, b0, b1, b2, b0, b1, b2, b0, b1, b2};
        if (prev) {
            int ya, yb;
            float wya, wyb;
            if (yy & 1) { ya = (yy - 1) >> 1; yb = (yy + 1) >> 1; wya = k2; wyb = k0; }
            else        { ya = (yy >> 1) - 1; yb = yy >> 1;       wya = k3; wyb = k1; }
            const int ys[2] = {ya, yb};
            const float wy[2] = {wya, wyb};
            const int xb = (x0 >> 1) - 1;                       // window columns xb .. xb + 3
            float col[4][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // vertically filtered
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                if (ys[a] < 0 || ys[a] >= h2) continue;
                const float* row = prev + ((size_t)n * h2 + ys[a]) * w2 * 3;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int xc = xb + j;
                    if (xc < 0 || xc >= w2) continue;
                    col[j][0] += wy[a] * __ldg(row + xc * 3); col[j][1] += wy[a] * __ldg(row + xc * 3 + 1);
                    col[j][2] += wy[a] * __ldg(row + xc * 3 + 2);
                }
            }
            // output x0 + p: even p -> columns (p/2, p/2 + 1) of the window with (k3, k1); odd p -> ((p+1)/2, (p+1)/2 + 1) with (k2, k0)
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int ja = (p & 1) ? (p + 1) / 2 : p / 2;
                const float wa = (p & 1) ? k2 : k3, wb = (p & 1) ? k0 : k1;
#pragma unroll
                for (int c = 0; c < 3; ++c) o[p * 3 + c] += wa * col[ja][c] + wb * col[ja + 1][c];
            }
        }
        float4* d = reinterpret_cast<float4*>(rgb + (((size_t)n * H + yy) * W + x0) * 3);
        d[0] = make_float4(o[0], o[1], o[2], o[3]);
        d[1] = make_float4(o[4], o[5], o[6], o[7]);
        d[2] = make_float4(o[8], o[9], o[10], o[11]);
    }
}

// wm[n][o][c] = wscale * W[o][c] * s[n][c]   (modulated ToRGB weights consumed by the conv epilogue)
__global__ void rgb_weights_kernel(const float* __restrict__ W, const float* __restrict__ s, long long s_ld,
                                   float* __restrict__ wm, int N, int C, float wscale) {
    const int total = N * 3 * C;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int c = i % C, o = (i / C) % 3, n = i / (3 * C);
        wm[i] = wscale * __ldg(W + o * C + c) * __ldg(s + (size_t)n * s_ld + c);
    }
}

// ------------------------------------------------------------------------------------------------
// ToRGB (model.py:270-282): 1x1 modulated conv without demodulation + bias + FIR-upsampled skip.
//   rgb[n,y,x,o] = sum_c a[n,y,x,c] * (wscale * W[o,c] * s[n,c]) + bias[o] + up2(prev)[n,y,x,o]
// a: [N,H,W,C]; s: [N, C] (row stride s_ld); W: [3, C]; prev: [N,H/2,W/2,3] or NULL; rgb: [N,H,W,3].
// up2 = upfirdn2d(up=2, taps*2 per axis, pad (2,1)) -> per axis: even y: k3*p[y/2-1] + k1*p[y/2];
// odd y: k2*p[(y-1)/2] + k0*p[(y+1)/2]  (taps (k0..k3) already include the factor 2).
// LPP lanes cooperate on one pixel (LPP = min(32, C/4)); 32/LPP pixels per warp.
__global__ void __launch_bounds__(256)
torgb_kernel(const float* __restrict__ a, const float* __restrict__ s, long long s_ld, const float* __restrict__ W,
             const float* __restrict__ bias, const float* __restrict__ prev, float* __restrict__ rgb, int N, int H,
             int Wd, int C, float wscale, float k0, float k1, float k2, float k3, int lpp) {
    extern __shared__ float wm[];                               // [3][C] modulated weights of image n
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < 3 * C; i += blockDim.x) {
        const int c = i % C;
        wm[i] = wscale * __ldg(W + i) * __ldg(s + (size_t)n * s_ld + c);
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int sub = lane % lpp, pix_in_warp = lane / lpp, ppw = 32 / lpp;
    const long long warps_total = (long long)gridDim.x * (blockDim.x >> 5);
    const long long warp_id = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long npix = (long long)H * Wd;
    const int h2 = H >> 1, w2 = Wd >> 1;
    for (long long p0 = warp_id * ppw; p0 < npix; p0 += warps_total * ppw) {
        const long long pix = p0 + pix_in_warp;
        float r0 = 0.f, r1 = 0.f, r2 = 0.f;
        if (pix < npix) {
            const float* ap = a + ((size_t)n * npix + pix) * C;
            for (int c = sub * 4; c < C; c += lpp * 4) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(ap + c));
                const float4 w0 = *reinterpret_cast<const float4*>(wm + c);
                const float4 w1 = *reinterpret_cast<const float4*>(wm + C + c);
                const float4 w2v = *reinterpret_cast<const float4*>(wm + 2 * C + c);
                r0 += v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
                r1 += v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
                r2 += v.x * w2v.x + v.y * w2v.y + v.z * w2v.z + v.w * w2v.w;
            }
        }
        for (int o = lpp >> 1; o > 0; o >>= 1) {
            r0 += __shfl_xor_sync(0xffffffffu, r0, o);
            r1 += __shfl_xor_sync(0xffffffffu, r1, o);
            r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        }
        if (sub == 0 && pix < npix) {
            float o3[3] = {r0 + __ldg(bias), r1 + __ldg(bias + 1), r2 + __ldg(bias + 2)};
            if (prev) {
                const int yy = (int)(pix / Wd), xx = (int)(pix % Wd);
                int ya, yb, xa, xb;
                float wya, wyb, wxa, wxb;
                if (yy & 1) { ya = (yy - 1) >> 1; yb = (yy + 1) >> 1; wya = k2; wyb = k0; }
                else        { ya = (yy >> 1) - 1; yb = yy >> 1;       wya = k3; wyb = k1; }
                if (xx & 1) { xa = (xx - 1) >> 1; xb = (xx + 1) >> 1; wxa = k2; wxb = k0; }
                else        { xa = (xx >> 1) - 1; xb = xx >> 1;       wxa = k3; wxb = k1; }
                const int ys[2] = {ya, yb}, xs[2] = {xa, xb};
                const float wy[2] = {wya, wyb}, wx[2] = {wxa, wxb};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (ys[i] < 0 || ys[i] >= h2) continue;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (xs[j] < 0 || xs[j] >= w2) continue;
                        const float* pp = prev + (((size_t)n * h2 + ys[i]) * w2 + xs[j]) * 3;
                        const float wgt = wy[i] * wx[j];
                        o3[0] += wgt * __ldg(pp); o3[1] += wgt * __ldg(pp + 1); o3[2] += wgt * __ldg(pp + 2);
                    }
                }
            }
            float* dst = rgb + ((size_t)n * npix + pix) * 3;
            dst[0] = o3[0]; dst[1] = o3[1]; dst[2] = o3[2];
        }
    }
}

}  // namespace wgs

using namespace wgs;

static int launch_linear_group(const wgs_linear_problem* problems, int count, int B, void* stream) {
    WGS_REQUIRE(problems != nullptr && count >= 1 && count <= WGS_MAX_LINEAR_GROUP, "linear_group: 1..WGS_MAX_LINEAR_GROUP problems");
    WGS_REQUIRE(B >= 0, "linear_group: bad batch");
    if (B == 0) return 0;
    LinearGroup g;
    memset(&g, 0, sizeof(g));
    g.count = count; g.B = B;
    int blocks = 0;
    for (int i = 0; i < count; ++i) {
        const wgs_linear_problem& q = problems[i];
        WGS_REQUIRE(q.I > 0 && q.O > 0, "linear: bad sizes");
        WGS_REQUIRE(q.I % 4 == 0 && q.x_ld % 4 == 0 && q.w_ld % 4 == 0, "linear: I, x_ld, w_ld must be multiples of 4");
        WGS_REQUIRE(((uintptr_t)q.x & 15) == 0 && ((uintptr_t)q.W & 15) == 0, "linear: operands must be 16-byte aligned");
        WGS_REQUIRE(q.in_mode >= 0 && q.in_mode <= 3 && q.epi >= 0 && q.epi <= 2, "linear: bad in_mode / epi");
        WGS_REQUIRE(q.in_mode < 2 || (q.x2 != nullptr && q.x2_ld % 4 == 0 && ((uintptr_t)q.x2 & 15) == 0),
                    "linear: in_mode 2/3 needs a 16-byte aligned x2 with x2_ld % 4 == 0");
        WGS_REQUIRE(q.x != nullptr && q.W != nullptr && q.out != nullptr, "linear: null operand");
        g.prob[i] = q;
        g.block_start[i] = blocks;
        blocks += ceil_div(q.O, LIN_WARPS);
    }
    g.block_start[count] = blocks;
    linear_group_kernel<<<blocks, LIN_WARPS * 32, 0, (cudaStream_t)stream>>>(g);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_linear_problem_size(void) { return (int)sizeof(wgs_linear_problem); }

extern "C" int wgs_linear_group(const wgs_linear_problem* problems, int count, int B, void* stream) {
    return launch_linear_group(problems, count, B, stream);
}

extern "C" int wgs_linear_small(const float* x, long long x_ld, const float* W, long long w_ld, const float* bias,
                                float* out, long long out_ld, int B, int I, int O, float wscale, float bscale,
                                int in_square, int epi, float eps, int accumulate, void* stream) {
    wgs_linear_problem q;
    memset(&q, 0, sizeof(q));
    q.x = x; q.x_ld = x_ld; q.W = W; q.w_ld = w_ld; q.bias = bias; q.out = out; q.out_ld = out_ld;
    q.I = I; q.O = O; q.wscale = wscale; q.bscale = bscale; q.eps = eps;
    q.in_mode = in_square ? 1 : 0; q.epi = epi; q.accumulate = accumulate;
    return launch_linear_group(&q, 1, B, stream);
}

extern "C" int wgs_pixelnorm_rows(const float* x, float* out, int B, int d, void* stream) {
    WGS_REQUIRE(B >= 0 && d > 0, "pixelnorm_rows: bad sizes");
    if (B == 0) return 0;
    pixelnorm_rows_kernel<<<B, 128, 0, (cudaStream_t)stream>>>(x, out, d);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_fir4_act(const float* y, float* out, int N, int Hin, int Win, int Hout, int Wout, int C, int pad0,
                            const float* taps4, const float* alpha, const float* beta, const float* noise,
                            float noise_w, int act, void* out_split, const float* split_scale, long long split_scale_ld,
                            int out_from_n, void* stream) {
    WGS_REQUIRE(N > 0 && C > 0 && C % 4 == 0, "fir4_act: channels must be a multiple of 4");
    WGS_REQUIRE(taps4 != nullptr, "fir4_act: taps4 is a HOST pointer to 4 floats");
    {   // large layers: TMA-fed shared-memory variant (fir_tma.cu); returns 0 when the shape is not its business
        const int took = wgs::fir4_act_tma_launch(y, out, N, Hin, Win, Hout, Wout, C, pad0, taps4, alpha, beta, noise, noise_w, act,
                                                  out_split, split_scale, split_scale_ld, out_from_n, stream);
        if (took != 0) return took > 0 ? 0 : took;
    }
    static int rows = 0;
    if (!rows) {
        const char* e = getenv("WGS_FIR_ROWS");
        rows = (e && atoi(e) == 4) ? 4 : 8;                  // measured: 8 rows 196.4 pairs/s, 4 rows 194.4
    }
    const long long total = (long long)N * ((Hout + rows - 1) / rows) * Wout * (C / 4);
    const int blocks = (int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 48);
    if (rows == 8)
        fir4_act_kernel<8, 2><<<blocks, 256, 0, (cudaStream_t)stream>>>(y, out, N, Hin, Win, Hout, Wout, C, pad0, taps4[0], taps4[1],
                                                                        taps4[2], taps4[3], alpha, beta, noise, noise_w, act,
                                                                        (__nv_bfloat16*)out_split, split_scale, split_scale_ld, out_from_n);
    else
        fir4_act_kernel<4, 3><<<blocks, 256, 0, (cudaStream_t)stream>>>(y, out, N, Hin, Win, Hout, Wout, C, pad0, taps4[0], taps4[1],
                                                                        taps4[2], taps4[3], alpha, beta, noise, noise_w, act,
                                                                        (__nv_bfloat16*)out_split, split_scale, split_scale_ld, out_from_n);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_torgb(const float* a, const float* s, long long s_ld, const float* W, const float* bias,
                             const float* prev, float* rgb, int N, int H, int Wd, int C, float wscale,
                             const float* taps4, void* stream) {
    WGS_REQUIRE(N > 0 && C >= 4 && C % 4 == 0, "torgb: channels must be a multiple of 4");
    WGS_REQUIRE(!prev || (H % 2 == 0 && Wd % 2 == 0), "torgb: skip needs even output size");
    int lpp = 32;
    while (lpp > 1 && lpp * 4 > C) lpp >>= 1;
    const long long npix = (long long)H * Wd;
    const int ppw = 32 / lpp;
    const long long warps = (npix + ppw - 1) / ppw;
    const int blocks = (int)std::max<long long>(1, std::min<long long>((warps + 7) / 8, (long long)num_sms() * 8 / std::max(1, N) + 1));
    const float k0 = taps4 ? taps4[0] : 0.25f, k1 = taps4 ? taps4[1] : 0.75f, k2 = taps4 ? taps4[2] : 0.75f,
                k3 = taps4 ? taps4[3] : 0.25f;
    torgb_kernel<<<dim3(blocks, N), 256, (size_t)3 * C * sizeof(float), (cudaStream_t)stream>>>(
        a, s, s_ld, W, bias, prev, rgb, N, H, Wd, C, wscale, k0, k1, k2, k3, lpp);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_rgb_init(const float* bias, const float* prev, float* rgb, int N, int H, int W,
                                const float* h_taps4, void* stream) {
    WGS_REQUIRE(N > 0 && H > 0 && W > 0, "rgb_init: bad sizes");
    WGS_REQUIRE(!prev || (H % 2 == 0 && W % 2 == 0), "rgb_init: skip needs even output size");
    const float k0 = h_taps4 ? h_taps4[0] : 0.25f, k1 = h_taps4 ? h_taps4[1] : 0.75f, k2 = h_taps4 ? h_taps4[2] : 0.75f,
                k3 = h_taps4 ? h_taps4[3] : 0.25f;
    WGS_REQUIRE(W % 4 == 0 && (reinterpret_cast<uintptr_t>(rgb) & 15) == 0, "rgb_init: width must be a multiple of 4, output 16-byte aligned");
    const long long total = (long long)N * H * (W / 4);
    rgb_init_kernel<<<(int)std::min<long long>((total + 255) / 256, (long long)num_sms() * 16), 256, 0, (cudaStream_t)stream>>>(
        bias, prev, rgb, N, H, W, k0, k1, k2, k3);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}

extern "C" int wgs_sg2_rgb_weights(const float* W, const float* s, long long s_ld, float* wm, int N, int C, float wscale,
                                   void* stream) {
    WGS_REQUIRE(N > 0 && C > 0, "rgb_weights: bad sizes");
    rgb_weights_kernel<<<ceil_div(N * 3 * C, 256), 256, 0, (cudaStream_t)stream>>>(W, s, s_ld, wm, N, C, wscale);
    count_launch();
    WGS_LAUNCH_CHECK();
    return 0;
}
