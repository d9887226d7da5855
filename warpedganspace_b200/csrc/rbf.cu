// SupportSets RBF warp: forward, backward, and the batched traversal chains.
//
// Replaces the ~12 ATen launches of /root/reference/lib/support_sets.py:81-101 (three one-hot
// "gather" GEMMs over the whole [K, 2D*d] matrix plus broadcast / norm / exp / sum / normalise
// kernels materialising [B, 2D, d] twice) with one kernel: one CTA per latent, the selected
// support set is row-gathered by index, each warp streams whole support vectors with 128-bit
// coalesced loads, squared distances are warp-shuffle reductions, and the weighted sum is kept in
// registers until one cross-warp reduction at the end.
//
// HBM-bound: algorithmic bytes per latent = (2D*d + d + 2D + 1)*4 read + d*4 written.
#include "common.cuh"

namespace wgs {

constexpr int RBF_THREADS = 256;
constexpr int RBF_WARPS = RBF_THREADS / 32;

template <int NV>
struct LaneVec { float4 v[NV]; };

// lane owns elements [(i*32 + lane)*4, +4), i < NV, of a d-vector.  d % 4 == 0 (every row 16-byte aligned): one 128-bit
// load per slot; otherwise (BigGAN-256's dim_z = 119, models/BigGAN/BigGAN.py:103-108) four guarded scalar loads, zeros
// beyond d.  Shared-memory rows are padded to dp = round_up(d, 4) so the float4 staging below stays aligned either way.
template <int NV>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, int d, int lane, LaneVec<NV>& r) {
    if ((d & 3) == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = (i * 32 + lane) * 4;
            r.v[i] = (e < d) ? __ldg(reinterpret_cast<const float4*>(p + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = (i * 32 + lane) * 4;
            r.v[i].x = (e + 0 < d) ? __ldg(p + e + 0) : 0.f;
            r.v[i].y = (e + 1 < d) ? __ldg(p + e + 1) : 0.f;
            r.v[i].z = (e + 2 < d) ? __ldg(p + e + 2) : 0.f;
            r.v[i].w = (e + 3 < d) ? __ldg(p + e + 3) : 0.f;
        }
    }
}

__device__ __forceinline__ int pad4(int d) { return (d + 3) & ~3; }

// store one lane slot to a global d-vector (vector store when rows are 16-byte aligned)
__device__ __forceinline__ void store_slot(float* __restrict__ p, int d, int e, const float4& v) {
    if ((d & 3) == 0) {
        *reinterpret_cast<float4*>(p + e) = v;
    } else {
        if (e + 0 < d) p[e + 0] = v.x;
        if (e + 1 < d) p[e + 1] = v.y;
        if (e + 2 < d) p[e + 2] = v.z;
        if (e + 3 < d) p[e + 3] = v.w;
    }
}

// Pass shared by forward and backward: acc = sum_j w_j (z - s_j) over this warp's vectors.
template <int NV>
__device__ __forceinline__ void rbf_accumulate(const float* __restrict__ set, const float* __restrict__ alpha,
                                               float gamma, const LaneVec<NV>& z, int n_vec, int d, int warp,
                                               int lane, LaneVec<NV>& acc) {
#pragma unroll
    for (int i = 0; i < NV; ++i) acc.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int j = warp; j < n_vec; j += RBF_WARPS) {
        LaneVec<NV> s;
        load_vec<NV>(set + (size_t)j * d, d, lane, s);
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            s.v[i].x = z.v[i].x - s.v[i].x; s.v[i].y = z.v[i].y - s.v[i].y;
            s.v[i].z = z.v[i].z - s.v[i].z; s.v[i].w = z.v[i].w - s.v[i].w;
            q += s.v[i].x * s.v[i].x + s.v[i].y * s.v[i].y + s.v[i].z * s.v[i].z + s.v[i].w * s.v[i].w;
        }
        q = warp_sum(q);
        const float w = __ldg(alpha + j) * gamma * expf(-gamma * q);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            acc.v[i].x += w * s.v[i].x; acc.v[i].y += w * s.v[i].y;
            acc.v[i].z += w * s.v[i].z; acc.v[i].w += w * s.v[i].w;
        }
    }
}

// Cross-warp reduction of the per-warp partial sums; result (already * -2) lands in `g` for every
// thread's own float4 slots, `sm` is [RBF_WARPS][d] floats. Returns ||g||^2 via block_sum.
template <int NV>
__device__ __forceinline__ float rbf_reduce(LaneVec<NV>& acc, float* sm, float* red, int d, int warp, int lane) {
    const int dp = pad4(d);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int e = (i * 32 + lane) * 4;
        if (e < d) *reinterpret_cast<float4*>(sm + (size_t)warp * dp + e) = acc.v[i];
    }
    __syncthreads();
    float nrm = 0.f;
    for (int e = threadIdx.x; e < d; e += RBF_THREADS) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < RBF_WARPS; ++w) t += sm[(size_t)w * dp + e];
        t *= -2.f;
        sm[e] = t;                      // row 0 of sm now holds g (each e is touched by exactly one thread)
        nrm += t * t;
    }
    return block_sum(nrm, red);        // contains the __syncthreads that publishes sm[0..d)
}

template <int NV>
__global__ void __launch_bounds__(RBF_THREADS)
rbf_forward_kernel(const float* __restrict__ support_sets, const float* __restrict__ alphas,
                   const float* __restrict__ loggamma, float fixed_gamma, const long long* __restrict__ idx,
                   const float* __restrict__ z, const float* __restrict__ mag, float* __restrict__ out,
                   int K, int n_vec, int d) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long k = idx[b];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    const float gamma = loggamma ? expf(__ldg(loggamma + k)) : fixed_gamma;
    LaneVec<NV> zv, acc;
    load_vec<NV>(z + (size_t)b * d, d, lane, zv);
    rbf_accumulate<NV>(support_sets + (size_t)k * n_vec * d, alphas + (size_t)k * n_vec, gamma, zv, n_vec, d,
                       warp, lane, acc);
    const float n2 = rbf_reduce<NV>(acc, sm, red, d, warp, lane);
    // unit vector first, magnitude second: bit-identical to `mag * S(mask, z)` formed outside the kernel
    // (lib/trainer.py:236), so a module-level loop and the fused engine see the same shift.  No epsilon, as the reference.
    const float inv_n = 1.f / sqrtf(n2);
    const float m = mag ? __ldg(mag + b) : 1.f;
    for (int e = threadIdx.x; e < d; e += RBF_THREADS) out[(size_t)b * d + e] = (sm[e] * inv_n) * m;
}

// Backward of out = m * g/||g||,  g = -2 sum_j w_j D_j,  D_j = z - s_j,  w_j = a_j*gamma*exp(-gamma |D_j|^2).
// With dg = (du - u (u.du)) / ||g||, du = m*dout, t_j = dg.D_j :
//   dD_j = w_j (-2 dg + 4 gamma t_j D_j);  ds_j = -dD_j;  dz = sum_j dD_j
//   dloggamma_k += sum_j (-2 t_j) w_j (1 - gamma q_j);  dalpha_j += (-2 t_j) gamma exp(-gamma q_j)
template <int NV>
__global__ void __launch_bounds__(RBF_THREADS)
rbf_backward_kernel(const float* __restrict__ support_sets, const float* __restrict__ alphas,
                    const float* __restrict__ loggamma, float fixed_gamma, const long long* __restrict__ idx,
                    const float* __restrict__ z, const float* __restrict__ mag, const float* __restrict__ dout,
                    float* __restrict__ d_support_sets, float* __restrict__ d_loggamma,
                    float* __restrict__ d_alphas, float* __restrict__ dz, int K, int n_vec, int d) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    const int b = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long k = idx[b];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    const float gamma = loggamma ? expf(__ldg(loggamma + k)) : fixed_gamma;
    const float* set = support_sets + (size_t)k * n_vec * d;
    const float* alpha = alphas + (size_t)k * n_vec;
    LaneVec<NV> zv, acc;
    load_vec<NV>(z + (size_t)b * d, d, lane, zv);
    rbf_accumulate<NV>(set, alpha, gamma, zv, n_vec, d, warp, lane, acc);
    const float n2 = rbf_reduce<NV>(acc, sm, red, d, warp, lane);
    const float inv_n = 1.f / sqrtf(n2);
    const float m = mag ? __ldg(mag + b) : 1.f;
    // u.du
    float dot = 0.f;
    for (int e = threadIdx.x; e < d; e += RBF_THREADS) dot += sm[e] * inv_n * (m * __ldg(dout + (size_t)b * d + e));
    dot = block_sum(dot, red);
    // dg into sm row 1 (sm has RBF_WARPS >= 2 rows)
    const int dp = pad4(d);
    float* dg_s = sm + dp;
    for (int e = threadIdx.x; e < d; e += RBF_THREADS)
        dg_s[e] = (m * __ldg(dout + (size_t)b * d + e) - sm[e] * inv_n * dot) * inv_n;
    __syncthreads();
    LaneVec<NV> dg, dzacc;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const int e = (i * 32 + lane) * 4;
        dg.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (e < d) {                                    // (elements beyond d inside the last slot: D_j is 0 there)
            dg.v[i] = *reinterpret_cast<const float4*>(dg_s + e);
            if (e + 1 >= d) dg.v[i].y = 0.f;
            if (e + 2 >= d) dg.v[i].z = 0.f;
            if (e + 3 >= d) dg.v[i].w = 0.f;
        }
        dzacc.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();                                   // sm is reused below for the dz reduction
    float dlg = 0.f;
    float* ds_row = d_support_sets ? d_support_sets + (size_t)k * n_vec * d : nullptr;
    for (int j = warp; j < n_vec; j += RBF_WARPS) {
        LaneVec<NV> s;
        load_vec<NV>(set + (size_t)j * d, d, lane, s);
        float q = 0.f, t = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            s.v[i].x = zv.v[i].x - s.v[i].x; s.v[i].y = zv.v[i].y - s.v[i].y;
            s.v[i].z = zv.v[i].z - s.v[i].z; s.v[i].w = zv.v[i].w - s.v[i].w;
            q += s.v[i].x * s.v[i].x + s.v[i].y * s.v[i].y + s.v[i].z * s.v[i].z + s.v[i].w * s.v[i].w;
            t += s.v[i].x * dg.v[i].x + s.v[i].y * dg.v[i].y + s.v[i].z * dg.v[i].z + s.v[i].w * dg.v[i].w;
        }
        q = warp_sum(q);
        t = warp_sum(t);
        const float a = __ldg(alpha + j);
        const float ex = gamma * expf(-gamma * q);
        const float w = a * ex;
        const float c1 = -2.f * w, c2 = 4.f * gamma * t * w;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = (i * 32 + lane) * 4;
            float4 g4;
            g4.x = c1 * dg.v[i].x + c2 * s.v[i].x; g4.y = c1 * dg.v[i].y + c2 * s.v[i].y;
            g4.z = c1 * dg.v[i].z + c2 * s.v[i].z; g4.w = c1 * dg.v[i].w + c2 * s.v[i].w;
            dzacc.v[i].x += g4.x; dzacc.v[i].y += g4.y; dzacc.v[i].z += g4.z; dzacc.v[i].w += g4.w;
            if (ds_row && e < d) {
                float* p = ds_row + (size_t)j * d + e;     // rows may repeat inside a batch -> atomics
                atomicAdd(p + 0, -g4.x);
                if (e + 1 < d) atomicAdd(p + 1, -g4.y);
                if (e + 2 < d) atomicAdd(p + 2, -g4.z);
                if (e + 3 < d) atomicAdd(p + 3, -g4.w);
            }
        }
        if (lane == 0) {
            dlg += (-2.f * t) * w * (1.f - gamma * q);
            if (d_alphas) atomicAdd(d_alphas + (size_t)k * n_vec + j, (-2.f * t) * ex);
        }
    }
    if (d_loggamma && loggamma) {
        dlg = block_sum(dlg, red);
        if (threadIdx.x == 0) atomicAdd(d_loggamma + k, dlg);
    }
    if (dz) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = (i * 32 + lane) * 4;
            if (e < d) *reinterpret_cast<float4*>(sm + (size_t)warp * dp + e) = dzacc.v[i];
        }
        __syncthreads();
        for (int e = threadIdx.x; e < d; e += RBF_THREADS) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < RBF_WARPS; ++w) t += sm[(size_t)w * dp + e];
            dz[(size_t)b * d + e] = t;
        }
    }
}

// Traversal chains (traverse_latent_space.py:369-438): one CTA per (latent, path) chain walks
// `steps` sequential RBF steps in each direction, keeping the code in registers; the support set of
// the path is re-streamed from L2 each step.  codes/shifts: [chains, 2*steps+1, d], ordered from the
// most negative step to the most positive; centre = (start, 0).
template <int NV>
__global__ void __launch_bounds__(RBF_THREADS)
rbf_traverse_kernel(const float* __restrict__ support_sets, const float* __restrict__ alphas,
                    const float* __restrict__ loggamma, float fixed_gamma, const long long* __restrict__ path,
                    const float* __restrict__ start, float eps, int steps, float* __restrict__ codes,
                    float* __restrict__ shifts, int K, int n_vec, int d) {
    extern __shared__ float sm[];
    __shared__ float red[32];
    const int c = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long k = path[c];
    k = k < 0 ? 0 : (k >= K ? K - 1 : k);
    const float gamma = loggamma ? expf(__ldg(loggamma + k)) : fixed_gamma;
    const float* set = support_sets + (size_t)k * n_vec * d;
    const float* alpha = alphas + (size_t)k * n_vec;
    const int frames = 2 * steps + 1;
    float* codes_c = codes + (size_t)c * frames * d;
    float* shifts_c = shifts + (size_t)c * frames * d;
    for (int e = threadIdx.x; e < d; e += RBF_THREADS) {
        codes_c[(size_t)steps * d + e] = start[(size_t)c * d + e];
        shifts_c[(size_t)steps * d + e] = 0.f;
    }
    for (int dir = 0; dir < 2; ++dir) {
        const float sgn = dir == 0 ? eps : -eps;
        LaneVec<NV> zv, acc;
        load_vec<NV>(start + (size_t)c * d, d, lane, zv);
        for (int s = 1; s <= steps; ++s) {
            rbf_accumulate<NV>(set, alpha, gamma, zv, n_vec, d, warp, lane, acc);
            const float n2 = rbf_reduce<NV>(acc, sm, red, d, warp, lane);
            const float scale = sgn / sqrtf(n2);
            const int frame = dir == 0 ? steps + s : steps - s;
#pragma unroll
            for (int i = 0; i < NV; ++i) {
                const int e = (i * 32 + lane) * 4;
                if (e < d) {
                    float4 g4 = *reinterpret_cast<const float4*>(sm + e);
                    g4.x *= scale; g4.y *= scale; g4.z *= scale; g4.w *= scale;
                    if (e + 1 >= d) g4.y = 0.f;              // keep the padding lanes of the code at zero
                    if (e + 2 >= d) g4.z = 0.f;
                    if (e + 3 >= d) g4.w = 0.f;
                    zv.v[i].x += g4.x; zv.v[i].y += g4.y; zv.v[i].z += g4.z; zv.v[i].w += g4.w;
                    if (warp == 0) {
                        store_slot(shifts_c + (size_t)frame * d, d, e, g4);
                        store_slot(codes_c + (size_t)frame * d, d, e, zv.v[i]);
                    }
                }
            }
            __syncthreads();            // sm is rewritten by the next step
        }
    }
}

template <typename F>
static int dispatch_nv(int d, F&& f) {
    const int nv = (d + 127) / 128;
    if (nv <= 1) return f(std::integral_constant<int, 1>());
    if (nv <= 2) return f(std::integral_constant<int, 2>());
    if (nv <= 4) return f(std::integral_constant<int, 4>());
    if (nv <= 8) return f(std::integral_constant<int, 8>());
    return fail(__FILE__, __LINE__, "support vector dimension > 1024 is not supported");
}

}  // namespace wgs

using namespace wgs;

extern "C" int wgs_rbf_warp_forward(const float* support_sets, const float* alphas, const float* loggamma,
                                    float fixed_gamma, const long long* idx, const float* z, const float* mag,
                                    float* out, int B, int K, int n_vec, int d, void* stream) {
    WGS_REQUIRE(B >= 0 && K > 0 && n_vec > 0 && d > 0, "rbf_warp_forward: bad sizes");
    if (B == 0) return 0;
    const size_t smem = (size_t)RBF_WARPS * ((d + 3) & ~3) * sizeof(float);
    return dispatch_nv(d, [&](auto nv) -> int {
        constexpr int NV = decltype(nv)::value;
        rbf_forward_kernel<NV><<<B, RBF_THREADS, smem, (cudaStream_t)stream>>>(
            support_sets, alphas, loggamma, fixed_gamma, idx, z, mag, out, K, n_vec, d);
        count_launch();
        WGS_LAUNCH_CHECK();
        return 0;
    });
}

extern "C" int wgs_rbf_warp_backward(const float* support_sets, const float* alphas, const float* loggamma,
                                     float fixed_gamma, const long long* idx, const float* z, const float* mag,
                                     const float* dout, float* d_support_sets, float* d_loggamma, float* d_alphas,
                                     float* dz, int B, int K, int n_vec, int d, void* stream) {
    WGS_REQUIRE(B >= 0 && K > 0 && n_vec > 0 && d > 0, "rbf_warp_backward: bad sizes");
    if (B == 0) return 0;
    const size_t smem = (size_t)RBF_WARPS * ((d + 3) & ~3) * sizeof(float);
    return dispatch_nv(d, [&](auto nv) -> int {
        constexpr int NV = decltype(nv)::value;
        rbf_backward_kernel<NV><<<B, RBF_THREADS, smem, (cudaStream_t)stream>>>(
            support_sets, alphas, loggamma, fixed_gamma, idx, z, mag, dout, d_support_sets, d_loggamma, d_alphas,
            dz, K, n_vec, d);
        count_launch();
        WGS_LAUNCH_CHECK();
        return 0;
    });
}

extern "C" int wgs_rbf_traverse(const float* support_sets, const float* alphas, const float* loggamma,
                                float fixed_gamma, const long long* path, const float* start, float eps, int steps,
                                float* codes, float* shifts, int chains, int K, int n_vec, int d, void* stream) {
    WGS_REQUIRE(chains >= 0 && K > 0 && n_vec > 0 && d > 0 && steps >= 0, "rbf_traverse: bad sizes");
    if (chains == 0) return 0;
    const size_t smem = (size_t)RBF_WARPS * ((d + 3) & ~3) * sizeof(float);
    return dispatch_nv(d, [&](auto nv) -> int {
        constexpr int NV = decltype(nv)::value;
        rbf_traverse_kernel<NV><<<chains, RBF_THREADS, smem, (cudaStream_t)stream>>>(
            support_sets, alphas, loggamma, fixed_gamma, path, start, eps, steps, codes, shifts, K, n_vec, d);
        count_launch();
        WGS_LAUNCH_CHECK();
        return 0;
    });
}
