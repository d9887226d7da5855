// Probe: register <-> (lane, column) mapping of tcgen05.ld shapes .16x256b / .16x128b / .16x64b (sm_100a).
// TMEM is filled through tcgen05.st.32x32b (thread i <-> lane i, register j <-> column j, the layout the kernels already rely
// on) with value lane * 1000 + column, then read back with each shape and printed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_probe tmem_ld_layout.cu && ./tmem_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void probe(float* out) {
    __shared__ uint32_t slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot + ((uint32_t)(warp * 32) << 16);
    // fill 32 columns: thread = lane, 16 registers per store
    for (int c0 = 0; c0 < 32; c0 += 16) {
        uint32_t r[16];
        for (int j = 0; j < 16; ++j) r[j] = __float_as_uint((float)((warp * 32 + lane) * 1000 + c0 + j));
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                     ::"r"(base + c0), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
                       "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    if (warp == 0) {
        uint32_t a[4], b[2], c[2], d[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.ld.sync.aligned.16x128b.x1.b32 {%0,%1}, [%2];" : "=r"(b[0]), "=r"(b[1]) : "r"(base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.ld.sync.aligned.16x64b.x1.b32 {%0}, [%1];" : "=r"(c[0]) : "r"(base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(d[0]), "=r"(d[1]), "=r"(d[2]), "=r"(d[3]), "=r"(d[4]), "=r"(d[5]), "=r"(d[6]), "=r"(d[7]) : "r"(base));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        // second half of the warp's lanes: lane field + 16
        uint32_t e[4];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x1.b32 {%0,%1,%2,%3}, [%4];" : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]) : "r"(base + (16u << 16)));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float* o = out + lane * 32;
        for (int j = 0; j < 4; ++j) o[j] = __uint_as_float(a[j]);
        for (int j = 0; j < 2; ++j) o[4 + j] = __uint_as_float(b[j]);
        o[6] = __uint_as_float(c[0]);
        for (int j = 0; j < 8; ++j) o[8 + j] = __uint_as_float(d[j]);
        for (int j = 0; j < 4; ++j) o[16 + j] = __uint_as_float(e[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(slot) : "memory");
}

int main() {
    float* d; cudaMalloc(&d, 32 * 32 * 4); cudaMemset(d, 0, 32 * 32 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(err)); return 1; }
    float h[32 * 32]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("value = lane*1000 + column\nthread | 16x256b.x1 r0..r3 | 16x128b.x1 r0 r1 | 16x64b.x1 r0 | 16x256b.x2 r0..r7 | 16x256b.x1 @lane+16\n");
    for (int t = 0; t < 32; ++t) {
        printf("%2d |", t);
        for (int j = 0; j < 4; ++j) printf(" %6.0f", h[t * 32 + j]);
        printf(" |"); for (int j = 4; j < 6; ++j) printf(" %6.0f", h[t * 32 + j]);
        printf(" | %6.0f |", h[t * 32 + 6]);
        for (int j = 8; j < 16; ++j) printf(" %6.0f", h[t * 32 + j]);
        printf(" |"); for (int j = 16; j < 20; ++j) printf(" %6.0f", h[t * 32 + j]);
        printf("\n");
    }
    return 0;
}
