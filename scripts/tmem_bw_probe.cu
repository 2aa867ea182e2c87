// tmem_bw_probe.cu — read / write bandwidth of tensor memory used as a per-thread scratchpad (tcgen05.ld / tcgen05.st,
// shape 32x32b), per SM, as a function of the number of warps issuing and of the vector width (.x8 / .x32).
// The element kernel parks 78 doubles per thread in TMEM and re-reads 156 columns per Gauss point: this probe says
// what that costs.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tmem_bw_probe.bin tmem_bw_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld8(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(t) : "memory");
}
__device__ __forceinline__ void ld32(uint32_t t, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(t) : "memory");
}
__device__ __forceinline__ void st8(uint32_t t, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "r"(t), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// mode 0: ld.x8 x16 (128 columns) per iteration, one wait;  1: ld.x32 x4;  2: st.x8 x16;  3: LDS.64 of 64 doubles (shared-memory twin)
__global__ void __launch_bounds__(384, 1) probe(long long* cycles, uint32_t* sink, int active, int iters, int mode) {
    __shared__ uint32_t tbase;
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = threadIdx.x; i < 64 * 384; i += 384) sm[i] = i;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t my = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 168);
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < active) {
        uint32_t v[32];
        for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
        for (int c = 0; c < 128; c += 8) st8(my + c, v);
        wait_st();
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (mode == 0) {
#pragma unroll
                for (int c = 0; c < 128; c += 8) { ld8(my + c, v); acc ^= v[0] ^ v[7]; }
                wait_ld();
            } else if (mode == 1) {
#pragma unroll
                for (int c = 0; c < 128; c += 32) { ld32(my + c, v); acc ^= v[0] ^ v[31]; }
                wait_ld();
            } else if (mode == 2) {
#pragma unroll
                for (int c = 0; c < 128; c += 8) { v[0] = acc + c; st8(my + c, v); }
                wait_st();
                acc += it;
            } else {
                double a = 0;
#pragma unroll
                for (int c = 0; c < 64; ++c) a += sm[c * 384 + threadIdx.x];
                acc ^= (uint32_t)a;
            }
        }
        t1 = clock64();
    }
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) cycles[warp] = t1 - t0;
    if (acc == 0x12345678u) sink[threadIdx.x] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

int main() {
    long long* cyc; uint32_t* sink;
    cudaMalloc(&cyc, 12 * 8); cudaMalloc(&sink, 384 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 384 * 8);
    const char* names[4] = {"tcgen05.ld.x8 (16 per wait)", "tcgen05.ld.x32 (4 per wait)", "tcgen05.st.x8 (16 per wait)", "LDS.64 (64 per thread)"};
    const int iters = 2000;
    printf("{\"bytes_per_warp_iteration\": 16384, \"results\": [\n");
    for (int mode = 0; mode < 4; ++mode)
        for (int active = 1; active <= 12; active = active < 4 ? active * 2 : active + 4 > 12 && active < 11 ? 11 : active + 4) {
            if (active > 11) break;
            cudaMemset(cyc, 0, 12 * 8);
            probe<<<148, 384, 64 * 384 * 8>>>(cyc, sink, active, iters, mode);
            cudaError_t rc = cudaDeviceSynchronize();
            if (rc != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(rc)); return 1; }
            long long h[12];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int w = 0; w < active; ++w) mx = h[w] > mx ? h[w] : mx;
            const double bytes = (double)active * iters * 128 * 32 * 4;
            printf(" {\"op\": \"%s\", \"warps\": %d, \"cycles\": %lld, \"bytes_per_cycle_per_SM\": %.1f},\n", names[mode], active, mx, bytes / mx);
        }
    printf(" {\"op\": \"end\"}]}\n");
    return 0;
}
