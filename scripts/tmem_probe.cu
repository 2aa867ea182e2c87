// Probe: TMEM (tcgen05.st / tcgen05.ld, shape 32x32b) as per-thread scratch for a CTA of 12 warps.
// Each thread stores 160 doubles' worth of 32-bit columns... here: NCOL columns at its own column base, reads back.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(384, 1) probe(uint32_t* out, int* err) {
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    const int quarter = warp & 3, group = warp >> 2;           // lane quarter of this warp, column group
    const uint32_t my = base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(group * 160);
    // write 160 columns: value encodes (thread, column)
    for (int c = 0; c < 160; c += 4) {
        uint32_t v = threadIdx.x * 1000 + c;
        tmem_st4(my + c, v, v + 1, v + 2, v + 3);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    __syncthreads();
    int bad = 0;
    for (int c = 0; c < 160; c += 4) {
        uint32_t a, b, cc, d;
        tmem_ld4(my + c, a, b, cc, d);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t v = threadIdx.x * 1000 + c;
        bad += (a != v) + (b != v + 1) + (cc != v + 2) + (d != v + 3);
        if (c == 8) out[threadIdx.x] = a;
    }
    if (bad) atomicAdd(err, bad);
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
    (void)lane;
}

int main() {
    uint32_t* out; int* err;
    cudaMalloc(&out, 384 * 4); cudaMalloc(&err, 4); cudaMemset(err, 0, 4);
    probe<<<148, 384>>>(out, err);
    cudaError_t rc = cudaDeviceSynchronize();
    int herr = -1; uint32_t hout[384];
    cudaMemcpy(&herr, err, 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hout, out, 384 * 4, cudaMemcpyDeviceToHost);
    printf("rc=%s mismatches=%d sample out[0]=%u out[37]=%u out[383]=%u (expect 8, 37008, 383008)\n", cudaGetErrorString(rc), herr, hout[0], hout[37], hout[383]);
    return rc != cudaSuccess || herr != 0;
}
