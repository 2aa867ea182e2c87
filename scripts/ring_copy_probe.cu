// ring_copy_probe.cu — what the TMA ring of the element kernel can move with NO consumers at all.
//
// One persistent CTA per SM; lane 0 of one warp streams its share of a large buffer through a ring of S shared-memory
// stages of B bytes (cp.async.bulk load -> mbarrier -> cp.async.bulk store of the same block, in place), exactly the
// protocol of hk_element_ring_kernel's producer, with the "consumer" reduced to an mbarrier wait.  Prints GB/s
// (read + write) per (B, S) next to a plain grid-stride LDG/STG copy and cudaMemcpy D2D: the gap between this ceiling
// and the element kernel is what the consumers cost; the gap between this and the plain copy is what the access
// pattern (one 18-79 KB burst per SM at a time) costs.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ring_copy_probe ring_copy_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned ok, spins = 0;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(ok) : "r"(addr), "r"(parity) : "memory");
        if (!ok && ++spins > 100000000u) __trap();
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// items of B bytes; CTA b takes items b, b+grid, ...; `interleave`: 0 = that round-robin (neighbouring SMs work on
// neighbouring blocks, like the element kernel's tiles), 1 = each CTA owns one contiguous chunk of the buffer
__global__ void __launch_bounds__(32, 1) ring_copy(char* buf, long long n_items, int B, int S, int chunked) {
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned long long* full = reinterpret_cast<unsigned long long*>(smem + (size_t)S * B);
    if (threadIdx.x != 0) return;
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const long long per = (n_items + gridDim.x - 1) / gridDim.x;
    const long long first = chunked ? blockIdx.x * per : blockIdx.x;
    const long long stride = chunked ? 1 : gridDim.x;
    long long mine = chunked ? (n_items - first < per ? n_items - first : per) : (n_items - first + gridDim.x - 1) / gridDim.x;
    if (mine < 0) mine = 0;
    auto addr = [&](long long q) { return buf + (first + q * stride) * (long long)B; };
    auto load = [&](long long q) {
        const int st = (int)(q % S);
        mbar_expect_tx(&full[st], B);
        tma_load_1d(smem + (size_t)st * B, addr(q), B, &full[st]);
    };
    for (long long q = 0; q < S && q < mine; ++q) load(q);
    for (long long q = 0; q < mine; ++q) {
        const int st = (int)(q % S);
        mbar_wait(&full[st], (unsigned)((q / S) & 1));
        tma_store_1d(addr(q), smem + (size_t)st * B, B);
        tma_commit();
        const long long qn = q - 1 + S;
        if (q >= 1 && qn < mine) { tma_wait_read1(); load(qn); }
    }
    tma_wait_all();
}

__global__ void plain_copy(const double2* __restrict__ src, double2* __restrict__ dst, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s\n", cudaGetErrorString(e_), #x); return 1; } } while (0)

int main() {
    int n_sm = 0;
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0));
    const size_t bytes = (size_t)8 << 30;
    char *a = nullptr, *b = nullptr;
    CK(cudaMalloc(&a, bytes));
    CK(cudaMalloc(&b, bytes));
    CK(cudaMemset(a, 1, bytes));
    CK(cudaMemset(b, 2, bytes));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    printf("{\"n_sm\": %d, \"buffer_GB\": %.1f, \"results\": [\n", n_sm, bytes / 1e9);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice));
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf(" {\"kind\": \"cudaMemcpy D2D\", \"GBps\": %.0f},\n", 2.0 * bytes * 5 / (ms * 1e-3) / 1e9);
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) plain_copy<<<n_sm * 16, 512>>>((const double2*)a, (double2*)b, (long long)(bytes / 16));
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        printf(" {\"kind\": \"plain LDG.128/STG.128 copy\", \"GBps\": %.0f},\n", 2.0 * bytes * 5 / (ms * 1e-3) / 1e9);
    }
    const int Bs[] = {17920, 39424, 78848, 8192, 32768, 65536};
    for (int chunked = 0; chunked < 2; ++chunked)
        for (int bi = 0; bi < 6; ++bi)
            for (int S = 2; S <= 12; ++S) {
                const int B = Bs[bi];
                const size_t smem = (size_t)S * B + 8 * S + 64;
                if (smem > 227 * 1024) continue;
                if (S > 6 && S != 8 && S != 12) continue;
                CK(cudaFuncSetAttribute(ring_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                const long long n_items = (long long)(bytes / B);
                ring_copy<<<n_sm, 32, smem>>>(a, n_items, B, S, chunked);      // warm-up
                CK(cudaDeviceSynchronize());
                cudaEventRecord(e0);
                for (int i = 0; i < 3; ++i) ring_copy<<<n_sm, 32, smem>>>(a, n_items, B, S, chunked);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
                printf(" {\"kind\": \"tma ring\", \"chunked\": %d, \"block_bytes\": %d, \"stages\": %d, \"smem_KB\": %.0f, \"GBps\": %.0f},\n",
                       chunked, B, S, smem / 1024.0, 2.0 * n_items * B * 3 / (ms * 1e-3) / 1e9);
                fflush(stdout);
            }
    printf(" {\"kind\": \"end\"}]}\n");
    return 0;
}
