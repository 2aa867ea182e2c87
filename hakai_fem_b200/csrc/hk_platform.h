// hk_platform.h — memory / launch / atomic shims.
//
// The product build (libhakai_b200.so) is CUDA only: kernels run on the engine's stream and
// hk_create() fails when no device exists — there is no CPU fallback.
// With -DHK_EMU (tests/emu only: never shipped, never loaded by the package) the same per-thread
// kernel bodies are driven by serial host loops so that kernel logic and the engine plumbing can be
// debugged in the GPU-less build container.  It is a debugging aid, not a product path.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#ifndef HK_EMU
#include <cuda_runtime.h>
#define HK_HD __host__ __device__ __forceinline__
#define HK_D __device__ __forceinline__
#else
#define HK_HD inline
#define HK_D inline
typedef void* cudaStream_t;
#endif

namespace hkp {

#ifndef HK_EMU
inline int dev_malloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
inline int dev_free(void* p) { return (int)cudaFree(p); }
inline int h2d(void* d, const void* h, size_t n, cudaStream_t s) {
    int rc = (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s);
    if (rc) return rc;
    return (int)cudaStreamSynchronize(s);
}
inline int d2h(void* h, const void* d, size_t n, cudaStream_t s) {
    int rc = (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s);
    if (rc) return rc;
    return (int)cudaStreamSynchronize(s);
}
// asynchronous forms: the caller synchronises the stream before the host buffer is reused / read
inline int h2d_async(void* d, const void* h, size_t n, cudaStream_t s) { return (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s); }
inline int d2h_async(void* h, const void* d, size_t n, cudaStream_t s) { return (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s); }
inline int d2d(void* dst, const void* src, size_t n, cudaStream_t s) {
    return (int)cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s);
}
inline int dev_memset(void* d, int v, size_t n, cudaStream_t s) { return (int)cudaMemsetAsync(d, v, n, s); }
inline int sync(cudaStream_t s) { return (int)cudaStreamSynchronize(s); }
inline int last_error() { return (int)cudaGetLastError(); }
inline const char* error_string(int rc) { return cudaGetErrorString((cudaError_t)rc); }
#else
inline int dev_malloc(void** p, size_t n) { *p = std::malloc(n ? n : 1); return *p ? 0 : 2; }
inline int dev_free(void* p) { std::free(p); return 0; }
inline int h2d(void* d, const void* h, size_t n, cudaStream_t) { std::memcpy(d, h, n); return 0; }
inline int d2h(void* h, const void* d, size_t n, cudaStream_t) { std::memcpy(h, d, n); return 0; }
inline int h2d_async(void* d, const void* h, size_t n, cudaStream_t) { std::memcpy(d, h, n); return 0; }
inline int d2h_async(void* h, const void* d, size_t n, cudaStream_t) { std::memcpy(h, d, n); return 0; }
inline int d2d(void* dst, const void* src, size_t n, cudaStream_t) { std::memcpy(dst, src, n); return 0; }
inline int dev_memset(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return 0; }
inline int sync(cudaStream_t) { return 0; }
inline int last_error() { return 0; }
inline const char* error_string(int) { return "emu"; }
#endif

}  // namespace hkp

// ---- atomics used by kernels ----------------------------------------------------------------
#ifndef HK_EMU
HK_D unsigned long long hk_atomic_add_u64(unsigned long long* p, unsigned long long v) { return atomicAdd(p, v); }
HK_D int hk_atomic_add_i32(int* p, int v) { return atomicAdd(p, v); }
HK_D int hk_atomic_exch_i32(int* p, int v) { return atomicExch(p, v); }
HK_D unsigned long long hk_atomic_min_u64(unsigned long long* p, unsigned long long v) { return atomicMin(p, v); }
HK_D unsigned long long hk_atomic_max_u64(unsigned long long* p, unsigned long long v) { return atomicMax(p, v); }
#else
inline unsigned long long hk_atomic_add_u64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p = o + v; return o; }
inline int hk_atomic_add_i32(int* p, int v) { int o = *p; *p = o + v; return o; }
inline int hk_atomic_exch_i32(int* p, int v) { int o = *p; *p = v; return o; }
inline unsigned long long hk_atomic_min_u64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v < o) *p = v; return o; }
inline unsigned long long hk_atomic_max_u64(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; if (v > o) *p = v; return o; }
#endif

// ---- generic 1-D launch: body(i) for i in [0,n) ------------------------------------------------
#ifndef HK_EMU
template <class F>
__global__ void hk_generic_kernel(long long n, F f) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) f(i);
}
template <class F>
inline void hk_parallel_for(long long n, cudaStream_t s, F f, int block = 256) {
    if (n <= 0) return;
    hk_generic_kernel<<<(unsigned)((n + block - 1) / block), block, 0, s>>>(n, f);
}
#define HK_LAMBDA [=] __device__
#else
template <class F>
inline void hk_parallel_for(long long n, cudaStream_t, F f, int = 256) {
    for (long long i = 0; i < n; ++i) f(i);
}
#define HK_LAMBDA [=]
#endif
