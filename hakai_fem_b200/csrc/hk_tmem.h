// hk_tmem.h — Blackwell tensor memory (TMEM) used as a per-thread scratchpad.
//
// TMEM is 512 columns x 128 lanes x 32 bit per SM and is normally the accumulator store of tcgen05.mma.  The
// element kernel has no matrix product to offer the tensor cores, but it has 78 doubles of per-element state that must
// survive the Gauss-point loop.  With the 32x32b access shape every thread of a warp owns one TMEM lane (warp w of the
// CTA reaches lanes 32*(w%4)..+31), so a thread can park N consecutive 32-bit columns there with tcgen05.st and fetch
// them back with tcgen05.ld (LDTM/STTM in SASS) — a second register file of 2 KB per lane that costs neither shared
// memory nor occupancy.  Measured (profiles/r2_tmem_bw_probe.json): 67 B/clk per warp with .x8 loads, scaling
// linearly with the number of warps (487 B/clk/SM at 8 warps), i.e. faster than shared memory (128 B/clk/SM).
#pragma once
#include <cstdint>

__device__ __forceinline__ void tmem_ld_x2(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];"
                 : "=r"(v[0]), "=r"(v[1])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_x8(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};"
                 :
                 : "r"(taddr), "r"(v[0]), "r"(v[1])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_x8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :
                 : "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ND doubles -> 2*ND columns, ND in {1, 4}
template <int ND>
__device__ __forceinline__ void tmem_st_doubles(uint32_t taddr, const double* in) {
    static_assert(ND == 1 || ND == 4, "tmem_st_doubles: 1 or 4 doubles");
    uint32_t v[2 * ND];
#pragma unroll
    for (int i = 0; i < ND; ++i) { v[2 * i] = (uint32_t)__double2loint(in[i]); v[2 * i + 1] = (uint32_t)__double2hiint(in[i]); }
    if (ND == 1) tmem_st_x2(taddr, v);
    if (ND == 4) tmem_st_x8(taddr, v);
}
