// hk_element.cu — per-element hex8 internal-force kernel (FP64, FMA allowed).
//
// Replaces cal_stress_hexa + cal_BVbar_hexa + cal_Bfinal (J2:1033-1371, 1705-1784, 1415-1519),
// cal_triax_stress (J2:982-1022) and the fracture loop (J2:682-764) with ONE pass over the element:
// node gather, Jacobians at the 8 Gauss points, mean-dilatation (B-bar) strain increment, J2 radial
// return, state update, nodal force, triaxiality, ductile-damage deletion.  The math (trilinear mode
// form) is in hk_element_math.h.
//
// Kernels sharing that math:
//   hk_element_ring_kernel   (default) persistent CTAs; ip state streamed through shared-memory rings by the TMA
//       (cp.async.bulk + mbarrier), element state parked in tensor memory, phase-shifted consumer groups — see the
//       comment above the kernel.
//   hk_element_simple_kernel  thread per element with plain coalesced loads (HK_ELEMENT_KERNEL=simple): the A/B
//       baseline of the profiles and the body the host-compiled debugging build runs.
#include "hk_element_math.h"

struct ElemArgs {
    HkDev d;
    long long step;
    int write_triax;
    int fast;          // MatLite::fast
    int n_tiles;       // ring kernels: nEp / tile (host-computed so the tile loop needs no 64-bit division and no
                       // register that survives the Gauss-point loop: round 2's ncu showed the spilled trip count
                       // costing a DRAM-latency reload per tile)
};
HK_HD MatLite mat_lite(const HkMaterialDev* m) {
    MatLite l;
    l.D11 = m->D11; l.D12 = m->D12; l.D44 = m->D44; l.G3 = 3.0 * m->G; l.npp = m->npp;
    l.pe = m->plastic_e; l.hd = m->Hd;
    l.fast = 1;
    return l;
}

HK_HD double triax_of(const double s[6]) {
    // (I1/3)/sqrt(3 J2): equals mean(p)/sigma_eq of the principal stresses p (J2:1004-1016)
    const double oeq = sqrt(0.5 * ((s[0] - s[1]) * (s[0] - s[1]) + (s[1] - s[2]) * (s[1] - s[2]) +
                                   (s[0] - s[2]) * (s[0] - s[2]) + 6.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5])));
    if (oeq < 1E-10) return 0.0;
    return (s[0] + s[1] + s[2]) / 3.0 / oeq;
}

// ---- pieces shared by both kernels ------------------------------------------------------------------
HK_D bool element_dead(const HkDev& d, long long e, int fl_known = -1) {
    const unsigned char fl = fl_known >= 0 ? (unsigned char)fl_known : d.flag[e];   // ring kernels: flag already on chip
    if (fl == 1) return false;
    if (fl == 0) {      // deleted during the previous step: its last force has been consumed, clear it
#pragma unroll
        for (int r = 0; r < 24; ++r) d.Qe[(long long)r * d.nEp + e] = 0.0;
#pragma unroll
        for (int k = 0; k < 8; ++k) d.triax[(long long)k * d.nEp + e] = 0.0;   // zero stress -> triax 0 (J2:1012)
        d.flag[e] = 2;
    }
    return true;
}

HK_D void element_gather(const HkDev& d, long long e, HexModes& X, HexModes& U) {
    double x[8][3], du[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long long n = d.conn[(long long)a * d.nEp + e];
#if defined(__CUDA_ARCH__)
        const double2* r = reinterpret_cast<const double2*>(d.rec + 6 * n);
        const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
        x[a][0] = r0.x; x[a][1] = r0.y; x[a][2] = r1.x;
        du[a][0] = r1.y; du[a][1] = r2.x; du[a][2] = r2.y;
#else
        const double* r = d.rec + 6 * n;
        for (int c = 0; c < 3; ++c) { x[a][c] = r[c]; du[a][c] = r[3 + c]; }
#endif
    }
    hex_modes(x, X);
    hex_modes(du, U);
}

// V = sum_k det_k and trbar = (sum_k det_k tr L_k) / V from the closed-form adjugate sums
HK_HD void element_volume_terms(const HexModes& X, const HexModes& U, const double G[3][3][3], double& V, double& trbar) {
    V = dot3(X.c0, G[0][0]) + dot3(X.h01, G[0][1]) + dot3(X.h02, G[0][2]);
    const double tv = dot3(U.c0, G[0][0]) + dot3(U.h01, G[0][1]) + dot3(U.h02, G[0][2]) +
                      dot3(U.c1, G[1][0]) + dot3(U.h01, G[1][1]) + dot3(U.h12, G[1][2]) +
                      dot3(U.c2, G[2][0]) + dot3(U.h02, G[2][1]) + dot3(U.h12, G[2][2]);
    trbar = tv / V;
}

HK_HD void acc_init(ElemAcc& acc) {
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int m = 0; m < 4; ++m) acc.M[r][m][0] = acc.M[r][m][1] = acc.M[r][m][2] = 0.0;
    acc.pdet = 0.0; acc.v_e = 0.0; acc.t_e = 0.0; acc.negj = 0;
}

// ductile damage (J2:701-762): returns true when the element must be deleted
HK_HD bool ductile_check(const HkMaterialDev& M, double v_sum, double t_sum) {
    if (M.nd <= 0) return false;
    const double v_e = v_sum / 8, t_e = t_sum / 8;
    if (t_e < 0) return false;
    const int nd = M.nd;
    double fr_e = M.duct_e[nd - 1];
    for (int j = 0; j + 1 < nd; ++j)
        if (t_e >= M.duct_t[j] && t_e < M.duct_t[j + 1]) {
            fr_e = M.duct_e[j] + (M.duct_e[j + 1] - M.duct_e[j]) / (M.duct_t[j + 1] - M.duct_t[j]) * (t_e - M.duct_t[j]);
            break;
        }
    return v_e >= fr_e;
}

HK_D void element_finish(const ElemArgs& A, long long e, const HexModes& X, const ElemAcc& acc, double V) {
    const HkDev& d = A.d;
    double G[3][3][3];
    adj_mode_sums(X, G);
    double f[8][3];
    element_forces(acc, G, acc.pdet / V, f);
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) d.Qe[(long long)(a * 3 + c) * d.nEp + e] = f[a][c];
    if (acc.negj) hk_atomic_add_u64(&d.counters[0], (unsigned long long)acc.negj);
}

// deletion (J2:733-756): the element is only MARKED here (flag 3); hk_launch_deletion_pass, which follows every element
// kernel in stream order, logs the step's deletions in ascending element order, zeroes their stress/strain (in the ring
// kernels the rows are still in flight in bulk stores at this point) and updates the contact surfaces
HK_D void element_delete(const ElemArgs& A, long long e) { A.d.flag[e] = 3; }

// ---- simple kernel: thread per element, state straight from global memory ---------------------------------
HK_D void element_body_simple(const ElemArgs& A, long long e) {
    const HkDev& d = A.d;
    const long long nEp = d.nEp;
    if (element_dead(d, e)) return;
    const HkMaterialDev& M = d.mats[d.mat[e]];
    MatLite ML = mat_lite(&M);
    ML.fast = A.fast;
    HexModes X, U;
    element_gather(d, e, X, U);
    double V, trbar;
    {
        double G[3][3][3];
        adj_mode_sums(X, G);
        element_volume_terms(X, U, G, V, trbar);
    }
    ElemAcc acc;
    acc_init(acc);
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        const long long row = (long long)k * nEp + e;
        double* sp = d.ips + hk_ip(d, 0, k, e);
        const long long TL = d.TL;
        double st[14];
#pragma unroll
        for (int r = 0; r < 14; ++r) st[r] = sp[r * TL];
        const double ep_old = st[12];
        const double tx = gauss_point(X, U, ML, k, trbar, st, acc);
#pragma unroll
        for (int r = 0; r < 12; ++r) sp[r * TL] = st[r];
        if (st[12] != ep_old) { sp[12 * TL] = st[12]; sp[13 * TL] = st[13]; }
        if (A.write_triax) d.triax[row] = tx;
    }
    element_finish(A, e, X, acc, V);
    if (ductile_check(M, acc.v_e, acc.t_e)) element_delete(A, e);
}

#ifndef HK_EMU
__global__ void __launch_bounds__(128, 2) hk_element_simple_kernel(ElemArgs A) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e < A.d.nElement) element_body_simple(A, e);
}

// ---- ring kernel: TMA-staged ip state, TMEM-parked element state, phase-shifted consumer groups ---------------
// One persistent CTA per SM.  The CTA holds NG independent consumer groups of WG warps (one element per thread, tile =
// WG*32 elements) and one producer warp per group.  Per group:
//   * the producer's lane 0 streams the tile's ip state through a ring of S shared-memory stages: ONE
//     cp.async.bulk (TMA, SASS UBLKCP) of 14 rows x tile doubles per (tile, Gauss point) with mbarrier completion
//     (`full`), and one bulk store of the updated block after the group's warps arrived on the stage's `done`
//     barrier.  Consumers never issue a global load/store for ip state and never meet at a CTA-wide barrier.
//   * CP = 1: the producer also brings the NEXT tile's connectivity (8 x tile int32) into a double-buffered
//     shared-memory block while the current tile is being processed, so the node gather of a tile starts with its
//     node ids already on chip (the dependent DRAM round trip conn -> node record is gone from the prologue).
//   * all per-element state that must survive the Gauss-point loop (geometry modes X, displacement modes U, the 36
//     force-mode accumulators M: 78 doubles) lives in tensor memory, one TMEM lane per thread (hk_tmem.h), so
//     registers hold only one Gauss point's temporaries.
// Why groups: a tile has a prologue (gather 8 node records, Hadamard transforms, closed-form B-bar sums) and an
// epilogue (force modes -> 24 nodal forces, Qe stores) during which its ring does not drain.  With ONE group the whole
// SM pauses for ~3 us of every ~21 us tile (round 1: 0.71 of the copy roofline, and the same 6.1 ms with the math
// deleted).  Two groups run half a tile out of phase — group 1 starts after group 0 has finished 4 Gauss points —
// so one group's prologue/epilogue overlaps the other's Gauss-point loop and the SM's DRAM streams never all pause.
// Every lane runs the full math (dead or padded elements get a unit cube with zero displacement, which leaves their
// state rows bit-unchanged), so the warp-collective tcgen05.ld/st never execute under divergence.
#define HK_ROWS 14                        // state rows per Gauss point: stress 6, strain 6, eps, yield
#define HK_SMEM_MATS 8                    // materials whose hardening tables are cached in shared memory
#define HK_TCOLS 168                      // TMEM columns per thread: X [0,42) U [48,90) M [96,168)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned ok, spins = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!ok && ++spins > 100000000u) __trap();      // a protocol bug must fail loudly, never hang the GPU
    } while (!ok);
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
// the same with an L2 cache policy (createpolicy): the Gauss-point state is a pure stream — marked evict-first it stops
// pushing the node records (re-read by the z-neighbour tile ~180 tiles later) and its own write-backs out of L2
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_1d_hint(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar,
                                                 unsigned long long pol) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_store_1d_hint(void* gdst, const void* smem_src, unsigned bytes, unsigned long long pol) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// global address of the state of Gauss point k of the tile starting at element e0: 14 rows x TL doubles, contiguous
__device__ __forceinline__ double* stage_base(const HkDev& d, int k, long long e0) {
    return d.ips + ((e0 / d.TL) * 8 + k) * 14 * d.TL;
}

#include "hk_tmem.h"

__device__ __forceinline__ void tmem_store_modes(uint32_t t, const HexModes& m) {
    double v[24];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        v[0 + c] = m.c0[c]; v[3 + c] = m.c1[c]; v[6 + c] = m.c2[c]; v[9 + c] = m.h01[c];
        v[12 + c] = m.h02[c]; v[15 + c] = m.h12[c]; v[18 + c] = m.h012[c];
    }
    v[21] = v[22] = v[23] = 0.0;
#pragma unroll
    for (int i = 0; i < 5; ++i) tmem_st_doubles<4>(t + 8 * i, v + 4 * i);
    tmem_st_doubles<1>(t + 40, v + 20);
}
__device__ __forceinline__ void tmem_load_modes(uint32_t t, HexModes& m) {
    uint32_t w[42];
#pragma unroll
    for (int i = 0; i < 5; ++i) tmem_ld_x8(t + 8 * i, w + 8 * i);
    tmem_ld_x2(t + 40, w + 40);
    tmem_wait_ld();
    double v[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) v[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        m.c0[c] = v[0 + c]; m.c1[c] = v[3 + c]; m.c2[c] = v[6 + c]; m.h01[c] = v[9 + c];
        m.h02[c] = v[12 + c]; m.h12[c] = v[15 + c]; m.h012[c] = v[18 + c];
    }
}

// node records of the 8 nodes listed in shared memory (conn block [8][TLD] of the tile) -> modes
template <int TLD>
__device__ __forceinline__ void element_gather_ids(const HkDev& d, const int* ids, HexModes& X, HexModes& U) {
    double x[8][3], du[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long long n = ids[a * TLD];
        const double2* r = reinterpret_cast<const double2*>(d.rec + 6 * n);
        const double2 r0 = __ldg(r), r1 = __ldg(r + 1), r2 = __ldg(r + 2);
        x[a][0] = r0.x; x[a][1] = r0.y; x[a][2] = r1.x;
        du[a][0] = r1.y; du[a][1] = r2.x; du[a][2] = r2.y;
    }
    hex_modes(x, X);
    hex_modes(du, U);
}

template <int NG, int WG, int S, int CP>
struct RingCfg {
    static constexpr int TLD = WG * 32;                        // elements per tile
    static constexpr int STAGE = HK_ROWS * TLD;                // doubles per stage
    static constexpr int THREADS = (NG * WG + NG) * 32;        // consumers + one producer warp per group
    static constexpr int SMEM = NG * S * STAGE * 8 + (2 * NG * S + 2 * NG) * 8 + HK_SMEM_MATS * 2 * HK_MAX_TABLE * 8 +
                                (CP ? NG * 2 * 8 * TLD * 4 : 0) + (CP >= 2 ? NG * 2 * TLD * 4 : 0) + 64;
};

template <int NG, int WG, int S, int CP, int EF = 0>
__global__ void __launch_bounds__((NG * WG + NG) * 32, 1) hk_element_ring_kernel(ElemArgs A) {
    using Cfg = RingCfg<NG, WG, S, CP>;
    constexpr int TLD = Cfg::TLD, STAGE = Cfg::STAGE;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double* stage_buf = reinterpret_cast<double*>(smem_raw);                                  // [NG][S][STAGE]
    unsigned long long* full = reinterpret_cast<unsigned long long*>(stage_buf + NG * S * STAGE);   // [NG][S]
    unsigned long long* done = full + NG * S;                                                 // [NG][S]
    unsigned long long* cfull = done + NG * S;                                                // [NG][2]
    double* mat_tab = reinterpret_cast<double*>(cfull + NG * 2);
    int* conn_buf = reinterpret_cast<int*>(mat_tab + HK_SMEM_MATS * 2 * HK_MAX_TABLE);        // [NG][2][8][TLD]
    unsigned short* mat_buf = reinterpret_cast<unsigned short*>(conn_buf + (CP ? NG * 2 * 8 * TLD : 0));   // [NG][2][TLD]
    unsigned char* flag_buf = reinterpret_cast<unsigned char*>(mat_buf + (CP >= 2 ? NG * 2 * TLD : 0));      // [NG][2][TLD]
    uint32_t* tbase_s = reinterpret_cast<uint32_t*>(flag_buf + (CP >= 2 ? NG * 2 * TLD * 2 : 0));
    const HkDev& d = A.d;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int n_tiles = A.n_tiles;

    if (tid == 0) {
        for (int s = 0; s < NG * S; ++s) { mbar_init(&full[s], 1); mbar_init(&done[s], WG); }
        for (int s = 0; s < NG * 2; ++s) mbar_init(&cfull[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tbase_s)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    const bool tabs_in_smem = d.n_mat <= HK_SMEM_MATS;
    if (tabs_in_smem)
        for (int i = tid; i < d.n_mat * 2 * HK_MAX_TABLE; i += Cfg::THREADS) {
            const int m = i / (2 * HK_MAX_TABLE), r = i % (2 * HK_MAX_TABLE);
            mat_tab[i] = r < HK_MAX_TABLE ? d.mats[m].plastic_e[r] : d.mats[m].Hd[r - HK_MAX_TABLE];
        }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tbase = *tbase_s;

    // group of this warp, its virtual CTA id and its share of the tiles (round-robin over all groups of the grid)
    const int g = warp < NG * WG ? warp / WG : warp - NG * WG;
    const int n_v = (int)gridDim.x * NG;
    const int vcta = (int)blockIdx.x * NG + g;
    double* gstage = stage_buf + (long long)g * S * STAGE;
    unsigned long long* gfull = full + g * S;
    unsigned long long* gdone = done + g * S;
    unsigned long long* gcfull = cfull + g * 2;
    int* gconn = conn_buf + (CP ? g * 2 * 8 * TLD : 0);
    unsigned short* gmat = mat_buf + (CP >= 2 ? g * 2 * TLD : 0);
    unsigned char* gflag = flag_buf + (CP >= 2 ? g * 2 * TLD : 0);

    if (warp >= NG * WG) {
        // ===== producer warp of group g =====
        if ((tid & 31) == 0) {
            const long long my_tiles = vcta < n_tiles ? (n_tiles - vcta + n_v - 1) / n_v : 0;
            const long long total_q = my_tiles * 8;             // (tile, Gauss point) work items of this group
            auto issue_conn = [&](long long it) {             // connectivity of this group's tile number `it`
                if (!CP || it >= my_tiles) return;
                const long long e0 = ((long long)vcta + it * n_v) * TLD;
                unsigned long long* bar = &gcfull[it & 1];
                mbar_expect_tx(bar, 8 * TLD * 4 + (CP >= 2 ? TLD * 3 : 0));
#pragma unroll
                for (int a = 0; a < 8; ++a)
                    tma_load_1d(gconn + ((it & 1) * 8 + a) * TLD, d.conn + (long long)a * d.nEp + e0, TLD * 4, bar);
                if (CP >= 2) {                                // material ids and element flags of the tile ride along
                    tma_load_1d(gmat + (it & 1) * TLD, d.mat + e0, TLD * 2, bar);
                    tma_load_1d(gflag + (it & 1) * TLD, d.flag + e0, TLD, bar);
                }
            };
            const unsigned long long pol = EF ? l2_policy_evict_first() : 0ull;
            auto issue_load = [&](long long q) {
                const int st = (int)(q % S);
                const long long e0 = ((long long)vcta + (q >> 3) * n_v) * TLD;
                mbar_expect_tx(&gfull[st], STAGE * 8);
                if (EF) tma_load_1d_hint(gstage + st * STAGE, stage_base(d, (int)(q & 7), e0), STAGE * 8, &gfull[st], pol);
                else tma_load_1d(gstage + st * STAGE, stage_base(d, (int)(q & 7), e0), STAGE * 8, &gfull[st]);
            };
            issue_conn(0);
            for (long long q = 0; q < S && q < total_q; ++q) issue_load(q);
            for (long long q = 0; q < total_q; ++q) {
                const int st = (int)(q % S);
                mbar_wait(&gdone[st], (unsigned)((q / S) & 1));      // every warp of the group finished item q
                const long long e0 = ((long long)vcta + (q >> 3) * n_v) * TLD;
                if (EF) tma_store_1d_hint(stage_base(d, (int)(q & 7), e0), gstage + st * STAGE, STAGE * 8, pol);
                else tma_store_1d(stage_base(d, (int)(q & 7), e0), gstage + st * STAGE, STAGE * 8);
                tma_commit();
                // the group is past the prologue of tile q/8: the buffer of tile q/8 - 1 is free for tile q/8 + 1
                if ((q & 7) == 0) issue_conn((q >> 3) + 1);
                // early refill: wait until THIS store has left the stage (a few hundred ns; the next `done` is a whole item
                // time away, the producer has nothing else to do) and reuse the stage at once.  Round 1 refilled the stage
                // of the PREVIOUS item here (wait_group.read 1, no blocking): one block fewer in flight — measured 3 %
                // slower, and 3 stages with the early refill equal 4 stages with the lagged one
                const long long qn = q + S;
                if (qn < total_q) {
                    tma_wait_read<0>();
                    issue_load(qn);
                }
            }
            tma_wait_all<0>();
        }
    } else {
        // ===== consumers =====
        const int gt = tid - g * TLD;                          // thread within the group = element within the tile
        const uint32_t tcol = tbase + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * HK_TCOLS);
        const uint32_t tX = tcol, tU = tcol + 48, tM = tcol + 96;
        // phase shift: group g > 0 starts when group 0 has finished g*8/NG Gauss points of its first tile
        if (NG > 1 && g > 0) {
            const int tiles0 = (int)blockIdx.x * NG < n_tiles ? 1 : 0;
            const int item = g * 8 / NG - 1;                    // group 0's work item whose completion releases group g
            if (tiles0) mbar_wait(&done[item % S], (unsigned)((item / S) & 1));
        }
        unsigned q = 0;                                        // work item = 8 * (tile number of this group) + Gauss point
        unsigned it = 0;
        for (int tile = vcta; tile < A.n_tiles; tile += n_v, ++it) {      // trip bound from the constant bank
            const long long e0 = (long long)tile * TLD;
            const long long e = e0 + gt;
            if (CP) mbar_wait(&gcfull[it & 1], (unsigned)((it >> 1) & 1));
            const bool live = !element_dead(d, e, CP >= 2 ? (int)gflag[(it & 1) * TLD + gt] : -1);
            const HkMaterialDev* Mt = &d.mats[0];
            double V = 0.125, trbar = 0.0;
            int mi = 0;
            {
                HexModes X, U;
                if (live) {
                    mi = CP >= 2 ? (int)gmat[(it & 1) * TLD + gt] : (int)d.mat[e];
                    if (CP) element_gather_ids<TLD>(d, gconn + (it & 1) * 8 * TLD + gt, X, U);
                    else element_gather(d, e, X, U);
                } else {                                     // unit cube at rest: finite math, state rows unchanged
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        X.c0[c] = X.c1[c] = X.c2[c] = X.h01[c] = X.h02[c] = X.h12[c] = X.h012[c] = 0.0;
                        U.c0[c] = U.c1[c] = U.c2[c] = U.h01[c] = U.h02[c] = U.h12[c] = U.h012[c] = 0.0;
                    }
                    X.c0[0] = X.c1[1] = X.c2[2] = 0.5;
                }
                double G[3][3][3];
                adj_mode_sums(X, G);
                element_volume_terms(X, U, G, V, trbar);
                __syncwarp();
                tmem_store_modes(tX, X);
                tmem_store_modes(tU, U);
            }
            {
                double z[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int i = 0; i < 9; ++i) tmem_st_doubles<4>(tM + 8 * i, z);
            }
            tmem_wait_st();
            Mt = &d.mats[mi];
            MatLite ML = mat_lite(Mt);
            ML.fast = A.fast;
            if (tabs_in_smem) { ML.pe = mat_tab + mi * 2 * HK_MAX_TABLE; ML.hd = ML.pe + HK_MAX_TABLE; }
            const bool need_triax = A.write_triax || Mt->nd > 0;
            double pdet = 0.0, v_e = 0.0, t_e = 0.0;
            int negj = 0;
#pragma unroll 1
            for (int k = 0; k < 8; ++k, ++q) {
                const int st = (int)(q % S);
                double* sb = gstage + st * STAGE + gt;
                mbar_wait(&gfull[st], (unsigned)((q / S) & 1));
                __syncwarp();
#ifdef HK_PROFILE_NOMATH                                     // profiling builds only: memory pipeline without the math
                {
#pragma unroll
                    for (int r = 0; r < HK_ROWS; ++r) sb[r * TLD] = sb[r * TLD] + 0.0;
                    fence_async_smem();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(&gdone[st]);
                    continue;
                }
#endif
                const double s0 = (k & 4) ? 1.0 : -1.0, s1 = (k & 2) ? 1.0 : -1.0, s2 = (k & 1) ? 1.0 : -1.0;
                double Aj[3][3], det;
                {
                    HexModes X;
                    tmem_load_modes(tX, X);
                    gp_geometry(X, s0, s1, s2, Aj, det);
                }
                if (det < 0) negj++;
                const double idet = hk_rcp(det);
                double de[6];
                {
                    HexModes U;
                    tmem_load_modes(tU, U);
                    gp_strain(U, Aj, idet, trbar, s0, s1, s2, de);
                }
                GpStress gs;
                {
                    double sv[14];
#pragma unroll
                    for (int r = 0; r < HK_ROWS; ++r) sv[r] = sb[r * TLD];
                    gp_stress(ML, sv, de, need_triax, gs);
#pragma unroll
                    for (int r = 0; r < HK_ROWS; ++r) sb[r * TLD] = sv[r];
                }
                fence_async_smem();                          // generic-proxy smem writes -> visible to the TMA
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&gdone[st]);   // the stage can go back to HBM while the forces are formed
                pdet += gs.mean * det;
                v_e += gs.ep;
                t_e += gs.tx;
                if (A.write_triax && live) d.triax[(long long)k * d.nEp + e] = gs.tx;
                const double sa[3] = {s1, s0, s0};
                const double sb_[3] = {s2, s2, s1};
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    double T[3];
                    gp_T(gs, Aj[r], T);
                    uint32_t w[24];
                    tmem_ld_x8(tM + 24 * r, w);
                    tmem_ld_x8(tM + 24 * r + 8, w + 8);
                    tmem_ld_x8(tM + 24 * r + 16, w + 16);
                    tmem_wait_ld();
                    const double coef[4] = {1.0, sa[r], sb_[r], sa[r] * sb_[r]};
                    double mo[12];
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int i = m * 3 + c;
                            mo[i] = fma(coef[m], T[c], __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]));
                        }
                    tmem_st_doubles<4>(tM + 24 * r, mo);
                    tmem_st_doubles<4>(tM + 24 * r + 8, mo + 4);
                    tmem_st_doubles<4>(tM + 24 * r + 16, mo + 8);
                }
                tmem_wait_st();
            }
            // ---- element epilogue: forces from the modes parked in TMEM
            __syncwarp();
            ElemAcc acc;
            {
                uint32_t w[72];
#pragma unroll
                for (int i = 0; i < 9; ++i) tmem_ld_x8(tM + 8 * i, w + 8 * i);
                tmem_wait_ld();
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int m = 0; m < 4; ++m)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int i = r * 12 + m * 3 + c;
                            acc.M[r][m][c] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
                        }
            }
            acc.pdet = pdet; acc.v_e = v_e; acc.t_e = t_e; acc.negj = negj;
            HexModes X;
            tmem_load_modes(tX, X);
            if (live) {
                element_finish(A, e, X, acc, V);
                // the material id is read again (shared memory / L1) rather than kept in a register across the Gauss-point
                // loop: it used to be spilled, and its reload was a DRAM-latency stall per tile (ncu, round 2)
                const HkMaterialDev& Md = d.mats[CP >= 2 ? (int)gmat[(it & 1) * TLD + gt] : (int)d.mat[e]];
                if (ductile_check(Md, acc.v_e, acc.t_e)) element_delete(A, e);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tbase));
}

template <int NG, int WG, int S, int CP, int EF = 0>
static int launch_ring(const ElemArgs& A, int n_sm, cudaStream_t s) {
    using Cfg = RingCfg<NG, WG, S, CP>;
    static_assert(Cfg::SMEM <= 227 * 1024, "ring does not fit in shared memory");
    static_assert(((NG * WG + 3) / 4) * HK_TCOLS <= 512, "TMEM columns");
    cudaError_t rc = cudaFuncSetAttribute(hk_element_ring_kernel<NG, WG, S, CP, EF>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          Cfg::SMEM);               // per device: cheap, so set on every launch
    if (rc != cudaSuccess) return (int)rc;
    const long long n_tiles = A.d.nEp / Cfg::TLD;
    long long grid = n_sm;
    if (grid * NG > n_tiles) grid = (n_tiles + NG - 1) / NG;
    ElemArgs B = A;
    B.n_tiles = (int)n_tiles;
    hk_element_ring_kernel<NG, WG, S, CP, EF><<<(unsigned)grid, Cfg::THREADS, Cfg::SMEM, s>>>(B);
    return 0;
}

#endif

// element-kernel variant table: the configurations kept for A/B (profiles/r2_element_kernel_variants.md has the
// measurements of these and of the ones that were tried and dropped).  variant = groups, warps per group, ring stages
// per group, prefetch (0 none, 1 connectivity, 2 connectivity + material ids + flags)
struct RingVariant { int id, ng, wg, stages, cp; };
static const RingVariant kVariants[] = {
    {13, 1, 11, 4, 2},      // default: one group of 11 warps (tile 352), conn + mat + flag prefetched, evict-first L2 policy
#ifdef HK_AB_VARIANTS       // `make ab`: the comparison kernels of scripts/ab_element.py; not in the shipped library
    {11, 1, 11, 4, 0},      // no prefetch
    {12, 1, 11, 4, 1},
    {14, 1, 11, 4, 2},      // default without the L2 policy
    {20, 2, 5, 4, 1},       // two phase-shifted groups of 5 warps (tile 160)
    {25, 2, 5, 4, 2},
#endif
};
#define HK_DEFAULT_VARIANT 13

int hk_element_variant_from_env() {
    const char* v = getenv("HK_ELEMENT_VARIANT");
    int id = v ? atoi(v) : HK_DEFAULT_VARIANT;
    if (const char* k = getenv("HK_ELEMENT_KERNEL")) if (strcmp(k, "simple") == 0) return 1;
    for (const RingVariant& r : kVariants) if (r.id == id) return id;
    return HK_DEFAULT_VARIANT;
}

int hk_launch_element(const HkDev& d, long long step, int write_triax, cudaStream_t s) {
    if (d.element_mode == 1) { hk_launch_element_exact(d, step, write_triax, s); return 0; }
    ElemArgs A{d, step, write_triax, 1, 0};
#ifndef HK_EMU
    switch (d.variant) {
        case 1: {
            const int block = 128;
            hk_element_simple_kernel<<<(unsigned)((d.nElement + block - 1) / block), block, 0, s>>>(A);
            return 0;
        }
#ifdef HK_AB_VARIANTS
        case 11: return launch_ring<1, 11, 4, 0>(A, d.n_sm, s);
        case 12: return launch_ring<1, 11, 4, 1>(A, d.n_sm, s);
        case 14: return launch_ring<1, 11, 4, 2>(A, d.n_sm, s);
        case 20: return launch_ring<2, 5, 4, 1>(A, d.n_sm, s);
        case 25: return launch_ring<2, 5, 4, 2>(A, d.n_sm, s);
#endif
        default: return launch_ring<1, 11, 4, 2, 1>(A, d.n_sm, s);
    }
#else
    for (long long e = 0; e < d.nElement; ++e) element_body_simple(A, e);
    return 0;
#endif
}

// integ_triax_stress recomputed from the current stress (used when no step has written it yet)
void hk_launch_triax(const HkDev& dd, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nElement * 8, s, HK_LAMBDA(long long i) {
        const long long e = i % d.nElement;
        const long long k = i / d.nElement;
        const long long row = k * d.nEp + e;
        double sg[6];
        for (int c = 0; c < 6; ++c) sg[c] = d.ips[hk_ip(d, c, (int)k, e)];
        d.triax[row] = triax_of(sg);
    });
}

// elementVolume[e] = sum_k det_k at the current position (J2:1168-1169)
void hk_launch_element_volume(const HkDev& dd, double* V_out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nElement, s, HK_LAMBDA(long long e) {
        double x[8][3];
        for (int a = 0; a < 8; ++a) {
            const long long n = d.conn[(long long)a * d.nEp + e];
            for (int c = 0; c < 3; ++c) x[a][c] = d.rec[6 * n + c];
        }
        HexModes X;
        hex_modes(x, X);
        double G[3][3][3];
        adj_mode_sums(X, G);
        V_out[e] = dot3(X.c0, G[0][0]) + dot3(X.h01, G[0][1]) + dot3(X.h02, G[0][2]);
    });
}


long long hk_element_tile(int variant) {      // layout tile TL of the ip state = tile of the element kernel that will run
#ifndef HK_EMU
    for (const RingVariant& r : kVariants) if (r.id == variant) return r.wg * 32;
    return 128;                                // simple kernel: any multiple of its block works
#else
    (void)variant;
    return 32;
#endif
}
