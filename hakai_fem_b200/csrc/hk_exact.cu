// hk_exact.cu — kernels whose floating-point operation order mirrors the reference exactly.
// Built with -fmad=false: no FMA contraction, so given identical inputs these kernels reproduce the
// reference's (and the oracle's) IEEE results bit for bit:
//   * nodal kernel: force gather (assembly, J2:668-675) + central difference (J2:562-567)
//                   + boundary conditions (J2:585-617) + kinematics (J2:624-652)
//   * contact: bounding boxes, cell buckets, node-to-triangle penalty forces (J2:2248-2706)
//   * layout transposes and small on-demand taps.
#include "hk_common.h"

// ------------------------------------------------------------------ 128-bit fixed-point accumulators
// Contact forces are summed in the reference in Float128 (J2:435), i.e. effectively exactly, and rounded
// to Float64 once (J2:536-538).  Here each contribution is converted to a 128-bit two's-complement integer
// in units of 2^lsb_exp and added with two 64-bit atomics (carry derived from the low word's old value).
// Integer addition is associative, so the sum is exact AND independent of the order in which threads
// arrive: bitwise reproducible without sorting.
HK_HD void fx_from_double(double v, int lsb_exp, unsigned long long& lo, unsigned long long& hi, bool& ovf) {
    const double two64 = 18446744073709551616.0;
    double mag = ldexp(fabs(v), -lsb_exp);
    ovf = !(mag < 8.5070591730234616e37);        // 2^126; also true for NaN
    if (ovf) { lo = 0; hi = 0; return; }
    double hi_m = floor(mag * (1.0 / two64));
    double lo_m = mag - hi_m * two64;            // exact, in [0, 2^64)
#if defined(__CUDA_ARCH__)
    lo = __double2ull_rz(lo_m);
    hi = __double2ull_rz(hi_m);
#else
    lo = (unsigned long long)lo_m;
    hi = (unsigned long long)hi_m;
#endif
    if (v < 0.0) {
        lo = ~lo + 1ull;
        hi = ~hi + (lo == 0ull ? 1ull : 0ull);
    }
}

HK_HD double fx_to_double(unsigned long long lo, unsigned long long hi, int lsb_exp) {
    bool neg = (hi >> 63) != 0;
    if (neg) {
        lo = ~lo + 1ull;
        hi = ~hi + (lo == 0ull ? 1ull : 0ull);
    }
    if (hi == 0ull && lo == 0ull) return 0.0;
    int lz;
#if defined(__CUDA_ARCH__)
    lz = hi ? __clzll((long long)hi) : 64 + __clzll((long long)lo);
#else
    lz = hi ? __builtin_clzll(hi) : 64 + __builtin_clzll(lo);
#endif
    unsigned long long m, rest;
    if (lz == 0) { m = hi; rest = lo; }
    else if (lz < 64) { m = (hi << lz) | (lo >> (64 - lz)); rest = lo << lz; }
    else if (lz == 64) { m = lo; rest = 0ull; }
    else { m = lo << (lz - 64); rest = 0ull; }
    if (rest) m |= 1ull;                          // sticky bit: 64-bit -> 53-bit RN stays correct
#if defined(__CUDA_ARCH__)
    double r = __ull2double_rn(m);
#else
    double r = (double)m;
#endif
    r = ldexp(r, 64 - lz + lsb_exp);
    return neg ? -r : r;
}

HK_D void fx_atomic_add(unsigned long long* acc, double v, int lsb_exp, unsigned long long* ovf_counter) {
    unsigned long long lo, hi;
    bool ovf;
    fx_from_double(v, lsb_exp, lo, hi, ovf);
    if (ovf) { hk_atomic_add_u64(ovf_counter, 1ull); return; }
    unsigned long long old = hk_atomic_add_u64(&acc[0], lo);
    unsigned long long carry = (old + lo < old) ? 1ull : 0ull;
    hk_atomic_add_u64(&acc[1], hi + carry);
}

// order-preserving encoding of doubles for atomicMin/Max on u64
HK_HD unsigned long long enc_double(double d) {
    unsigned long long b;
    memcpy(&b, &d, 8);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
HK_HD double dec_double(unsigned long long e) {
    unsigned long long b = (e >> 63) ? (e & 0x7fffffffffffffffull) : ~e;
    double d;
    memcpy(&d, &b, 8);
    return d;
}

// ------------------------------------------------------------------ nodal kernel
struct NodalArgs {
    HkDev d;
    double current_time, d_time, dt2, dt2p;
    int lsb_exp, contact_on, use_Q0;
    int mode;                 // 0: every node; 1: all but the halo (interface) nodes; 2: only the nodes in `list`
    const int* list;
    long long n_list;
    unsigned long long* dmax_out;   // clamp mode: bit pattern of max_n sqrt(dx^2+dy^2+dz^2) of d_disp (J1:611-618)
};

HK_HD double eval_amp(const HkDev& d, int amp_id, double current_time) {   // J2:586-600
    const HkAmpTable tb = d.amp_tab[amp_id];
    const double* a_t = d.amp_time + tb.offset;
    const double* a_v = d.amp_value + tb.offset;
    int ti = 0;
    for (int j = 0; j + 1 < tb.n; ++j)
        if (current_time >= a_t[j] && current_time <= a_t[j + 1]) { ti = j; break; }
    return a_v[ti] + (a_v[ti + 1] - a_v[ti]) * (current_time - a_t[ti]) / (a_t[ti + 1] - a_t[ti]);
}

#if defined(__CUDA_ARCH__)
#define HK_LDG(p) __ldg(p)
#else
#define HK_LDG(p) (*(p))
#endif

HK_HD void nodal_body(const NodalArgs& A, long long n) {
    const HkDev& d = A.d;
    // ---- issue every independent load first (the kernel is a pure latency/bandwidth problem) -------------
    int ent[8];
#pragma unroll
    for (int w = 0; w < 8; ++w) ent[w] = (w < d.ell_width) ? HK_LDG(&d.ell[(long long)w * d.nNode + n]) : -1;
    const int si = HK_LDG(&d.spec_idx[n]);
    if (A.mode == 1 && si >= 0 && d.spec[si].halo_slot >= 0) return;    // updated after the halo exchange (mode 2)
    const double M = HK_LDG(&d.mass[n]);
    double u[3], up[3], X[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) { u[c] = d.u[3 * n + c]; up[c] = d.u_pre[3 * n + c]; X[c] = HK_LDG(&d.X[3 * n + c]); }

    // internal force of the node: Q[n] = sum over incident elements in ascending element order (J2:668-675).
    // A missing entry contributes +0.0, which leaves the running sum bit-identical to skipping it.
    double q0 = 0.0, q1 = 0.0, q2 = 0.0;
    if (A.use_Q0) {
        q0 = d.Q0[3 * n]; q1 = d.Q0[3 * n + 1]; q2 = d.Q0[3 * n + 2];
    } else {
        double v[8][3];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            const bool ok = ent[w] >= 0;
            const long long e = ok ? (ent[w] >> 3) : 0;
            const int a = ok ? (ent[w] & 7) : 0;
#pragma unroll
            for (int c = 0; c < 3; ++c) v[w][c] = ok ? HK_LDG(&d.Qe[(long long)(a * 3 + c) * d.nEp + e]) : 0.0;
        }
#pragma unroll
        for (int w = 0; w < 8; ++w) { q0 += v[w][0]; q1 += v[w][1]; q2 += v[w][2]; }
        for (int w = 8; w < d.ell_width; ++w) {          // irregular meshes: more than 8 elements at a node
            const int en = d.ell[(long long)w * d.nNode + n];
            if (en < 0) break;
            const long long e = en >> 3;
            const int a = en & 7;
            q0 += d.Qe[(long long)(a * 3 + 0) * d.nEp + e];
            q1 += d.Qe[(long long)(a * 3 + 1) * d.nEp + e];
            q2 += d.Qe[(long long)(a * 3 + 2) * d.nEp + e];
        }
    }
    double F[3] = {0.0, 0.0, 0.0};
    double bcv[3] = {0.0, 0.0, 0.0};
    bool has_bc[3] = {false, false, false};
    if (si >= 0) {
        const HkSpecialNode sp = d.spec[si];
        if (sp.halo_slot >= 0) {                 // interface node (multi-GPU): the complete sum over all holders, formed in
            q0 = d.halo_recv[3 * sp.halo_slot];  // ascending global-rank order with this rank's own partial in its place
            q1 = d.halo_recv[3 * sp.halo_slot + 1];   // (hk_engine.cu: halo_total) — the same bits on every holder
            q2 = d.halo_recv[3 * sp.halo_slot + 2];
        }
        if (A.contact_on && sp.contact_slot >= 0) {     // external_force += c_force3, J2:536-538
            const unsigned long long* acc = d.cacc + (long long)sp.contact_slot * 6;
            F[0] = fx_to_double(acc[0], acc[1], A.lsb_exp);
            F[1] = fx_to_double(acc[2], acc[3], A.lsb_exp);
            F[2] = fx_to_double(acc[4], acc[5], A.lsb_exp);
        }
        for (int c = 0; c < 3; ++c) {
            int be = sp.bc_entry[c];
            if (be >= 0) {
                double amp = 1.0;
                int aid = d.bc_amp[be];
                // a step replayed from a CUDA graph reads its number from the device: same product as the host's (double)t * dt
                if (aid >= 0) amp = eval_amp(d, aid, d.t_dev ? (double)(*d.t_dev) * A.d_time : A.current_time);
                bcv[c] = d.bc_value[be] * amp;    // disp_new[dof] .= v * amp, J2:612
                has_bc[c] = true;
            }
        }
    }
    // central difference, J2:564 (diag_C == 0: its terms vanish exactly)
    const double a = M / A.dt2;
    const double a2 = M / A.dt2p;
    const double inva = 1.0 / a;
    const double q[3] = {q0, q1, q2};
    double un[3], dd[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        un[c] = inva * (F[c] - q[c] + a2 * (2.0 * u[c] - up[c]));
        if (has_bc[c]) un[c] = bcv[c];
        dd[c] = un[c] - u[c];                    // d_disp, J2:625
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long i = 3 * n + c;
        d.u_pre[i] = u[c];
        d.u[i] = un[c];
        if (A.contact_on) d.velo[i] = dd[c] / A.d_time;   // velo, J2:628
    }
    if (A.dmax_out) {                            // d_disp_norm, J1:611-616; the max over nodes is order independent
        const double nrm = sqrt(dd[0] * dd[0] + dd[1] * dd[1] + dd[2] * dd[2]);
        unsigned long long bits;
        memcpy(&bits, &nrm, 8);
#if defined(__CUDA_ARCH__) || defined(HK_EMU)
        if (bits > *A.dmax_out) hk_atomic_max_u64(A.dmax_out, bits);       // racy pre-check only skips losers
#endif
    }
    // node record {position (J2:650-652), d_disp}: 48 contiguous bytes
#if defined(__CUDA_ARCH__)
    double2* r = reinterpret_cast<double2*>(d.rec + 6 * n);
    r[0] = make_double2(X[0] + un[0], X[1] + un[1]);
    r[1] = make_double2(X[2] + un[2], dd[0]);
    r[2] = make_double2(dd[1], dd[2]);
#else
    for (int c = 0; c < 3; ++c) { d.rec[6 * n + c] = X[c] + un[c]; d.rec[6 * n + 3 + c] = dd[c]; }
#endif
}

#ifndef HK_EMU
__global__ void __launch_bounds__(256) hk_nodal_kernel(NodalArgs A) {
    long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (A.mode == 2) {
        if (n < A.n_list) nodal_body(A, A.list[n]);
    } else if (n < A.d.nNode) {
        nodal_body(A, n);
    }
}
#endif

void hk_launch_nodal(const HkDev& d, double current_time, double d_time, double dt2, double dt2p, int lsb_exp,
                     int contact_on, int use_Q0, int mode, const int* list, long long n_list, unsigned long long* dmax_out,
                     cudaStream_t s) {
    NodalArgs A{d, current_time, d_time, dt2, dt2p, lsb_exp, contact_on, use_Q0, mode, list, n_list, dmax_out};
    const long long n = mode == 2 ? n_list : d.nNode;
    if (n <= 0) return;
#ifndef HK_EMU
    const int block = 256;
    hk_nodal_kernel<<<(unsigned)((n + block - 1) / block), block, 0, s>>>(A);
#else
    for (long long i = 0; i < n; ++i) nodal_body(A, mode == 2 ? (long long)list[i] : i);
#endif
}

// ------------------------------------------------------------------ contact (J2:2248-2706)
HK_HD double my3norm(double b1, double b2, double b3) { return sqrt(b1 * b1 + b2 * b2 + b3 * b3); }

struct ContactArgs {
    HkDev d;
    HkPairDev p;
    HkContactParams cp;
};

// bounding boxes of the i nodes and j nodes (J2:2284-2296); serial form used by the emu build only
#ifdef HK_EMU
HK_HD void contact_bbox_body(const ContactArgs& A, long long t) {
    const HkPairDev& p = A.p;
    const bool is_i = t < p.dyn->nn_i;
    const int node = is_i ? p.nodes_i[t] : p.nodes_j[t - p.dyn->nn_i];
    unsigned long long* bb = p.bbox + (is_i ? 0 : 6);
    for (int a = 0; a < 3; ++a) {
        unsigned long long e = enc_double(A.d.rec[6ll * node + a]);
        hk_atomic_min_u64(&bb[a], e);
        hk_atomic_max_u64(&bb[3 + a], e);
    }
}
#endif

struct PairBox {
    double range_min[3], range_max[3], all_min[3];
    bool skip;
};
HK_HD PairBox pair_box(const HkPairDev& p) {                      // J2:2298-2315
    PairBox b;
    b.skip = false;
    for (int a = 0; a < 3; ++a) {
        double mn_i = dec_double(p.bbox[a]), mx_i = dec_double(p.bbox[3 + a]);
        double mn_j = dec_double(p.bbox[6 + a]), mx_j = dec_double(p.bbox[9 + a]);
        b.range_min[a] = mn_i > mn_j ? mn_i : mn_j;
        b.range_max[a] = mx_i < mx_j ? mx_i : mx_j;
        b.all_min[a] = mn_i < mn_j ? mn_i : mn_j;
        if (b.range_min[a] > b.range_max[a]) b.skip = true;
    }
    return b;
}

HK_HD unsigned cell_hash(int cx, int cy, int cz) {
    return ((unsigned)cx * 73856093u) ^ ((unsigned)cy * 19349663u) ^ ((unsigned)cz * 83492791u);
}

// cell coordinates of the i nodes (J2:2337-2349) and insertion into hashed buckets
HK_D void contact_cells_body(const ContactArgs& A, long long k) {
    const HkPairDev& p = A.p;
    const PairBox b = pair_box(p);
    if (b.skip) return;
    const double ddiv = p.self ? A.cp.ddiv_s : A.cp.ddiv_o;
    const int node = p.nodes_i[k];
    int c[3];
    for (int a = 0; a < 3; ++a) c[a] = (int)ceil((A.d.rec[6ll * node + a] - b.all_min[a]) / ddiv);
    p.cell_i[k] = c[0];
    p.cell_i[p.cap_i + k] = c[1];
    p.cell_i[2ll * p.cap_i + k] = c[2];
    unsigned h = cell_hash(c[0], c[1], c[2]) & (unsigned)(p.dyn->n_bucket - 1);
    p.next[k] = hk_atomic_exch_i32(&p.head[h], (int)k);
}

// one master triangle against the slave nodes of the 27 neighbouring cells (J2:2370-2692), in two parts: everything
// that depends on the triangle only (flag, culls against the overlap box, normal, area, the node-independent half of
// my3SolveAb, the cell of vertex j0) and the walk through ONE of the 27 cells.  The serial form does both for all cells;
// on the GPU one thread per triangle culls and a warp per surviving triangle walks its cells in parallel — the force sums
// are exact integers, so the order of the hits does not matter.
struct TriCtx {
    int eleid, j0, j1, j2;
    double q0x, q0y, q0z, cx, cy, cz, Rmax, Lmax, S, nx, ny, nz, detA;
    double im11, im12, im13, im21, im22, im23, im31, im32, im33;
    int cj[3];
    PairBox b;
};

HK_D bool contact_tri_prepare(const ContactArgs& A, long long j, TriCtx& T) {
    const HkDev& d = A.d;
    const HkPairDev& p = A.p;
    T.eleid = p.tele[j];
    if (d.flag[T.eleid] != 1) return false;                       // J2:2373-2376
    T.b = pair_box(p);
    const PairBox& b = T.b;
    if (b.skip) return false;
    const double ddiv = p.self ? A.cp.ddiv_s : A.cp.ddiv_o;
    const int j0 = p.t0[j], j1 = p.t1[j], j2 = p.t2[j];
    T.j0 = j0; T.j1 = j1; T.j2 = j2;
    const double q0x = d.rec[6ll * j0], q0y = d.rec[6ll * j0 + 1], q0z = d.rec[6ll * j0 + 2];
    const double q1x = d.rec[6ll * j1], q1y = d.rec[6ll * j1 + 1], q1z = d.rec[6ll * j1 + 2];
    const double q2x = d.rec[6ll * j2], q2y = d.rec[6ll * j2 + 1], q2z = d.rec[6ll * j2 + 2];
    if (q0x < b.range_min[0] && q1x < b.range_min[0] && q2x < b.range_min[0]) return false;
    if (q0y < b.range_min[1] && q1y < b.range_min[1] && q2y < b.range_min[1]) return false;
    if (q0z < b.range_min[2] && q1z < b.range_min[2] && q2z < b.range_min[2]) return false;
    if (q0x > b.range_max[0] && q1x > b.range_max[0] && q2x > b.range_max[0]) return false;
    if (q0y > b.range_max[1] && q1y > b.range_max[1] && q2y > b.range_max[1]) return false;
    if (q0z > b.range_max[2] && q1z > b.range_max[2] && q2z > b.range_max[2]) return false;
    T.q0x = q0x; T.q0y = q0y; T.q0z = q0z;
    const double cx = (q0x + q1x + q2x) / 3.0, cy = (q0y + q1y + q2y) / 3.0, cz = (q0z + q1z + q2z) / 3.0;
    T.cx = cx; T.cy = cy; T.cz = cz;
    const double R0 = my3norm(q0x - cx, q0y - cy, q0z - cz);
    const double R1 = my3norm(q1x - cx, q1y - cy, q1z - cz);
    const double R2 = my3norm(q2x - cx, q2y - cy, q2z - cz);
    T.Rmax = fmax(fmax(R0, R1), R2);
    const double v1x = q1x - q0x, v1y = q1y - q0y, v1z = q1z - q0z;
    const double v2x = q2x - q0x, v2y = q2y - q0y, v2z = q2z - q0z;
    const double L1 = my3norm(v1x, v1y, v1z), L2 = my3norm(v2x, v2y, v2z);
    T.Lmax = fmax(L1, L2);
    double nx = v1y * v2z - v1z * v2y;                            // my3crossNNz, J2:3209
    double ny = v1z * v2x - v1x * v2z;
    double nz = v1x * v2y - v1y * v2x;
    const double mag_n = sqrt(nx * nx + ny * ny + nz * nz);
    nx = nx / mag_n; ny = ny / mag_n; nz = nz / mag_n;
    T.nx = nx; T.ny = ny; T.nz = nz;
    const double d12 = v1x * v2x + v1y * v2y + v1z * v2z;
    T.S = 0.5 * sqrt(L1 * L1 * L2 * L2 - d12 * d12);
    const double A11 = v1x, A21 = v1y, A31 = v1z, A12 = v2x, A22 = v2y, A32 = v2z;
    const double A13 = -nx, A23 = -ny, A33 = -nz;
    // my3SolveAb, J2:3342: the parts that do not depend on the node
    T.detA = (A11 * A22 * A33 + A12 * A23 * A31 + A13 * A21 * A32 - A11 * A23 * A32 - A12 * A21 * A33 -
              A13 * A22 * A31);
    T.im11 = A22 * A33 - A23 * A32; T.im21 = A23 * A31 - A21 * A33; T.im31 = A21 * A32 - A22 * A31;
    T.im12 = A13 * A32 - A12 * A33; T.im22 = A11 * A33 - A13 * A31; T.im32 = A12 * A31 - A11 * A32;
    T.im13 = A12 * A23 - A13 * A22; T.im23 = A13 * A21 - A11 * A23; T.im33 = A11 * A22 - A12 * A21;
    // cell of vertex j0 (J2:2351-2363, 2462-2472): same formula, evaluated directly on j0's position
    T.cj[0] = (int)ceil((q0x - b.all_min[0]) / ddiv);
    T.cj[1] = (int)ceil((q0y - b.all_min[1]) / ddiv);
    T.cj[2] = (int)ceil((q0z - b.all_min[2]) / ddiv);
    return true;
}

// slave nodes of cell (cj + (dx, dy, dz)) against the prepared triangle
HK_D void contact_tri_cell(const ContactArgs& A, const TriCtx& T, int dx, int dy, int dz, unsigned long long& n_tests,
                           unsigned long long& n_hits) {
    const HkDev& d = A.d;
    const HkPairDev& p = A.p;
    const HkContactParams& cp = A.cp;
    const PairBox& b = T.b;
    const double kc = p.self ? cp.kc_s : cp.kc_o;
    const double Cr = p.self ? cp.cr_s : cp.cr_o;
    const int j0 = T.j0;
    const double nx = T.nx, ny = T.ny, nz = T.nz;
    const unsigned bucket_mask = (unsigned)(p.dyn->n_bucket - 1);
    const int ccx = T.cj[0] + dx, ccy = T.cj[1] + dy, ccz = T.cj[2] + dz;
    const unsigned h = cell_hash(ccx, ccy, ccz) & bucket_mask;
    for (int k = p.head[h]; k >= 0; k = p.next[k]) {
        if (p.cell_i[k] != ccx || p.cell_i[p.cap_i + k] != ccy || p.cell_i[2ll * p.cap_i + k] != ccz) continue;
        const int i = p.nodes_i[k];
        if (p.self) {
            bool own = false;
            for (int q = 0; q < 8; ++q) own = own || (i == d.conn[(long long)q * d.nEp + T.eleid]);
            if (own) continue;
        }
        const double px = d.rec[6ll * i], py = d.rec[6ll * i + 1], pz = d.rec[6ll * i + 2];
        if (px < b.range_min[0] || py < b.range_min[1] || pz < b.range_min[2]) continue;
        if (px > b.range_max[0] || py > b.range_max[1] || pz > b.range_max[2]) continue;
        const double dpc = my3norm(px - T.cx, py - T.cy, pz - T.cz);
        if (dpc >= T.Rmax) continue;
        const double bx = px - T.q0x, by = py - T.q0y, bz = pz - T.q0z;
        ++n_tests;
        const double x1 = (T.im11 * bx + T.im12 * by + T.im13 * bz) / T.detA;
        const double x2 = (T.im21 * bx + T.im22 * by + T.im23 * bz) / T.detA;
        const double dd = (T.im31 * bx + T.im32 * by + T.im33 * bz) / T.detA;
        if (0.0 <= x1 && 0.0 <= x2 && x1 + x2 <= 1.0 && dd > 0.0 && dd <= cp.d_lim) {
            ++n_hits;
            const int slot_i = d.spec[d.spec_idx[i]].contact_slot;
            double dcl = dd;
            if (cp.clamp) {                           // J1:2756-2758
                double dmax;
                const unsigned long long mb = *cp.dmax;
                memcpy(&dmax, &mb, 8);
                const double pre = cp.dnode_pre[slot_i];
                if (dcl - pre > dmax) dcl = pre + dmax;
            }
            const double vx = d.velo[3ll * i] - d.velo[3ll * j0];
            const double vy = d.velo[3ll * i + 1] - d.velo[3ll * j0 + 1];
            const double vz = d.velo[3ll * i + 2] - d.velo[3ll * j0 + 2];
            const double mag_v = my3norm(vx, vy, vz);
            double vex = 0.0, vey = 0.0, vez = 0.0;
            if (mag_v > 0.0) { vex = vx / mag_v; vey = vy / mag_v; vez = vz / mag_v; }
            const double k_ = p.young * T.S / T.Lmax * kc;
            const double F = k_ * dcl;
            double fx = F * nx, fy = F * ny, fz = F * nz;
            // damping: diag_M[i] is indexed with the NODE id in the reference (J2:2593)
            const double C = 2 * sqrt(d.mass[i / 3] * k_) * Cr;
            const double fc_x = -C * vx, fc_y = -C * vy, fc_z = -C * vz;
            const double dot_ve_n = vex * nx + vey * ny + vez * nz;
            const double vsx = vex - dot_ve_n * nx, vsy = vey - dot_ve_n * ny, vsz = vez - dot_ve_n * nz;
            const double fric_x = -cp.myu * F * vsx, fric_y = -cp.myu * F * vsy, fric_z = -cp.myu * F * vsz;
            fx += fric_x + fc_x;
            fy += fric_y + fc_y;
            fz += fric_z + fc_z;
            const double f[3] = {fx, fy, fz};
            const double f3[3] = {-fx / 3.0, -fy / 3.0, -fz / 3.0};
            unsigned long long* ovf = &d.counters[3];
            for (int c = 0; c < 3; ++c) fx_atomic_add(d.cacc + 6ll * slot_i + 2 * c, f[c], cp.lsb_exp, ovf);
            if (cp.clamp) {                           // d_node[i] = max(d_node[i], d), J1:2898-2900
                unsigned long long bits;
                memcpy(&bits, &dcl, 8);
                hk_atomic_max_u64(reinterpret_cast<unsigned long long*>(cp.dnode + slot_i), bits);
            }
            const int jn[3] = {j0, T.j1, T.j2};
            for (int q = 0; q < 3; ++q) {
                const int slot = d.spec[d.spec_idx[jn[q]]].contact_slot;
                for (int c = 0; c < 3; ++c) fx_atomic_add(d.cacc + 6ll * slot + 2 * c, f3[c], cp.lsb_exp, ovf);
            }
        }
    }
}

// serial form: one triangle, all 27 cells (the host-compiled build; the order the reference visits them in)
HK_D void contact_tri_body(const ContactArgs& A, long long j) {
    TriCtx T;
    if (!contact_tri_prepare(A, j, T)) return;
    unsigned long long n_tests = 0, n_hits = 0;
    for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
            for (int dx = -1; dx <= 1; ++dx) contact_tri_cell(A, T, dx, dy, dz, n_tests, n_hits);
    if (n_tests) hk_atomic_add_u64(&A.d.counters[2], n_tests);
    if (n_hits) hk_atomic_add_u64(&A.d.counters[1], n_hits);
}

#ifndef HK_EMU
// All contact kernels are grid-stride over list lengths that live on the device (HkPairDyn): the lists grow when
// deleted elements expose new faces (erode_element), and the host never has to know by how much.
__global__ void hk_contact_bbox_kernel(ContactArgs A) {
    const long long nn_i = A.p.dyn->nn_i;
    const long long n = nn_i + A.p.dyn->nn_j;
    const unsigned FULL = 0xffffffffu;
    const long long stride = (long long)gridDim.x * blockDim.x;
    unsigned long long mn[2][3], mx[2][3];           // [side][axis] running min / max of this thread
    for (int sd = 0; sd < 2; ++sd)
        for (int a = 0; a < 3; ++a) { mn[sd][a] = ~0ull; mx[sd][a] = 0ull; }
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += stride) {
        const int sd = t < nn_i ? 0 : 1;
        const int node = sd == 0 ? A.p.nodes_i[t] : A.p.nodes_j[t - nn_i];
        for (int a = 0; a < 3; ++a) {
            const unsigned long long en = enc_double(A.d.rec[6ll * node + a]);
            mn[sd][a] = en < mn[sd][a] ? en : mn[sd][a];
            mx[sd][a] = en > mx[sd][a] ? en : mx[sd][a];
        }
    }
    // warp-level min/max (warp-shuffle reduction), then one atomic set per (warp, side)
    for (int sd = 0; sd < 2; ++sd) {
        for (int off = 16; off; off >>= 1)
            for (int a = 0; a < 3; ++a) {
                const unsigned long long o1 = __shfl_xor_sync(FULL, mn[sd][a], off);
                const unsigned long long o2 = __shfl_xor_sync(FULL, mx[sd][a], off);
                mn[sd][a] = o1 < mn[sd][a] ? o1 : mn[sd][a];
                mx[sd][a] = o2 > mx[sd][a] ? o2 : mx[sd][a];
            }
        if ((threadIdx.x & 31) == 0 && mn[sd][0] != ~0ull) {
            unsigned long long* bb = A.p.bbox + (sd == 0 ? 0 : 6);
            for (int a = 0; a < 3; ++a) { atomicMin(&bb[a], mn[sd][a]); atomicMax(&bb[3 + a], mx[sd][a]); }
        }
    }
}
__global__ void hk_contact_cells_kernel(ContactArgs A) {
    const long long n = A.p.dyn->nn_i, stride = (long long)gridDim.x * blockDim.x;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < n; k += stride) contact_cells_body(A, k);
}
// narrow phase in two kernels.  cull: one thread per master triangle — dead element, overlap-box culls (J2:2373-2440);
// the survivors (the triangles of the contact zone: thousands out of a million) are compacted into p.cand.
// narrow: one WARP per surviving triangle, lane c < 27 walks cell c's bucket chain.  One thread walking 27 chains of
// dependent loads per triangle took 100-200 us per pair at 4 warps active per SM (ncu, round 2): the few triangles in
// the contact zone set the kernel's time.
__global__ void __launch_bounds__(256) hk_contact_cull_kernel(ContactArgs A) {
    const long long n = A.p.dyn->nTri, stride = (long long)gridDim.x * blockDim.x;
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += stride) {
        const HkDev& d = A.d;
        const HkPairDev& p = A.p;
        if (d.flag[p.tele[j]] != 1) continue;
        const PairBox b = pair_box(p);
        if (b.skip) continue;
        const int j0 = p.t0[j], j1 = p.t1[j], j2 = p.t2[j];
        bool out = false;
        for (int a = 0; a < 3; ++a) {
            const double q0 = d.rec[6ll * j0 + a], q1 = d.rec[6ll * j1 + a], q2 = d.rec[6ll * j2 + a];
            out = out || (q0 < b.range_min[a] && q1 < b.range_min[a] && q2 < b.range_min[a]) ||
                  (q0 > b.range_max[a] && q1 > b.range_max[a] && q2 > b.range_max[a]);
        }
        if (!out) p.cand[atomicAdd(&p.dyn->n_cand, 1)] = (int)j;
    }
}
__global__ void __launch_bounds__(128) hk_contact_narrow_kernel(ContactArgs A) {
    const int n = A.p.dyn->n_cand;
    const int lane = threadIdx.x & 31;
    const int warps = (int)(gridDim.x * (blockDim.x >> 5));
    for (int w = (int)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)); w < n; w += warps) {
        TriCtx T;
        if (!contact_tri_prepare(A, A.p.cand[w], T)) continue;       // (never: the cull kernel applied the same tests)
        unsigned long long n_tests = 0, n_hits = 0;
        if (lane < 27) contact_tri_cell(A, T, lane % 3 - 1, (lane / 3) % 3 - 1, lane / 9 - 1, n_tests, n_hits);
        if (n_tests) hk_atomic_add_u64(&A.d.counters[2], n_tests);
        if (n_hits) hk_atomic_add_u64(&A.d.counters[1], n_hits);
    }
}
__global__ void hk_contact_reset_kernel(HkPairDev p) {
    const long long n = p.dyn->n_bucket, stride = (long long)gridDim.x * blockDim.x;
    for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < n; t += stride) p.head[t] = -1;
    if (blockIdx.x == 0 && threadIdx.x < 12) p.bbox[threadIdx.x] = ((threadIdx.x % 6) < 3) ? ~0ull : 0ull;
    if (blockIdx.x == 0 && threadIdx.x == 12) p.dyn->n_cand = 0;
}
static unsigned contact_grid(long long cap, int block, int n_sm) {
    long long g = (cap + block - 1) / block, mx = (long long)(n_sm > 0 ? n_sm : 148) * 32;
    if (g > mx) g = mx;                     // larger lists: grid-stride
    return (unsigned)(g < 1 ? 1 : g);
}
#endif

void hk_launch_contact(const HkDev& d, const HkPairDev& p, const HkContactParams& cp, cudaStream_t s) {
    if (p.cap_i == 0 || p.cap_j == 0 || p.cap_tri == 0) return;
    ContactArgs A{d, p, cp};
#ifndef HK_EMU
    hk_contact_reset_kernel<<<contact_grid(p.cap_bucket, 256, d.n_sm), 256, 0, s>>>(p);
    {   // few blocks, several nodes per thread: the warp reduction (120 shuffles) and the 12 atomics are per warp
        unsigned g = contact_grid((long long)p.cap_i + p.cap_j, 256, d.n_sm);
        const unsigned cap = (unsigned)(d.n_sm > 0 ? d.n_sm : 148) * 2;
        hk_contact_bbox_kernel<<<g < cap ? g : cap, 256, 0, s>>>(A);
    }
    hk_contact_cells_kernel<<<contact_grid(p.cap_i, 256, d.n_sm), 256, 0, s>>>(A);
    hk_contact_cull_kernel<<<contact_grid(p.cap_tri, 256, d.n_sm), 256, 0, s>>>(A);
    // a warp per candidate; the grid covers every triangle the pair could hold (warps without a candidate exit at once),
    // so that a small deck's few dozen candidates all walk their cells at the same time
    hk_contact_narrow_kernel<<<contact_grid((long long)p.cap_tri * 32, 128, d.n_sm), 128, 0, s>>>(A);
#else
    for (int t = 0; t < p.dyn->n_bucket; ++t) p.head[t] = -1;
    for (int t = 0; t < 12; ++t) p.bbox[t] = ((t % 6) < 3) ? ~0ull : 0ull;
    for (long long t = 0; t < (long long)p.dyn->nn_i + p.dyn->nn_j; ++t) contact_bbox_body(A, t);
    for (long long k = 0; k < p.dyn->nn_i; ++k) contact_cells_body(A, k);
    for (long long j = 0; j < p.dyn->nTri; ++j) contact_tri_body(A, j);
#endif
}

// zero the contact-force accumulators in use (their number grows on the device with the contact surface)
void hk_launch_cacc_zero(const HkDev& dd, const int* n_slots, int slot_cap, cudaStream_t s) {
    const HkDev d = dd;
#ifndef HK_EMU
    const long long cap = (long long)slot_cap * 6;
    long long grid = (cap + 255) / 256, mx = (long long)(d.n_sm > 0 ? d.n_sm : 148) * 16;
    if (grid > mx) grid = mx;
    if (grid < 1) return;
    hk_generic_kernel<<<(unsigned)grid, 256, 0, s>>>(grid * 256, HK_LAMBDA(long long t) {
        const long long n = (long long)(*n_slots) * 6, stride = (long long)gridDim.x * blockDim.x;
        for (long long i = t; i < n; i += stride) d.cacc[i] = 0ull;
    });
#else
    (void)slot_cap; (void)s;
    for (long long i = 0; i < (long long)(*n_slots) * 6; ++i) d.cacc[i] = 0ull;
#endif
}

// ------------------------------------------------------------------ deletion pass + exposed faces on the device (A9/A10)
// Step 1 (count):  per block of HK_DEL_BLOCK elements, how many the element kernel marked for deletion (flag 3).
// Step 2 (scan + emit): an exclusive scan of the block counts (one CTA), then blocks holding marks append their elements
//                  to the deletion log in ASCENDING id order (offset = log length before this step + the block's scan
//                  value), zero stress/strain (J2:742-756) and set flag 0.  The log is therefore in the reference's
//                  deletion order (ascending step, ascending id within a step, J2:701-735) without any sort, and the
//                  element kernel needs no atomic.
// Step 3 (finish): ONE thread; with contact it replays the reference's serial loop J2:767-804 over the step's entries:
//                  for each face of a deleted element the twin face (precomputed at the first step) becomes two
//                  master triangles of every pair whose j instance is the element's instance, and its nodes join
//                  c_nodes_i / c_nodes_j, in the reference's order; then it advances the log length.  Deletions per
//                  step are few, the replay is inherently ordered, and nothing leaves the device.
#define HK_DEL_BLOCK 1024

HK_HD void flush_element(const HkDev& d, long long e) {
    const long long TL = d.TL, t = e / TL;                 // one division per element (hk_ip would do 96)
    double* p = d.ips + t * 8 * 14 * TL + (e - t * TL);     // (row 0, Gauss point 0) of element e: hk_ip(d, 0, 0, e)
    for (int k = 0; k < 8; ++k)
        for (int r = 0; r < 12; ++r) p[(long long)(k * 14 + r) * TL] = 0.0;
    d.flag[e] = 0;
}

HK_HD void erode_contact_slot(const HkDev& d, const HkErodeDev& E, int node) {
    int si = d.spec_idx[node];
    if (si < 0) {
        if (*E.n_spec >= E.spec_cap) { *E.overflow = 1; return; }
        si = (*E.n_spec)++;
        HkSpecialNode sn;
        sn.bc_entry[0] = sn.bc_entry[1] = sn.bc_entry[2] = -1;
        sn.contact_slot = -1; sn.halo_slot = -1; sn.pad = 0;
        d.spec[si] = sn;
        d.spec_idx[node] = si;
    }
    if (d.spec[si].contact_slot < 0) {
        if (*E.n_slots >= E.slot_cap) { *E.overflow = 1; return; }
        d.spec[si].contact_slot = (*E.n_slots)++;
    }
}

// add_surface_triangle (J2:2167-2245) + the pair loop J2:778-801 for ONE deleted element
HK_HD void erode_element(const HkDev& d, const HkErodeDev& E, int e) {
    const int inst = E.einst[e];
    if (inst < 1 || inst > E.n_inst) return;
    const HkInstDev& I = E.inst[inst - 1];
    if (!I.twin) return;
    const long long F = 6 * I.nElement;
    int tri[36], tele[12], nodes[24];
    int nt = 0, nn = 0;
    for (int j = 0; j < 6; ++j) {
        const long long k = I.twin[6 * (e - I.element_offset) + j];
        if (k < 0) continue;
        const int s0 = I.surf[k], s1 = I.surf[k + F], s2 = I.surf[k + 2 * F], s3 = I.surf[k + 3 * F];
        tri[3 * nt] = s0; tri[3 * nt + 1] = s1; tri[3 * nt + 2] = s2; tele[nt++] = I.feleid[k];
        tri[3 * nt] = s2; tri[3 * nt + 1] = s3; tri[3 * nt + 2] = s0; tele[nt++] = I.feleid[k];
        const int q[4] = {s0, s1, s2, s3};
        for (int a = 0; a < 4; ++a) {                 // sorted unique insert (`nodes = unique(sort(tri))`), global-id order
            int pos = 0;
            const int kq = E.node_key ? E.node_key[q[a]] : q[a];
            while (pos < nn && (E.node_key ? E.node_key[nodes[pos]] : nodes[pos]) < kq) ++pos;
            if (pos < nn && nodes[pos] == q[a]) continue;
            for (int m = nn; m > pos; --m) nodes[m] = nodes[m - 1];
            nodes[pos] = q[a];
            ++nn;
        }
    }
    if (nt == 0) return;
    for (int c = 0; c < E.n_pair; ++c) {
        const HkPairDev& p = E.pairs[c];
        if (!p.in_i) continue;                        // this pair's surface cannot erode
        HkPairDyn& D = *p.dyn;
        if (p.i_instance == inst) {                   // J2:784-787
            for (int m = 0; m < nn; ++m) {
                const int g = nodes[m];
                if (p.in_i[g]) continue;
                if (D.nn_i >= p.cap_i) { *E.overflow = 1; continue; }
                p.in_i[g] = 1;
                p.nodes_i[D.nn_i++] = g;
                erode_contact_slot(d, E, g);
            }
            while (D.n_bucket < 2 * D.nn_i && D.n_bucket < p.cap_bucket) D.n_bucket <<= 1;
        } else if (p.j_instance == inst) {            // J2:789-797
            for (int m = 0; m < nn; ++m) {
                const int g = nodes[m];
                if (p.in_j[g]) continue;
                if (D.nn_j >= p.cap_j) { *E.overflow = 1; continue; }
                p.in_j[g] = 1;
                p.nodes_j[D.nn_j++] = g;
                erode_contact_slot(d, E, g);
            }
            for (int r = 0; r < nt; ++r) {
                if (tele[r] < 0) continue;            // the rank that owns that element adds this triangle
                if (D.nTri >= p.cap_tri) { *E.overflow = 1; break; }
                const int o = D.nTri++;
                p.t0[o] = tri[3 * r]; p.t1[o] = tri[3 * r + 1]; p.t2[o] = tri[3 * r + 2]; p.tele[o] = tele[r];
            }
        }
    }
}

#ifndef HK_EMU
__global__ void __launch_bounds__(256) hk_delete_count_kernel(HkDev d) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const long long e0 = (long long)blockIdx.x * HK_DEL_BLOCK;
    int mine = 0;
    if (e0 + HK_DEL_BLOCK <= d.nElement) {                  // whole block inside the mesh: four flags per load
        const unsigned w = reinterpret_cast<const unsigned*>(d.flag + e0)[threadIdx.x];
        mine = ((w & 0xffu) == 3u) + (((w >> 8) & 0xffu) == 3u) + (((w >> 16) & 0xffu) == 3u) + ((w >> 24) == 3u);
    } else {
        for (int i = threadIdx.x; i < HK_DEL_BLOCK; i += 256) {
            const long long e = e0 + i;
            if (e < d.nElement && d.flag[e] == 3) ++mine;
        }
    }
    if (mine) atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) d.del_block[blockIdx.x] = cnt;
}
// exclusive scan of the per-block counts (one CTA; the mesh has nElement / 1024 blocks) -> del_block[b] becomes the
// offset of block b inside this step's entries, *del_fresh the step's total
__global__ void __launch_bounds__(1024) hk_delete_scan_kernel(HkDev d, int nb) {
    __shared__ int warp_sum[32];
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nb ? d.del_block[i] : 0;
        int x = v;
        for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, x, off); if (lane >= off) x += y; }
        if (lane == 31) warp_sum[w] = x;
        __syncthreads();
        if (w == 0) {
            int s = warp_sum[lane];
            for (int off = 1; off < 32; off <<= 1) { const int y = __shfl_up_sync(0xffffffffu, s, off); if (lane >= off) s += y; }
            warp_sum[lane] = s;
        }
        __syncthreads();
        const int carry = carry_s;
        const int incl = x + (w ? warp_sum[w - 1] : 0);
        if (i < nb) d.del_block[i] = carry + incl - v;          // exclusive offset; the count is recovered by the emit kernel
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *d.del_fresh = carry_s;
}
__global__ void __launch_bounds__(256) hk_delete_emit_kernel(HkDev d, long long step, int nb) {
    if (*d.del_fresh == 0) return;                             // nothing was deleted in this step (the common case)
    if (d.t_dev) step = *d.t_dev;
    const int off0 = d.del_block[blockIdx.x];
    const int off1 = (int)blockIdx.x + 1 < nb ? d.del_block[blockIdx.x + 1] : *d.del_fresh;
    if (off1 == off0) return;
    __shared__ int warp_tot[8];
    const long long e0 = (long long)blockIdx.x * HK_DEL_BLOCK;
    int run = *d.del_count + off0;                            // the log length is advanced by the finish kernel
    for (int chunk = 0; chunk < HK_DEL_BLOCK; chunk += 256) {      // ordered compaction, 256 elements at a time
        const long long e = e0 + chunk + threadIdx.x;
        const bool m = e < d.nElement && d.flag[e] == 3;
        const unsigned bal = __ballot_sync(0xffffffffu, m);
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (lane == 0) warp_tot[w] = __popc(bal);
        __syncthreads();
        int before = 0, total = 0;
        for (int i = 0; i < 8; ++i) { if (i < w) before += warp_tot[i]; total += warp_tot[i]; }
        if (m) {
            const int slot = run + before + __popc(bal & ((1u << lane) - 1u));
            if (slot < d.del_cap) d.del_list[slot] = (step << 32) | e;
            flush_element(d, e);
        }
        run += total;
        __syncthreads();
    }
}
__global__ void hk_delete_finish_kernel(HkDev d, HkErodeDev E, int erode, long long* send, const int* e_l2g, int cap) {
    const int n = *d.del_fresh, first = *d.del_count;
    if (erode)
        for (int i = 0; i < n && first + i < d.del_cap; ++i) erode_element(d, E, (int)(d.del_list[first + i] & 0xffffffffll));
    if (send) {                                               // this step's deletions as global ids, for the all-gather
        send[0] = n;
        for (int i = 0; i < n && i < cap && first + i < d.del_cap; ++i) send[1 + i] = e_l2g[d.del_list[first + i] & 0xffffffffll];
    }
    *d.del_count = first + n;
}
__global__ void hk_erode_replay_kernel(HkDev d, HkErodeDev E, const long long* gathered, int world, int cap) {
    for (int r = 0; r < world; ++r) {
        const long long* g = gathered + (long long)r * (cap + 1);
        const long long n = g[0];
        if (n > cap) { *E.overflow = 1; }
        for (long long i = 0; i < n && i < cap; ++i) erode_element(d, E, (int)g[1 + i]);
    }
}
#endif

#ifndef HK_EMU
__global__ void hk_step_set_kernel(long long* t_dev, long long t) { *t_dev = t; }
__global__ void hk_step_advance_kernel(long long* t_dev) { *t_dev = *t_dev + 1; }
#endif
void hk_launch_step_set(long long* t_dev, long long t, cudaStream_t s) {
#ifndef HK_EMU
    hk_step_set_kernel<<<1, 1, 0, s>>>(t_dev, t);
#else
    (void)s; *t_dev = t;
#endif
}
void hk_launch_step_advance(long long* t_dev, cudaStream_t s) {
#ifndef HK_EMU
    hk_step_advance_kernel<<<1, 1, 0, s>>>(t_dev);
#else
    (void)s; *t_dev = *t_dev + 1;
#endif
}

void hk_launch_erode_replay(const HkDev& dd, const HkErodeDev& E, const long long* gathered, int world, int cap, cudaStream_t s) {
#ifndef HK_EMU
    hk_erode_replay_kernel<<<1, 1, 0, s>>>(dd, E, gathered, world, cap);
#else
    (void)s;
    for (int r = 0; r < world; ++r) {
        const long long* g = gathered + (long long)r * (cap + 1);
        const long long n = g[0];
        if (n > cap) *E.overflow = 1;
        for (long long i = 0; i < n && i < cap; ++i) erode_element(dd, E, (int)g[1 + i]);
    }
#endif
}

void hk_launch_deletion_pass(const HkDev& dd, const HkErodeDev* er, long long step, cudaStream_t s, long long* n_launch,
                             long long* send, const int* e_l2g, int cap) {
    const HkDev d = dd;
    HkErodeDev E;
    memset(&E, 0, sizeof(E));
    if (er) E = *er;
#ifndef HK_EMU
    const unsigned nb = (unsigned)((d.nElement + HK_DEL_BLOCK - 1) / HK_DEL_BLOCK);
    hk_delete_count_kernel<<<nb, 256, 0, s>>>(d);
    hk_delete_scan_kernel<<<1, 1024, 0, s>>>(d, (int)nb);
    hk_delete_emit_kernel<<<nb, 256, 0, s>>>(d, step, (int)nb);
    hk_delete_finish_kernel<<<1, 1, 0, s>>>(d, E, er ? 1 : 0, send, e_l2g, cap);
    if (n_launch) *n_launch += 4;
#else
    (void)s; (void)n_launch;
    const int first = *d.del_count;
    int n = 0;
    for (long long e = 0; e < d.nElement; ++e)
        if (d.flag[e] == 3) {
            if (first + n < d.del_cap) d.del_list[first + n] = (step << 32) | e;
            ++n;
            flush_element(d, e);
        }
    if (er)
        for (int i = 0; i < n && first + i < d.del_cap; ++i) erode_element(d, E, (int)(d.del_list[first + i] & 0xffffffffll));
    if (send) {
        send[0] = n;
        for (int i = 0; i < n && i < cap && first + i < d.del_cap; ++i) send[1 + i] = e_l2g[d.del_list[first + i] & 0xffffffffll];
    }
    *d.del_count = first + n;
#endif
}

// ------------------------------------------------------------------ small taps and transposes
void hk_launch_velo_from_rec(const HkDev& d, double d_time, cudaStream_t s) {
    double* velo = d.velo;
    const double* rec = d.rec;
    hk_parallel_for(d.nNode * 3, s, HK_LAMBDA(long long i) {
        long long n = i / 3;
        int c = (int)(i - 3 * n);
        velo[i] = rec[6 * n + 3 + c] / d_time;
    });
}

void hk_launch_gather_Q(const HkDev& dd, double* Q_out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nNode, s, HK_LAMBDA(long long n) {
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        for (int w = 0; w < d.ell_width; ++w) {
            int ent = d.ell[(long long)w * d.nNode + n];
            if (ent < 0) break;
            long long e = ent >> 3;
            int a = ent & 7;
            q0 += d.Qe[(long long)(a * 3 + 0) * d.nEp + e];
            q1 += d.Qe[(long long)(a * 3 + 1) * d.nEp + e];
            q2 += d.Qe[(long long)(a * 3 + 2) * d.nEp + e];
        }
        Q_out[3 * n] = q0; Q_out[3 * n + 1] = q1; Q_out[3 * n + 2] = q2;
    });
}

// partial internal force of the listed (interface) nodes, same gather order as the nodal kernel; Q0 != NULL: the
// partial force was supplied through hk_upload_state
void hk_launch_halo_pack(const HkDev& dd, const int* nodes, long long n, double* out, const double* Q0, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n, s, HK_LAMBDA(long long i) {
        const long long nd = nodes[i];
        double q0 = 0.0, q1 = 0.0, q2 = 0.0;
        if (Q0) {
            q0 = Q0[3 * nd]; q1 = Q0[3 * nd + 1]; q2 = Q0[3 * nd + 2];
        } else {
            for (int w = 0; w < d.ell_width; ++w) {
                const int ent = d.ell[(long long)w * d.nNode + nd];
                if (ent < 0) break;
                const long long e = ent >> 3;
                const int a = ent & 7;
                q0 += d.Qe[(long long)(a * 3 + 0) * d.nEp + e];
                q1 += d.Qe[(long long)(a * 3 + 1) * d.nEp + e];
                q2 += d.Qe[(long long)(a * 3 + 2) * d.nEp + e];
            }
        }
        out[3 * i] = q0; out[3 * i + 1] = q1; out[3 * i + 2] = q2;
    });
}

// send[i] = own[slots[i]]: the part of this rank's interface partials that neighbour shares
void hk_launch_halo_gather(const double* own, const int* slots, long long n, double* send, cudaStream_t s) {
    hk_parallel_for(n * 3, s, HK_LAMBDA(long long j) {
        const long long i = j / 3;
        send[j] = own[3ll * slots[i] + (j - 3 * i)];
    });
}

// halo_recv[slot] += recv[i] (slots == NULL: dense, slot = i): contributions are added one holder at a time in
// ascending global-rank order, so every holder of a node forms the same sum bit for bit
void hk_launch_halo_accumulate(const HkDev& dd, const int* slots, long long n, const double* recv, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 3, s, HK_LAMBDA(long long j) {
        const long long i = j / 3;
        const int c = (int)(j - 3 * i);
        const long long k = 3ll * (slots ? slots[i] : i) + c;
        d.halo_recv[k] = d.halo_recv[k] + recv[j];
    });
}

// {position, velocity} of listed nodes -> out (6 doubles per node); and the inverse for ghost copies
void hk_launch_nodes_export(const HkDev& dd, const int* nodes, long long n, double* out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 6, s, HK_LAMBDA(long long j) {
        const long long i = j / 6;
        const int c = (int)(j - 6 * i);
        const long long nd = nodes[i];
        out[j] = c < 3 ? d.rec[6 * nd + c] : d.velo[3 * nd + c - 3];
    });
}
void hk_launch_nodes_import(const HkDev& dd, const int* nodes, const long long* src, long long n, const double* in,
                            cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 6, s, HK_LAMBDA(long long j) {
        const long long i = j / 6;
        const int c = (int)(j - 6 * i);
        const long long nd = nodes[i];
        const double v = in[6 * src[i] + c];
        if (c < 3) d.rec[6 * nd + c] = v; else d.velo[3 * nd + c - 3] = v;
    });
}
// ghost-element partitions: {disp, disp_pre} of the listed nodes -> out (6 doubles per node); the import overwrites a
// ghost node's whole kinematic state with the owner's values, rebuilding position / d_disp / velo by the same
// expressions as the nodal kernel (J2:625-652), so the copy is bit-identical to the owner's node
void hk_launch_state_export(const HkDev& dd, const int* nodes, long long n, double* out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 6, s, HK_LAMBDA(long long j) {
        const long long i = j / 6;
        const int c = (int)(j - 6 * i);
        const long long nd = nodes[i];
        out[j] = c < 3 ? d.u[3 * nd + c] : d.u_pre[3 * nd + c - 3];
    });
}
void hk_launch_state_import(const HkDev& dd, const int* nodes, long long n, const double* in, double d_time, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 3, s, HK_LAMBDA(long long j) {
        const long long i = j / 3;
        const int c = (int)(j - 3 * i);
        const long long nd = nodes[i];
        const double un = in[6 * i + c], up = in[6 * i + 3 + c];
        const double dd_ = un - up;
        d.u[3 * nd + c] = un;
        d.u_pre[3 * nd + c] = up;
        d.rec[6 * nd + c] = d.X[3 * nd + c] + un;
        d.rec[6 * nd + 3 + c] = dd_;
        d.velo[3 * nd + c] = dd_ / d_time;
    });
}

// contact accumulators of the listed nodes -> out (6 x u64 per node); import = exact 128-bit sum over n_ranks records
void hk_launch_cacc_export(const HkDev& dd, const int* nodes, long long n, unsigned long long* out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 6, s, HK_LAMBDA(long long j) {
        const long long i = j / 6;
        const int w = (int)(j - 6 * i);
        const int si = d.spec_idx[nodes[i]];
        const int slot = si < 0 ? -1 : d.spec[si].contact_slot;
        out[j] = slot < 0 ? 0ull : d.cacc[6ll * slot + w];
    });
}
void hk_launch_cacc_import(const HkDev& dd, const int* nodes, long long n, const unsigned long long* in, long long n_ranks,
                           cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 3, s, HK_LAMBDA(long long j) {
        const long long i = j / 3;
        const int c = (int)(j - 3 * i);
        unsigned long long lo = 0ull, hi = 0ull;
        for (long long r = 0; r < n_ranks; ++r) {
            const unsigned long long* p = in + (r * n + i) * 6 + 2 * c;
            const unsigned long long nlo = lo + p[0];
            hi += p[1] + (nlo < lo ? 1ull : 0ull);
            lo = nlo;
        }
        const int si = d.spec_idx[nodes[i]];
        const int slot = si < 0 ? -1 : d.spec[si].contact_slot;
        if (slot < 0) return;
        d.cacc[6ll * slot + 2 * c] = lo;
        d.cacc[6ll * slot + 2 * c + 1] = hi;
    });
}

// The same exchange as a plain integer all-reduce: each 128-bit accumulator X (two's complement, mod 2^128) travels as
// three limbs X[0:43], X[43:86], X[86:128] in int64 lanes.  The lane-wise sums over R ranks stay below R * 2^43 (no
// overflow for R < 2^20), and S0 + S1*2^43 + S2*2^86 mod 2^128 is the exact sum of the X — no carry is ever lost.
void hk_launch_cacc_export_limbs(const HkDev& dd, const int* nodes, long long n, long long* out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 3, s, HK_LAMBDA(long long j) {
        const long long i = j / 3;
        const int c = (int)(j - 3 * i);
        const int si = d.spec_idx[nodes[i]];
        const int slot = si < 0 ? -1 : d.spec[si].contact_slot;      // candidate node not (yet) on a contact surface: zero
        const unsigned long long lo = slot < 0 ? 0ull : d.cacc[6ll * slot + 2 * c], hi = slot < 0 ? 0ull : d.cacc[6ll * slot + 2 * c + 1];
        const unsigned long long m43 = (1ull << 43) - 1;
        out[9 * i + 3 * c + 0] = (long long)(lo & m43);
        out[9 * i + 3 * c + 1] = (long long)(((lo >> 43) | (hi << 21)) & m43);      // bits 43..85
        out[9 * i + 3 * c + 2] = (long long)(hi >> 22);                             // bits 86..127
    });
}
void hk_launch_cacc_import_limbs(const HkDev& dd, const int* nodes, long long n, const long long* in, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(n * 3, s, HK_LAMBDA(long long j) {
        const long long i = j / 3;
        const int c = (int)(j - 3 * i);
        const unsigned long long s0 = (unsigned long long)in[9 * i + 3 * c + 0];
        const unsigned long long s1 = (unsigned long long)in[9 * i + 3 * c + 1];
        const unsigned long long s2 = (unsigned long long)in[9 * i + 3 * c + 2];
        // X = s0 + s1 * 2^43 + s2 * 2^86  (mod 2^128), 128-bit adds with carry
        unsigned long long lo = s0, hi = 0ull;
        const unsigned long long a_lo = s1 << 43, a_hi = s1 >> 21;
        unsigned long long nlo = lo + a_lo;
        hi += a_hi + (nlo < lo ? 1ull : 0ull);
        lo = nlo;
        hi += s2 << 22;
        const int si = d.spec_idx[nodes[i]];
        const int slot = si < 0 ? -1 : d.spec[si].contact_slot;
        if (slot < 0) return;
        d.cacc[6ll * slot + 2 * c] = lo;
        d.cacc[6ll * slot + 2 * c + 1] = hi;
    });
}

// external_force of the last step (J2:497-538): zero plus the rounded contact sums
void hk_launch_external_force(const HkDev& dd, double* F_out, int lsb_exp, int contact_on, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nNode, s, HK_LAMBDA(long long n) {
        double F[3] = {0.0, 0.0, 0.0};
        const int si = d.spec_idx[n];
        if (contact_on && si >= 0 && d.spec[si].contact_slot >= 0) {
            const unsigned long long* acc = d.cacc + (long long)d.spec[si].contact_slot * 6;
            for (int c = 0; c < 3; ++c) F[c] = fx_to_double(acc[2 * c], acc[2 * c + 1], lsb_exp);
        }
        for (int c = 0; c < 3; ++c) F_out[3 * n + c] = F[c];
    });
}

// reference layout (ncomp, nip) chunk for elements [e0,e0+ne)  <->  rows [row0,row0+ncomp) of the blocked ip state
void hk_launch_ip_to_dev(const double* aos, const HkDev& dd, int row0, int ncomp, long long e0, long long ne, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(ne * 8 * ncomp, s, HK_LAMBDA(long long i) {
        long long el = i % ne;           // thread index runs along elements so the device-side accesses coalesce
        long long r = i / ne;            // r = c*8 + k
        int k = (int)(r % 8);
        int c = (int)(r / 8);
        d.ips[hk_ip(d, row0 + c, k, e0 + el)] = aos[(el * 8 + k) * ncomp + c];
    });
}
void hk_launch_ip_to_aos(const HkDev& dd, double* aos, int row0, int ncomp, long long e0, long long ne, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(ne * 8 * ncomp, s, HK_LAMBDA(long long i) {
        long long el = i % ne;
        long long r = i / ne;
        int k = (int)(r % 8);
        int c = (int)(r / 8);
        aos[(el * 8 + k) * ncomp + c] = d.ips[hk_ip(d, row0 + c, k, e0 + el)];
    });
}
void hk_launch_triax_to_aos(const HkDev& dd, double* aos, long long e0, long long ne, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(ne * 8, s, HK_LAMBDA(long long i) {
        long long el = i % ne;
        int k = (int)(i / ne);
        aos[el * 8 + k] = d.triax[(long long)k * d.nEp + e0 + el];
    });
}

// ------------------------------------------------------------------ state summary (hk_state_summary)
// out[0] live elements, out[1] / out[2] order-encoded min / max of integ_eq_plastic_strain over the Gauss points of
// live elements, out[3] Gauss points of live elements with eps > 0.  `out` must be initialised {0, ~0, 0, 0}.
HK_HD void summary_element(const HkDev& d, long long e, unsigned long long& live, unsigned long long& mn,
                           unsigned long long& mx, unsigned long long& npl) {
    live = 0ull; mn = ~0ull; mx = 0ull; npl = 0ull;
    if (e >= d.nElement || d.flag[e] != 1) return;
    live = 1ull;
    for (int k = 0; k < 8; ++k) {
        const double ep = d.ips[hk_ip(d, 12, k, e)];
        const unsigned long long en = enc_double(ep);
        mn = en < mn ? en : mn;
        mx = en > mx ? en : mx;
        npl += ep > 0.0 ? 1ull : 0ull;
    }
}
#ifndef HK_EMU
__global__ void __launch_bounds__(256) hk_state_summary_kernel(HkDev d, unsigned long long* out) {
    const unsigned FULL = 0xffffffffu;
    unsigned long long live = 0ull, mn = ~0ull, mx = 0ull, npl = 0ull;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < d.nElement; e += (long long)gridDim.x * blockDim.x) {
        unsigned long long l, a, b, c;
        summary_element(d, e, l, a, b, c);
        live += l; npl += c;
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    for (int off = 16; off; off >>= 1) {
        live += __shfl_xor_sync(FULL, live, off);
        npl += __shfl_xor_sync(FULL, npl, off);
        const unsigned long long a = __shfl_xor_sync(FULL, mn, off), b = __shfl_xor_sync(FULL, mx, off);
        mn = a < mn ? a : mn;
        mx = b > mx ? b : mx;
    }
    if ((threadIdx.x & 31) == 0) {
        if (live) atomicAdd(&out[0], live);
        if (npl) atomicAdd(&out[3], npl);
        atomicMin(&out[1], mn);
        atomicMax(&out[2], mx);
    }
}
#endif
void hk_launch_state_summary(const HkDev& d, unsigned long long* out, cudaStream_t s) {
#ifndef HK_EMU
    const int grid = d.n_sm > 0 ? d.n_sm * 8 : 256;
    hk_state_summary_kernel<<<grid, 256, 0, s>>>(d, out);
#else
    for (long long e = 0; e < d.nElement; ++e) {
        unsigned long long l, a, b, c;
        summary_element(d, e, l, a, b, c);
        out[0] += l; out[3] += c;
        if (a < out[1]) out[1] = a;
        if (b > out[2]) out[2] = b;
    }
#endif
}
double hk_decode_double(unsigned long long v) { return dec_double(v); }

// ------------------------------------------------------------------ cal_node_stress_strain on the device (J2:3408-3486)
// element means (J2:3428-3440): rows 0-5 stress, 6-11 strain, 12 eps, 13 triax; the 8 Gauss points are summed in order
void hk_launch_element_means(const HkDev& dd, double* emean, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nElement * 14, s, HK_LAMBDA(long long i) {
        const long long e = i % d.nElement;
        const int row = (int)(i / d.nElement);
        double a = 0.0;
        if (row < 13) for (int k = 0; k < 8; ++k) a += d.ips[hk_ip(d, row, k, e)];
        else for (int k = 0; k < 8; ++k) a += d.triax[(long long)k * d.nEp + e];
        emean[(long long)row * d.nEp + e] = a / 8.0;
    });
}
// nodal means (J2:3442-3482): the node-centric table lists a node's elements in ascending order, which is the order
// the reference's element loop adds them in.  out: [16][nNode] = stress 6, strain 6, eps, mises, triax, inc_num
void hk_launch_node_means(const HkDev& dd, const double* emean, double* out, int raw, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nNode, s, HK_LAMBDA(long long n) {
        double acc[14];
        for (int r = 0; r < 14; ++r) acc[r] = 0.0;
        double inc = 0.0;
        for (int w = 0; w < d.ell_width; ++w) {
            const int ent = d.ell[(long long)w * d.nNode + n];
            if (ent < 0) break;
            const long long e = ent >> 3;
            for (int r = 0; r < 14; ++r) acc[r] += emean[(long long)r * d.nEp + e];
            inc += 1.0;
        }
        if (!raw) for (int r = 0; r < 14; ++r) acc[r] /= inc;          // 0/0 = NaN for unreferenced nodes, as J2:3464-3469
        for (int r = 0; r < 13; ++r) out[(long long)r * d.nNode + n] = acc[r];
        out[14ll * d.nNode + n] = acc[13];
        out[15ll * d.nNode + n] = inc;
        if (!raw) {
            const double ox = acc[0], oy = acc[1], oz = acc[2], txy = acc[3], tyz = acc[4], txz = acc[5];
            out[13ll * d.nNode + n] = sqrt(0.5 * ((ox - oy) * (ox - oy) + (oy - oz) * (oy - oz) + (ox - oz) * (ox - oz) +
                                                  6 * (txy * txy + tyz * tyz + txz * txz)));      // J2:3471-3480
        }
    });
}


// ------------------------------------------------------------------ reference-order element kernel (element_mode 1)
// cal_stress_hexa + cal_BVbar_hexa + cal_Bfinal + cal_triax_stress + the fracture loop exactly as the reference
// evaluates them (J2:1033-1371, 1705-1784, 1415-1519, 982-1022, 701-762): dense 6x24 Bfinal, left-to-right sums,
// no FMA (this translation unit is built with -fmad=false).  One thread per element; ~10x slower than the fast
// kernel.  Its purpose is parity: with it the whole engine is bit-identical to the CPU oracle, including the
// contact ties that depend on the last bit of a nodal position.
#ifndef HK_EMU
__device__ double x_P[8][3][8];
#else
static double x_P[8][3][8];
#endif

int hk_upload_pusai(const double* P) {          // per device: called by every engine at hk_finalize
#ifndef HK_EMU
    return (int)cudaMemcpyToSymbol(x_P, P, sizeof(double) * 192);
#else
    memcpy(x_P, P, sizeof(double) * 192);
    return 0;
#endif
}

struct ExactArgs {
    HkDev d;
    long long step;
    int write_triax;
};

HK_D void exact_jac(const double P1[3][8], const double ep[3][8], double J[3][3]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[r][c] = 0.0;
    for (int i = 0; i < 8; ++i) {
        J[0][0] += P1[0][i] * ep[0][i]; J[0][1] += P1[0][i] * ep[1][i]; J[0][2] += P1[0][i] * ep[2][i];
        J[1][0] += P1[1][i] * ep[0][i]; J[1][1] += P1[1][i] * ep[1][i]; J[1][2] += P1[1][i] * ep[2][i];
        J[2][0] += P1[2][i] * ep[0][i]; J[2][1] += P1[2][i] * ep[1][i]; J[2][2] += P1[2][i] * ep[2][i];
    }
}
HK_D double exact_det3(const double J[3][3]) {
    return (J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
            J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] - J[0][2] * J[1][1] * J[2][0]);
}
HK_D void exact_inv3(const double J[3][3], double div_v, double iJ[3][3]) {
    iJ[0][0] = (J[1][1] * J[2][2] - J[1][2] * J[2][1]) * div_v;
    iJ[1][0] = (J[1][2] * J[2][0] - J[1][0] * J[2][2]) * div_v;
    iJ[2][0] = (J[1][0] * J[2][1] - J[1][1] * J[2][0]) * div_v;
    iJ[0][1] = (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * div_v;
    iJ[1][1] = (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * div_v;
    iJ[2][1] = (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * div_v;
    iJ[0][2] = (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * div_v;
    iJ[1][2] = (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * div_v;
    iJ[2][2] = (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * div_v;
}

HK_D void element_body_exact(const ExactArgs& A, long long e) {
    const HkDev& d = A.d;
    const unsigned char fl = d.flag[e];
    if (fl != 1) {
        if (fl == 0) {
            for (int r = 0; r < 24; ++r) d.Qe[(long long)r * d.nEp + e] = 0.0;
            for (int k = 0; k < 8; ++k) d.triax[(long long)k * d.nEp + e] = 0.0;
            d.flag[e] = 2;
        }
        return;
    }
    const HkMaterialDev& M = d.mats[d.mat[e]];
    double Dm[6][6];
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) Dm[r][c] = 0.0;
    Dm[0][0] = Dm[1][1] = Dm[2][2] = M.D11;
    Dm[0][1] = Dm[0][2] = Dm[1][0] = Dm[1][2] = Dm[2][0] = Dm[2][1] = M.D12;
    Dm[3][3] = Dm[4][4] = Dm[5][5] = M.D44;
    double d_u[24], ep[3][8];
    for (int i = 0; i < 8; ++i) {
        const long long n = d.conn[(long long)i * d.nEp + e];
        for (int c = 0; c < 3; ++c) { ep[c][i] = d.rec[6 * n + c]; d_u[i * 3 + c] = d.rec[6 * n + 3 + c]; }
    }
    // cal_BVbar_hexa
    double BVbar[144];
    for (int i = 0; i < 144; ++i) BVbar[i] = 0.0;
    double V = 0.0;
    int negj = 0;
    for (int k = 0; k < 8; ++k) {
        double J[3][3], iJ[3][3];
        exact_jac(x_P[k], ep, J);
        double detJi = exact_det3(J);
        if (detJi < 0) { detJi = fabs(detJi); negj++; }
        V += detJi;
        const double div_v = 1.0 / detJi;
        exact_inv3(J, div_v, iJ);
        for (int i = 0; i < 8; ++i) {
            const double Pix = iJ[0][0] * x_P[k][0][i] + iJ[0][1] * x_P[k][1][i] + iJ[0][2] * x_P[k][2][i];
            const double Piy = iJ[1][0] * x_P[k][0][i] + iJ[1][1] * x_P[k][1][i] + iJ[1][2] * x_P[k][2][i];
            const double Piz = iJ[2][0] * x_P[k][0][i] + iJ[2][1] * x_P[k][1][i] + iJ[2][2] * x_P[k][2][i];
            for (int r = 0; r < 3; ++r) {
                BVbar[r + 6 * (i * 3 + 0)] += Pix / 3.0 * detJi;
                BVbar[r + 6 * (i * 3 + 1)] += Piy / 3.0 * detJi;
                BVbar[r + 6 * (i * 3 + 2)] += Piz / 3.0 * detJi;
            }
        }
    }
    for (int i = 0; i < 144; ++i) BVbar[i] = BVbar[i] / V;
    if (negj) hk_atomic_add_u64(&d.counters[0], (unsigned long long)negj);

    double Qe[24];
    for (int j = 0; j < 24; ++j) Qe[j] = 0.0;
    double v_e = 0.0, t_e = 0.0;
    for (int k = 0; k < 8; ++k) {
        double Bf[144];
        for (int q = 0; q < 144; ++q) Bf[q] = 0.0;
        double J[3][3], iJ[3][3];
        exact_jac(x_P[k], ep, J);
        const double detJ = exact_det3(J);
        const double div_v = 1.0 / detJ;
        exact_inv3(J, div_v, iJ);
        for (int i = 0; i < 8; ++i) {
            const double Pix = iJ[0][0] * x_P[k][0][i] + iJ[0][1] * x_P[k][1][i] + iJ[0][2] * x_P[k][2][i];
            const double Piy = iJ[1][0] * x_P[k][0][i] + iJ[1][1] * x_P[k][1][i] + iJ[1][2] * x_P[k][2][i];
            const double Piz = iJ[2][0] * x_P[k][0][i] + iJ[2][1] * x_P[k][1][i] + iJ[2][2] * x_P[k][2][i];
            const int c0 = i * 3, c1 = i * 3 + 1, c2 = i * 3 + 2;
            Bf[0 + 6 * c0] += Pix; Bf[1 + 6 * c1] += Piy; Bf[2 + 6 * c2] += Piz;
            Bf[3 + 6 * c0] += Piy; Bf[3 + 6 * c1] += Pix;
            Bf[4 + 6 * c1] += Piz; Bf[4 + 6 * c2] += Piy;
            Bf[5 + 6 * c0] += Piz; Bf[5 + 6 * c2] += Pix;
            for (int r = 0; r < 3; ++r) {
                Bf[r + 6 * c0] += -Pix / 3.0 + BVbar[r + 6 * c0];
                Bf[r + 6 * c1] += -Piy / 3.0 + BVbar[r + 6 * c1];
                Bf[r + 6 * c2] += -Piz / 3.0 + BVbar[r + 6 * c2];
            }
        }
        double de[6], dov[6];
        for (int r = 0; r < 6; ++r) {
            double s = Bf[r] * d_u[0];
            for (int c = 1; c < 24; ++c) s += Bf[r + 6 * c] * d_u[c];
            de[r] = s;
        }
        for (int r = 0; r < 6; ++r) {
            double s = Dm[r][0] * de[0];
            for (int c = 1; c < 6; ++c) s += Dm[r][c] * de[c];
            dov[r] = s;
        }
        double pre[6], fin[6];
        for (int r = 0; r < 6; ++r) { pre[r] = d.ips[hk_ip(d, r, k, e)]; fin[r] = pre[r] + dov[r]; }
        double ep_ = d.ips[hk_ip(d, 12, k, e)];
        if (M.npp > 0) {
            double tri[6];
            for (int r = 0; r < 6; ++r) tri[r] = pre[r] + dov[r];
            const double mean_stress = (tri[0] + tri[1] + tri[2]) / 3.0;
            const double tds[6] = {tri[0] - mean_stress, tri[1] - mean_stress, tri[2] - mean_stress, tri[3], tri[4], tri[5]};
            const double mises = sqrt(1.5 * (tds[0] * tds[0] + tds[1] * tds[1] + tds[2] * tds[2] + 2 * (tds[3] * tds[3]) +
                                             2 * (tds[4] * tds[4]) + 2 * (tds[5] * tds[5])));
            const double y = d.ips[hk_ip(d, 13, k, e)];
            if (mises > y) {
                int p_index = 1;
                for (int j = 2; j <= M.npp; ++j) {
                    if (ep_ <= M.plastic_e[j - 1]) { p_index = j - 1; break; }
                    if (j == M.npp) p_index = j - 1;
                }
                const double H = M.Hd[p_index - 1];
                const double d_ep = (mises - y) / (3 * M.G + H);
                const double fac_num = (y + H * d_ep);
                for (int r = 0; r < 6; ++r) {
                    const double fds = tds[r] * fac_num / mises;
                    fin[r] = fds + (r < 3 ? mean_stress : 0.0);
                }
                ep_ = ep_ + d_ep;
                d.ips[hk_ip(d, 12, k, e)] = ep_;
                d.ips[hk_ip(d, 13, k, e)] = y + H * d_ep;
            }
        }
        for (int r = 0; r < 6; ++r) {
            d.ips[hk_ip(d, 6 + r, k, e)] += de[r];
            d.ips[hk_ip(d, r, k, e)] = fin[r];
        }
        for (int j = 0; j < 24; ++j) {
            double s = Bf[6 * j] * fin[0];
            for (int r = 1; r < 6; ++r) s += Bf[r + 6 * j] * fin[r];
            Qe[j] += 1.0 * 1.0 * 1.0 * detJ * s;
        }
        // cal_triax_stress, invariant form (the oracle's triax_route 0)
        const double ox = fin[0], oy = fin[1], oz = fin[2], txy = fin[3], tyz = fin[4], txz = fin[5];
        const double oeq = sqrt(0.5 * ((ox - oy) * (ox - oy) + (oy - oz) * (oy - oz) + (ox - oz) * (ox - oz) +
                                       6 * (txy * txy + tyz * tyz + txz * txz)));
        double tx = 0.0;
        if (!(oeq < 1E-10)) tx = (ox + oy + oz) / 3.0 / oeq;
        if (A.write_triax) d.triax[(long long)k * d.nEp + e] = tx;
        v_e += ep_;
        t_e += tx;
    }
    for (int j = 0; j < 24; ++j) d.Qe[(long long)j * d.nEp + e] = Qe[j];
    // fracture, J2:701-762
    if (M.nd > 0) {
        v_e /= 8;
        t_e /= 8;
        if (!(t_e < 0)) {
            const int nd = M.nd;
            double fr_e = M.duct_e[nd - 1];
            for (int j = 0; j + 1 < nd; ++j)
                if (t_e >= M.duct_t[j] && t_e < M.duct_t[j + 1]) {
                    fr_e = M.duct_e[j] + (M.duct_e[j + 1] - M.duct_e[j]) / (M.duct_t[j + 1] - M.duct_t[j]) * (t_e - M.duct_t[j]);
                    break;
                }
            if (v_e >= fr_e) {
                d.flag[e] = 3;                     // logged and zeroed by hk_launch_deletion_pass (stream order)
            }
        }
    }
}

#ifndef HK_EMU
__global__ void __launch_bounds__(64) hk_element_exact_kernel(ExactArgs A) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e < A.d.nElement) element_body_exact(A, e);
}
#endif

void hk_launch_element_exact(const HkDev& d, long long step, int write_triax, cudaStream_t s) {
    ExactArgs A{d, step, write_triax};
#ifndef HK_EMU
    hk_element_exact_kernel<<<(unsigned)((d.nElement + 63) / 64), 64, 0, s>>>(A);
#else
    for (long long e = 0; e < d.nElement; ++e) element_body_exact(A, e);
#endif
}
