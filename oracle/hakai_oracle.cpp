// hakai_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A line-by-line C++ restatement of the reference's per-step hot path,
// HAKAI-v0.0.2/Julia/HAKAI_j.jl (cited as J2:<line>), FP64, built with -ffp-contract=off so no
// FMA contraction happens (Julia does not contract either).  It is NOT part of the product:
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
// load it.  It exports the ABI of include/hakai_b200.h with the prefix hko_ instead of hk_.
//
// PARITY UNPINNED: the reference ships no golden vectors or tests and neither Julia nor MATLAB
// exists in the build image, so this oracle could not be checked against outputs of the
// reference itself.  It is anchored instead on (1) closed-form facts of Tensile5e.inp,
// (2) element / contact identities, (3) agreement with an independent NumPy restatement of the
// v0.0.0 matrix form (oracle/hakai_np.py) — see tests/test_oracle_*.py and DESIGN.md.
//
// Third-party arithmetic restated here (versions unpinned, no Manifest.toml in the reference):
//   * StaticArrays.eigvals on a symmetric 3x3 SMatrix (J2:1004-1007): closed-form trigonometric
//     solution (StaticArrays src/eigen.jl, `_eigvals(::Size{(3,3)}, ::Hermitian)`); triax_route=1.
//     triax_route=0 uses the invariants (I1/3)/sqrt(3 J2), mathematically identical.
//   * Quadmath.Float128 contact accumulators (J2:435): __float128 (libquadmath semantics).
//   * FLoops.@floop: OpenMP parallel for over the same loops (J2:562,624,644,995,1114,2370).

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/hakai_b200.h"

typedef __float128 f128;

namespace {

struct Material {
    double young = 0, poisson = 0, density = 0, G = 0;
    double Dmat[36];                 // column-major 6x6 (symmetric)             J2:143-159
    int64_t npp = 0;
    std::vector<double> plastic;     // (npp,2) column-major
    std::vector<double> Hd;          // (npp-1)
    int64_t nd = 0;
    std::vector<double> ductile;     // (nd,3) column-major
};

struct BCEntry {
    std::vector<std::vector<int64_t>> dof;   // 1-based
    std::vector<double> value;
    std::vector<double> a_t, a_v;
    bool has_amp = false;
};

struct ICEntry {
    std::vector<std::vector<int64_t>> dof;
    std::vector<double> value;
};

struct InstanceO {
    int64_t node_offset = 0, nNode = 0, element_offset = 0, nElement = 0;
    std::vector<int64_t> surfaces;          // (6nE,4) column-major, part-local ids
    std::vector<int64_t> sorted_surfaces;   // (6nE,4) column-major
    std::vector<int64_t> surfaces_eleid;    // (6nE)
};

struct ContactTriangleO {                   // ContactTriangle, J2:72-78
    int64_t i_instance = 0, j_instance = 0;
    std::vector<int64_t> c_nodes_i, c_nodes_j, c_triangles_eleid;
    std::vector<int64_t> t0, t1, t2;        // c_triangles[:,1..3]
    double young = 0;
};

}  // namespace

struct hk_engine {
    hk_params prm;
    std::string err;
    bool finalized = false;
    int64_t nNode = 0, nElement = 0, fn = 0, nip = 0;
    std::vector<double> coordmat, diag_M, diag_C;
    std::vector<int64_t> elementmat, element_material, element_instance;
    std::vector<Material> MATERIAL;
    std::vector<BCEntry> BC;
    std::vector<ICEntry> IC;
    std::vector<InstanceO> INSTANCE;
    std::vector<ContactTriangleO> CT;
    int flag_fracture = 1;                  // always 1: length(failure_stress::Float64)==1, J2:162-165
    double Pusai[8][3][8];                  // Pusai_mat[k][dir][node], J2:1895-1943
    // loop-carried state (Julia layouts)
    std::vector<double> position, disp, disp_new, disp_pre, d_disp, velo, external_force, Q, Qe;
    std::vector<double> integ_stress, integ_strain, integ_eq_plastic_strain, integ_triax_stress,
        integ_yield_stress, elementVolume, d_disp_norm;
    // evaluation order of the three StaticArrays products of cal_stress_hexa (J2:1204-1205, 1330), a THIRD-PARTY choice the
    // reference tree does not pin (no Manifest): 0 = left-to-right sums of products (StaticArrays' unrolled `+` of `*`,
    // no FMA: Julia never contracts by itself), 1 = a muladd chain (what newer StaticArrays kernels emit; fused on any
    // CPU with FMA).  Set by the environment variable HKO_MATVEC=muladd at hko_create; scripts/oracle_matvec_orders.py
    // reports what the choice changes.
    int matvec_mode = 0;
    // v0.0.1 penetration-rate clamp (hk_params.contact_dmax_clamp): J1:413-415, 492, 513, 618
    std::vector<double> d_node, d_node_pre;
    double d_max = 0.0;
    std::vector<int64_t> element_flag;
    std::vector<f128> c_force3;             // (fn, Nth)
    int Nth = 1;
    std::vector<int64_t> deleted_all, deleted_step;
    int64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

static std::string g_create_err;

static int fail(hk_engine* e, int code, const std::string& msg) {
    if (e) e->err = msg; else g_create_err = msg;
    return code;
}

// ---------------------------------------------------------------- small math helpers (J2:3162-3373)
static inline double my3norm(double b1, double b2, double b3) {           // J2:3167
    return std::sqrt(b1 * b1 + b2 * b2 + b3 * b3);
}
static inline void my3crossNNz(double a1, double a2, double a3, double b1, double b2, double b3,
                               double& n1, double& n2, double& n3) {       // J2:3209
    n1 = a2 * b3 - a3 * b2;
    n2 = a3 * b1 - a1 * b3;
    n3 = a1 * b2 - a2 * b1;
    double mag_n = std::sqrt(n1 * n1 + n2 * n2 + n3 * n3);
    n1 = n1 / mag_n;
    n2 = n2 / mag_n;
    n3 = n3 / mag_n;
}
static inline void my3SolveAb(double A11, double A21, double A31, double A12, double A22, double A32,
                              double A13, double A23, double A33, double bx, double by, double bz,
                              double& x1, double& x2, double& x3) {        // J2:3342
    double v = (A11 * A22 * A33 + A12 * A23 * A31 + A13 * A21 * A32 - A11 * A23 * A32 -
                A12 * A21 * A33 - A13 * A22 * A31);
    double im11 = A22 * A33 - A23 * A32;
    double im21 = A23 * A31 - A21 * A33;
    double im31 = A21 * A32 - A22 * A31;
    double im12 = A13 * A32 - A12 * A33;
    double im22 = A11 * A33 - A13 * A31;
    double im32 = A12 * A31 - A11 * A32;
    double im13 = A12 * A23 - A13 * A22;
    double im23 = A13 * A21 - A11 * A23;
    double im33 = A11 * A22 - A12 * A21;
    x1 = (im11 * bx + im12 * by + im13 * bz) / v;
    x2 = (im21 * bx + im22 * by + im23 * bz) / v;
    x3 = (im31 * bx + im32 * by + im33 * bz) / v;
}

// ---------------------------------------------------------------- cal_Pusai_hexa (J2:1895-1943)
static void cal_Pusai_hexa(double P[8][3][8]) {
    static const double delta_mat[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1},
                                           {-1, -1, 1},  {1, -1, 1},  {1, 1, 1},  {-1, 1, 1}};
    const double g = 1.0 / std::sqrt(3.0);
    const double gc[8][3] = {{-g, -g, -g}, {-g, -g, g}, {-g, g, -g}, {-g, g, g},
                             {g, -g, -g},  {g, -g, g},  {g, g, -g},  {g, g, g}};
    for (int k = 0; k < 8; ++k) {
        double gzai = gc[k][0], eta = gc[k][1], tueta = gc[k][2];
        for (int i = 0; i < 8; ++i) {
            P[k][0][i] = 1.0 / 8.0 * delta_mat[i][0] * (1.0 + eta * delta_mat[i][1]) * (1.0 + tueta * delta_mat[i][2]);
            P[k][1][i] = 1.0 / 8.0 * delta_mat[i][1] * (1.0 + gzai * delta_mat[i][0]) * (1.0 + tueta * delta_mat[i][2]);
            P[k][2][i] = 1.0 / 8.0 * delta_mat[i][2] * (1.0 + gzai * delta_mat[i][0]) * (1.0 + eta * delta_mat[i][1]);
        }
    }
}

// Jacobian J = Pusai1 * e_position' accumulated as J2:1424-1434 / 1717-1727
static inline void jac(const double P1[3][8], const double ep[3][8], double J[3][3]) {
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) J[r][c] = 0.0;
    for (int i = 0; i < 8; ++i) {
        J[0][0] += P1[0][i] * ep[0][i];
        J[0][1] += P1[0][i] * ep[1][i];
        J[0][2] += P1[0][i] * ep[2][i];
        J[1][0] += P1[1][i] * ep[0][i];
        J[1][1] += P1[1][i] * ep[1][i];
        J[1][2] += P1[1][i] * ep[2][i];
        J[2][0] += P1[2][i] * ep[0][i];
        J[2][1] += P1[2][i] * ep[1][i];
        J[2][2] += P1[2][i] * ep[2][i];
    }
}
static inline double det3(const double J[3][3]) {                          // J2:1436-1441
    return (J[0][0] * J[1][1] * J[2][2] + J[0][1] * J[1][2] * J[2][0] + J[0][2] * J[1][0] * J[2][1] -
            J[0][0] * J[1][2] * J[2][1] - J[0][1] * J[1][0] * J[2][2] - J[0][2] * J[1][1] * J[2][0]);
}
static inline void inv3(const double J[3][3], double div_v, double iJ[3][3]) {   // J2:1445-1455
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

// cal_BVbar_hexa (J2:1705-1784).  BVbar is 6x24 column-major: BVbar[r + 6*c].
static double cal_BVbar_hexa(const double P[8][3][8], const double ep[3][8], double* BVbar, int64_t* negJ) {
    double V = 0.0;
    for (int k = 0; k < 8; ++k) {
        double J[3][3], iJ[3][3];
        jac(P[k], ep, J);
        double detJi = det3(J);
        if (detJi < 0) {
            detJi = std::fabs(detJi);
            if (negJ) {
#pragma omp atomic
                (*negJ)++;
            }
        }
        V += detJi;
        double div_v = 1.0 / detJi;
        inv3(J, div_v, iJ);
        for (int i = 0; i < 8; ++i) {
            double Pix = iJ[0][0] * P[k][0][i] + iJ[0][1] * P[k][1][i] + iJ[0][2] * P[k][2][i];
            double Piy = iJ[1][0] * P[k][0][i] + iJ[1][1] * P[k][1][i] + iJ[1][2] * P[k][2][i];
            double Piz = iJ[2][0] * P[k][0][i] + iJ[2][1] * P[k][1][i] + iJ[2][2] * P[k][2][i];
            for (int r = 0; r < 3; ++r) {
                BVbar[r + 6 * (i * 3 + 0)] += Pix / 3.0 * detJi;
                BVbar[r + 6 * (i * 3 + 1)] += Piy / 3.0 * detJi;
                BVbar[r + 6 * (i * 3 + 2)] += Piz / 3.0 * detJi;
            }
        }
    }
    for (int i = 0; i < 144; ++i) BVbar[i] = BVbar[i] / V;
    return V;
}

// cal_Bfinal (J2:1415-1519): Bfinal = B - BV + BVbar, returns signed detJ.
static double cal_Bfinal(double* Bfinal, const double* BVbar, const double P1[3][8], const double ep[3][8]) {
    double J[3][3], iJ[3][3];
    jac(P1, ep, J);
    double v = det3(J);
    double div_v = 1.0 / v;
    inv3(J, div_v, iJ);
#define BF(r, c) Bfinal[(r) + 6 * (c)]
#define BB(r, c) BVbar[(r) + 6 * (c)]
    for (int i = 0; i < 8; ++i) {
        double Pix = iJ[0][0] * P1[0][i] + iJ[0][1] * P1[1][i] + iJ[0][2] * P1[2][i];
        double Piy = iJ[1][0] * P1[0][i] + iJ[1][1] * P1[1][i] + iJ[1][2] * P1[2][i];
        double Piz = iJ[2][0] * P1[0][i] + iJ[2][1] * P1[1][i] + iJ[2][2] * P1[2][i];
        int c0 = i * 3, c1 = i * 3 + 1, c2 = i * 3 + 2;
        BF(0, c0) += Pix;
        BF(1, c1) += Piy;
        BF(2, c2) += Piz;
        BF(3, c0) += Piy;
        BF(3, c1) += Pix;
        BF(4, c1) += Piz;
        BF(4, c2) += Piy;
        BF(5, c0) += Piz;
        BF(5, c2) += Pix;
        for (int r = 0; r < 3; ++r) {
            BF(r, c0) += -Pix / 3.0 + BB(r, c0);
            BF(r, c1) += -Piy / 3.0 + BB(r, c1);
            BF(r, c2) += -Piz / 3.0 + BB(r, c2);
        }
    }
#undef BF
#undef BB
    return v;
}

// ---------------------------------------------------------------- cal_stress_hexa (J2:1033-1371)
static void cal_stress_hexa(hk_engine* E) {
    const int64_t nElement = E->nElement;
    const int integ_num = 8;
    const double W = 1.0;
    int64_t negJ = 0;
    const bool fused = E->matvec_mode == 1;
#pragma omp parallel for schedule(static) reduction(+ : negJ)
    for (int64_t e = 0; e < nElement; ++e) {
        if (E->element_flag[e] == 0) continue;
        const Material& M = E->MATERIAL[E->element_material[e] - 1];
        const double G = M.G;
        const int64_t npp = M.npp;
        const double* Dmat = M.Dmat;
        double d_u[24], ep[3][8];
        for (int i = 0; i < 8; ++i) {
            int64_t nd = E->elementmat[i + 8 * e] - 1;
            d_u[i * 3 + 0] = E->d_disp[nd * 3 + 0];
            d_u[i * 3 + 1] = E->d_disp[nd * 3 + 1];
            d_u[i * 3 + 2] = E->d_disp[nd * 3 + 2];
            ep[0][i] = E->position[0 + 3 * nd];
            ep[1][i] = E->position[1 + 3 * nd];
            ep[2][i] = E->position[2 + 3 * nd];
        }
        double BVbar[144];
        for (int i = 0; i < 144; ++i) BVbar[i] = 0.0;
        int64_t nj = 0;
        double V = cal_BVbar_hexa(E->Pusai, ep, BVbar, &nj);
        negJ += nj;
        E->elementVolume[e] = V;
        double Bfinal[144];
        double* Qe = &E->Qe[24 * e];
        for (int i = 0; i < integ_num; ++i) {
            for (int q = 0; q < 144; ++q) Bfinal[q] = 0.0;
            double detJ = cal_Bfinal(Bfinal, BVbar, E->Pusai[i], ep);
            double d_e_vec[6], d_o_vec[6];
            for (int r = 0; r < 6; ++r) {                 // d_e_vec = Bfinal * d_u      J2:1204
                double s = Bfinal[r] * d_u[0];
                if (fused) for (int c = 1; c < 24; ++c) s = std::fma(Bfinal[r + 6 * c], d_u[c], s);
                else for (int c = 1; c < 24; ++c) s += Bfinal[r + 6 * c] * d_u[c];
                d_e_vec[r] = s;
            }
            for (int r = 0; r < 6; ++r) {                 // d_o_vec = Dmat * d_e_vec    J2:1205
                double s = Dmat[r] * d_e_vec[0];
                if (fused) for (int c = 1; c < 6; ++c) s = std::fma(Dmat[r + 6 * c], d_e_vec[c], s);
                else for (int c = 1; c < 6; ++c) s += Dmat[r + 6 * c] * d_e_vec[c];
                d_o_vec[r] = s;
            }
            const int64_t index_i = e * integ_num + i;
            double pre_stress[6], final_stress[6];
            for (int r = 0; r < 6; ++r) pre_stress[r] = E->integ_stress[r + 6 * index_i];
            for (int r = 0; r < 6; ++r) final_stress[r] = pre_stress[r] + d_o_vec[r];
            if (npp > 0) {                                 // length(plastic_property_) > 0  J2:1227
                double tri_stress[6];
                for (int r = 0; r < 6; ++r) tri_stress[r] = pre_stress[r] + d_o_vec[r];
                double mean_stress = (tri_stress[0] + tri_stress[1] + tri_stress[2]) / 3.0;
                double tds[6] = {tri_stress[0] - mean_stress, tri_stress[1] - mean_stress,
                                 tri_stress[2] - mean_stress, tri_stress[3], tri_stress[4], tri_stress[5]};
                double tri_mises_stress = std::sqrt(1.5 * (tds[0] * tds[0] + tds[1] * tds[1] + tds[2] * tds[2] +
                                                           2 * (tds[3] * tds[3]) + 2 * (tds[4] * tds[4]) +
                                                           2 * (tds[5] * tds[5])));
                double y = E->integ_yield_stress[index_i];
                if (tri_mises_stress > y) {
                    int64_t p_index = 1;
                    for (int64_t j = 2; j <= npp; ++j) {   // J2:1256-1264
                        if (E->integ_eq_plastic_strain[index_i] <= M.plastic[(j - 1) + npp * 1]) {
                            p_index = j - 1;
                            break;
                        }
                        if (j == npp) p_index = j - 1;
                    }
                    double H = M.Hd[p_index - 1];
                    double d_ep = (tri_mises_stress - y) / (3 * G + H);
                    double fac_num = (y + H * d_ep);
                    for (int r = 0; r < 6; ++r) {          // tri_dev*(y+H*d_ep)/tri_mises   J2:1274
                        double fds = tds[r] * fac_num / tri_mises_stress;
                        final_stress[r] = fds + (r < 3 ? mean_stress : 0.0);
                    }
                    E->integ_eq_plastic_strain[index_i] += d_ep;
                    E->integ_yield_stress[index_i] += H * d_ep;
                }
            }
            for (int r = 0; r < 6; ++r) E->integ_strain[r + 6 * index_i] += d_e_vec[r];
            for (int r = 0; r < 6; ++r) E->integ_stress[r + 6 * index_i] = final_stress[r];
            for (int j = 0; j < 24; ++j) {                 // q_vec_i = Bfinal' * final_stress  J2:1330
                double s = Bfinal[6 * j] * final_stress[0];
                if (fused) for (int r = 1; r < 6; ++r) s = std::fma(Bfinal[r + 6 * j], final_stress[r], s);
                else for (int r = 1; r < 6; ++r) s += Bfinal[r + 6 * j] * final_stress[r];
                Qe[j] += W * W * W * detJ * s;
            }
        }
    }
    E->counters[0] += negJ;
}

// ---------------------------------------------------------------- cal_triax_stress (J2:982-1022)
static inline void eigvals_sym3(double a11, double a22, double a33, double a12, double a23, double a13,
                                double p[3]) {
    // closed-form eigenvalues of a real symmetric 3x3 (StaticArrays eigen.jl, Smith's algorithm)
    double p1 = a12 * a12 + a13 * a13 + a23 * a23;
    if (p1 == 0) {
        p[0] = a11; p[1] = a22; p[2] = a33;
        std::sort(p, p + 3);
        return;
    }
    double q = (a11 + a22 + a33) / 3;
    double p2 = (a11 - q) * (a11 - q) + (a22 - q) * (a22 - q) + (a33 - q) * (a33 - q) + 2 * p1;
    double pp = std::sqrt(p2 / 6);
    double invp = 1.0 / pp;
    double b11 = (a11 - q) * invp, b22 = (a22 - q) * invp, b33 = (a33 - q) * invp;
    double b12 = a12 * invp, b13 = a13 * invp, b23 = a23 * invp;
    double detB = b11 * (b22 * b33 - b23 * b23) - b12 * (b12 * b33 - b23 * b13) + b13 * (b12 * b23 - b22 * b13);
    double r = detB / 2;
    double phi;
    const double pi = 3.14159265358979323846;
    if (r <= -1) phi = pi / 3;
    else if (r >= 1) phi = 0.0;
    else phi = std::acos(r) / 3;
    double eig3 = q + 2 * pp * std::cos(phi);
    double eig1 = q + 2 * pp * std::cos(phi + (2 * pi / 3));
    double eig2 = 3 * q - eig1 - eig3;
    p[0] = eig1; p[1] = eig2; p[2] = eig3;
}

static void cal_triax_stress(hk_engine* E) {
    const int64_t n = E->nip;
    const int route = E->prm.triax_route;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        E->integ_triax_stress[i] = 0.0;
        const double* s = &E->integ_stress[6 * i];
        double ox = s[0], oy = s[1], oz = s[2], txy = s[3], tyz = s[4], txz = s[5];
        double oeq, mean3;
        if (route == 1) {
            double p[3];
            eigvals_sym3(ox, oy, oz, txy, tyz, txz, p);
            oeq = std::sqrt(0.5 * ((p[0] - p[1]) * (p[0] - p[1]) + (p[1] - p[2]) * (p[1] - p[2]) +
                                   (p[2] - p[0]) * (p[2] - p[0])));
            mean3 = (p[0] + p[1] + p[2]) / 3.0;
        } else {
            // same quantity from invariants (the reference's own commented formula, J2:1002)
            oeq = std::sqrt(0.5 * ((ox - oy) * (ox - oy) + (oy - oz) * (oy - oz) + (ox - oz) * (ox - oz) +
                                   6 * (txy * txy + tyz * tyz + txz * txz)));
            mean3 = (ox + oy + oz) / 3.0;
        }
        if (oeq < 1E-10) continue;
        E->integ_triax_stress[i] = mean3 / oeq;
    }
}

// ---------------------------------------------------------------- add_surface_triangle (J2:2167-2245)
static void add_surface_triangle(const InstanceO& I, int64_t ele_id, std::vector<int64_t>& add_tri /* rows x3 */,
                                 std::vector<int64_t>& add_eleid, std::vector<int64_t>& add_nodes) {
    const int64_t F = I.nElement * 6;
    std::vector<int64_t> add_surf;  // rows x4
    for (int j = 0; j < 6; ++j) {
        int64_t fj = 6 * (ele_id - 1) + j;
        int64_t sj[4];
        for (int q = 0; q < 4; ++q) sj[q] = I.sorted_surfaces[fj + F * q];
        for (int64_t k = 0; k < F; ++k) {
            if (I.surfaces_eleid[k] == ele_id) continue;
            if (sj[0] == I.sorted_surfaces[k] && sj[1] == I.sorted_surfaces[k + F] &&
                sj[2] == I.sorted_surfaces[k + 2 * F] && sj[3] == I.sorted_surfaces[k + 3 * F]) {
                for (int q = 0; q < 4; ++q) add_surf.push_back(I.surfaces[k + F * q]);
                add_eleid.push_back(I.surfaces_eleid[k]);
                add_eleid.push_back(I.surfaces_eleid[k]);
                break;
            }
        }
    }
    size_t ns = add_surf.size() / 4;
    for (size_t j = 0; j < ns; ++j) {
        const int64_t* s = &add_surf[4 * j];
        add_tri.push_back(s[0]); add_tri.push_back(s[1]); add_tri.push_back(s[2]);
        add_tri.push_back(s[2]); add_tri.push_back(s[3]); add_tri.push_back(s[0]);
    }
    add_nodes = add_tri;
    std::sort(add_nodes.begin(), add_nodes.end());
    add_nodes.erase(std::unique(add_nodes.begin(), add_nodes.end()), add_nodes.end());
}

static void append_unique(std::vector<int64_t>& v, const std::vector<int64_t>& add, int64_t offset) {
    // append!(v, add .+ offset); unique!(v)  (keeps first occurrences, order preserved)
    for (int64_t a : add) {
        int64_t x = a + offset;
        if (std::find(v.begin(), v.end(), x) == v.end()) v.push_back(x);
    }
}

// update contact surface, J2:767-804
static void update_contact_surface(hk_engine* E, const std::vector<int64_t>& deleted_element) {
    if (E->prm.contact_flag > 0) {
        for (int64_t i : deleted_element) {
            int64_t instance_id = E->element_instance[i - 1];
            const InstanceO& I = E->INSTANCE[instance_id - 1];
            int64_t ele_id = i - I.element_offset;
            std::vector<int64_t> add_tri, add_eleid, add_nodes;
            add_surface_triangle(I, ele_id, add_tri, add_eleid, add_nodes);
            for (ContactTriangleO& ct : E->CT) {
                if (ct.i_instance == instance_id) {
                    append_unique(ct.c_nodes_i, add_nodes, I.node_offset);
                } else if (ct.j_instance == instance_id) {
                    append_unique(ct.c_nodes_j, add_nodes, I.node_offset);
                    for (int64_t x : add_eleid) ct.c_triangles_eleid.push_back(x + I.element_offset);
                    for (size_t r = 0; r < add_tri.size() / 3; ++r) {
                        ct.t0.push_back(add_tri[3 * r + 0] + I.node_offset);
                        ct.t1.push_back(add_tri[3 * r + 1] + I.node_offset);
                        ct.t2.push_back(add_tri[3 * r + 2] + I.node_offset);
                    }
                }
            }
        }
    }
}


// ---------------------------------------------------------------- cal_contact_force (J2:2248-2706)
static void cal_contact_force(hk_engine* E) {
    const double* position = E->position.data();
    const double* velo = E->velo.data();
    const double d_lim = E->prm.element_min_size * E->prm.contact_d_lim_factor;
    const double myu = E->prm.contact_myu;
    const double kc_o = E->prm.contact_kc_other, kc_s = E->prm.contact_kc_self;
    const double Cr_o = E->prm.contact_cr_other, Cr_s = E->prm.contact_cr_self;
    const int64_t fn = E->fn;
    int64_t hits = 0, tests = 0;
    const bool clamp = E->prm.contact_dmax_clamp != 0;

    for (size_t c = 0; c < E->CT.size(); ++c) {
        const ContactTriangleO& ct = E->CT[c];
        const int64_t i_instance = ct.i_instance, j_instance = ct.j_instance;
        const std::vector<int64_t>& c_nodes_i = ct.c_nodes_i;
        const std::vector<int64_t>& c_nodes_j = ct.c_nodes_j;
        const double young = ct.young;
        const int64_t nn_i = (int64_t)c_nodes_i.size(), nn_j = (int64_t)c_nodes_j.size();
        const int64_t nTri = (int64_t)ct.t0.size();
        if (nn_i == 0 || nn_j == 0) continue;   // Julia's minimum() would throw; never happens in the decks

        double mn_i[3], mx_i[3], mn_j[3], mx_j[3];
        for (int a = 0; a < 3; ++a) {
            mn_i[a] = mx_i[a] = position[a + 3 * (c_nodes_i[0] - 1)];
            mn_j[a] = mx_j[a] = position[a + 3 * (c_nodes_j[0] - 1)];
        }
        for (int64_t k = 0; k < nn_i; ++k)
            for (int a = 0; a < 3; ++a) {
                double v = position[a + 3 * (c_nodes_i[k] - 1)];
                mn_i[a] = std::min(mn_i[a], v); mx_i[a] = std::max(mx_i[a], v);
            }
        for (int64_t k = 0; k < nn_j; ++k)
            for (int a = 0; a < 3; ++a) {
                double v = position[a + 3 * (c_nodes_j[k] - 1)];
                mn_j[a] = std::min(mn_j[a], v); mx_j[a] = std::max(mx_j[a], v);
            }
        double range_min[3], range_max[3], all_range_min[3];
        bool skip = false;
        for (int a = 0; a < 3; ++a) {
            range_min[a] = std::max(mn_i[a], mn_j[a]);
            range_max[a] = std::min(mx_i[a], mx_j[a]);
            all_range_min[a] = std::min(mn_i[a], mn_j[a]);
            if (range_min[a] > range_max[a]) skip = true;
        }
        if (skip) continue;

        double ddiv = E->prm.element_max_size * E->prm.contact_ddiv_other;
        if (i_instance == j_instance) ddiv = E->prm.element_max_size * E->prm.contact_ddiv_self;

        std::vector<int64_t> node_map_i(3 * nn_i), node_map_j(3 * nn_j);
        for (int64_t k = 0; k < nn_i; ++k)
            for (int a = 0; a < 3; ++a)
                node_map_i[a + 3 * k] = (int64_t)std::ceil((position[a + 3 * (c_nodes_i[k] - 1)] - all_range_min[a]) / ddiv);
        for (int64_t k = 0; k < nn_j; ++k)
            for (int a = 0; a < 3; ++a)
                node_map_j[a + 3 * k] = (int64_t)std::ceil((position[a + 3 * (c_nodes_j[k] - 1)] - all_range_min[a]) / ddiv);

        // The reference finds j0 in c_nodes_j by a linear search per triangle (J2:2465-2472);
        // a first-occurrence lookup table gives the same index without the O(nn_j) scan.
        std::vector<int64_t> first_j(E->nNode + 1, -1);
        for (int64_t k = nn_j - 1; k >= 0; --k) first_j[c_nodes_j[k]] = k;

#pragma omp parallel for schedule(dynamic, 64) reduction(+ : hits, tests)
        for (int64_t j = 0; j < nTri; ++j) {
            const int64_t eleid_ = ct.c_triangles_eleid[j];
            if (E->element_flag[eleid_ - 1] == 0) continue;
            double kc = kc_o, Cr = Cr_o;
            if (i_instance == j_instance) { kc = kc_s; Cr = Cr_s; }
            int index_th = 0;
#ifdef _OPENMP
            index_th = omp_get_thread_num();
#endif
            f128* cf = &E->c_force3[(size_t)fn * index_th];
            const int64_t j0 = ct.t0[j], j1 = ct.t1[j], j2 = ct.t2[j];
            const double q0x = position[0 + 3 * (j0 - 1)], q0y = position[1 + 3 * (j0 - 1)], q0z = position[2 + 3 * (j0 - 1)];
            const double q1x = position[0 + 3 * (j1 - 1)], q1y = position[1 + 3 * (j1 - 1)], q1z = position[2 + 3 * (j1 - 1)];
            const double q2x = position[0 + 3 * (j2 - 1)], q2y = position[1 + 3 * (j2 - 1)], q2z = position[2 + 3 * (j2 - 1)];
            if (q0x < range_min[0] && q1x < range_min[0] && q2x < range_min[0]) continue;
            if (q0y < range_min[1] && q1y < range_min[1] && q2y < range_min[1]) continue;
            if (q0z < range_min[2] && q1z < range_min[2] && q2z < range_min[2]) continue;
            if (q0x > range_max[0] && q1x > range_max[0] && q2x > range_max[0]) continue;
            if (q0y > range_max[1] && q1y > range_max[1] && q2y > range_max[1]) continue;
            if (q0z > range_max[2] && q1z > range_max[2] && q2z > range_max[2]) continue;

            const double cx = (q0x + q1x + q2x) / 3.0, cy = (q0y + q1y + q2y) / 3.0, cz = (q0z + q1z + q2z) / 3.0;
            const double R0 = my3norm(q0x - cx, q0y - cy, q0z - cz);
            const double R1 = my3norm(q1x - cx, q1y - cy, q1z - cz);
            const double R2 = my3norm(q2x - cx, q2y - cy, q2z - cz);
            const double Rmax = std::max(std::max(R0, R1), R2);
            const double v1x = q1x - q0x, v1y = q1y - q0y, v1z = q1z - q0z;
            const double v2x = q2x - q0x, v2y = q2y - q0y, v2z = q2z - q0z;
            const double L1 = my3norm(v1x, v1y, v1z), L2 = my3norm(v2x, v2y, v2z);
            const double Lmax = std::max(L1, L2);
            double nx, ny, nz;
            my3crossNNz(v1x, v1y, v1z, v2x, v2y, v2z, nx, ny, nz);
            const double d12 = v1x * v2x + v1y * v2y + v1z * v2z;
            const double S = 0.5 * std::sqrt(L1 * L1 * L2 * L2 - d12 * d12);
            const double A11 = v1x, A21 = v1y, A31 = v1z, A12 = v2x, A22 = v2y, A32 = v2z;
            const double A13 = -nx, A23 = -ny, A33 = -nz;

            int64_t map_j0[3] = {1, 1, 1};
            if (first_j[j0] >= 0)
                for (int a = 0; a < 3; ++a) map_j0[a] = node_map_j[a + 3 * first_j[j0]];
            int64_t en[8];
            for (int q = 0; q < 8; ++q) en[q] = E->elementmat[q + 8 * (eleid_ - 1)];

            for (int64_t k = 0; k < nn_i; ++k) {
                if (std::llabs(map_j0[0] - node_map_i[0 + 3 * k]) > 1 ||
                    std::llabs(map_j0[1] - node_map_i[1 + 3 * k]) > 1 ||
                    std::llabs(map_j0[2] - node_map_i[2 + 3 * k]) > 1)
                    continue;
                const int64_t i = c_nodes_i[k];
                if (i_instance == j_instance) {
                    bool own = false;
                    for (int q = 0; q < 8; ++q) own = own || (i == en[q]);
                    if (own) continue;
                }
                const double px = position[0 + 3 * (i - 1)], py = position[1 + 3 * (i - 1)], pz = position[2 + 3 * (i - 1)];
                if (px < range_min[0] || py < range_min[1] || pz < range_min[2]) continue;
                if (px > range_max[0] || py > range_max[1] || pz > range_max[2]) continue;
                const double dpc = my3norm(px - cx, py - cy, pz - cz);
                if (dpc >= Rmax) continue;
                const double bx = px - q0x, by = py - q0y, bz = pz - q0z;
                double x1, x2, d;
                ++tests;
                my3SolveAb(A11, A21, A31, A12, A22, A32, A13, A23, A33, bx, by, bz, x1, x2, d);
                if (0.0 <= x1 && 0.0 <= x2 && x1 + x2 <= 1.0 && d > 0.0 && d <= d_lim) {
                    ++hits;
                    if (clamp && d - E->d_node_pre[i - 1] > E->d_max) d = E->d_node_pre[i - 1] + E->d_max;   // J1:2756-2758
                    const double vx = velo[i * 3 - 3] - velo[j0 * 3 - 3];
                    const double vy = velo[i * 3 - 2] - velo[j0 * 3 - 2];
                    const double vz = velo[i * 3 - 1] - velo[j0 * 3 - 1];
                    const double mag_v = my3norm(vx, vy, vz);
                    double vex = 0.0, vey = 0.0, vez = 0.0;
                    if (mag_v > 0.0) { vex = vx / mag_v; vey = vy / mag_v; vez = vz / mag_v; }
                    const double k_ = young * S / Lmax * kc;
                    const double F = k_ * d;
                    double fx = F * nx, fy = F * ny, fz = F * nz;
                    const double C = 2 * std::sqrt(E->diag_M[i - 1] * k_) * Cr;   // diag_M[i]: sic, J2:2593
                    const double fc_x = -C * vx, fc_y = -C * vy, fc_z = -C * vz;
                    const double dot_ve_n = vex * nx + vey * ny + vez * nz;
                    const double vsx = vex - dot_ve_n * nx, vsy = vey - dot_ve_n * ny, vsz = vez - dot_ve_n * nz;
                    const double fric_x = -myu * F * vsx, fric_y = -myu * F * vsy, fric_z = -myu * F * vsz;
                    fx += fric_x + fc_x;
                    fy += fric_y + fc_y;
                    fz += fric_z + fc_z;
                    cf[0 + (i - 1) * 3] += fx;
                    cf[1 + (i - 1) * 3] += fy;
                    cf[2 + (i - 1) * 3] += fz;
                    const int64_t jn[3] = {j0, j1, j2};
                    for (int q = 0; q < 3; ++q) {
                        cf[0 + (jn[q] - 1) * 3] += -fx / 3.0;
                        cf[1 + (jn[q] - 1) * 3] += -fy / 3.0;
                        cf[2 + (jn[q] - 1) * 3] += -fz / 3.0;
                    }
                    if (clamp) {                                             // J1:2898-2900 (a max: order independent)
#pragma omp critical(hk_d_node)
                        if (d > E->d_node[i - 1]) E->d_node[i - 1] = d;
                    }
                }
            }
        }
    }
    E->counters[1] += hits;
    E->counters[2] += tests;
}

// ---------------------------------------------------------------- one time step (J2:487-951)
static void one_step(hk_engine* E, int64_t t, int64_t* n_deleted) {
    const int64_t fn = E->fn, nNode = E->nNode, nElement = E->nElement;
    const double d_time = E->prm.d_time;
    const int integ_num = 8;

    std::fill(E->external_force.begin(), E->external_force.end(), 0.0);      // J2:497

    if (E->prm.contact_flag >= 1) {                                          // J2:500-548
        std::fill(E->c_force3.begin(), E->c_force3.end(), (f128)0);
        if (E->prm.contact_dmax_clamp) std::fill(E->d_node.begin(), E->d_node.end(), 0.0);      // J1:492
        cal_contact_force(E);
        if (E->prm.contact_dmax_clamp) E->d_node_pre = E->d_node;                              // J1:512-514
        if (E->Nth > 1) {
#pragma omp parallel for schedule(static)
            for (int64_t i = 0; i < fn; ++i)
                for (int j = 1; j < E->Nth; ++j) E->c_force3[i] += E->c_force3[i + (size_t)fn * j];
        }
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < fn; ++i)
            E->external_force[i] = (double)((f128)E->external_force[i] + E->c_force3[i]);
    }

    // central difference update, J2:562-567
    const double dt2 = d_time * d_time;            // d_time^2
    const double dt2p = std::pow(d_time, 2.0);     // d_time^2.0
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < fn; ++i) {
        const double M = E->diag_M[i], C = E->diag_C[i];
        E->disp_new[i] = 1.0 / (M / dt2 + C / 2.0 / d_time) *
                         (E->external_force[i] - E->Q[i] + M / dt2p * (2.0 * E->disp[i] - E->disp_pre[i]) +
                          C / 2.0 / d_time * E->disp_pre[i]);
    }

    // boundary conditions, J2:585-617
    for (const BCEntry& bc : E->BC) {
        double amp = 1.0;
        if (bc.has_amp) {
            size_t time_index = 0;
            const double current_time = (double)t * d_time;
            for (size_t j = 0; j + 1 < bc.a_t.size(); ++j)
                if (current_time >= bc.a_t[j] && current_time <= bc.a_t[j + 1]) { time_index = j; break; }
            amp = bc.a_v[time_index] + (bc.a_v[time_index + 1] - bc.a_v[time_index]) *
                                           (current_time - bc.a_t[time_index]) /
                                           (bc.a_t[time_index + 1] - bc.a_t[time_index]);
        }
        for (size_t j = 0; j < bc.dof.size(); ++j) {
            const double v = bc.value[j];
            for (int64_t dof : bc.dof[j]) E->disp_new[dof - 1] = v * amp;
        }
    }

    // kinematics, J2:624-657
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < fn; ++i) {
        E->d_disp[i] = E->disp_new[i] - E->disp[i];
        E->disp_pre[i] = E->disp[i];
        E->disp[i] = E->disp_new[i];
        E->velo[i] = E->d_disp[i] / d_time;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nNode; ++i) {
        double x = E->d_disp[i * 3], y = E->d_disp[i * 3 + 1], z = E->d_disp[i * 3 + 2];
        E->d_disp_norm[i] = std::sqrt(x * x + y * y + z * z);
        E->position[0 + 3 * i] = E->coordmat[0 + 3 * i] + E->disp[i * 3];
        E->position[1 + 3 * i] = E->coordmat[1 + 3 * i] + E->disp[i * 3 + 1];
        E->position[2 + 3 * i] = E->coordmat[2 + 3 * i] + E->disp[i * 3 + 2];
    }
    if (E->prm.contact_dmax_clamp) {                                         // d_max = maximum(d_disp_norm), J1:618
        double m = E->d_disp_norm[0];
        for (int64_t i = 1; i < nNode; ++i) m = E->d_disp_norm[i] > m ? E->d_disp_norm[i] : m;
        E->d_max = m;
    }

    // internal force, J2:662-675
    std::fill(E->Qe.begin(), E->Qe.end(), 0.0);
    cal_stress_hexa(E);
    std::fill(E->Q.begin(), E->Q.end(), 0.0);
    for (int64_t e = 0; e < nElement; ++e)
        for (int i = 0; i < 8; ++i) {
            int64_t nd = E->elementmat[i + 8 * e] - 1;
            E->Q[0 + nd * 3] += E->Qe[0 + i * 3 + 24 * e];
            E->Q[1 + nd * 3] += E->Qe[1 + i * 3 + 24 * e];
            E->Q[2 + nd * 3] += E->Qe[2 + i * 3 + 24 * e];
        }

    cal_triax_stress(E);                                                     // J2:677

    // fracture, J2:682-764
    std::vector<int64_t> deleted_element;
    if (E->flag_fracture == 1) {
        for (int64_t i = 0; i < nElement; ++i) {
            const Material& M = E->MATERIAL[E->element_material[i] - 1];
            const int64_t nd = M.nd;
            if (nd > 0) {
                double v_e = 0.0, t_e = 0.0;
                for (int j = 0; j < integ_num; ++j) {
                    v_e += E->integ_eq_plastic_strain[j + i * integ_num];
                    t_e += E->integ_triax_stress[j + i * integ_num];
                }
                v_e /= integ_num;
                t_e /= integ_num;
                if (t_e < 0) continue;
                const double* d = M.ductile.data();   // d[r + nd*c]
                double fr_e = d[(nd - 1) + nd * 0];
                for (int64_t j = 0; j + 1 < nd; ++j) {
                    if (t_e >= d[j + nd * 1] && t_e < d[j + 1 + nd * 1]) {
                        fr_e = d[j] + (d[j + 1] - d[j]) / (d[j + 1 + nd] - d[j + nd]) * (t_e - d[j + nd]);
                        break;
                    }
                }
                if (v_e >= fr_e && E->element_flag[i] == 1) {
                    E->element_flag[i] = 0;
                    deleted_element.push_back(i + 1);
                    for (int j = 0; j < integ_num; ++j)
                        for (int r = 0; r < 6; ++r) {
                            E->integ_stress[r + 6 * (j + i * integ_num)] = 0.0;
                            E->integ_strain[r + 6 * (j + i * integ_num)] = 0.0;
                        }
                }
            }
        }
    }

    update_contact_surface(E, deleted_element);                              // J2:767-804
    for (int64_t i : deleted_element) { E->deleted_all.push_back(i); E->deleted_step.push_back(t); }
    if (n_deleted) *n_deleted += (int64_t)deleted_element.size();
    E->counters[4] += 1;
}

template <class T>
static void cp(T* dst, const std::vector<T>& src) { if (dst) std::memcpy(dst, src.data(), src.size() * sizeof(T)); }
template <class T>
static void up(std::vector<T>& dst, const T* src) { if (src) std::memcpy(dst.data(), src, dst.size() * sizeof(T)); }

// ================================================================ exported ABI (prefix hko_)
extern "C" {

int hko_default_params(hk_params* p) {
    if (!p) return HK_ERR_ARG;
    std::memset(p, 0, sizeof(*p));
    p->struct_size = (int32_t)sizeof(hk_params);
    p->device = 0;
    p->d_time = 0.0;
    p->element_min_size = 0.0;
    p->element_max_size = 0.0;
    p->contact_flag = 0;
    p->triax_route = 0;
    p->contact_d_lim_factor = 0.3;
    p->contact_myu = 0.25;
    p->contact_kc_other = 1.0;
    p->contact_kc_self = 1.0;
    p->contact_cr_other = 0.0;
    p->contact_cr_self = 0.0;
    p->contact_ddiv_other = 1.1;
    p->contact_ddiv_self = 0.6;
    p->deterministic = 1;
    return HK_OK;
}

int hko_create(hk_engine** out, const hk_params* p) {
    if (!out || !p) return fail(nullptr, HK_ERR_ARG, "null argument");
    if (p->struct_size != (int32_t)sizeof(hk_params)) return fail(nullptr, HK_ERR_ARG, "hk_params size mismatch");
    hk_engine* e = new hk_engine();
    e->prm = *p;
    if (const char* m = std::getenv("HKO_MATVEC")) e->matvec_mode = std::strcmp(m, "muladd") == 0 ? 1 : 0;
    cal_Pusai_hexa(e->Pusai);
    *out = e;
    return HK_OK;
}

int hko_destroy(hk_engine* e) { delete e; return HK_OK; }

const char* hko_last_error(const hk_engine* e) { return e ? e->err.c_str() : g_create_err.c_str(); }

int hko_set_mesh(hk_engine* e, int64_t nNode, int64_t nElement, const double* coordmat, const int64_t* elementmat,
                 const int64_t* element_material, const int64_t* element_instance, const double* diag_M) {
    if (!e || !coordmat || !elementmat || !element_material || !diag_M) return fail(e, HK_ERR_ARG, "null argument");
    e->nNode = nNode; e->nElement = nElement; e->fn = 3 * nNode; e->nip = 8 * nElement;
    e->coordmat.assign(coordmat, coordmat + 3 * nNode);
    e->elementmat.assign(elementmat, elementmat + 8 * nElement);
    e->element_material.assign(element_material, element_material + nElement);
    if (element_instance) e->element_instance.assign(element_instance, element_instance + nElement);
    else e->element_instance.assign(nElement, 1);
    e->diag_M.assign(diag_M, diag_M + 3 * nNode);
    e->diag_C.assign(3 * nNode, 0.0);
    const double C = 0.0;                                           // J2:217-218
    for (int64_t i = 0; i < 3 * nNode; ++i) e->diag_C[i] = e->diag_M[i] * C;
    return HK_OK;
}

int hko_add_material(hk_engine* e, double young, double poisson, double density, int64_t npp, const double* plastic,
                     const double* Hd, int64_t nd, const double* ductile) {
    if (!e) return HK_ERR_ARG;
    Material m;
    m.young = young; m.poisson = poisson; m.density = density;
    m.G = young / 2. / (1.0 + poisson);                             // J2:146
    const double d1 = (1.0 - poisson), d2 = poisson, d3 = (1.0 - 2.0 * poisson) / 2.0;
    const double f = young / (1.0 + poisson) / (1.0 - 2.0 * poisson);
    const double D[6][6] = {{d1, d2, d2, 0, 0, 0}, {d2, d1, d2, 0, 0, 0}, {d2, d2, d1, 0, 0, 0},
                            {0, 0, 0, d3, 0, 0},   {0, 0, 0, 0, d3, 0},   {0, 0, 0, 0, 0, d3}};
    for (int r = 0; r < 6; ++r)
        for (int c = 0; c < 6; ++c) m.Dmat[r + 6 * c] = f * D[r][c];
    m.npp = npp;
    if (npp > 0) { m.plastic.assign(plastic, plastic + 2 * npp); if (npp > 1) m.Hd.assign(Hd, Hd + npp - 1); }
    m.nd = nd;
    if (nd > 0) m.ductile.assign(ductile, ductile + 3 * nd);
    e->MATERIAL.push_back(m);
    return HK_OK;
}

int hko_add_bc(hk_engine* e, int64_t n_lists, const int64_t* dof_ptr, const int64_t* dofs, const double* values,
               int64_t n_amp, const double* amp_time, const double* amp_value) {
    if (!e) return HK_ERR_ARG;
    BCEntry b;
    for (int64_t j = 0; j < n_lists; ++j) {
        b.dof.emplace_back(dofs + dof_ptr[j], dofs + dof_ptr[j + 1]);
        b.value.push_back(values[j]);
    }
    if (n_amp > 0) {
        if (n_amp < 2) return fail(e, HK_ERR_ARG, "amplitude table needs >= 2 points");
        b.has_amp = true;
        b.a_t.assign(amp_time, amp_time + n_amp);
        b.a_v.assign(amp_value, amp_value + n_amp);
    }
    e->BC.push_back(b);
    return HK_OK;
}

int hko_add_ic(hk_engine* e, int64_t n_lists, const int64_t* dof_ptr, const int64_t* dofs, const double* values) {
    if (!e) return HK_ERR_ARG;
    ICEntry b;
    for (int64_t j = 0; j < n_lists; ++j) {
        b.dof.emplace_back(dofs + dof_ptr[j], dofs + dof_ptr[j + 1]);
        b.value.push_back(values[j]);
    }
    e->IC.push_back(b);
    return HK_OK;
}

int hko_add_instance(hk_engine* e, int64_t node_offset, int64_t nNode, int64_t element_offset, int64_t nElement,
                     const int64_t* surfaces, const int64_t* surfaces_eleid) {
    if (!e) return HK_ERR_ARG;
    InstanceO I;
    I.node_offset = node_offset; I.nNode = nNode; I.element_offset = element_offset; I.nElement = nElement;
    const int64_t F = 6 * nElement;
    if (surfaces) {
        I.surfaces.assign(surfaces, surfaces + 4 * F);
        I.surfaces_eleid.assign(surfaces_eleid, surfaces_eleid + F);
        I.sorted_surfaces.resize(4 * F);
        for (int64_t j = 0; j < F; ++j) {                            // J2:1987-1989
            int64_t s[4] = {surfaces[j], surfaces[j + F], surfaces[j + 2 * F], surfaces[j + 3 * F]};
            std::sort(s, s + 4);
            for (int q = 0; q < 4; ++q) I.sorted_surfaces[j + F * q] = s[q];
        }
    }
    e->INSTANCE.push_back(I);
    return HK_OK;
}

int hko_add_contact_pair(hk_engine* e, int64_t i_instance, int64_t j_instance, int64_t nn_i, const int64_t* c_nodes_i,
                         int64_t nn_j, const int64_t* c_nodes_j, int64_t nTri, const int64_t* c_triangles,
                         const int64_t* c_triangles_eleid, double young) {
    if (!e) return HK_ERR_ARG;
    ContactTriangleO ct;
    ct.i_instance = i_instance; ct.j_instance = j_instance; ct.young = young;
    ct.c_nodes_i.assign(c_nodes_i, c_nodes_i + nn_i);
    ct.c_nodes_j.assign(c_nodes_j, c_nodes_j + nn_j);
    ct.t0.assign(c_triangles, c_triangles + nTri);
    ct.t1.assign(c_triangles + nTri, c_triangles + 2 * nTri);
    ct.t2.assign(c_triangles + 2 * nTri, c_triangles + 3 * nTri);
    ct.c_triangles_eleid.assign(c_triangles_eleid, c_triangles_eleid + nTri);
    e->CT.push_back(ct);
    return HK_OK;
}

int hko_finalize(hk_engine* e) {
    if (!e) return HK_ERR_ARG;
    if (e->nNode == 0) return fail(e, HK_ERR_STATE, "hk_set_mesh not called");
    const int64_t fn = e->fn, nip = e->nip, nE = e->nElement;
    e->position = e->coordmat;                                      // J2:222
    e->disp.assign(fn, 0.0); e->disp_new.assign(fn, 0.0); e->disp_pre.assign(fn, 0.0);
    e->d_disp.assign(fn, 0.0); e->velo.assign(fn, 0.0);
    for (const ICEntry& ic : e->IC)                                 // J2:233-239
        for (size_t j = 0; j < ic.dof.size(); ++j)
            for (int64_t d : ic.dof[j]) {
                e->disp_pre[d - 1] = -ic.value[j] * e->prm.d_time;
                e->velo[d - 1] = ic.value[j];
            }
    e->external_force.assign(fn, 0.0); e->Q.assign(fn, 0.0); e->Qe.assign(24 * nE, 0.0);
    e->integ_stress.assign(6 * nip, 0.0); e->integ_strain.assign(6 * nip, 0.0);
    e->integ_eq_plastic_strain.assign(nip, 0.0); e->integ_triax_stress.assign(nip, 0.0);
    e->integ_yield_stress.assign(nip, 0.0);
    e->element_flag.assign(nE, 1);
    e->elementVolume.assign(nE, 0.0);
    e->d_disp_norm.assign(e->nNode, 0.0);
    e->d_node.assign(e->nNode, 0.0);
    e->d_node_pre.assign(e->nNode, 0.0);
    e->d_max = 0.0;
    for (int64_t i = 0; i < nE; ++i) {                              // J2:456-465
        const Material& M = e->MATERIAL.at(e->element_material[i] - 1);
        if (M.npp > 0)
            for (int k = 0; k < 8; ++k) e->integ_yield_stress[i * 8 + k] = M.plastic[0];
    }
    e->Nth = 1;
#ifdef _OPENMP
    e->Nth = omp_get_max_threads();
#endif
    if (e->prm.contact_flag >= 1) e->c_force3.assign((size_t)fn * e->Nth, (f128)0);   // J2:435
    e->finalized = true;
    return HK_OK;
}

int hko_step(hk_engine* e, int64_t t_first, int64_t n_steps, int64_t* n_deleted_out) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    int64_t nd = 0;
    for (int64_t t = t_first; t < t_first + n_steps; ++t) one_step(e, t, &nd);
    if (n_deleted_out) *n_deleted_out = nd;
    return HK_OK;
}

static int64_t g_pending_deleted = 0;
int hko_step_enqueue(hk_engine* e, int64_t t_first, int64_t n_steps) {
    int64_t nd = 0;
    int rc = hko_step(e, t_first, n_steps, &nd);
    if (e) e->counters[7] += nd;     // reported by the next hko_sync
    (void)g_pending_deleted;
    return rc;
}
int hko_sync(hk_engine* e, int64_t* n_deleted_out) {
    if (!e) return HK_ERR_ARG;
    if (n_deleted_out) *n_deleted_out = e->counters[7];
    e->counters[7] = 0;
    return HK_OK;
}

int hko_download(hk_engine* e, double* disp, double* velo, double* integ_stress, double* integ_strain,
                 double* integ_eq_plastic_strain, double* integ_triax_stress, int64_t* element_flag) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    cp(disp, e->disp); cp(velo, e->velo); cp(integ_stress, e->integ_stress); cp(integ_strain, e->integ_strain);
    cp(integ_eq_plastic_strain, e->integ_eq_plastic_strain); cp(integ_triax_stress, e->integ_triax_stress);
    cp(element_flag, e->element_flag);
    return HK_OK;
}

// cal_node_stress_strain, J2:3408-3486, loop for loop (Julia (nNode,6) column-major outputs)
int hko_node_output(hk_engine* e, double* node_stress, double* node_strain, double* node_eq_plastic_strain,
                    double* node_mises_stress, double* node_triax_stress, double* inc_num, int32_t raw) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    const int64_t nN = e->nNode, nE = e->nElement;
    std::vector<double> es(nE * 6), en(nE * 6), ep(nE), et(nE);
    for (int64_t i = 0; i < nE; ++i) {                                     // J2:3428-3440
        for (int c = 0; c < 6; ++c) {
            double a = 0.0, b = 0.0;
            for (int k = 0; k < 8; ++k) { a += e->integ_stress[(i * 8 + k) * 6 + c]; b += e->integ_strain[(i * 8 + k) * 6 + c]; }
            es[i * 6 + c] = a / 8; en[i * 6 + c] = b / 8;
        }
        double a = e->integ_eq_plastic_strain[i * 8], b = e->integ_triax_stress[i * 8];
        for (int k = 1; k < 8; ++k) { a += e->integ_eq_plastic_strain[i * 8 + k]; b += e->integ_triax_stress[i * 8 + k]; }
        ep[i] = a / 8; et[i] = b / 8;
    }
    std::vector<double> ns(nN * 6, 0.0), nn(nN * 6, 0.0), np_(nN, 0.0), nt(nN, 0.0), inc(nN, 0.0), mises(nN, 0.0);
    for (int64_t i = 0; i < nE; ++i)                                       // J2:3442-3454
        for (int k = 0; k < 8; ++k) {
            const int64_t nd = e->elementmat[i * 8 + k] - 1;
            for (int c = 0; c < 6; ++c) { ns[c * nN + nd] += es[i * 6 + c]; nn[c * nN + nd] += en[i * 6 + c]; }
            np_[nd] += ep[i];
            nt[nd] += et[i];
        }
    for (int64_t i = 0; i < nE; ++i)                                       // J2:3456-3460
        for (int k = 0; k < 8; ++k) inc[e->elementmat[i * 8 + k] - 1] += 1.0;
    if (!raw)
        for (int64_t i = 0; i < nN; ++i) {                                 // J2:3463-3481
            for (int c = 0; c < 6; ++c) { ns[c * nN + i] /= inc[i]; nn[c * nN + i] /= inc[i]; }
            np_[i] /= inc[i];
            nt[i] /= inc[i];
            const double ox = ns[i], oy = ns[nN + i], oz = ns[2 * nN + i], txy = ns[3 * nN + i], tyz = ns[4 * nN + i],
                         txz = ns[5 * nN + i];
            mises[i] = std::sqrt(0.5 * ((ox - oy) * (ox - oy) + (oy - oz) * (oy - oz) + (ox - oz) * (ox - oz) +
                                        6 * (txy * txy + tyz * tyz + txz * txz)));
        }
    cp(node_stress, ns); cp(node_strain, nn); cp(node_eq_plastic_strain, np_); cp(node_triax_stress, nt); cp(inc_num, inc);
    if (!raw) cp(node_mises_stress, mises);
    return HK_OK;
}

int hko_download_ex(hk_engine* e, double* disp_pre, double* Q, double* external_force, double* position,
                    double* integ_yield_stress, double* elementVolume) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    cp(disp_pre, e->disp_pre); cp(Q, e->Q); cp(external_force, e->external_force); cp(position, e->position);
    cp(integ_yield_stress, e->integ_yield_stress); cp(elementVolume, e->elementVolume);
    return HK_OK;
}

int hko_upload_state(hk_engine* e, const double* disp, const double* disp_pre, const double* velo, const double* Q,
                     const double* integ_stress, const double* integ_strain, const double* integ_eq_plastic_strain,
                     const double* integ_yield_stress, const int64_t* element_flag) {
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    up(e->disp, disp); up(e->disp_pre, disp_pre); up(e->velo, velo); up(e->Q, Q);
    up(e->integ_stress, integ_stress); up(e->integ_strain, integ_strain);
    up(e->integ_eq_plastic_strain, integ_eq_plastic_strain); up(e->integ_yield_stress, integ_yield_stress);
    up(e->element_flag, element_flag);
    if (disp)                                                       // position = coordmat + disp, J2:650-652
        for (int64_t i = 0; i < e->fn; ++i) e->position[i] = e->coordmat[i] + e->disp[i];
    return HK_OK;
}

int hko_deleted_ids(hk_engine* e, int64_t* ids, int64_t cap, int64_t* n_out) {
    if (!e) return HK_ERR_ARG;
    int64_t n = (int64_t)e->deleted_all.size();
    if (n_out) *n_out = n;
    if (ids) for (int64_t i = 0; i < std::min(n, cap); ++i) ids[i] = e->deleted_all[i];
    return HK_OK;
}

int hko_deleted_steps(hk_engine* e, int64_t* steps, int64_t cap, int64_t* n_out) {
    if (!e) return HK_ERR_ARG;
    int64_t n = (int64_t)e->deleted_step.size();
    if (n_out) *n_out = n;
    if (steps) for (int64_t i = 0; i < std::min(n, cap); ++i) steps[i] = e->deleted_step[i];
    return HK_OK;
}

int hko_contact_pair_info(hk_engine* e, int64_t c, int64_t* nn_i, int64_t* nn_j, int64_t* nTri, int64_t* c_nodes_i,
                          int64_t* c_nodes_j, int64_t* c_triangles, int64_t* c_triangles_eleid) {
    if (!e || c < 0 || c >= (int64_t)e->CT.size()) return fail(e, HK_ERR_ARG, "bad contact pair index");
    const ContactTriangleO& ct = e->CT[c];
    const int64_t nt = (int64_t)ct.t0.size();
    if (nn_i) *nn_i = (int64_t)ct.c_nodes_i.size();
    if (nn_j) *nn_j = (int64_t)ct.c_nodes_j.size();
    if (nTri) *nTri = nt;
    cp(c_nodes_i, ct.c_nodes_i); cp(c_nodes_j, ct.c_nodes_j); cp(c_triangles_eleid, ct.c_triangles_eleid);
    if (c_triangles) {
        std::memcpy(c_triangles, ct.t0.data(), nt * 8);
        std::memcpy(c_triangles + nt, ct.t1.data(), nt * 8);
        std::memcpy(c_triangles + 2 * nt, ct.t2.data(), nt * 8);
    }
    return HK_OK;
}

int hko_counters(hk_engine* e, int64_t out[8]) {
    if (!e) return HK_ERR_ARG;
    for (int i = 0; i < 8; ++i) out[i] = e->counters[i];
    return HK_OK;
}

// hk_state_summary twin: plain loops over the reference's arrays
int hko_state_summary(hk_engine* e, double out[8]) {
    if (!e || !e->finalized || !out) return fail(e, HK_ERR_STATE, "engine not finalised");
    for (int i = 0; i < 8; ++i) out[i] = 0.0;
    bool any = false;
    for (int64_t el = 0; el < e->nElement; ++el) {
        if (e->element_flag[el] != 1) continue;
        out[0] += 1.0;
        for (int k = 0; k < 8; ++k) {
            const double ep = e->integ_eq_plastic_strain[el * 8 + k];
            if (!any || ep < out[1]) out[1] = ep;
            if (!any || ep > out[2]) out[2] = ep;
            any = true;
            if (ep > 0.0) out[3] += 1.0;
        }
    }
    return HK_OK;
}

int hko_profile(hk_engine*, int32_t) { return HK_OK; }
int hko_profile_read(hk_engine*, double ms[4], int64_t launches[4]) {
    for (int i = 0; i < 4; ++i) { ms[i] = 0; launches[i] = 0; }
    return HK_OK;
}
int hko_profile_read_ex(hk_engine*, double ms[8], int64_t launches[8]) {
    for (int i = 0; i < 8; ++i) { ms[i] = 0; launches[i] = 0; }
    return HK_OK;
}
int hko_set_stream(hk_engine*, void*) { return HK_OK; }
// the oracle is single-domain (the reference has no distributed code): halo calls are rejected
int hko_set_halo(hk_engine* e, int64_t, const int64_t*, const int64_t*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_halo_bind(hk_engine* e, int64_t, void*, void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_halo_pack(hk_engine* e) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_set_halo_ranks(hk_engine* e, int64_t, int64_t, const int64_t*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_comm_unique_id(void*) { return HK_ERR_UNSUPPORTED; }
int hko_comm_init(hk_engine* e, const void*, int32_t, int32_t) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_build_contact(hk_engine* e, int64_t, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const double*, int64_t,
                      const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*, const int64_t*) {
    return fail(e, HK_ERR_UNSUPPORTED, "the oracle takes the host-built contact tables (hko_add_instance / hko_add_contact_pair)");
}
int hko_comm_contact(hk_engine* e, int64_t, const int64_t*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_comm_erosion(hk_engine* e, int32_t) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_set_node_list(hk_engine* e, int32_t, int64_t, const int64_t*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_nodes_export(hk_engine* e, void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_nodes_import(hk_engine* e, const void*, const int64_t*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_contact_enqueue(hk_engine* e) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_contact_export(hk_engine* e, void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_contact_import(hk_engine* e, const void*, int64_t) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_state_export(hk_engine* e, void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_state_import(hk_engine* e, const void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_contact_export_limbs(hk_engine* e, void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_contact_import_limbs(hk_engine* e, const void*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_set_global_maps(hk_engine* e, int64_t, const int64_t*, int64_t, const int64_t*, const int64_t*) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_apply_deleted(hk_engine* e, int64_t n, const int64_t* ids) {           // restart: replay + record (local ids)
    if (!e || !e->finalized) return fail(e, HK_ERR_STATE, "engine not finalised");
    std::vector<int64_t> v(ids, ids + n);
    for (int64_t g : v) if (g < 1 || g > e->nElement) return fail(e, HK_ERR_ARG, "element id out of range");
    update_contact_surface(e, v);
    for (int64_t g : v) { e->deleted_all.push_back(g); e->deleted_step.push_back(0); }
    return HK_OK;
}
int hko_mark_frame(hk_engine* e) { return e && e->finalized ? HK_OK : fail(e, HK_ERR_STATE, "engine not finalised"); }   // triax is always stored
int hko_step_begin(hk_engine* e, int64_t) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }
int hko_step_finish(hk_engine* e, int64_t) { return fail(e, HK_ERR_UNSUPPORTED, "oracle is single-domain"); }

}  // extern "C"
