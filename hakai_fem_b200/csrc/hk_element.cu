// hk_element.cu — per-element hex8 internal-force kernel (FP64, FMA allowed).
//
// Replaces cal_stress_hexa + cal_BVbar_hexa + cal_Bfinal (J2:1033-1371, 1705-1784, 1415-1519),
// cal_triax_stress (J2:982-1022) and the fracture loop (J2:682-764) with ONE pass over the element:
// node gather, Jacobians at the 8 Gauss points, mean-dilatation (B-bar) strain increment, J2 radial
// return, state update, nodal force, triaxiality, ductile-damage deletion.
//
// The reference builds the dense 6x24 Bfinal = B - BV + BVbar and multiplies; this kernel evaluates
// the algebraically identical 3x3 tensor form:
//     g_a(k)  = adj(J_k) * Pusai_k[:,a]            (= detJ_k * grad N_a at Gauss point k)
//     Gbar_a  = sum_k g_a(k),  V = sum_k |detJ_k|  (BVbar rows = Gbar_a / (3V))
//     L_k     = (sum_a du_a (x) g_a(k)) / detJ_k,  trbar = (sum_a du_a . Gbar_a) / V
//     d_eps   = sym(L_k) + (trbar - tr L_k)/3 * I  (engineering shear)
//     f_a     = sum_k s_k * g_a(k) + (sum_k p_k detJ_k)/V * Gbar_a,   s = dev(sigma), p = tr(sigma)/3
// which is ~4x fewer flops and needs no 6x24 temporaries.  Results agree with the dense form to
// rounding (tests/test_element_parity.py states the tolerance).
#include "hk_common.h"

// Pusai_mat[k][dir][node] (J2:1895-1943), computed on the host with the reference's expression
#ifndef HK_EMU
__constant__ double c_P[8][3][8];
#else
static double c_P[8][3][8];
#endif

void hk_upload_pusai(const double* P) {
#ifndef HK_EMU
    cudaMemcpyToSymbol(c_P, P, sizeof(double) * 192);
#else
    memcpy(c_P, P, sizeof(double) * 192);
#endif
}

struct ElemArgs {
    HkDev d;
    long long step;
    int write_triax;
};

// Jacobian, its adjugate (cofactor transpose) and determinant at Gauss point k
HK_HD void jac_adj(const double x[8][3], int k, double adj[3][3], double& det) {
    double J[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
#pragma unroll
            for (int a = 0; a < 8; ++a) s += c_P[k][r][a] * x[a][c];
            J[r][c] = s;
        }
    // adj[i][j] such that inv(J) = adj / det  (same cofactors as J2:1445-1455)
    adj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    adj[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    adj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    adj[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    adj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    adj[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    adj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    adj[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    adj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    det = J[0][0] * adj[0][0] + J[0][1] * adj[1][0] + J[0][2] * adj[2][0];
}

HK_HD double triax_of(const double s[6]) {
    // (I1/3)/sqrt(3 J2): equals mean(p)/sigma_eq of the principal stresses p (J2:1004-1016)
    const double oeq = sqrt(0.5 * ((s[0] - s[1]) * (s[0] - s[1]) + (s[1] - s[2]) * (s[1] - s[2]) +
                                   (s[0] - s[2]) * (s[0] - s[2]) + 6.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5])));
    if (oeq < 1E-10) return 0.0;
    return (s[0] + s[1] + s[2]) / 3.0 / oeq;
}

HK_D void element_body(const ElemArgs& A, long long e) {
    const HkDev& d = A.d;
    const long long nEp = d.nEp;
    const unsigned char fl = d.flag[e];
    if (fl != 1) {
        if (fl == 0) {      // deleted during the previous step: its last force has been consumed, clear it
#pragma unroll
            for (int r = 0; r < 24; ++r) d.Qe[(long long)r * nEp + e] = 0.0;
#pragma unroll
            for (int k = 0; k < 8; ++k) d.triax[(long long)k * nEp + e] = 0.0;   // zero stress -> triax 0 (J2:1012)
            d.flag[e] = 2;
        }
        return;
    }
    const HkMaterialDev& M = d.mats[d.mat[e]];

    double x[8][3], du[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a) {
        const long long n = d.conn[(long long)a * nEp + e];
        const double* r = d.rec + 6 * n;
#pragma unroll
        for (int c = 0; c < 3; ++c) { x[a][c] = r[c]; du[a][c] = r[3 + c]; }
    }

    // ---- pass A: volume and mean gradients (cal_BVbar_hexa, J2:1705-1784)
    double Gbar[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a) Gbar[a][0] = Gbar[a][1] = Gbar[a][2] = 0.0;
    double V = 0.0;
    int negj = 0;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        double adj[3][3], det;
        jac_adj(x, k, adj, det);
        if (det < 0) negj++;
        V += fabs(det);
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                Gbar[a][i] += adj[i][0] * c_P[k][0][a] + adj[i][1] * c_P[k][1][a] + adj[i][2] * c_P[k][2][a];
    }
    if (negj) hk_atomic_add_u64(&d.counters[0], (unsigned long long)negj);
    const double invV = 1.0 / V;
    double trbar = 0.0;
#pragma unroll
    for (int a = 0; a < 8; ++a) trbar += du[a][0] * Gbar[a][0] + du[a][1] * Gbar[a][1] + du[a][2] * Gbar[a][2];
    trbar *= invV;

    // ---- pass B: Gauss points
    double f[8][3];
#pragma unroll
    for (int a = 0; a < 8; ++a) f[a][0] = f[a][1] = f[a][2] = 0.0;
    double pdet = 0.0, v_e = 0.0, t_e = 0.0;
    const double G = M.G;
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
        double adj[3][3], det;
        jac_adj(x, k, adj, det);
        double g[8][3];
        double L[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
        for (int a = 0; a < 8; ++a) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
                g[a][i] = adj[i][0] * c_P[k][0][a] + adj[i][1] * c_P[k][1][a] + adj[i][2] * c_P[k][2][a];
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) L[i][j] += du[a][i] * g[a][j];
        }
        const double idet = 1.0 / det;
        const double trL = (L[0][0] + L[1][1] + L[2][2]) * idet;
        const double vol = (trbar - trL) * (1.0 / 3.0);
        double de[6];
        de[0] = L[0][0] * idet + vol;
        de[1] = L[1][1] * idet + vol;
        de[2] = L[2][2] * idet + vol;
        de[3] = (L[0][1] + L[1][0]) * idet;
        de[4] = (L[1][2] + L[2][1]) * idet;
        de[5] = (L[0][2] + L[2][0]) * idet;

        const long long row = (long long)k * nEp + e;
        double s[6];
        // trial stress = old + D*de   (J2:1205-1220)
        s[0] = d.stress[0 * 8 * nEp + row] + (M.D11 * de[0] + M.D12 * de[1] + M.D12 * de[2]);
        s[1] = d.stress[1 * 8 * nEp + row] + (M.D12 * de[0] + M.D11 * de[1] + M.D12 * de[2]);
        s[2] = d.stress[2 * 8 * nEp + row] + (M.D12 * de[0] + M.D12 * de[1] + M.D11 * de[2]);
        s[3] = d.stress[3 * 8 * nEp + row] + M.D44 * de[3];
        s[4] = d.stress[4 * 8 * nEp + row] + M.D44 * de[4];
        s[5] = d.stress[5 * 8 * nEp + row] + M.D44 * de[5];
        double ep = d.eps[row];
        if (M.npp > 0) {                          // J2 radial return, J2:1227-1285
            const double mean = (s[0] + s[1] + s[2]) / 3.0;
            const double t0 = s[0] - mean, t1 = s[1] - mean, t2 = s[2] - mean;
            const double mises = sqrt(1.5 * (t0 * t0 + t1 * t1 + t2 * t2 + 2 * (s[3] * s[3]) + 2 * (s[4] * s[4]) +
                                             2 * (s[5] * s[5])));
            const double y = d.yield[row];
            if (mises > y) {
                int p_index = M.npp - 2;          // last segment extrapolates (J2:1261-1263)
                for (int j = 1; j < M.npp; ++j)
                    if (ep <= M.plastic_e[j]) { p_index = j - 1; break; }
                const double H = M.Hd[p_index];
                const double d_ep = (mises - y) / (3 * G + H);
                const double ynew = y + H * d_ep;
                const double fac = ynew / mises;
                s[0] = t0 * fac + mean;
                s[1] = t1 * fac + mean;
                s[2] = t2 * fac + mean;
                s[3] *= fac; s[4] *= fac; s[5] *= fac;
                ep += d_ep;
                d.eps[row] = ep;
                d.yield[row] = ynew;
            }
        }
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const long long idx = (long long)c * 8 * nEp + row;
            d.strain[idx] += de[c];
            d.stress[idx] = s[c];
        }
        // nodal force: s_dev * g_a (p-part is applied once after the loop)
        const double p = (s[0] + s[1] + s[2]) * (1.0 / 3.0);
        const double sx = s[0] - p, sy = s[1] - p, sz = s[2] - p;
#pragma unroll
        for (int a = 0; a < 8; ++a) {
            f[a][0] += sx * g[a][0] + s[3] * g[a][1] + s[5] * g[a][2];
            f[a][1] += s[3] * g[a][0] + sy * g[a][1] + s[4] * g[a][2];
            f[a][2] += s[5] * g[a][0] + s[4] * g[a][1] + sz * g[a][2];
        }
        pdet += p * det;
        const double tx = triax_of(s);
        if (A.write_triax) d.triax[row] = tx;
        v_e += ep;
        t_e += tx;
    }
    const double pbar = pdet * invV;
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c) d.Qe[(long long)(a * 3 + c) * nEp + e] = f[a][c] + pbar * Gbar[a][c];

    // ---- ductile damage / element deletion (J2:701-762)
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
                d.flag[e] = 0;
#pragma unroll 1
                for (int r = 0; r < 48; ++r) {
                    d.stress[(long long)r * nEp + e] = 0.0;
                    d.strain[(long long)r * nEp + e] = 0.0;
                }
                const int slot = hk_atomic_add_i32(d.del_count, 1);
                if (slot < d.del_cap) d.del_list[slot] = (A.step << 32) | e;
            }
        }
    }
}

#ifndef HK_EMU
__global__ void __launch_bounds__(128) hk_element_kernel(ElemArgs A) {
    long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (e < A.d.nElement) element_body(A, e);
}
#endif

void hk_launch_element(const HkDev& d, long long step, int write_triax, cudaStream_t s) {
    ElemArgs A{d, step, write_triax};
#ifndef HK_EMU
    const int block = 128;
    hk_element_kernel<<<(unsigned)((d.nElement + block - 1) / block), block, 0, s>>>(A);
#else
    for (long long e = 0; e < d.nElement; ++e) element_body(A, e);
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
        for (int c = 0; c < 6; ++c) sg[c] = d.stress[(long long)c * 8 * d.nEp + row];
        d.triax[row] = triax_of(sg);
    });
}

// elementVolume[e] = sum_k |detJ_k| at the current position (J2:1168-1169)
void hk_launch_element_volume(const HkDev& dd, double* V_out, cudaStream_t s) {
    const HkDev d = dd;
    hk_parallel_for(d.nElement, s, HK_LAMBDA(long long e) {
        double x[8][3];
        for (int a = 0; a < 8; ++a) {
            const long long n = d.conn[(long long)a * d.nEp + e];
            for (int c = 0; c < 3; ++c) x[a][c] = d.rec[6 * n + c];
        }
        double V = 0.0;
        for (int k = 0; k < 8; ++k) {
            double adj[3][3], det;
            jac_adj(x, k, adj, det);
            V += fabs(det);
        }
        V_out[e] = V;
    });
}
