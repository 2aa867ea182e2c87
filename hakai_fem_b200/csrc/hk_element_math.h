// hk_element_math.h — hex8 B-bar element math in trilinear "mode" form (per-thread, registers only).
//
// The reference evaluates, per Gauss point, a dense 6x24 Bfinal = B - BV + BVbar (J2:1415-1519,
// 1705-1784).  The hex8 shape functions are N_a = (1/8) prod_d (1 + delta_ad xi_d), so every nodal field w_a
// is the trilinear polynomial  w(xi) = c_ + sum_d c_d xi_d + c_01 xi0 xi1 + c_02 xi0 xi2 + c_12 xi1 xi2
// + c_012 xi0 xi1 xi2  with c_m = (1/8) sum_a chi_m(a) w_a (an 8-point Hadamard transform).  With the Gauss
// points at xi_d = g s_d (s_d = +-1, g = 1/sqrt 3) the rows of the Jacobian are
//     R_0 = c_0 + s1 h_01 + s2 h_02 + s1 s2 h_012      (h_01 = g c_01, h_012 = g^2 c_012, ...)
//     R_1 = c_1 + s0 h_01 + s2 h_12 + s0 s2 h_012
//     R_2 = c_2 + s0 h_02 + s1 h_12 + s0 s1 h_012
// (R_r[c] = J[r][c] = d x_c / d xi_r, identical to Pusai_k * e_position', J2:1424-1434), the same formula
// with the coefficients of d_disp gives D_r = d(du)/d xi_r, and with A_r = R_{r+1} x R_{r+2} (columns of the
// adjugate; inv(J) = adj/det as J2:1445-1455):
//     det   = R_0 . A_0
//     L det = sum_r D_r (x) A_r                         (velocity-gradient increment times det)
//     f_a   = sum_k sum_r T_r(k) P_k[r][a],  T_r = sigma' A_r     (nodal force, sigma' = s + pbar I)
// The force is the ADJOINT of the gradient operator, so it is accumulated in mode form
// M[r][.] += {1, s_a, s_b, s_a s_b} T_r and mapped to the nodes by one inverse Hadamard transform.
// The B-bar terms need sum_k adj_k P_k (BVbar) and V = sum_k det_k; both are exact 2x2x2 quadratures of
// polynomials and have the closed forms G[r][.] below (products of the h's), so no separate pass over the
// Gauss points is needed.  ~2600 FP64 operations per element instead of ~5300 for the g_a form and
// ~12000 for the dense 6x24 form; results agree with the dense form to rounding (tests state 1e-13).
//
// Difference from the reference kept on purpose: V = sum_k det_k, whereas the reference sums |det_k| and
// prints a warning (J2:1736-1739).  They coincide unless a Gauss point is inverted; inverted points are
// counted (hk_counters()[0]) exactly like the reference's warning.
#pragma once
#include "hk_common.h"

#define HK_G 0.57735026918962576451      /* 1/sqrt(3) */

// Reciprocal and reciprocal square root without the IEEE slow paths: hardware seed (MUFU.RCP64H / RSQ64H,
// ~20 good bits) + two Newton steps = full double precision to ~1 ulp, ~6 instructions and no branch, versus
// ~25 instructions plus a divergent special-case branch for `1.0/x` and `sqrt(x)`.  Arguments here are
// Jacobians, von Mises stresses and yield stresses: finite, normal, non-zero (zero is guarded by the caller).
HK_HD double hk_rcp(double a) {
#if defined(__CUDA_ARCH__)
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    x = fma(x, fma(-a, x, 1.0), x);
    x = fma(x, fma(-a, x, 1.0), x);
    return x;
#else
    return 1.0 / a;
#endif
}
HK_HD double hk_rsqrt(double a) {
#if defined(__CUDA_ARCH__)
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    x = fma(x * fma(-a * x, x, 1.0), 0.5, x);        // x += x*(1 - a x^2)/2
    x = fma(x * fma(-a * x, x, 1.0), 0.5, x);
    return x;
#else
    return 1.0 / sqrt(a);
#endif
}

struct HexModes {                 // coefficients of a nodal 3-vector field, pre-scaled
    double c0[3], c1[3], c2[3];   // c_d / 1   (already * 1/8)
    double h01[3], h02[3], h12[3];// g * c_dd'
    double h012[3];               // g^2 * c_012
};

// 8-point Hadamard transform of w[a][c] (node order = delta_mat, J2:1900-1907)
HK_HD void hex_modes(const double w[8][3], HexModes& m) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // nodes by sign pattern (d0,d1,d2): 0(---) 1(+--) 3(-+-) 2(++-) 4(--+) 5(+-+) 7(-++) 6(+++)
        const double s00 = w[1][c] + w[0][c], d00 = w[1][c] - w[0][c];     // (d1,d2) = (-,-)
        const double s10 = w[2][c] + w[3][c], d10 = w[2][c] - w[3][c];     // (+,-)
        const double s01 = w[5][c] + w[4][c], d01 = w[5][c] - w[4][c];     // (-,+)
        const double s11 = w[6][c] + w[7][c], d11 = w[6][c] - w[7][c];     // (+,+)
        const double ss0 = s10 + s00, sd0 = s10 - s00;                     // d2 = -
        const double ss1 = s11 + s01, sd1 = s11 - s01;                     // d2 = +
        const double ds0 = d10 + d00, dd0 = d10 - d00;
        const double ds1 = d11 + d01, dd1 = d11 - d01;
        m.c2[c] = (ss1 - ss0) * 0.125;
        m.c1[c] = (sd1 + sd0) * 0.125;
        m.h12[c] = (sd1 - sd0) * (0.125 * HK_G);
        m.c0[c] = (ds1 + ds0) * 0.125;
        m.h02[c] = (ds1 - ds0) * (0.125 * HK_G);
        m.h01[c] = (dd1 + dd0) * (0.125 * HK_G);
        m.h012[c] = (dd1 - dd0) * (0.125 * HK_G * HK_G);
    }
}

HK_HD void cross3(const double a[3], const double b[3], double o[3]) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
HK_HD void cross3_acc(const double a[3], const double b[3], double o[3]) {
    o[0] += a[1] * b[2] - a[2] * b[1];
    o[1] += a[2] * b[0] - a[0] * b[2];
    o[2] += a[0] * b[1] - a[1] * b[0];
}
HK_HD double dot3(const double a[3], const double b[3]) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// G[r][0..2] = sum_k {1, s_a, s_b} A_r(k)  (a < b the two directions other than r); sum_k s_a s_b A_r = 0.
HK_HD void adj_mode_sums(const HexModes& x, double G[3][3][3]) {
    cross3(x.c1, x.c2, G[0][0]);   cross3_acc(x.h01, x.h02, G[0][0]);
    cross3(x.c1, x.h12, G[0][1]);  cross3_acc(x.h01, x.h012, G[0][1]);     // s1
    cross3(x.h12, x.c2, G[0][2]);  cross3_acc(x.h012, x.h02, G[0][2]);     // s2
    cross3(x.c2, x.c0, G[1][0]);   cross3_acc(x.h12, x.h01, G[1][0]);
    cross3(x.h02, x.c0, G[1][1]);  cross3_acc(x.h012, x.h01, G[1][1]);     // s0
    cross3(x.c2, x.h02, G[1][2]);  cross3_acc(x.h12, x.h012, G[1][2]);     // s2
    cross3(x.c0, x.c1, G[2][0]);   cross3_acc(x.h02, x.h12, G[2][0]);
    cross3(x.c0, x.h01, G[2][1]);  cross3_acc(x.h02, x.h012, G[2][1]);     // s0
    cross3(x.h01, x.c1, G[2][2]);  cross3_acc(x.h012, x.h12, G[2][2]);     // s1
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int m = 0; m < 3; ++m)
#pragma unroll
            for (int c = 0; c < 3; ++c) G[r][m][c] *= 8.0;
}

// rows R_r (or D_r) at the Gauss point with signs s0,s1,s2 (+-1.0)
HK_HD void mode_rows(const HexModes& m, double s0, double s1, double s2, double R[3][3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        R[0][c] = (m.c0[c] + s1 * m.h01[c]) + s2 * (m.h02[c] + s1 * m.h012[c]);
        R[1][c] = (m.c1[c] + s0 * m.h01[c]) + s2 * (m.h12[c] + s0 * m.h012[c]);
        R[2][c] = (m.c2[c] + s0 * m.h02[c]) + s1 * (m.h12[c] + s0 * m.h012[c]);
    }
}

struct MatLite {                  // the per-Gauss-point scalars of a material, held in registers
    double D11, D12, D44, G3;     // Dmat entries (J2:143-159) and 3G
    int npp;
    int fast;                     // 1: MUFU-seeded rcp/rsqrt (default); 0: IEEE division / sqrt (A/B switch)
    const double* pe;             // plastic_e[] and Hd[] tables (shared memory in the TMA kernel), touched only
    const double* hd;             // while yielding
};

struct ElemAcc {                  // per-element accumulators over the Gauss points
    double M[3][4][3];            // force modes: M[r][{1,s_a,s_b,s_a s_b}][component]
    double pdet;                  // sum_k p_k det_k
    double v_e, t_e;              // sums of eps / triax over the Gauss points (J2:711-716)
    int negj;
};

HK_HD double triax_from(double mean, double oeq, double inv_oeq) {       // J2:1010-1017
    if (oeq < 1E-10) return 0.0;
    return mean * inv_oeq;
}

// ---- one Gauss point, in pieces (the TMEM kernel interleaves them with tensor-memory loads) -----------------
// geometry: adjugate columns A_r and det J at the Gauss point with signs (s0,s1,s2)
HK_HD void gp_geometry(const HexModes& X, double s0, double s1, double s2, double A[3][3], double& det) {
    double R[3][3];
    mode_rows(X, s0, s1, s2, R);
    cross3(R[1], R[2], A[0]);
    cross3(R[2], R[0], A[1]);
    cross3(R[0], R[1], A[2]);
    det = dot3(R[0], A[0]);
}

// strain increment (engineering shear) with the mean-dilatation correction
HK_HD void gp_strain(const HexModes& U, const double A[3][3], double idet, double trbar, double s0, double s1, double s2,
                     double de[6]) {
    double D[3][3];
    mode_rows(U, s0, s1, s2, D);
    // L[i][j] * det = sum_r D_r[i] A_r[j]
    const double l00 = D[0][0] * A[0][0] + D[1][0] * A[1][0] + D[2][0] * A[2][0];
    const double l11 = D[0][1] * A[0][1] + D[1][1] * A[1][1] + D[2][1] * A[2][1];
    const double l22 = D[0][2] * A[0][2] + D[1][2] * A[1][2] + D[2][2] * A[2][2];
    const double l01 = D[0][0] * A[0][1] + D[1][0] * A[1][1] + D[2][0] * A[2][1];
    const double l10 = D[0][1] * A[0][0] + D[1][1] * A[1][0] + D[2][1] * A[2][0];
    const double l12 = D[0][1] * A[0][2] + D[1][1] * A[1][2] + D[2][1] * A[2][2];
    const double l21 = D[0][2] * A[0][1] + D[1][2] * A[1][1] + D[2][2] * A[2][1];
    const double l02 = D[0][0] * A[0][2] + D[1][0] * A[1][2] + D[2][0] * A[2][2];
    const double l20 = D[0][2] * A[0][0] + D[1][2] * A[1][0] + D[2][2] * A[2][0];
    const double vol = (trbar - (l00 + l11 + l22) * idet) * (1.0 / 3.0);     // mean-dilatation correction
    de[0] = l00 * idet + vol;
    de[1] = l11 * idet + vol;
    de[2] = l22 * idet + vol;
    de[3] = (l01 + l10) * idet;
    de[4] = (l12 + l21) * idet;
    de[5] = (l02 + l20) * idet;
}

struct GpStress {                 // deviatoric stress t0,t1,t2 (normal) + shear, mean stress, eps, triaxiality
    double t0, t1, t2, s3, s4, s5, mean, ep, tx;
};

// elastic predictor, J2 radial return, state update.  st[0..5] stress, st[6..11] strain, st[12] eps, st[13] yield
HK_HD void gp_stress(const MatLite& Mt, double st[14], const double de[6], bool need_triax, GpStress& o) {
    // trial stress = old + D*de  (J2:1205-1220)
    double s[6];
    s[0] = st[0] + (Mt.D11 * de[0] + Mt.D12 * de[1] + Mt.D12 * de[2]);
    s[1] = st[1] + (Mt.D12 * de[0] + Mt.D11 * de[1] + Mt.D12 * de[2]);
    s[2] = st[2] + (Mt.D12 * de[0] + Mt.D12 * de[1] + Mt.D11 * de[2]);
    s[3] = st[3] + Mt.D44 * de[3];
    s[4] = st[4] + Mt.D44 * de[4];
    s[5] = st[5] + Mt.D44 * de[5];
    const double mean = (s[0] + s[1] + s[2]) * (1.0 / 3.0);
    double t0 = s[0] - mean, t1 = s[1] - mean, t2 = s[2] - mean;
    const double j2x = 1.5 * (t0 * t0 + t1 * t1 + t2 * t2 + 2.0 * (s[3] * s[3] + s[4] * s[4] + s[5] * s[5]));
    double inv_oeq, mises;
    if (Mt.fast) {
        inv_oeq = hk_rsqrt(fmax(j2x, 1e-300));              // 1/sqrt(3 J2); j2x == 0 (unstressed) -> mises = 0
        mises = j2x * inv_oeq;
        mises = fma(0.5 * inv_oeq, fma(-mises, mises, j2x), mises);     // one Newton step on the square root itself
    } else {
        mises = sqrt(j2x);
        inv_oeq = 1.0 / fmax(mises, 1e-300);
    }
    double oeq = mises;            // sqrt(3 J2) of the FINAL stress: the trial value, or the new yield stress
    double ep = st[12];
    if (Mt.npp > 0) {              // J2 radial return, J2:1227-1285
        const double y = st[13];
        if (mises > y) {
            // segment of the hardening table: first j with ep <= plastic[j,2] -> j-1, last segment extrapolates
            // (J2:1255-1264).  The table is increasing (checked in hk_add_material), so the index is a count,
            // which needs no dependent chain of loads.
            int p_index = 0;
            for (int j = 1; j + 1 < Mt.npp; ++j) p_index += (ep > Mt.pe[j]) ? 1 : 0;
            const double H = Mt.hd[p_index];
            const double d_ep = Mt.fast ? (mises - y) * hk_rcp(Mt.G3 + H) : (mises - y) / (Mt.G3 + H);
            const double ynew = y + H * d_ep;
            const double fac = ynew * inv_oeq;
            t0 *= fac; t1 *= fac; t2 *= fac;
            s[3] *= fac; s[4] *= fac; s[5] *= fac;
            s[0] = t0 + mean; s[1] = t1 + mean; s[2] = t2 + mean;
            ep += d_ep;
            st[12] = ep;
            st[13] = ynew;
            oeq = ynew;
            if (need_triax) inv_oeq = Mt.fast ? hk_rcp(ynew) : 1.0 / ynew;
        }
    }
#pragma unroll
    for (int c = 0; c < 6; ++c) {
        st[6 + c] += de[c];
        st[c] = s[c];
    }
    o.t0 = t0; o.t1 = t1; o.t2 = t2; o.s3 = s[3]; o.s4 = s[4]; o.s5 = s[5];
    o.mean = mean; o.ep = ep;
    // triaxiality feeds only the ductile-damage criterion and the output frames: skipped otherwise
    o.tx = need_triax ? triax_from(mean, oeq, inv_oeq) : 0.0;
}

// T_r = s_dev * A_r  (the mean part is added once per element as pbar * G)
HK_HD void gp_T(const GpStress& g, const double Ar[3], double T[3]) {
    T[0] = g.t0 * Ar[0] + g.s3 * Ar[1] + g.s5 * Ar[2];
    T[1] = g.s3 * Ar[0] + g.t1 * Ar[1] + g.s4 * Ar[2];
    T[2] = g.s5 * Ar[0] + g.s4 * Ar[1] + g.t2 * Ar[2];
}

// One Gauss point: strain increment, radial return, state update, force-mode accumulation.
HK_HD double gauss_point(const HexModes& X, const HexModes& U, const MatLite& Mt, int k, double trbar,
                         double st[14], ElemAcc& acc, bool need_triax = true) {
    const double s0 = (k & 4) ? 1.0 : -1.0, s1 = (k & 2) ? 1.0 : -1.0, s2 = (k & 1) ? 1.0 : -1.0;
    double A[3][3], det;
    gp_geometry(X, s0, s1, s2, A, det);
    if (det < 0) acc.negj++;
    const double idet = Mt.fast ? hk_rcp(det) : 1.0 / det;
    double de[6];
    gp_strain(U, A, idet, trbar, s0, s1, s2, de);
    GpStress g;
    gp_stress(Mt, st, de, need_triax, g);
    acc.pdet += g.mean * det;
    const double sa[3] = {s1, s0, s0};        // sign of the lower / higher "other" direction for r = 0,1,2
    const double sb[3] = {s2, s2, s1};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double T[3];
        gp_T(g, A[r], T);
        const double sab = sa[r] * sb[r];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            acc.M[r][0][c] += T[c];
            acc.M[r][1][c] += sa[r] * T[c];
            acc.M[r][2][c] += sb[r] * T[c];
            acc.M[r][3][c] += sab * T[c];
        }
    }
    acc.v_e += g.ep;
    acc.t_e += g.tx;
    return g.tx;
}

// nodal forces from the accumulated modes:  f = adjoint( M + pbar * G )
HK_HD void element_forces(const ElemAcc& acc, const double G[3][3][3], double pbar, double f[8][3]) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const double m00 = acc.M[0][0][c] + pbar * G[0][0][c], m01 = acc.M[0][1][c] + pbar * G[0][1][c];
        const double m02 = acc.M[0][2][c] + pbar * G[0][2][c], m03 = acc.M[0][3][c];
        const double m10 = acc.M[1][0][c] + pbar * G[1][0][c], m11 = acc.M[1][1][c] + pbar * G[1][1][c];
        const double m12 = acc.M[1][2][c] + pbar * G[1][2][c], m13 = acc.M[1][3][c];
        const double m20 = acc.M[2][0][c] + pbar * G[2][0][c], m21 = acc.M[2][1][c] + pbar * G[2][1][c];
        const double m22 = acc.M[2][2][c] + pbar * G[2][2][c], m23 = acc.M[2][3][c];
        // force modes (f_ = 0: the element's nodal forces sum to zero identically)
        const double f0 = m00 * 0.125, f1 = m10 * 0.125, f2 = m20 * 0.125;
        const double f01 = (m01 + m11) * (0.125 * HK_G);       // r=0: s_a = s1 ; r=1: s_a = s0
        const double f02 = (m02 + m21) * (0.125 * HK_G);       // r=0: s_b = s2 ; r=2: s_a = s0
        const double f12 = (m12 + m22) * (0.125 * HK_G);       // r=1: s_b = s2 ; r=2: s_b = s1
        const double f012 = (m03 + m13 + m23) * (0.125 * HK_G * HK_G);
        // inverse Hadamard f_a = sum_m chi_m(a) fhat_m with node signs (d0,d1,d2):
        //   f_a = base(d1,d2) + d0 * slope(d1,d2)
        // group by (d1,d2): base(d1,d2) = f1 d1 + f2 d2 + f12 d1 d2 ; slope(d1,d2) = f0 + f01 d1 + f02 d2 + f012 d1 d2
        const double b_mm = -f1 - f2 + f12, s_mm = f0 - f01 - f02 + f012;
        const double b_pm = f1 - f2 - f12, s_pm = f0 + f01 - f02 - f012;
        const double b_mp = -f1 + f2 - f12, s_mp = f0 - f01 + f02 - f012;
        const double b_pp = f1 + f2 + f12, s_pp = f0 + f01 + f02 + f012;
        f[0][c] = b_mm - s_mm;   // (-,-,-)
        f[1][c] = b_mm + s_mm;   // (+,-,-)
        f[2][c] = b_pm + s_pm;   // (+,+,-)
        f[3][c] = b_pm - s_pm;   // (-,+,-)
        f[4][c] = b_mp - s_mp;   // (-,-,+)
        f[5][c] = b_mp + s_mp;   // (+,-,+)
        f[6][c] = b_pp + s_pp;   // (+,+,+)
        f[7][c] = b_pp - s_pp;   // (-,+,+)
    }
}
