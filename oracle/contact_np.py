"""Second, independent restatement of the reference's contact pass — TEST INFRASTRUCTURE ONLY (see
hakai_oracle.cpp's header for the rules; nothing under hakai_fem_b200/ may import this).

cal_contact_force of HAKAI-v0.0.2 (Julia/HAKAI_j.jl:2248-2706) written straight from the Julia text as plain
Python loops over (master triangle, slave node), with none of the C++ oracle's code shared: it exists so that the
C++ oracle's contact forces (and through it the CUDA kernels') are checked against a second reading of the reference
(tests/test_oracle_anchors.py).  Parity is still unpinned — no Julia here — but a transcription slip would have to
be made twice, in two languages, to go unnoticed.

Sums: the reference adds every contribution into a Float128 column per thread and adds the columns (J2:435,
497-548); here each dof collects its contributions and math.fsum adds them exactly, so the result is the
correctly rounded sum whatever the order.
"""
import math

import numpy as np


def _norm3(a, b, c):                                            # my3norm, J2:3167
    return math.sqrt(a * a + b * b + c * c)


def _solve3(A11, A21, A31, A12, A22, A32, A13, A23, A33, bx, by, bz):       # my3SolveAb, J2:3342-3405
    v = (A11 * A22 * A33 + A12 * A23 * A31 + A13 * A21 * A32 - A11 * A23 * A32 - A12 * A21 * A33 - A13 * A22 * A31)
    im11 = A22 * A33 - A23 * A32
    im21 = A23 * A31 - A21 * A33
    im31 = A21 * A32 - A22 * A31
    im12 = A13 * A32 - A12 * A33
    im22 = A11 * A33 - A13 * A31
    im32 = A12 * A31 - A11 * A32
    im13 = A12 * A23 - A13 * A22
    im23 = A13 * A21 - A11 * A23
    im33 = A11 * A22 - A12 * A21
    x1 = (im11 * bx + im12 * by + im13 * bz) / v
    x2 = (im21 * bx + im22 * by + im23 * bz) / v
    x3 = (im31 * bx + im32 * by + im33 * bz) / v
    return x1, x2, x3


def contact_force(pairs, position, velo, diag_M, element_min_size, element_max_size, element_flag, elementmat,
                  myu=0.25, d_lim_factor=0.3):
    """pairs: list of dicts {i_instance, j_instance, c_nodes_i, c_nodes_j, c_triangles (nTri,3), c_triangles_eleid,
    young} with 1-based ids; position (3,nNode); velo, diag_M (fn); elementmat (8,nE) 1-based.
    Returns (c_force (fn), n_hits)."""
    fn = position.shape[1] * 3
    terms = [[] for _ in range(fn)]
    hits = 0
    d_lim = element_min_size * d_lim_factor                     # J2:2254
    kc_o = kc_s = 1.0
    Cr_o = Cr_s = 0.0
    P = position
    for ct in pairs:
        same = ct["i_instance"] == ct["j_instance"]
        ni = np.asarray(ct["c_nodes_i"]) - 1
        nj = np.asarray(ct["c_nodes_j"]) - 1
        tri = np.asarray(ct["c_triangles"]) - 1
        tele = np.asarray(ct["c_triangles_eleid"]) - 1
        young = ct["young"]
        mn_i, mx_i = P[:, ni].min(axis=1), P[:, ni].max(axis=1)  # J2:2284-2298
        mn_j, mx_j = P[:, nj].min(axis=1), P[:, nj].max(axis=1)
        rmin, rmax = np.maximum(mn_i, mn_j), np.minimum(mx_i, mx_j)
        if np.any(rmin > rmax):                                 # J2:2307-2309
            continue
        amin = np.minimum(mn_i, mn_j)
        ddiv = element_max_size * (0.6 if same else 1.1)        # J2:2331-2334
        map_i = np.ceil((P[:, ni] - amin[:, None]) / ddiv).astype(np.int64)      # J2:2337-2349
        map_j = np.ceil((P[:, nj] - amin[:, None]) / ddiv).astype(np.int64)
        first_j = {}
        for idx, n in enumerate(nj):                            # first match wins, J2:2462-2472
            first_j.setdefault(int(n), idx)
        kc, Cr = (kc_s, Cr_s) if same else (kc_o, Cr_o)
        for t in range(tri.shape[0]):
            el = tele[t]
            if element_flag[el] == 0:                           # J2:2373-2376
                continue
            j0, j1, j2 = (int(v) for v in tri[t])
            q0, q1, q2 = P[:, j0], P[:, j1], P[:, j2]
            skip = False
            for c in range(3):                                  # J2:2400-2419
                if (q0[c] < rmin[c] and q1[c] < rmin[c] and q2[c] < rmin[c]) or \
                   (q0[c] > rmax[c] and q1[c] > rmax[c] and q2[c] > rmax[c]):
                    skip = True
            if skip:
                continue
            cx = (q0[0] + q1[0] + q2[0]) / 3.0
            cy = (q0[1] + q1[1] + q2[1]) / 3.0
            cz = (q0[2] + q1[2] + q2[2]) / 3.0
            Rmax = max(max(_norm3(q0[0] - cx, q0[1] - cy, q0[2] - cz), _norm3(q1[0] - cx, q1[1] - cy, q1[2] - cz)),
                       _norm3(q2[0] - cx, q2[1] - cy, q2[2] - cz))
            v1 = q1 - q0
            v2 = q2 - q0
            L1, L2 = _norm3(*v1), _norm3(*v2)
            Lmax = max(L1, L2)
            n1 = v1[1] * v2[2] - v1[2] * v2[1]                  # my3crossNNz, J2:3209-3231
            n2 = v1[2] * v2[0] - v1[0] * v2[2]
            n3 = v1[0] * v2[1] - v1[1] * v2[0]
            mag = math.sqrt(n1 * n1 + n2 * n2 + n3 * n3)
            nx, ny, nz = n1 / mag, n2 / mag, n3 / mag
            d12 = v1[0] * v2[0] + v1[1] * v2[1] + v1[2] * v2[2]
            S = 0.5 * math.sqrt(L1 * L1 * L2 * L2 - d12 * d12)  # J2:2450 (no max(.,0) guard)
            if j0 in first_j:
                mj = map_j[:, first_j[j0]]
            else:
                mj = np.array([1, 1, 1])                        # J2:2458-2460 defaults
            own = set(int(v) - 1 for v in elementmat[:, el])
            near = np.flatnonzero(np.all(np.abs(mj[:, None] - map_i) <= 1, axis=0))     # J2:2484-2489
            for k in near:
                i = int(ni[k])
                if same and i in own:                           # J2:2496-2507
                    continue
                px, py, pz = P[0, i], P[1, i], P[2, i]
                if px < rmin[0] or py < rmin[1] or pz < rmin[2]:
                    continue
                if px > rmax[0] or py > rmax[1] or pz > rmax[2]:
                    continue
                if _norm3(px - cx, py - cy, pz - cz) >= Rmax:   # J2:2521-2524
                    continue
                x1, x2, d = _solve3(v1[0], v1[1], v1[2], v2[0], v2[1], v2[2], -nx, -ny, -nz,
                                    px - q0[0], py - q0[1], pz - q0[2])
                if not (0.0 <= x1 and 0.0 <= x2 and x1 + x2 <= 1.0 and d > 0.0 and d <= d_lim):
                    continue
                hits += 1
                vx = velo[3 * i] - velo[3 * j0]
                vy = velo[3 * i + 1] - velo[3 * j0 + 1]
                vz = velo[3 * i + 2] - velo[3 * j0 + 2]
                mag_v = _norm3(vx, vy, vz)
                vex = vey = vez = 0.0
                if mag_v > 0.0:
                    vex, vey, vez = vx / mag_v, vy / mag_v, vz / mag_v
                kk = young * S / Lmax * kc                      # J2:2572
                F = kk * d
                fx, fy, fz = F * nx, F * ny, F * nz
                C = 2 * math.sqrt(diag_M[i] * kk) * Cr          # J2:2583 indexes the DOF vector by the node id (sic); Cr = 0
                dot = vex * nx + vey * ny + vez * nz
                vsx, vsy, vsz = vex - dot * nx, vey - dot * ny, vez - dot * nz
                fx += -myu * F * vsx + -C * vx
                fy += -myu * F * vsy + -C * vy
                fz += -myu * F * vsz + -C * vz
                for c, f in enumerate((fx, fy, fz)):
                    terms[3 * i + c].append(f)
                    for j in (j0, j1, j2):
                        terms[3 * j + c].append(-f / 3.0)
    return np.array([math.fsum(t) for t in terms]), hits
