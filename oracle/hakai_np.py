"""Second, INDEPENDENT CPU restatement of the hot path — NumPy, following the v0.0.0 matrix form.

TEST INFRASTRUCTURE ONLY (same status as hakai_oracle.cpp).  Where the C++ oracle follows
HAKAI-v0.0.2's scalar code line by line, this file follows HAKAI-v0.0.0/Julia/HAKAI_j.jl (cited
J0:<line>): dense 6x24 `B + BVbar - BV` (J0:489-641), `inv(J)`/`det(J)` from LAPACK (J0:649-692),
principal stresses from a symmetric eigen-solver (J0:451-486, here numpy.linalg.eigvalsh), the step
order of J0:254-426.  It shares no code with the C++ oracle or the CUDA engine; agreement of the
three (tests/test_oracle_np.py) is the anchor that replaces the golden vectors the reference lacks.
Contact is not restated here (v0.0.0's contact is a different, frictionless algorithm).
"""
from __future__ import annotations

import numpy as np


def pusai():
    """cal_Pusai_hexa, J0:695-741 -> P[k, dir, node]."""
    delta = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                      [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    g = 1.0 / np.sqrt(3.0)
    gc = np.array([[-g, -g, -g], [-g, -g, g], [-g, g, -g], [-g, g, g], [g, -g, -g], [g, -g, g], [g, g, -g], [g, g, g]])
    P = np.zeros((8, 3, 8))
    for k in range(8):
        for i in range(8):
            P[k, 0, i] = 0.125 * delta[i, 0] * (1 + gc[k, 1] * delta[i, 1]) * (1 + gc[k, 2] * delta[i, 2])
            P[k, 1, i] = 0.125 * delta[i, 1] * (1 + gc[k, 0] * delta[i, 0]) * (1 + gc[k, 2] * delta[i, 2])
            P[k, 2, i] = 0.125 * delta[i, 2] * (1 + gc[k, 0] * delta[i, 0]) * (1 + gc[k, 1] * delta[i, 1])
    return P


def dmat(young, poisson):
    d1, d2, d3 = 1.0 - poisson, poisson, (1.0 - 2.0 * poisson) / 2.0
    D = np.zeros((6, 6))
    D[:3, :3] = d2
    D[0, 0] = D[1, 1] = D[2, 2] = d1
    D[3, 3] = D[4, 4] = D[5, 5] = d3
    return young / (1.0 + poisson) / (1.0 - 2.0 * poisson) * D


class NumpyHakai:
    """State + step of the v0.0.0 solver loop (J0:254-426), vectorised over elements."""

    def __init__(self, setup):
        m = setup.model
        self.m = m
        self.dt = setup.d_time
        self.nN, self.nE = m.nNode, m.nElement
        self.X = np.ascontiguousarray(m.coordmat.T)              # (nN,3) row-major like v0.0.0
        self.em = np.ascontiguousarray(m.elementmat.T) - 1       # (nE,8)
        self.M = setup.diag_M.copy()
        fn = 3 * self.nN
        self.disp = np.zeros(fn)
        self.disp_pre = np.zeros(fn)
        self.velo = np.zeros(fn)
        self.d_disp = np.zeros(fn)
        for ic in m.IC:
            for dof, v in zip(ic.dof, ic.value):
                self.disp_pre[dof - 1] = -v * self.dt
                self.velo[dof - 1] = v
        self.Q = np.zeros(fn)
        nip = 8 * self.nE
        self.stress = np.zeros((nip, 6))
        self.strain = np.zeros((nip, 6))
        self.eps = np.zeros(nip)
        self.triax = np.zeros(nip)
        self.yld = np.zeros(nip)
        self.flag = np.ones(self.nE, np.int64)
        self.P = pusai()
        self.mat = m.element_material - 1
        for i, mt in enumerate(m.MATERIAL):
            if mt.plastic.shape[0]:
                sel = np.repeat(self.mat == i, 8)
                self.yld[sel] = mt.plastic[0, 0]
        self.deleted = []

    # -- one step ------------------------------------------------------------------------------
    def step(self, t):
        dt = self.dt
        m = self.m
        M = self.M
        disp_new = 1.0 / (M / dt ** 2) * (0.0 - self.Q + M / dt ** 2 * (2.0 * self.disp - self.disp_pre))   # J0:274
        for bc in m.BC:                                                                                    # J0:280-308
            amp = 1.0
            if len(bc.amp_name) > 0:
                a_t, a_v = bc.amplitude.time, bc.amplitude.value
                ct = t * dt
                ti = 0
                for j in range(len(a_t) - 1):
                    if a_t[j] <= ct <= a_t[j + 1]:
                        ti = j
                        break
                amp = a_v[ti] + (a_v[ti + 1] - a_v[ti]) * (ct - a_t[ti]) / (a_t[ti + 1] - a_t[ti])
            for dof, v in zip(bc.dof, bc.value):
                disp_new[dof - 1] = v * amp
        self.d_disp = disp_new - self.disp                                                                  # J0:311-316
        self.disp_pre = self.disp
        self.disp = disp_new
        self.velo = self.d_disp / dt
        pos = self.X + self.disp.reshape(-1, 3)
        self._stress(pos)
        self._triax()
        self._fracture()

    def _stress(self, pos):
        """cal_stress_hexa, J0:489-647 (matrix form)."""
        live = np.flatnonzero(self.flag == 1)
        em = self.em[live]
        ep = pos[em]                                             # (n,8,3)
        du = self.d_disp.reshape(-1, 3)[em].reshape(len(live), 24)
        P = self.P
        J = np.einsum("kdi,nic->nkdc", P, ep)                    # J = Pusai1 * e_position      J0:673
        detJ = np.linalg.det(J)                                  # (n,8)                       J0:559
        V = detJ.sum(axis=1)
        P2 = np.einsum("nkab,kbi->nkai", np.linalg.inv(J), P)    # inv(J)*Pusai1               J0:674
        n = len(live)
        B = np.zeros((n, 8, 6, 24))
        for i in range(8):                                       # cal_B_hexa                  J0:677-689
            B[:, :, 0, 3 * i + 0] = P2[:, :, 0, i]
            B[:, :, 1, 3 * i + 1] = P2[:, :, 1, i]
            B[:, :, 2, 3 * i + 2] = P2[:, :, 2, i]
            B[:, :, 3, 3 * i + 0] = P2[:, :, 1, i]
            B[:, :, 3, 3 * i + 1] = P2[:, :, 0, i]
            B[:, :, 4, 3 * i + 1] = P2[:, :, 2, i]
            B[:, :, 4, 3 * i + 2] = P2[:, :, 1, i]
            B[:, :, 5, 3 * i + 0] = P2[:, :, 2, i]
            B[:, :, 5, 3 * i + 2] = P2[:, :, 0, i]
        BV = np.zeros((n, 8, 6, 24))                             # cal_BVbar                   J0:649-668
        N = P2.transpose(0, 1, 3, 2).reshape(n, 8, 24)           # reshape(P2,1,24): column-major = node-major
        BV[:, :, 0, :] = N / 3.0
        BV[:, :, 1, :] = N / 3.0
        BV[:, :, 2, :] = N / 3.0
        BVbar = (BV * detJ[:, :, None, None]).sum(axis=1) / V[:, None, None]
        Bf = B + BVbar[:, None, :, :] - BV                        # J0:568
        d_e = np.einsum("nkrc,nc->nkr", Bf, du)
        Q = np.zeros(3 * self.nN)
        qe = np.zeros((n, 24))
        for mi, mt in enumerate(self.m.MATERIAL):
            sel = np.flatnonzero(self.mat[live] == mi)
            if len(sel) == 0:
                continue
            D = dmat(mt.young, mt.poisson)
            G = mt.young / 2.0 / (1.0 + mt.poisson)
            ipidx = (live[sel][:, None] * 8 + np.arange(8)[None, :])          # (ns,8)
            pre = self.stress[ipidx]                                          # (ns,8,6)
            d_o = np.einsum("rc,nkc->nkr", D, d_e[sel])
            tri = pre + d_o
            final = tri.copy()
            if mt.plastic.shape[0] > 0:
                mean = tri[..., :3].sum(axis=-1) / 3
                dev = tri.copy()
                dev[..., :3] -= mean[..., None]
                mises = np.sqrt(3 / 2 * (dev[..., 0] ** 2 + dev[..., 1] ** 2 + dev[..., 2] ** 2 +
                                         2 * dev[..., 3] ** 2 + 2 * dev[..., 4] ** 2 + 2 * dev[..., 5] ** 2))
                y = self.yld[ipidx]
                e_p = self.eps[ipidx]
                yielding = mises > y
                pl = mt.plastic
                # p_index: first j (2..npp) with eps <= plastic[j,2] -> j-1, else npp-1          J0:592-601
                seg = np.searchsorted(pl[1:, 1], e_p, side="left")
                seg = np.minimum(seg, pl.shape[0] - 2)
                H = (pl[seg + 1, 0] - pl[seg, 0]) / (pl[seg + 1, 1] - pl[seg, 1])
                with np.errstate(divide="ignore", invalid="ignore"):
                    d_ep = (mises - y) / (3 * G + H)
                    fdev = dev * ((y + H * d_ep) / mises)[..., None]
                fin_pl = fdev.copy()
                fin_pl[..., :3] += mean[..., None]
                final = np.where(yielding[..., None], fin_pl, tri)
                self.eps[ipidx] = np.where(yielding, e_p + d_ep, e_p)
                self.yld[ipidx] = np.where(yielding, y + H * d_ep, y)
            self.stress[ipidx] = final
            self.strain[ipidx] += d_e[sel]
            q_i = np.einsum("nkrc,nkr->nkc", Bf[sel], final)                  # B' * o_vec        J0:631
            qe[sel] = (detJ[sel][:, :, None] * q_i).sum(axis=1)
        for i in range(8):                                                    # J0:637-641
            for c in range(3):
                np.add.at(Q, 3 * em[:, i] + c, qe[:, 3 * i + c])
        self.Q = Q
        self.qe_live = (live, qe)

    def _triax(self):
        """cal_triax_stress, J0:451-486, eigenvalue route (LAPACK syevd here)."""
        s = self.stress
        T = np.zeros((s.shape[0], 3, 3))
        T[:, 0, 0], T[:, 1, 1], T[:, 2, 2] = s[:, 0], s[:, 1], s[:, 2]
        T[:, 0, 1] = T[:, 1, 0] = s[:, 3]
        T[:, 1, 2] = T[:, 2, 1] = s[:, 4]
        T[:, 0, 2] = T[:, 2, 0] = s[:, 5]
        p = np.linalg.eigvalsh(T)
        oeq = np.sqrt(0.5 * ((p[:, 0] - p[:, 1]) ** 2 + (p[:, 1] - p[:, 2]) ** 2 + (p[:, 2] - p[:, 0]) ** 2))
        with np.errstate(divide="ignore", invalid="ignore"):
            v = (p[:, 0] + p[:, 1] + p[:, 2]) / 3 / oeq
        self.triax = np.where(oeq < 1e-10, 0.0, v)

    def _fracture(self):
        """J0:332-385."""
        for mi, mt in enumerate(self.m.MATERIAL):
            nd = mt.ductile.shape[0]
            if nd == 0:
                continue
            els = np.flatnonzero(self.mat == mi)
            v_e = self.eps.reshape(-1, 8)[els].sum(axis=1) / 8
            t_e = self.triax.reshape(-1, 8)[els].sum(axis=1) / 8
            d = mt.ductile
            fr = np.full(len(els), d[nd - 1, 0])
            done = np.zeros(len(els), bool)
            for j in range(nd - 1):
                hit = (~done) & (t_e >= d[j, 1]) & (t_e < d[j + 1, 1])
                fr[hit] = d[j, 0] + (d[j + 1, 0] - d[j, 0]) / (d[j + 1, 1] - d[j, 1]) * (t_e[hit] - d[j, 1])
                done |= hit
            kill = (t_e >= 0) & (v_e >= fr) & (self.flag[els] == 1)
            for e in els[kill]:
                self.flag[e] = 0
                self.deleted.append(int(e) + 1)
                self.stress[8 * e:8 * e + 8] = 0.0
                self.strain[8 * e:8 * e + 8] = 0.0
