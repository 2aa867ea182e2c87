"""Synthetic structured hex8 decks: B1 bar, F16 block, W slabs, I8-style impact.

Every deck can be produced two ways that yield identical arrays (tests/test_mesh.py checks it):
`write_inp()` emits a real Abaqus `.inp` that the reference's reader accepts, and `build_model()`
fills the `Model` arrays directly (needed at 16 M elements, where the text deck would be GBs).

Node numbering: x fastest, then y, z slowest, so every z-layer is a contiguous id range
(`*Nset, generate`).  Element node order is C3D8 / delta_mat (HAKAI_j.jl:1900-1907).
Materials are the blocks of HAKAI-v0.0.0/input/Tensile5e.inp (mm-t-s units).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from .inp import Model, Part, Instance, Material, Amplitude, BC, IC

STEEL_PLASTIC = np.array([[755., 0.], [809., 0.01], [829., 0.02], [842., 0.1], [895., 0.15],
                          [922., 0.4], [953., 1.], [1100., 4.]])


def steel(name: str = "steel_Elastoplast", ductile: Optional[np.ndarray] = None, plastic=True) -> Material:
    m = Material(name=name, density=7.8e-09, young=210000., poisson=0.3)
    if plastic:
        m.plastic = STEEL_PLASTIC.copy()
        m.Hd = (m.plastic[1:, 0] - m.plastic[:-1, 0]) / (m.plastic[1:, 1] - m.plastic[:-1, 1])
    if ductile is not None:
        m.ductile = np.asarray(ductile, dtype=np.float64).reshape(-1, 3)
        m.fracture_flag = 1
    return m


def block_arrays(nx: int, ny: int, nz: int, h: float = 1.0, origin=(0.0, 0.0, 0.0)):
    """coordmat (3,nN) and 1-based elementmat (8,nE) of an nx x ny x nz brick of cubes of edge h."""
    nnx, nny, nnz = nx + 1, ny + 1, nz + 1
    x = origin[0] + h * np.arange(nnx, dtype=np.float64)
    y = origin[1] + h * np.arange(nny, dtype=np.float64)
    z = origin[2] + h * np.arange(nnz, dtype=np.float64)
    coord = np.empty((3, nnx * nny * nnz))
    coord[0] = np.tile(x, nny * nnz)
    coord[1] = np.tile(np.repeat(y, nnx), nnz)
    coord[2] = np.repeat(z, nnx * nny)
    i = np.arange(nx, dtype=np.int64)
    j = np.arange(ny, dtype=np.int64)
    k = np.arange(nz, dtype=np.int64)
    n0 = (i[None, None, :] + nnx * (j[None, :, None] + nny * k[:, None, None])).reshape(-1) + 1   # x fastest
    sx, sy, sz = 1, nnx, nnx * nny
    em = np.stack([n0, n0 + sx, n0 + sx + sy, n0 + sy,
                   n0 + sz, n0 + sx + sz, n0 + sx + sy + sz, n0 + sy + sz]).astype(np.int64)
    return coord, em


def _write_nodes_elements(f, coord: np.ndarray, em: np.ndarray):
    """`*Node` / `*Element` blocks; %.17g round-trips every double exactly."""
    nN, nE = coord.shape[1], em.shape[1]
    f.write("*Node\n")
    ids = np.arange(1, nN + 1, dtype=np.float64)[:, None]
    np.savetxt(f, np.concatenate([ids, coord.T], axis=1), fmt=["%7d", "%.17g", "%.17g", "%.17g"], delimiter=", ")
    f.write("*Element, type=C3D8R\n")
    np.savetxt(f, np.concatenate([np.arange(1, nE + 1, dtype=np.int64)[:, None], em.T], axis=1), fmt="%d", delimiter=", ")


@dataclass
class StretchDeck:
    """Uniform-stretch brick: z-layer initial velocities v = rate*z and amplitude-driven end layers."""
    nx: int
    ny: int
    nz: int
    h: float = 1.0
    material: Material = field(default_factory=steel)
    d_time: float = 8.0e-08
    n_steps: float = 99.5                 # end_time = n_steps*d_time (half-step margin)
    strain_per_step: float = 2.0e-4       # rate*d_time
    jitter: float = 0.0                   # interior node jitter, fraction of h (F16: 0.05)
    seed: int = 20240601
    name: str = "stretch"
    layer_offset: int = 0                 # multi-GPU slabs: this block starts at global element layer `layer_offset`
    global_nz: Optional[int] = None       # ... of a global mesh with `global_nz` element layers (None: nz)
    jitter_by_layer: bool = False         # noise seeded per GLOBAL node layer ([seed, layer]) instead of one stream over
                                          # the whole array: a z-slab of a larger mesh then carries exactly the nodes the
                                          # single-domain mesh has (the weak-scaling decks and their parity twin)

    def _layers(self):
        per = (self.nx + 1) * (self.ny + 1)
        return per, self.nz + 1

    def coord_elem(self):
        coord, em = block_arrays(self.nx, self.ny, self.nz, self.h, origin=(0.0, 0.0, self.layer_offset * self.h))
        if self.jitter > 0:
            nnx, nny, nnz = self.nx + 1, self.ny + 1, self.nz + 1
            gnz = self.global_nz if self.global_nz is not None else self.nz
            ii = np.tile(np.arange(nnx), nny * nnz)
            jj = np.tile(np.repeat(np.arange(nny), nnx), nnz)
            kk = np.repeat(np.arange(nnz), nnx * nny) + self.layer_offset        # GLOBAL node layer
            interior = (ii > 0) & (ii < nnx - 1) & (jj > 0) & (jj < nny - 1) & (kk > 0) & (kk < gnz)
            amp = self.jitter * self.h
            if self.jitter_by_layer:
                per = nnx * nny
                noise = np.empty(coord.shape)
                for k in range(nnz):
                    rng = np.random.default_rng([self.seed, k + self.layer_offset])
                    noise[:, k * per:(k + 1) * per] = rng.uniform(-amp, amp, size=(3, per))
            else:
                noise = np.random.default_rng(self.seed).uniform(-amp, amp, size=coord.shape)
            coord = coord + noise * interior[None, :]
        return coord, em

    def build_model(self) -> Model:
        coord, em = self.coord_elem()
        nN, nE = coord.shape[1], em.shape[1]
        per, nl = self._layers()
        rate = self.strain_per_step / self.d_time
        end_time = self.n_steps * self.d_time
        part = Part(name="Part-1", nNode=nN, coordmat=coord, nElement=nE, elementmat=em,
                    material_name=self.material.name, material_id=1)
        inst = Instance(name="Part-1-1", part_name="Part-1", part_id=1, material_id=1, node_offset=0, nNode=nN,
                        element_offset=0, nElement=nE, elements=np.arange(1, nE + 1))
        amp = Amplitude(name="Amp-1", time=np.array([0.0, end_time]), value=np.array([0.0, 1.0]))
        gnz = self.global_nz if self.global_nz is not None else self.nz
        lz = gnz * self.h
        layer = lambda k: np.arange(k * per + 1, (k + 1) * per + 1, dtype=np.int64)
        bc = BC(Nset_name="L%d" % (nl - 1), amp_name="Amp-1", amplitude=amp)
        if self.layer_offset == 0:                       # global bottom layer: u_z = 0
            bc.dof.append(layer(0) * 3)
            bc.value.append(0.0)
        if self.layer_offset + self.nz == gnz:           # global top layer: driven
            bc.dof.append(layer(nl - 1) * 3)
            bc.value.append(rate * lz * end_time)
        ic = IC(Nset_name="L%d" % (nl - 1), type="VELOCITY")
        for k in range(nl):
            kg = k + self.layer_offset
            if kg == 0:
                continue
            ic.dof.append(layer(k) * 3)
            ic.value.append(rate * (kg * self.h))
        return Model(PART=[part], INSTANCE=[inst], NSET=[], ELSET=[], SURFACE=[], AMPLITUDE=[amp],
                     MATERIAL=[self.material], BC=[bc], IC=[ic], CP=[], nNode=nN, coordmat=coord, nElement=nE,
                     elementmat=em, element_material=np.ones(nE, np.int64), element_instance=np.ones(nE, np.int64),
                     d_time=self.d_time, end_time=end_time, mass_scaling=1.0, contact_flag=0)

    def write_inp(self, path: str):
        coord, em = self.coord_elem()
        nN, nE = coord.shape[1], em.shape[1]
        per, nl = self._layers()
        rate = self.strain_per_step / self.d_time
        end_time = self.n_steps * self.d_time
        lz = self.nz * self.h
        m = self.material
        with open(path, "w") as f:
            w = f.write
            w("*Heading\n** synthetic stretch deck %s %dx%dx%d\n" % (self.name, self.nx, self.ny, self.nz))
            w("*Part, name=Part-1\n")
            _write_nodes_elements(f, coord, em)
            w("*Nset, nset=Set-all, generate\n  1, %d, 1\n" % nN)
            w("*Elset, elset=Set-all, generate\n 1, %d, 1\n" % nE)
            w("*Solid Section, elset=Set-all, material=%s\n,\n*End Part\n" % m.name)
            w("**\n*Assembly, name=Assembly\n**\n*Instance, name=Part-1-1, part=Part-1\n*End Instance\n**\n")
            for k in range(nl):
                w("*Nset, nset=L%d, instance=Part-1-1, generate\n %d, %d, 1\n" % (k, k * per + 1, (k + 1) * per))
            w("*End Assembly\n")
            w("*Amplitude, name=Amp-1\n %s, 0., %s, 1.\n" % (repr(0.0), repr(float(end_time))))
            w("**\n*Material, name=%s\n*Density\n %s,\n*Elastic\n%s, %s\n" % (m.name, repr(m.density), repr(m.young), repr(m.poisson)))
            if m.plastic.shape[0]:
                w("*Plastic\n")
                for r in m.plastic:
                    w(" %s, %s\n" % (repr(float(r[0])), repr(float(r[1]))))
            if m.ductile.shape[0]:
                w("*Damage Initiation, criterion=DUCTILE\n")
                for r in m.ductile:
                    w(" %s, %s, %s\n" % (repr(float(r[0])), repr(float(r[1])), repr(float(r[2]))))
            w("**\n*Initial Conditions, type=VELOCITY\n")
            for k in range(1, nl):
                w("L%d, 3, %s\n" % (k, repr(float(rate * (k * self.h)))))
            w("**\n*Step, name=Step-1, nlgeom=YES\n*Dynamic, Explicit\n%s, %s\n" % (repr(self.d_time), repr(float(end_time))))
            w("**\n*Boundary, amplitude=Amp-1\nL0, 3, 3\nL%d, 3, 3, %s\n**\n*End Step\n" % (nl - 1, repr(float(rate * lz * end_time))))


def B1(scale: float = 1.0, **kw) -> StretchDeck:
    """1 M-hex elastoplastic bar 50x50x400 (scale<1 shrinks every edge count for tests)."""
    s = lambda n: max(1, int(round(n * scale)))
    return StretchDeck(s(50), s(50), s(400), name="B1", **kw)


def F16(n: int = 252, nz: Optional[int] = None, ductile=True, **kw) -> StretchDeck:
    """16 M-hex block (252^3) with jittered interior nodes; ductile table scaled so deletions start early."""
    mat = steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]) if ductile else steel()
    kw.setdefault("jitter", 0.05)
    return StretchDeck(n, n, nz if nz is not None else n, material=mat, name="F16", **kw)


# ------------------------------------------------------------------------------ two-instance impact (I8 style)
@dataclass
class ImpactDeck:
    """Hex projectile above a hex plate, frictionless-or-not penalty contact, ALL EXTERIOR.
    SI units; materials `alum` / `lead` of HAKAI-v0.0.0/input/bullet-impact.inp."""
    plate: tuple = (400, 400, 48)
    proj: tuple = (68, 68, 68)
    h: float = 0.5e-3
    gap: float = 0.1
    v0: float = -500.0
    d_time: float = 1.0e-08
    n_steps: float = 99.5
    name: str = "impact"
    plate_ductile: Optional[list] = None      # rows [eps_f, triax, rate] replacing alum's table (brittle plates in tests)
    contact_pair: bool = False                # write_inp: `*Surface` element sets (top layer of the plate, bottom layer of
                                              # the projectile) + `*Contact Pair` instead of `*Contact Inclusions, ALL
                                              # EXTERIOR` (readInpFile_j.jl:1062-1103, HAKAI_j.jl:2094-2119)

    def materials(self):
        alum = Material(name="alum", density=2800., young=7e+10, poisson=0.33)
        alum.plastic = np.array([[1.6e+8, 0.], [3.4e+8, 0.3]])
        alum.Hd = (alum.plastic[1:, 0] - alum.plastic[:-1, 0]) / (alum.plastic[1:, 1] - alum.plastic[:-1, 1])
        alum.ductile = np.array([[1.0, 0., 30.], [0.7, 0.4, 30.]] if self.plate_ductile is None else self.plate_ductile,
                                dtype=np.float64)
        alum.fracture_flag = 1
        lead = Material(name="lead", density=11340., young=1.4e+10, poisson=0.425)
        return [alum, lead]

    def build_model(self) -> Model:
        h = self.h
        px, py, pz = self.plate
        qx, qy, qz = self.proj
        c1, e1 = block_arrays(px, py, pz, h)
        ox = (px - qx) * h / 2.0
        oy = (py - qy) * h / 2.0
        oz = pz * h + self.gap * h
        c2, e2 = block_arrays(qx, qy, qz, h)
        mats = self.materials()
        parts = [Part(name="plate", nNode=c1.shape[1], coordmat=c1, nElement=e1.shape[1], elementmat=e1,
                      material_name="alum", material_id=1),
                 Part(name="proj", nNode=c2.shape[1], coordmat=c2, nElement=e2.shape[1], elementmat=e2,
                      material_name="lead", material_id=2)]
        c2g = c2 + np.array([[ox], [oy], [oz]])
        n1, m1 = c1.shape[1], e1.shape[1]
        n2, m2 = c2.shape[1], e2.shape[1]
        insts = [Instance(name="plate-1", part_name="plate", part_id=1, material_id=1, node_offset=0, nNode=n1,
                          element_offset=0, nElement=m1, elements=np.arange(1, m1 + 1)),
                 Instance(name="proj-1", part_name="proj", part_id=2, material_id=2, translate=[],
                          node_offset=n1, nNode=n2, element_offset=m1, nElement=m2, elements=np.arange(1, m2 + 1))]
        coord = np.concatenate([c1, c2g], axis=1)
        em = np.concatenate([e1, e2 + n1], axis=1)
        # plate edge faces ENCASTRE
        nnx, nny, nnz = px + 1, py + 1, pz + 1
        ii = np.tile(np.arange(nnx), nny * nnz)
        jj = np.tile(np.repeat(np.arange(nny), nnx), nnz)
        edge = np.flatnonzero((ii == 0) | (ii == nnx - 1) | (jj == 0) | (jj == nny - 1)).astype(np.int64) + 1
        bc = BC(Nset_name="edge")
        bc.dof = [np.concatenate([edge * 3 - 2, edge * 3 - 1, edge * 3])]
        bc.value = [0.0]
        pn = np.arange(n1 + 1, n1 + n2 + 1, dtype=np.int64)
        ic = IC(Nset_name="proj", type="VELOCITY", dof=[pn * 3], value=[self.v0])
        end_time = self.n_steps * self.d_time
        return Model(PART=parts, INSTANCE=insts, NSET=[], ELSET=[], SURFACE=[], AMPLITUDE=[], MATERIAL=mats,
                     BC=[bc], IC=[ic], CP=[], nNode=n1 + n2, coordmat=coord, nElement=m1 + m2, elementmat=em,
                     element_material=np.concatenate([np.full(m1, 1, np.int64), np.full(m2, 2, np.int64)]),
                     element_instance=np.concatenate([np.full(m1, 1, np.int64), np.full(m2, 2, np.int64)]),
                     d_time=self.d_time, end_time=end_time, mass_scaling=1.0, contact_flag=1)

    def write_inp(self, path: str):
        """The same model as an Abaqus deck in the dialect of HAKAI-v0.0.0/input/bullet-impact.inp (two parts, two
        instances, assembly-level node set for the clamped plate edges, part-level `generate` sets, `*Contact` +
        `*Contact Inclusions, ALL EXTERIOR`), so that the reference and `read_inp_file` see the arrays of build_model()."""
        m = self.build_model()
        mats = {x.name: x for x in m.MATERIAL}
        with open(path, "w") as f:
            w = f.write
            w("*Heading\n** synthetic impact deck %s plate %s proj %s\n" % (self.name, self.plate, self.proj))
            for part in m.PART:
                w("*Part, name=%s\n" % part.name)
                _write_nodes_elements(f, part.coordmat, part.elementmat)
                w("*Nset, nset=Set-all, generate\n  1, %d, 1\n" % part.nNode)
                w("*Elset, elset=Set-all, generate\n 1, %d, 1\n" % part.nElement)
                w("*Solid Section, elset=Set-all, material=%s\n,\n*End Part\n**\n" % part.material_name)
            px, py, pz = self.plate
            qx, qy, _ = self.proj
            h = self.h
            w("*Assembly, name=Assembly\n**\n*Instance, name=plate-1, part=plate\n*End Instance\n**\n")
            w("*Instance, name=proj-1, part=proj\n%s, %s, %s\n*End Instance\n**\n" % (
                repr(float((px - qx) * h / 2.0)), repr(float((py - qy) * h / 2.0)), repr(float(pz * h + self.gap * h))))
            edge = (m.BC[0].dof[0][: len(m.BC[0].dof[0]) // 3] + 2) // 3
            w("*Nset, nset=edge, instance=plate-1\n")
            for i in range(0, len(edge), 16):
                w(", ".join(str(int(v)) for v in edge[i:i + 16]) + "\n")
            if self.contact_pair:
                pz, qy = self.plate[2], self.proj[1]
                w("*Elset, elset=_plate-top_S2, internal, instance=plate-1, generate\n %d, %d, 1\n" % ((pz - 1) * px * py + 1, pz * px * py))
                w("*Surface, type=ELEMENT, name=plate-top\n_plate-top_S2, S2\n")
                w("*Elset, elset=_proj-bottom_S1, internal, instance=proj-1, generate\n 1, %d, 1\n" % (qx * qy))
                w("*Surface, type=ELEMENT, name=proj-bottom\n_proj-bottom_S1, S1\n")
            w("*End Assembly\n**\n")
            for name in ("alum", "lead"):
                x = mats[name]
                w("*Material, name=%s\n*Density\n %s,\n*Elastic\n%s, %s\n" % (name, repr(x.density), repr(x.young), repr(x.poisson)))
                if x.plastic.shape[0]:
                    w("*Plastic\n")
                    for r in x.plastic:
                        w(" %s, %s\n" % (repr(float(r[0])), repr(float(r[1]))))
                if x.ductile.shape[0]:
                    w("*Damage Initiation, criterion=DUCTILE\n")
                    for r in x.ductile:
                        w(" %s, %s, %s\n" % (repr(float(r[0])), repr(float(r[1])), repr(float(r[2]))))
            w("**\n*Boundary\nedge, ENCASTRE\n")
            w("**\n*Initial Conditions, type=VELOCITY\nproj-1.Set-all, 3, %s\n" % repr(float(self.v0)))
            if self.contact_pair:
                w("**\n*Surface Interaction, name=IntProp-1\n1.,\n")
                w("*Contact Pair, interaction=IntProp-1, mechanical constraint=KINEMATIC, cpset=CP-1\nproj-bottom, plate-top\n")
            else:
                w("**\n*Contact, op=NEW\n*Contact Inclusions, ALL EXTERIOR\n")
            w("**\n*Step, name=Step-1, nlgeom=YES\n*Dynamic, Explicit\n%s, %s\n**\n*End Step\n" % (
                repr(self.d_time), repr(float(self.n_steps * self.d_time))))
