"""Host-side set-up that defines the inputs of the time-step engine.

Python twin of the part of ``hakai()`` that runs once before the loop
(HAKAI-v0.0.2/Julia/HAKAI_j.jl:81-480, cited J2:<line>): material constants, lumped mass,
initial conditions, contact surfaces / pairs, element sizes.  Everything is vectorised NumPy so
the same code serves 5-element decks and 16 M-element synthetic meshes; the O(F^2) face
matching of the reference (J2:2040-2084) is replaced by a sort that yields the SAME face list,
including the "last face never emitted" quirk.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .inp import Model, CP as CPType

# Gauss points, z fastest (J2:1913-1920) and the C3D8 corner signs delta_mat (J2:1900-1907)
DELTA = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1],
                  [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=np.float64)


def cal_Pusai_hexa() -> np.ndarray:
    """Shape-function derivatives at the 8 Gauss points: P[k, dir, node] (J2:1895-1943)."""
    g = 1.0 / np.sqrt(3.0)
    gc = np.array([[-g, -g, -g], [-g, -g, g], [-g, g, -g], [-g, g, g],
                   [g, -g, -g], [g, -g, g], [g, g, -g], [g, g, g]])
    P = np.zeros((8, 3, 8))
    for k in range(8):
        gz, et, tu = gc[k]
        for i in range(8):
            d = DELTA[i]
            P[k, 0, i] = 1.0 / 8.0 * d[0] * (1.0 + et * d[1]) * (1.0 + tu * d[2])
            P[k, 1, i] = 1.0 / 8.0 * d[1] * (1.0 + gz * d[0]) * (1.0 + tu * d[2])
            P[k, 2, i] = 1.0 / 8.0 * d[2] * (1.0 + gz * d[0]) * (1.0 + et * d[1])
    return P


def element_volumes(coordmat: np.ndarray, elementmat: np.ndarray, chunk: int = 1 << 19) -> np.ndarray:
    """elementVolume[e] = sum_k my3det(Pusai_k * e_position')  (J2:183-198, my3det J2:3235-3243: Sarrus, same term
    order; Gauss points added in order)."""
    P24 = cal_Pusai_hexa().reshape(24, 8)                      # rows k*3+dir
    nE = elementmat.shape[1]
    V = np.zeros(nE)
    X = np.ascontiguousarray(coordmat.T)                       # (nN,3)
    for s in range(0, nE, chunk):
        em = elementmat[:, s:s + chunk].T - 1                  # (n,8)
        J = np.matmul(P24, X[em])                              # (n,24,3): J[n, k*3+d, c]
        acc = np.zeros(J.shape[0])
        for k in range(8):
            a, b, c = J[:, 3 * k], J[:, 3 * k + 1], J[:, 3 * k + 2]            # rows of the 3x3 Jacobian
            acc += (a[:, 0] * b[:, 1] * c[:, 2] + a[:, 1] * b[:, 2] * c[:, 0] + a[:, 2] * b[:, 0] * c[:, 1]
                    - a[:, 0] * b[:, 2] * c[:, 1] - a[:, 1] * b[:, 0] * c[:, 2] - a[:, 2] * b[:, 1] * c[:, 0])
        V[s:s + chunk] = acc
    return V


def lumped_mass(model: Model, elementVolume: np.ndarray) -> np.ndarray:
    """diag_M (J2:201-215): rho*V/8 to each of the 8 nodes, same value on the 3 dofs, times mass_scaling."""
    dens = np.array([m.density for m in model.MATERIAL])
    node_mass = dens[model.element_material - 1] * elementVolume / 8.0
    # the reference adds element by element, node by node (J2:201-209): bincount accumulates in input order, so an
    # element-major / node-minor index stream reproduces that summation order (and its rounding) exactly
    m = np.bincount(model.elementmat.T.reshape(-1) - 1, weights=np.repeat(node_mass, 8), minlength=model.nNode)
    return np.repeat(m, 3) * model.mass_scaling


def element_sizes(coordmat, elementmat):
    """elementMinSize / elementMaxSize from edges 1-2, 1-4, 1-5 only (J2:405-421)."""
    X = coordmat
    e = elementmat - 1
    mn, mx = np.inf, 0.0
    for a in (1, 3, 4):
        d = X[:, e[0]] - X[:, e[a]]
        L = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
        mn = min(mn, float(L.min()))
        mx = max(mx, float(L.max()))
    return mn, mx


# ----------------------------------------------------------------------------- contact surfaces
def get_element_face(part_coordmat: np.ndarray, part_elementmat: np.ndarray):
    """get_element_face (J2:1946-1992): 6 faces per element, re-ordered to point outward using the
    undeformed PART coordinates.  Returns faces (6nE,4), faces_eleid (6nE), sorted_faces (6nE,4); 1-based."""
    em = part_elementmat                                       # (8,nE) 1-based
    nE = em.shape[1]
    loc = np.array([[0, 1, 2, 3], [4, 5, 6, 7], [0, 1, 5, 4], [1, 2, 6, 5], [2, 3, 7, 6], [3, 0, 4, 7]])
    faces = em[loc, :].transpose(2, 0, 1).reshape(nE * 6, 4).copy()          # element-major, face k
    eleid = np.repeat(np.arange(1, nE + 1, dtype=np.int64), 6)
    X = part_coordmat
    ctr = X[:, em - 1].sum(axis=1) / 8                                       # (3,nE): sum over the 8 nodes
    ctr = np.repeat(ctr, 6, axis=1)                                          # (3,6nE)
    p1 = X[:, faces[:, 0] - 1]
    v1 = X[:, faces[:, 1] - 1] - p1
    v2 = X[:, faces[:, 3] - 1] - p1
    nv = np.stack([v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]])
    vc = ctr - p1
    flip = (nv * vc).sum(axis=0) > 0.0
    faces[flip] = faces[flip][:, [0, 3, 2, 1]]
    return faces, eleid, np.sort(faces, axis=1)


def _unique_face_mask(sorted_faces: np.ndarray) -> np.ndarray:
    """Which rows the loop J2:2040-2084 emits: rows whose sorted 4-tuple has no partner (pairs are
    matched first-with-next in index order, so an odd group emits its last member), and never the very
    last row (loop bound `1 : nE*6-1`, J2:2040)."""
    F = sorted_faces.shape[0]
    order = np.lexsort((np.arange(F), sorted_faces[:, 3], sorted_faces[:, 2], sorted_faces[:, 1], sorted_faces[:, 0]))
    sf = sorted_faces[order]
    newgrp = np.ones(F, bool)
    newgrp[1:] = np.any(sf[1:] != sf[:-1], axis=1)
    gid = np.cumsum(newgrp) - 1
    gsize = np.bincount(gid)
    gstart = np.flatnonzero(newgrp)
    rank = np.arange(F) - gstart[gid]
    emit_sorted = (gsize[gid] % 2 == 1) & (rank == gsize[gid] - 1)
    mask = np.zeros(F, bool)
    mask[order] = emit_sorted
    if F > 0:
        mask[F - 1] = False
    return mask


def get_surface_triangle(surfaces, sorted_surfaces, surfaces_eleid, nElement_instance, contact_element):
    """get_surface_triangle (J2:1996-2164) for array_element = all elements of the instance.
    Returns c_triangles (nTri,3), c_triangles_eleid (nTri), c_nodes (sorted unique); part-local 1-based."""
    mask = _unique_face_mask(sorted_surfaces)
    c_surfaces = surfaces[mask]
    c_eleid = surfaces_eleid[mask]
    if nElement_instance != len(contact_element):               # J2:2094-2119
        keep = np.isin(c_eleid, contact_element)
        c_surfaces = c_surfaces[keep]
        c_eleid = c_eleid[keep]
    n = c_surfaces.shape[0]
    tri = np.zeros((2 * n, 3), np.int64)
    tri[0::2] = c_surfaces[:, [0, 1, 2]]                         # J2:2140-2145
    tri[1::2] = c_surfaces[:, [2, 3, 0]]
    tri_eleid = np.repeat(c_eleid, 2)
    c_nodes = np.unique(tri)
    return tri, tri_eleid, c_nodes


@dataclass
class ContactTriangle:           # J2:72-78 (+ the ordered instance pair it belongs to)
    i_instance: int
    j_instance: int
    c_nodes_i: np.ndarray
    c_nodes_j: np.ndarray
    c_triangles: np.ndarray      # (nTri,3) global 1-based
    c_triangles_eleid: np.ndarray
    young: float


@dataclass
class Setup:
    model: Model
    d_time: float
    time_num: float
    elementVolume: np.ndarray
    diag_M: np.ndarray
    elementMinSize: float
    elementMaxSize: float
    CT: List[ContactTriangle] = field(default_factory=list)
    instance_pair: List[List[int]] = field(default_factory=list)
    all_exterior_flag: int = 0
    contact_on_device: bool = False   # prepare(contact="device"): hk_build_contact builds the tables on the GPU


def build_contact(model: Model, setup: Setup):
    """Contact set-up, J2:250-398."""
    for ins in model.INSTANCE:
        part = model.PART[ins.part_id - 1]
        ins.surfaces, ins.surfaces_eleid, ins.sorted_surfaces = get_element_face(part.coordmat, part.elementmat)
    if len(model.CP) == 0:                                       # ALL EXTERIOR, J2:272-314
        setup.all_exterior_flag = 1
        ni = len(model.INSTANCE)
        cps = []
        if ni > 1:
            for i in range(1, ni + 1):
                js = i if model.contact_flag == 2 else i + 1
                for j in range(js, ni + 1):
                    cp = CPType()
                    cp.instance_id_1, cp.instance_id_2 = i, j
                    cp.elements_1 = np.arange(1, model.INSTANCE[i - 1].nElement + 1)
                    cp.elements_2 = np.arange(1, model.INSTANCE[j - 1].nElement + 1)
                    cps.append(cp)
        else:
            cp = CPType()
            cp.instance_id_1 = cp.instance_id_2 = 1
            cp.elements_1 = np.arange(1, model.INSTANCE[0].nElement + 1)
            cp.elements_2 = cp.elements_1.copy()
            cps.append(cp)
        model.CP = cps
    for cp in model.CP:                                          # J2:321-336
        I1 = model.INSTANCE[cp.instance_id_1 - 1]
        cp.c_triangles_1, cp.c_triangles_eleid_1, cp.c_nodes_1 = get_surface_triangle(
            I1.surfaces, I1.sorted_surfaces, I1.surfaces_eleid, I1.nElement, cp.elements_1)
        I2 = model.INSTANCE[cp.instance_id_2 - 1]
        cp.c_triangles_2, cp.c_triangles_eleid_2, cp.c_nodes_2 = get_surface_triangle(
            I2.surfaces, I2.sorted_surfaces, I2.surfaces_eleid, I2.nElement, cp.elements_2)
    instance_pair, cp_index = [], []
    for cc, cp in enumerate(model.CP):                           # J2:339-354
        if cp.instance_id_1 == cp.instance_id_2:
            instance_pair.append([cp.instance_id_1, cp.instance_id_2]); cp_index.append(cc)
        else:
            instance_pair.append([cp.instance_id_1, cp.instance_id_2]); cp_index.append(cc)
            instance_pair.append([cp.instance_id_2, cp.instance_id_1]); cp_index.append(cc)
    CT = []
    for c, cc in enumerate(cp_index):                            # J2:361-398
        i_inst, j_inst = instance_pair[c]
        cp = model.CP[cc]
        Ii, Ij = model.INSTANCE[i_inst - 1], model.INSTANCE[j_inst - 1]
        young = model.MATERIAL[Ij.material_id - 1].young
        if cp.instance_id_1 == i_inst:
            ni_, nj_, tr_, te_ = cp.c_nodes_1, cp.c_nodes_2, cp.c_triangles_2, cp.c_triangles_eleid_2
        else:
            ni_, nj_, tr_, te_ = cp.c_nodes_2, cp.c_nodes_1, cp.c_triangles_1, cp.c_triangles_eleid_1
        CT.append(ContactTriangle(i_inst, j_inst, ni_ + Ii.node_offset, nj_ + Ij.node_offset,
                                  tr_ + Ij.node_offset, te_ + Ij.element_offset, young))
    setup.CT = CT
    setup.instance_pair = instance_pair


def prepare(model: Model, elementVolume: Optional[np.ndarray] = None, contact: str = "host") -> Setup:
    """Everything hakai() computes before `for t = 1 : time_num` (J2:100-465).
    contact="device": the contact tables (get_element_face, get_surface_triangle, pair lists: J2:250-398) are NOT built
    here; configure_engine hands the instance ranges and the *Contact Pair list to hk_build_contact, which builds them
    on the GPU."""
    d_time = model.d_time * np.sqrt(model.mass_scaling)          # J2:114
    time_num = model.end_time / d_time
    for m in model.MATERIAL:                                     # J2:143-172 (Dmat is rebuilt in the engine)
        m.G = m.young / 2.0 / (1.0 + m.poisson)
    if elementVolume is None:
        elementVolume = element_volumes(model.coordmat, model.elementmat)
    diag_M = lumped_mass(model, elementVolume)
    emin, emax = element_sizes(model.coordmat, model.elementmat)
    st = Setup(model, float(d_time), float(time_num), elementVolume, diag_M, emin, emax)
    if contact not in ("host", "device"):
        raise ValueError("contact: host | device")
    st.contact_on_device = contact == "device"
    if model.contact_flag >= 1 and not st.contact_on_device:
        build_contact(model, st)
    return st


def configure_engine(engine_cls, setup: Setup, **param_overrides):
    """Creates an engine and feeds it the model through the C ABI (order = the header's contract)."""
    model = setup.model
    params = dict(d_time=setup.d_time, element_min_size=setup.elementMinSize,
                  element_max_size=setup.elementMaxSize, contact_flag=int(model.contact_flag))
    params.update(param_overrides)
    eng = engine_cls(**params)
    eng.set_mesh(model.coordmat, model.elementmat, model.element_material, model.element_instance, setup.diag_M)
    for m in model.MATERIAL:
        eng.add_material(m.young, m.poisson, m.density,
                         m.plastic if m.plastic.shape[0] else None, m.Hd if m.plastic.shape[0] > 1 else None,
                         m.ductile if m.ductile.shape[0] else None)
    for bc in model.BC:
        has_amp = len(bc.amp_name) > 0
        eng.add_bc(bc.dof, bc.value, bc.amplitude.time if has_amp else None,
                   bc.amplitude.value if has_amp else None)
    for ic in model.IC:
        eng.add_ic(ic.dof, ic.value)
    if model.contact_flag >= 1 and setup.contact_on_device:
        inst = [(i.node_offset, i.nNode, i.element_offset, i.nElement) for i in model.INSTANCE]
        young = [model.MATERIAL[i.material_id - 1].young for i in model.INSTANCE]
        pairs = [(cp.instance_id_1, cp.instance_id_2, cp.elements_1, cp.elements_2) for cp in model.CP] or None
        eng.build_contact(inst, young, pairs)
    elif model.contact_flag >= 1:
        for ins in model.INSTANCE:
            eng.add_instance(ins.node_offset, ins.nNode, ins.element_offset, ins.nElement,
                             ins.surfaces, ins.surfaces_eleid)
        for ct in setup.CT:
            eng.add_contact_pair(ct.i_instance, ct.j_instance, ct.c_nodes_i, ct.c_nodes_j, ct.c_triangles,
                                 ct.c_triangles_eleid, ct.young)
    eng.finalize()
    return eng
