"""Abaqus ``.inp`` deck reader for the HAKAI host side.

Mirror of the reference's ``readInpFile`` (HAKAI-v0.0.2/Julia/readInpFile_j.jl:152-1112,
types at :23-150).  Julia is not available in the build image, so the host side that the
north star leaves in Julia is restated here in Python; the deck -> array mapping (including
the reader's quirks, listed below) is what defines the inputs of the time-step engine.

Quirks reproduced on purpose (they change the arrays the hot path sees):
  * only the LAST data line of an ``*Amplitude`` block survives      (readInpFile_j.jl:649-665)
  * part-level ``*Nset`` is read only when it has ``generate``        (:261-290)
  * ``*Boundary`` data: direction = 3rd field, 2nd field ignored, dirs > 3 dropped (:931-953)
  * ``ENCASTRE`` = the 3 translations, value list reset to [0.]       (:923-930)
  * instance transform lines applied last-to-first; 3 numbers = translate,
    7 numbers = rotation by the axis DIRECTION only (about the origin) (:581-605)
  * ``*Fixed Mass Scaling, factor=`` is read from the keyword line     (:829-841)
  * contact_flag = 1 on any line containing ``*Contact``; 2 with
    ``HAKAIoption=self-contact`` on ``*Contact Inclusions``           (:1047-1060)

All node / element / dof ids stored in the returned model are 1-based, exactly like the
reference; conversion to 0-based happens once, inside the engine's ``hk_set_*`` calls.
"""
from __future__ import annotations

import bisect

import math
from dataclasses import dataclass, field
from typing import List

import numpy as np


@dataclass
class Nset:                      # NsetType, readInpFile_j.jl:23-30
    name: str = ""
    instance_name: str = ""
    instance_id: int = 0
    part_name: str = ""
    part_id: int = 0
    nodes: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))


@dataclass
class Elset:                     # ELsetType, :32-39
    name: str = ""
    instance_name: str = ""
    instance_id: int = 0
    part_name: str = ""
    part_id: int = 0
    elements: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))


@dataclass
class Surface:                   # SurfaceType, :41-46
    name: str = ""
    elset_name: List[str] = field(default_factory=list)
    instance_id: int = 0
    elements: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))


@dataclass
class Part:                      # PartType, :48-57
    name: str = ""
    nNode: int = 0
    coordmat: np.ndarray = None          # (3, nNode) column-major semantics
    nElement: int = 0
    elementmat: np.ndarray = None        # (8, nElement), 1-based
    NSET: List[Nset] = field(default_factory=list)
    material_name: str = ""
    material_id: int = 0


@dataclass
class Instance:                  # InstanceType, :59-76
    name: str = ""
    part_name: str = ""
    part_id: int = 0
    material_id: int = 0
    translate: List[str] = field(default_factory=list)
    node_offset: int = 0
    nNode: int = 0
    element_offset: int = 0
    nElement: int = 0
    elements: np.ndarray = None
    surfaces: np.ndarray = None          # (6 nE, 4) outward-oriented faces, part-local ids
    sorted_surfaces: np.ndarray = None
    surfaces_eleid: np.ndarray = None


@dataclass
class Amplitude:                 # AmplitudeType, :78-82
    name: str = ""
    time: np.ndarray = field(default_factory=lambda: np.zeros(1))
    value: np.ndarray = field(default_factory=lambda: np.zeros(1))


@dataclass
class Material:                  # MaterialType, :84-96
    name: str = ""
    density: float = 0.0
    young: float = 0.0
    poisson: float = 0.0
    plastic: np.ndarray = field(default_factory=lambda: np.zeros((0, 2)))
    Hd: np.ndarray = field(default_factory=lambda: np.zeros(0))
    fracture_flag: int = 0
    failure_stress: float = 0.0
    ductile: np.ndarray = field(default_factory=lambda: np.zeros((0, 3)))
    G: float = 0.0
    Dmat: np.ndarray = None


@dataclass
class BC:                        # BCType, :98-104
    Nset_name: str = ""
    dof: List[np.ndarray] = field(default_factory=list)     # 1-based dof ids
    value: List[float] = field(default_factory=list)
    amp_name: str = ""
    amplitude: Amplitude = field(default_factory=Amplitude)


@dataclass
class IC:                        # ICType, :106-111
    Nset_name: str = ""
    type: str = ""
    dof: List[np.ndarray] = field(default_factory=list)
    value: List[float] = field(default_factory=list)


@dataclass
class CP:                        # CPType, :113-127
    name: str = ""
    surface_name_1: str = ""
    surface_name_2: str = ""
    instance_id_1: int = 0
    instance_id_2: int = 0
    elements_1: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    elements_2: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int64))
    c_triangles_1: np.ndarray = None
    c_triangles_2: np.ndarray = None
    c_triangles_eleid_1: np.ndarray = None
    c_triangles_eleid_2: np.ndarray = None
    c_nodes_1: np.ndarray = None
    c_nodes_2: np.ndarray = None


@dataclass
class Model:                     # ModelType, :129-150
    PART: List[Part]
    INSTANCE: List[Instance]
    NSET: List[Nset]
    ELSET: List[Elset]
    SURFACE: List[Surface]
    AMPLITUDE: List[Amplitude]
    MATERIAL: List[Material]
    BC: List[BC]
    IC: List[IC]
    CP: List[CP]
    nNode: int
    coordmat: np.ndarray         # (3, nNode) float64
    nElement: int
    elementmat: np.ndarray       # (8, nElement) int64, 1-based global node ids
    element_material: np.ndarray  # (nElement,) 1-based
    element_instance: np.ndarray  # (nElement,) 1-based
    d_time: float
    end_time: float
    mass_scaling: float
    contact_flag: int


def _nosp(s: str) -> str:
    return s.replace(" ", "")


def _fields(s: str) -> List[str]:
    """split(s, ",", keepempty=false) on a space-stripped line."""
    return [x for x in _nosp(s).split(",") if x != ""]


def _after(token: str, key: str) -> str:
    return token[token.index(key) + len(key):]


def _jl_range(a: int, step: int, b: int) -> np.ndarray:
    """Julia a:step:b (inclusive)."""
    if step > 0:
        return np.arange(a, b + 1, step, dtype=np.int64)
    return np.arange(a, b - 1, step, dtype=np.int64)


def _read_table(lines, start, ncol, dtype):
    """Rows after ``start`` up to (not incl.) the first line containing '*'."""
    n = len(lines)
    end = start + 1
    while end < n and "*" not in lines[end]:
        end += 1
    rows = end - start - 1
    if rows == 0:
        return np.zeros((0, ncol), dtype)
    try:                                                        # regular table: NumPy's C parser (same correctly
        arr = np.loadtxt(lines[start + 1:end], delimiter=",", dtype=np.float64, ndmin=2)    # rounded doubles)
        if arr.shape[1] == len(_fields(lines[start + 1])) and np.all(np.isfinite(arr)):
            return arr
    except ValueError:                                          # ragged rows / trailing commas: the general path
        pass
    txt = ",".join(_nosp(l).rstrip(",") for l in lines[start + 1:end])
    flat = np.array(txt.split(","), dtype=object)
    per = len(_fields(lines[start + 1]))
    arr = flat.reshape(rows, per)
    return arr


def read_inp_file(fname: str) -> Model:
    """readInpFile (readInpFile_j.jl:152)."""
    with open(fname, "r", newline=None) as f:
        lines = [l.rstrip("\r\n") for l in f.read().split("\n")]
    if lines and lines[-1] == "":
        lines.pop()
    return parse_inp_lines(lines)


def parse_inp_lines(lines: List[str]) -> Model:
    n = len(lines)
    # every keyword test below looks for a token starting with '*': scan only the lines that contain one (a 16 M-
    # element deck has 3e7 data lines and a few hundred keyword lines; readInpFile re-scans the whole file per keyword)
    star = [i for i, l in enumerate(lines) if "*" in l]

    def star_from(i0):
        return star[bisect.bisect_left(star, i0):]

    # --- Part ---  (:165-308)
    part_index = [i for i in star if "*Part, name=" in lines[i]]
    PART: List[Part] = []
    for pi in part_index:
        p = Part()
        p.name = _after(_fields(lines[pi])[1], "name=")

        index = 0
        for i in star_from(pi):
            if "*Node" in lines[i]:
                index = i
                break
        tab = _read_table(lines, index, 4, object)
        p.nNode = tab.shape[0]
        p.coordmat = np.ascontiguousarray(tab[:, 1:4].astype(np.float64).T)

        index = 0
        for i in star_from(pi):
            if "*Element" in lines[i]:
                index = i
                break
        tab = _read_table(lines, index, 9, object)
        p.nElement = tab.shape[0]
        p.elementmat = np.ascontiguousarray(tab[:, 1:9].astype(np.int64).T)

        for i in star_from(pi):
            if "*Nset" in lines[i] and "generate" in lines[i]:
                ns = Nset()
                ns.name = _after(_fields(lines[i])[1], "nset=")
                ss = _nosp(lines[i + 1]).split(",")
                ns.nodes = _jl_range(int(ss[0]), int(ss[2]), int(ss[1]))
                p.NSET.append(ns)
            if "*End Part" in lines[i]:
                break

        for i in star_from(pi):
            if "*Solid Section" in lines[i]:
                for tok in _fields(lines[i]):
                    if "material=" in tok:
                        p.material_name = _after(tok, "material=")
                        break
                break
        PART.append(p)
    nPart = len(PART)

    # --- Instance ---  (:311-362)
    INSTANCE: List[Instance] = []
    for ii in [i for i in star if "*Instance" in lines[i]]:
        ins = Instance()
        ff = _fields(lines[ii])
        ins.name = _after(ff[1], "name=")
        ins.part_name = _after(ff[2], "part=")
        for k in range(nPart):
            if PART[k].name == ins.part_name:
                ins.part_id = k + 1
                break
        for i in range(ii + 1, n):
            if "*End Instance" in lines[i]:
                break
            ins.translate.append(_nosp(lines[i]))
        INSTANCE.append(ins)
    instance_num = len(INSTANCE)

    def _int_list_until_star(start):
        a = []
        for i in range(start + 1, n):
            if "*" in lines[i]:
                break
            a.extend(int(x) for x in _fields(lines[i]))
        return np.array(a, dtype=np.int64)

    # --- Nset (assembly level) ---  (:365-432)
    NSET: List[Nset] = []
    for idx in [i for i in star if "*Nset" in lines[i] and "instance=" in lines[i]]:
        ns = Nset()
        ff = _fields(lines[idx])
        ns.name = _after(ff[1], "nset=")
        ns.instance_name = _after(ff[2], "instance=")
        for i in range(instance_num):
            if ns.instance_name == INSTANCE[i].name:
                ns.part_name = INSTANCE[i].part_name
                ns.part_id = INSTANCE[i].part_id
                ns.instance_id = i + 1
        if len(ff) == 4 and ff[3] == "generate":
            ss = _fields(lines[idx + 1])
            ns.nodes = _jl_range(int(ss[0]), int(ss[2]), int(ss[1]))
        else:
            ns.nodes = _int_list_until_star(idx)
        NSET.append(ns)
    nset_num = len(NSET)

    # --- Elset ---  (:435-514)
    ELSET: List[Elset] = []
    for idx in [i for i in star if "*Elset" in lines[i] and "instance=" in lines[i]]:
        es = Elset()
        ff = _fields(lines[idx])
        es.name = _after(ff[1], "elset=")
        if "instance=" in ff[2]:
            es.instance_name = _after(ff[2], "instance=")
        elif "instance=" in ff[3]:
            es.instance_name = _after(ff[3], "instance=")
        for i in range(instance_num):
            if es.instance_name == INSTANCE[i].name:
                es.part_name = INSTANCE[i].part_name
                es.part_id = INSTANCE[i].part_id
                es.instance_id = i + 1
        a = np.zeros(0, np.int64)
        if (len(ff) == 4 and ff[3] == "generate") or \
           (len(ff) == 5 and ff[2] == "internal" and ff[4] == "generate"):
            ss = _fields(lines[idx + 1])
            a = _jl_range(int(ss[0]), int(ss[2]), int(ss[1]))
        elif len(ff) == 4 and ff[2] == "internal":
            a = _int_list_until_star(idx)
        es.elements = a
        ELSET.append(es)

    # --- Surface ---  (:517-563)
    SURFACE: List[Surface] = []
    for idx in [i for i in star if "*Surface," in lines[i]]:
        sf = Surface()
        sf.name = _after(_fields(lines[idx])[2], "name=")
        a = []
        for i in range(idx + 1, n):
            if "*" in lines[i]:
                break
            nm = _fields(lines[i])[0]
            sf.elset_name.append(nm)
            for es in ELSET:
                if nm == es.name:
                    sf.instance_id = es.instance_id
                    a.extend(es.elements.tolist())
        sf.elements = np.unique(np.array(a, dtype=np.int64))
        SURFACE.append(sf)
    surface_num = len(SURFACE)

    # --- Global model ---  (:567-621)
    nNode = 0
    nElement = 0
    coord_blocks, elem_blocks = [], []
    for ins in INSTANCE:
        part = PART[ins.part_id - 1]
        coordmat_i = part.coordmat.copy()
        ins.node_offset = nNode
        ins.element_offset = nElement
        ins.nNode = part.nNode
        ins.nElement = part.nElement
        ins.elements = np.arange(1, ins.nElement + 1, dtype=np.int64)
        for s in reversed(ins.translate):
            ss = [x for x in s.split(",") if x != ""]
            if len(ss) == 3:
                off = np.array([float(ss[0]), float(ss[1]), float(ss[2])]).reshape(3, 1)
                coordmat_i = coordmat_i + off * np.ones((1, coordmat_i.shape[1]))
            elif len(ss) == 7:
                nv = np.array([float(ss[3]) - float(ss[0]),
                               float(ss[4]) - float(ss[1]),
                               float(ss[5]) - float(ss[2])])
                nv = nv / math.sqrt(float(nv @ nv))
                n1, n2, n3 = nv
                d = float(ss[6]) / 180.0 * math.pi
                c, sn = math.cos(d), math.sin(d)
                T = np.array([
                    [n1 * n1 * (1 - c) + c, n1 * n2 * (1 - c) - n3 * sn, n1 * n3 * (1 - c) + n2 * sn],
                    [n1 * n2 * (1 - c) + n3 * sn, n2 * n2 * (1 - c) + c, n2 * n3 * (1 - c) - n1 * sn],
                    [n1 * n3 * (1 - c) - n2 * sn, n2 * n3 * (1 - c) + n1 * sn, n3 * n3 * (1 - c) + c]])
                coordmat_i = T @ coordmat_i
        coord_blocks.append(coordmat_i)
        elem_blocks.append(part.elementmat + nNode)
        nNode += part.nNode
        nElement += part.nElement
    coordmat = np.ascontiguousarray(np.concatenate(coord_blocks, axis=1))
    elementmat = np.ascontiguousarray(np.concatenate(elem_blocks, axis=1))

    # --- Amplitude ---  (:624-668)
    AMPLITUDE: List[Amplitude] = []
    for idx in [i for i in star if "*Amplitude" in lines[i]]:
        am = Amplitude()
        am.name = _after(_fields(lines[idx])[1], "name=")
        for i in range(idx + 1, n):
            if "*" in lines[i]:
                break
            ss = _fields(lines[i])
            m = len(ss) // 2
            am.time = np.array([float(ss[2 * j]) for j in range(m)])
            am.value = np.array([float(ss[2 * j + 1]) for j in range(m)])
        AMPLITUDE.append(am)

    # --- Material ---  (:671-793)
    MATERIAL: List[Material] = []
    for idx in [i for i in star if "*Material" in lines[i]]:
        mt = Material()
        mt.name = _after(_fields(lines[idx])[1], "name=")
        plastic_index = -1
        ductile_index = -1
        for i in range(idx + 1, n):
            if "*Material" in lines[i] or "**" in lines[i]:
                break
            if "*Density" in lines[i]:
                mt.density = float(_fields(lines[i + 1])[0])
            if "*Elastic" in lines[i]:
                ss = _fields(lines[i + 1])
                mt.young = float(ss[0])
                mt.poisson = float(ss[1])
            if "*Plastic" in lines[i]:
                plastic_index = i
            if "*Damage Initiation" in lines[i] and "criterion=DUCTILE" in lines[i]:
                ductile_index = i
                mt.fracture_flag = 1
            if "*Tensile Failure" in lines[i]:
                mt.failure_stress = float(_fields(lines[i + 1])[0])
                mt.fracture_flag = 1
        if plastic_index > idx:
            rows = []
            for i in range(plastic_index + 1, n):
                if "*" in lines[i]:
                    break
                ss = _fields(lines[i])
                rows.append([float(ss[0]), float(ss[1])])
            mt.plastic = np.array(rows, dtype=np.float64).reshape(-1, 2)
        npp = mt.plastic.shape[0]
        if npp > 1:
            with np.errstate(divide="ignore", invalid="ignore"):
                mt.Hd = (mt.plastic[1:, 0] - mt.plastic[:-1, 0]) / (mt.plastic[1:, 1] - mt.plastic[:-1, 1])
        if ductile_index > idx:
            rows = []
            for i in range(ductile_index + 1, n):
                if "*" in lines[i]:
                    break
                ss = _fields(lines[i])
                rows.append([float(ss[0]), float(ss[1]), float(ss[2])])
            mt.ductile = np.array(rows, dtype=np.float64).reshape(-1, 3)
        MATERIAL.append(mt)

    element_material = []
    element_instance = []
    for i, ins in enumerate(INSTANCE):
        part = PART[ins.part_id - 1]
        for j, mt in enumerate(MATERIAL):
            if part.material_name == mt.name:
                part.material_id = j + 1
                ins.material_id = j + 1
        element_material.append(np.full(part.nElement, part.material_id, np.int64))
        element_instance.append(np.full(part.nElement, i + 1, np.int64))
    element_material = np.concatenate(element_material)
    element_instance = np.concatenate(element_instance)

    # --- Step / mass scaling ---  (:816-840)
    d_time = 0.0
    end_time = 0.0
    for i in star:
        if "*Dynamic, Explicit" in lines[i]:
            ss = _fields(lines[i + 1])
            d_time = float(ss[0])
            end_time = float(ss[1])
            break
    mass_scaling = 1.0
    for i in star:
        if "*Fixed Mass Scaling" in lines[i]:
            mass_scaling = float(_after(_fields(lines[i])[1], "factor="))
            break

    def _resolve_nodes(name, first_only):
        """Nset lookup shared by *Boundary (:888-920) and *Initial Conditions (:996-1027)."""
        nodes = np.zeros(0, np.int64)
        if "." in name:
            sss = [x for x in name.split(".") if x != ""]
            instance_id = 0
            part_id = 0
            for j, ins in enumerate(INSTANCE):
                if ins.name == sss[0]:
                    instance_id = j + 1
                    part_id = ins.part_id
                    break
            for ns in PART[part_id - 1].NSET:
                if ns.name == sss[1]:
                    nodes = ns.nodes + INSTANCE[instance_id - 1].node_offset
                    break
        else:
            acc = []
            for ns in NSET:
                if name == ns.name:
                    acc.append(ns.nodes + INSTANCE[ns.instance_id - 1].node_offset)
                    if first_only:
                        break
            if acc:
                nodes = np.concatenate(acc)
        return nodes.astype(np.int64)

    # --- BC ---  (:843-957)
    BCs: List[BC] = []
    for idx in [i for i in star if "*Boundary" in lines[i]]:
        bc = BC()
        ff = _fields(lines[idx])
        if len(ff) == 2 and "amplitude=" in ff[1]:
            bc.amp_name = _after(ff[1], "amplitude=")
            for am in AMPLITUDE:
                if am.name == bc.amp_name:
                    bc.amplitude = am
                    break
        for i in range(idx + 1, n):
            if "*Boundary" in lines[i] or "**" in lines[i]:
                break
            ss = _fields(lines[i])
            bc.Nset_name = ss[0]
            nodes = _resolve_nodes(bc.Nset_name, first_only=False)
            if len(ss) == 2 and "ENCASTRE" in ss[1]:
                dof = np.concatenate([nodes * 3 - 2, nodes * 3 - 1, nodes * 3])
                bc.dof.append(dof)
                bc.value = [0.0]
            elif len(ss) == 3:
                direction = int(ss[2])
                if direction <= 3:
                    bc.dof.append(nodes * 3 - (3 - direction))
                    bc.value.append(0.0)
            elif len(ss) == 4:
                direction = int(ss[2])
                value = float(ss[3])
                if direction <= 3:
                    bc.dof.append(nodes * 3 - (3 - direction))
                    bc.value.append(value)
        BCs.append(bc)

    # --- Initial conditions ---  (:960-1043)
    ICs: List[IC] = []
    for idx in [i for i in star if "*Initial Conditions" in lines[i]]:
        ic = IC()
        ic.type = _after(_fields(lines[idx])[1], "type=")
        for i in range(idx + 1, n):
            if "*Initial Conditions" in lines[i] or "**" in lines[i]:
                break
            ss = _fields(lines[i])
            ic.Nset_name = ss[0]
            nodes = _resolve_nodes(ic.Nset_name, first_only=True)
            direction = int(ss[1])
            ic.dof.append(nodes * 3 - (3 - direction))
            ic.value.append(float(ss[2]))
        ICs.append(ic)

    # --- Contact ---  (:1046-1102)
    contact_flag = 0
    for i in star:
        if "*Contact" in lines[i]:
            contact_flag = 1
            break
    for i in star:
        if "*Contact Inclusions" in lines[i] and "HAKAIoption=self-contact" in lines[i]:
            contact_flag = 2
            break
    CPs: List[CP] = []
    for idx in [i for i in star if "*Contact Pair," in lines[i]]:
        cp = CP()
        cp.name = _after(_fields(lines[idx])[3], "cpset=")
        ss = _fields(lines[idx + 1])
        cp.surface_name_1 = ss[0]
        cp.surface_name_2 = ss[1]
        for sf in SURFACE:
            if cp.surface_name_1 == sf.name:
                cp.instance_id_1 = sf.instance_id
                cp.elements_1 = sf.elements
            if cp.surface_name_2 == sf.name:
                cp.instance_id_2 = sf.instance_id
                cp.elements_2 = sf.elements
        CPs.append(cp)

    return Model(PART, INSTANCE, NSET, ELSET, SURFACE, AMPLITUDE, MATERIAL, BCs, ICs, CPs,
                 nNode, coordmat, nElement, elementmat, element_material, element_instance,
                 d_time, end_time, mass_scaling, contact_flag)
