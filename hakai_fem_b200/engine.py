"""ctypes binding of the C ABI in include/hakai_b200.h (the Python twin of the Julia `ccall` stub).

`Engine` binds hakai_fem_b200/libhakai_b200.so — the CUDA engine.  There is no CPU fallback:
if the library is missing or no CUDA device exists the constructor raises.  The generic
`EngineBase(lib, prefix)` exists so that tests can drive the CPU oracle (prefix ``hko_``)
through the very same calls; the product never loads the oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhakai_b200.so")

c_i64 = C.c_int64
c_f64 = C.c_double
P_i64 = C.POINTER(C.c_int64)
P_f64 = C.POINTER(C.c_double)


class HkParams(C.Structure):
    """struct hk_params (include/hakai_b200.h)."""
    _fields_ = [
        ("struct_size", C.c_int32), ("device", C.c_int32),
        ("d_time", c_f64), ("element_min_size", c_f64), ("element_max_size", c_f64),
        ("contact_flag", C.c_int32), ("triax_route", C.c_int32),
        ("contact_d_lim_factor", c_f64), ("contact_myu", c_f64),
        ("contact_kc_other", c_f64), ("contact_kc_self", c_f64),
        ("contact_cr_other", c_f64), ("contact_cr_self", c_f64),
        ("contact_ddiv_other", c_f64), ("contact_ddiv_self", c_f64),
        ("deterministic", C.c_int32), ("element_mode", C.c_int32),
        ("contact_dmax_clamp", C.c_int32), ("reserved0", C.c_int32),
    ]


EXPORTS = [
    "default_params", "create", "destroy", "last_error", "set_mesh", "add_material", "add_bc", "add_ic",
    "add_instance", "add_contact_pair", "finalize", "step", "download", "download_ex", "upload_state",
    "deleted_ids", "contact_pair_info", "counters", "profile", "profile_read", "set_stream",
    "set_halo", "halo_bind", "halo_pack", "step_enqueue", "sync", "step_begin", "step_finish",
    "set_node_list", "nodes_export", "nodes_import", "contact_enqueue", "contact_export", "contact_import",
    "set_global_maps", "apply_deleted", "node_output", "mark_frame", "contact_export_limbs", "contact_import_limbs",
    "state_export", "state_import", "state_summary", "deleted_steps", "set_halo_ranks", "comm_unique_id", "comm_init", "profile_read_ex", "comm_contact", "build_contact", "comm_erosion",
]


class HakaiError(RuntimeError):
    pass


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def _pf(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(P_f64)


def _pi(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(P_i64)


def _csr(lists: Sequence[np.ndarray]):
    ptr = np.zeros(len(lists) + 1, np.int64)
    for j, l in enumerate(lists):
        ptr[j + 1] = ptr[j] + len(l)
    flat = _i64(np.concatenate([_i64(l) for l in lists])) if len(lists) else np.zeros(0, np.int64)
    return ptr, flat


class EngineBase:
    builds_contact = True      # hk_build_contact: contact tables built by the engine (the CPU oracle takes host-built ones)
    """Thin, stateful wrapper: one instance = one hk_engine*."""

    def __init__(self, lib: C.CDLL, prefix: str, **params):
        self._lib = lib
        self._pfx = prefix
        self._h = C.c_void_p()
        self._fn("last_error").restype = C.c_char_p
        self._fn("last_error").argtypes = [C.c_void_p]
        self.params = HkParams()
        self._chk(self._fn("default_params")(C.byref(self.params)), created=False)
        for k, v in params.items():
            if not hasattr(self.params, k):
                raise AttributeError(f"hk_params has no field {k}")
            setattr(self.params, k, v)
        rc = self._fn("create")(C.byref(self._h), C.byref(self.params))
        self._chk(rc, created=False)
        self.nNode = 0
        self.nElement = 0

    # -- plumbing ------------------------------------------------------------------
    def _fn(self, name):
        return getattr(self._lib, self._pfx + name)

    def _chk(self, rc: int, created: bool = True):
        if rc != 0:
            msg = self._fn("last_error")(self._h if created else None)
            raise HakaiError(f"{self._pfx}* call failed (code {rc}): {msg.decode() if msg else ''}")

    def close(self):
        if self._h:
            self._fn("destroy")(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- set-up ---------------------------------------------------------------------
    def set_mesh(self, coordmat, elementmat, element_material, element_instance, diag_M):
        coordmat = np.asarray(coordmat)
        elementmat = np.asarray(elementmat)
        self.nNode = coordmat.shape[1]
        self.nElement = elementmat.shape[1]
        # (3,nNode) / (8,nElement) column-major == C-order of the transposes
        cm = _f64(coordmat.T)
        em = _i64(elementmat.T)
        emat = _i64(element_material)
        eins = None if element_instance is None else _i64(element_instance)
        dm = _f64(diag_M)
        self._chk(self._fn("set_mesh")(self._h, c_i64(self.nNode), c_i64(self.nElement), _pf(cm), _pi(em),
                                       _pi(emat), _pi(eins), _pf(dm)))

    def add_material(self, young, poisson, density, plastic=None, Hd=None, ductile=None):
        npp = 0 if plastic is None else int(np.asarray(plastic).shape[0])
        nd = 0 if ductile is None else int(np.asarray(ductile).shape[0])
        pl = _f64(np.asarray(plastic).T) if npp else None          # (npp,2) column-major
        hd = _f64(Hd) if npp > 1 else None
        du = _f64(np.asarray(ductile).T) if nd else None
        self._chk(self._fn("add_material")(self._h, c_f64(young), c_f64(poisson), c_f64(density),
                                           c_i64(npp), _pf(pl), _pf(hd), c_i64(nd), _pf(du)))

    def add_bc(self, dof_lists, values, amp_time=None, amp_value=None):
        if len(values) != len(dof_lists):            # the reference raises BoundsError here (e.g. numbered lines + ENCASTRE)
            raise HakaiError(f"add_bc: {len(dof_lists)} dof lists but {len(values)} values")
        ptr, flat = _csr(dof_lists)
        vals = _f64(values)
        n_amp = 0 if amp_time is None else len(amp_time)
        at = _f64(amp_time) if n_amp else None
        av = _f64(amp_value) if n_amp else None
        self._chk(self._fn("add_bc")(self._h, c_i64(len(dof_lists)), _pi(ptr), _pi(flat), _pf(vals),
                                     c_i64(n_amp), _pf(at), _pf(av)))

    def add_ic(self, dof_lists, values):
        if len(values) != len(dof_lists):
            raise HakaiError(f"add_ic: {len(dof_lists)} dof lists but {len(values)} values")
        ptr, flat = _csr(dof_lists)
        vals = _f64(values)
        self._chk(self._fn("add_ic")(self._h, c_i64(len(dof_lists)), _pi(ptr), _pi(flat), _pf(vals)))

    def add_instance(self, node_offset, nNode, element_offset, nElement, surfaces=None, surfaces_eleid=None):
        sf = None if surfaces is None else _i64(np.asarray(surfaces).T)       # (6nE,4) column-major
        se = None if surfaces_eleid is None else _i64(surfaces_eleid)
        self._chk(self._fn("add_instance")(self._h, c_i64(node_offset), c_i64(nNode), c_i64(element_offset),
                                           c_i64(nElement), _pi(sf), _pi(se)))

    def add_contact_pair(self, i_instance, j_instance, c_nodes_i, c_nodes_j, c_triangles, c_triangles_eleid, young):
        ni, nj = _i64(c_nodes_i), _i64(c_nodes_j)
        tri = _i64(np.asarray(c_triangles).reshape(-1, 3).T)                   # (nTri,3) column-major
        te = _i64(c_triangles_eleid)
        self._chk(self._fn("add_contact_pair")(self._h, c_i64(i_instance), c_i64(j_instance), c_i64(len(ni)), _pi(ni),
                                               c_i64(len(nj)), _pi(nj), c_i64(len(te)), _pi(tri), _pi(te),
                                               c_f64(young)))

    def build_contact(self, instances, young, pairs=None):
        """Contact set-up on the device instead of add_instance / add_contact_pair.
        instances: (node_offset, nNode, element_offset, nElement) per instance; young: Young's modulus of each instance's
        material; pairs: None (ALL EXTERIOR) or a list of (inst1, inst2, elements1, elements2) with 1-based instance ids
        and part-local 1-based element ids of the *Surface sets (None / empty: all elements)."""
        a = _i64(np.asarray(instances, np.int64).reshape(-1, 4))
        cols = [np.ascontiguousarray(a[:, k]) for k in range(4)]
        yg = _f64(young)
        if pairs:
            i1, i2 = _i64([p[0] for p in pairs]), _i64([p[1] for p in pairs])
            p1, e1 = _csr([np.zeros(0, np.int64) if p[2] is None else p[2] for p in pairs])
            p2, e2 = _csr([np.zeros(0, np.int64) if p[3] is None else p[3] for p in pairs])
            args = (c_i64(len(pairs)), _pi(i1), _pi(i2), _pi(p1), _pi(e1), _pi(p2), _pi(e2))
        else:
            args = (c_i64(0), None, None, None, None, None, None)
        self._chk(self._fn("build_contact")(self._h, c_i64(len(a)), _pi(cols[0]), _pi(cols[1]), _pi(cols[2]), _pi(cols[3]),
                                             _pf(yg), *args))

    def finalize(self):
        self._chk(self._fn("finalize")(self._h))

    # -- stepping ---------------------------------------------------------------------
    def step(self, t_first: int, n_steps: int = 1) -> int:
        nd = c_i64(0)
        self._chk(self._fn("step")(self._h, c_i64(t_first), c_i64(n_steps), C.byref(nd)))
        return nd.value

    def step_enqueue(self, t_first: int, n_steps: int = 1):
        self._chk(self._fn("step_enqueue")(self._h, c_i64(t_first), c_i64(n_steps)))

    def step_begin(self, t: int):
        self._chk(self._fn("step_begin")(self._h, c_i64(t)))

    def step_finish(self, t: int):
        self._chk(self._fn("step_finish")(self._h, c_i64(t)))

    def sync(self) -> int:
        nd = c_i64(0)
        self._chk(self._fn("sync")(self._h, C.byref(nd)))
        return nd.value

    # -- taps ---------------------------------------------------------------------------
    def download(self, fields=("disp", "velo", "integ_stress", "integ_strain", "integ_eq_plastic_strain",
                               "integ_triax_stress", "element_flag"), out=None):
        """Returns a dict of arrays in the reference's shapes: (fn,), (6,nip) column-major -> returned as
        numpy (nip,6) C-order viewed transposed, i.e. out['integ_stress'][c, ip]."""
        nN, nE = self.nNode, self.nElement
        fn, nip = 3 * nN, 8 * nE
        shapes = dict(disp=(fn,), velo=(fn,), integ_stress=(nip, 6), integ_strain=(nip, 6),
                      integ_eq_plastic_strain=(nip,), integ_triax_stress=(nip,), element_flag=(nE,))
        bufs = {}
        for k in shapes:
            if k in fields:
                if out is not None and k in out:
                    bufs[k] = out[k]
                else:
                    bufs[k] = np.empty(shapes[k], np.int64 if k == "element_flag" else np.float64)
            else:
                bufs[k] = None
        self._chk(self._fn("download")(self._h, _pf(bufs["disp"]), _pf(bufs["velo"]), _pf(bufs["integ_stress"]),
                                       _pf(bufs["integ_strain"]), _pf(bufs["integ_eq_plastic_strain"]),
                                       _pf(bufs["integ_triax_stress"]), _pi(bufs["element_flag"])))
        res = {}
        for k, v in bufs.items():
            if v is None:
                continue
            res[k] = v.T if k in ("integ_stress", "integ_strain") else v
        return res

    def download_ex(self, fields=("disp_pre", "Q", "external_force", "position", "integ_yield_stress",
                                  "elementVolume")):
        nN, nE = self.nNode, self.nElement
        fn, nip = 3 * nN, 8 * nE
        shapes = dict(disp_pre=(fn,), Q=(fn,), external_force=(fn,), position=(nN, 3),
                      integ_yield_stress=(nip,), elementVolume=(nE,))
        bufs = {k: (np.empty(s) if k in fields else None) for k, s in shapes.items()}
        self._chk(self._fn("download_ex")(self._h, _pf(bufs["disp_pre"]), _pf(bufs["Q"]), _pf(bufs["external_force"]),
                                          _pf(bufs["position"]), _pf(bufs["integ_yield_stress"]),
                                          _pf(bufs["elementVolume"])))
        res = {k: v for k, v in bufs.items() if v is not None}
        if "position" in res:
            res["position"] = res["position"].T
        return res

    def upload_state(self, disp=None, disp_pre=None, velo=None, Q=None, integ_stress=None, integ_strain=None,
                     integ_eq_plastic_strain=None, integ_yield_stress=None, element_flag=None):
        def f(a, mat=False):
            if a is None:
                return None
            a = np.asarray(a)
            return _f64(a.T) if mat else _f64(a)
        a = [f(disp), f(disp_pre), f(velo), f(Q), f(integ_stress, True), f(integ_strain, True),
             f(integ_eq_plastic_strain), f(integ_yield_stress)]
        fl = None if element_flag is None else _i64(element_flag)
        self._chk(self._fn("upload_state")(self._h, *[_pf(x) for x in a], _pi(fl)))

    def node_output(self, raw: bool = False, out=None):
        """cal_node_stress_strain (J2:3408-3486) on the device; same dict as host.cal_node_stress_strain
        (+ inc_num).  raw=True: undivided sums, no von Mises (partitioned meshes).  `out`: optional dict of
        caller-owned C-contiguous float64 buffers (e.g. pinned) keyed like the result; node_stress / node_strain
        buffers have shape (6, nNode) — Julia's (nNode,6) column-major — and come back as (nNode,6) views."""
        nN = self.nNode
        out = out or {}

        def buf(key, shape):
            a = out.get(key)
            if a is None:
                return np.zeros(shape)
            if a.dtype != np.float64 or a.shape != shape or not a.flags.c_contiguous:
                raise ValueError(f"node_output: out[{key!r}] must be C-contiguous float64 {shape}")
            return a
        ns, ne = buf("node_stress", (6, nN)), buf("node_strain", (6, nN))
        ep, tx, inc = buf("node_eq_plastic_strain", (nN,)), buf("node_triax_stress", (nN,)), buf("inc_num", (nN,))
        mi = None if raw else buf("node_mises_stress", (nN,))
        self._chk(self._fn("node_output")(self._h, _pf(ns), _pf(ne), _pf(ep), _pf(mi), _pf(tx), _pf(inc),
                                          C.c_int32(1 if raw else 0)))
        copy = (lambda a: a) if out else np.ascontiguousarray
        res = dict(node_stress=copy(ns.T), node_strain=copy(ne.T), node_eq_plastic_strain=ep, node_triax_stress=tx,
                   inc_num=inc)
        if not raw:
            res["node_mises_stress"] = mi
        return res

    def deleted_ids(self) -> np.ndarray:
        n = c_i64(0)
        self._chk(self._fn("deleted_ids")(self._h, None, c_i64(0), C.byref(n)))
        ids = np.zeros(n.value, np.int64)
        if n.value:
            self._chk(self._fn("deleted_ids")(self._h, _pi(ids), c_i64(n.value), C.byref(n)))
        return ids

    def deleted_steps(self) -> np.ndarray:
        """Step in which each entry of deleted_ids() was deleted (0: replayed through apply_deleted)."""
        n = c_i64(0)
        self._chk(self._fn("deleted_steps")(self._h, None, c_i64(0), C.byref(n)))
        st = np.zeros(n.value, np.int64)
        if n.value:
            self._chk(self._fn("deleted_steps")(self._h, _pi(st), c_i64(n.value), C.byref(n)))
        return st

    def contact_pair(self, c: int):
        a, b, t = c_i64(0), c_i64(0), c_i64(0)
        self._chk(self._fn("contact_pair_info")(self._h, c_i64(c), C.byref(a), C.byref(b), C.byref(t),
                                                None, None, None, None))
        ni, nj = np.zeros(a.value, np.int64), np.zeros(b.value, np.int64)
        tri, te = np.zeros((3, t.value), np.int64), np.zeros(t.value, np.int64)
        self._chk(self._fn("contact_pair_info")(self._h, c_i64(c), None, None, None, _pi(ni), _pi(nj), _pi(tri),
                                                _pi(te)))
        return dict(c_nodes_i=ni, c_nodes_j=nj, c_triangles=tri.T.copy(), c_triangles_eleid=te)

    def counters(self) -> np.ndarray:
        out = np.zeros(8, np.int64)
        self._chk(self._fn("counters")(self._h, _pi(out)))
        return out

    def state_summary(self) -> dict:
        """Device-side reduction: live elements, min / max eq. plastic strain of live Gauss points, yielded points."""
        out = np.zeros(8)
        self._chk(self._fn("state_summary")(self._h, _pf(out)))
        return dict(live_elements=int(out[0]), eps_min=float(out[1]), eps_max=float(out[2]), yielded_points=int(out[3]))

    def profile(self, enable: bool = True):
        self._chk(self._fn("profile")(self._h, C.c_int32(1 if enable else 0)))

    def profile_read(self):
        ms = np.zeros(4)
        n = np.zeros(4, np.int64)
        self._chk(self._fn("profile_read")(self._h, _pf(ms), _pi(n)))
        return ms, n

    def profile_read_ex(self):
        """profile_read plus kinds 4 (halo exchange on the engine's side stream) and 5 (deletion pass)."""
        ms = np.zeros(8)
        n = np.zeros(8, np.int64)
        self._chk(self._fn("profile_read_ex")(self._h, _pf(ms), _pi(n)))
        return ms, n

    # -- multi-GPU halo ---------------------------------------------------------------
    def set_halo(self, node_lists):
        """node_lists[i]: local 1-based ids of the nodes shared with neighbour i (call before finalize)."""
        ptr, flat = _csr(node_lists)
        self._chk(self._fn("set_halo")(self._h, c_i64(len(node_lists)), _pi(ptr), _pi(flat)))

    def set_halo_ranks(self, my_rank: int, ranks):
        """Global rank of this engine and of each set_halo neighbour: rank-ordered (holder-count independent) sums."""
        a = _i64(ranks)
        self._chk(self._fn("set_halo_ranks")(self._h, c_i64(my_rank), c_i64(len(a)), _pi(a)))

    def comm_unique_id(self) -> bytes:
        """128-byte ncclUniqueId (call on one rank, broadcast the bytes)."""
        buf = C.create_string_buffer(128)
        self._chk(self._fn("comm_unique_id")(buf), created=False)
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        """The engine creates its own NCCL communicator; step_enqueue(t, n) then runs n multi-GPU steps by itself."""
        if len(unique_id) != 128:
            raise ValueError("unique_id: 128 bytes")
        self._chk(self._fn("comm_init")(self._h, C.c_char_p(unique_id), C.c_int32(rank), C.c_int32(world)))

    def comm_erosion(self, max_deleted_per_step: int = 4096):
        """Deletions of all ranks replayed on the device (hk_comm_erosion): after set_global_maps, before the first step;
        the node lists 0 / 1 / 2 must then cover every candidate surface node and never change."""
        self._chk(self._fn("comm_erosion")(self._h, C.c_int32(max_deleted_per_step)))

    def comm_contact(self, maxlen: int, src_index):
        """The engine runs the contact exchange of every step itself (all-gather of surface-node states, exact all-reduce
        of the force accumulators) — after set_node_list(0/1/2) and comm_init."""
        a = _i64(src_index)
        self._chk(self._fn("comm_contact")(self._h, c_i64(maxlen), _pi(a)))

    def halo_bind(self, neighbor: int, send_ptr: int, recv_ptr: int):
        self._chk(self._fn("halo_bind")(self._h, c_i64(neighbor), C.c_void_p(send_ptr), C.c_void_p(recv_ptr)))

    def halo_pack(self):
        self._chk(self._fn("halo_pack")(self._h))

    # -- multi-GPU contact ---------------------------------------------------------------
    def set_node_list(self, which: int, nodes):
        a = _i64(nodes)
        self._chk(self._fn("set_node_list")(self._h, C.c_int32(which), c_i64(len(a)), _pi(a)))

    def nodes_export(self, out_ptr: int):
        self._chk(self._fn("nodes_export")(self._h, C.c_void_p(out_ptr)))

    def nodes_import(self, in_ptr: int, src_index=None):
        a = None if src_index is None else _i64(src_index)
        self._chk(self._fn("nodes_import")(self._h, C.c_void_p(in_ptr), _pi(a)))

    def contact_enqueue(self):
        self._chk(self._fn("contact_enqueue")(self._h))

    def contact_export(self, out_ptr: int):
        self._chk(self._fn("contact_export")(self._h, C.c_void_p(out_ptr)))

    def contact_import(self, in_ptr: int, n_ranks: int):
        self._chk(self._fn("contact_import")(self._h, C.c_void_p(in_ptr), c_i64(n_ranks)))

    def state_export(self, out_ptr: int):
        self._chk(self._fn("state_export")(self._h, C.c_void_p(out_ptr)))

    def state_import(self, in_ptr: int):
        self._chk(self._fn("state_import")(self._h, C.c_void_p(in_ptr)))

    def contact_export_limbs(self, out_ptr: int):
        self._chk(self._fn("contact_export_limbs")(self._h, C.c_void_p(out_ptr)))

    def contact_import_limbs(self, in_ptr: int):
        self._chk(self._fn("contact_import_limbs")(self._h, C.c_void_p(in_ptr)))

    def mark_frame(self):
        """The next asynchronous step is followed by an output frame (stores integ_triax_stress)."""
        self._chk(self._fn("mark_frame")(self._h))

    def set_global_maps(self, node_map, elem_map, element_instance):
        nm, em, ei = _i64(node_map), _i64(elem_map), _i64(element_instance)
        self._chk(self._fn("set_global_maps")(self._h, c_i64(len(nm)), _pi(nm), c_i64(len(em)), _pi(em), _pi(ei)))

    def apply_deleted(self, global_ids):
        a = _i64(global_ids)
        self._chk(self._fn("apply_deleted")(self._h, c_i64(len(a)), _pi(a)))

    def set_stream(self, stream_ptr: int):
        self._chk(self._fn("set_stream")(self._h, C.c_void_p(stream_ptr)))


_lib_cache = {}


def load_library(path: str = LIB_PATH) -> C.CDLL:
    """Loads libhakai_b200.so; raises (never falls back) when it is missing."""
    if path not in _lib_cache:
        if not os.path.exists(path):
            raise HakaiError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)")
        _lib_cache[path] = C.CDLL(path)
    return _lib_cache[path]


class Engine(EngineBase):
    """The CUDA engine (libhakai_b200.so)."""

    def __init__(self, **params):
        super().__init__(load_library(), "hk_", **params)
