"""Multi-GPU driver: element-block (slab) partition, shared-node force halos over torch.distributed.

One process per GPU.  Every rank owns a contiguous block of elements and ALL nodes those
elements touch; nodes on a partition interface are duplicated.  Per time step
    engine.halo_pack()                      partial internal force of the interface nodes -> send buffers
    batch_isend_irecv with the neighbours   NCCL send/recv over NVLink (gloo in the CPU tests)
    engine.step(t, 1)                       adds the received partials, then the usual nodal + element kernels
Both sides of an interface then update the shared nodes redundantly from bit-identical inputs
(a + b == b + a), so positions never need to be exchanged.  Fracture is rank-local.

Contact across ranks (ContactExchanger): every rank tests the GLOBAL slave-node lists against the master
triangles of its own elements; surface-node states travel by all-gather, the fixed-point force accumulators by
an exact integer all-gather + sum.  When elements can be deleted (ductile material + contact), the faces they
expose join the contact surface on every rank (ErosionMaps, hk_apply_deleted): one host synchronisation and one
small all-gather of the fresh deleted ids per step.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import inp as I
from .model_setup import Setup, configure_engine


@dataclass
class LocalDomain:
    rank: int
    setup: Setup                       # local model + local (already halo-summed) diag_M
    node_l2g: np.ndarray               # local -> global node id (1-based values)
    elem_l2g: np.ndarray
    neighbors: List[int] = field(default_factory=list)
    halo_nodes: List[np.ndarray] = field(default_factory=list)     # local 1-based ids per neighbour
    contact: object = None             # ContactLists when the model has contact
    contact_all: object = None         # ContactLists over every node that may ever join a surface (device-side erosion)


def _restrict_dofs(dof_lists, values, g2l):
    out_d, out_v = [], []
    for dof, v in zip(dof_lists, values):
        node = (dof - 1) // 3 + 1
        comp = (dof - 1) % 3
        loc = g2l[node]
        keep = loc > 0
        out_d.append(((loc[keep] - 1) * 3 + comp[keep] + 1).astype(np.int64))
        out_v.append(v)
    return out_d, out_v


@dataclass
class ErosionMaps:
    """What a rank needs to replay add_surface_triangle (J2:2167-2245) for elements deleted on ANY rank."""
    node_map: np.ndarray               # global node (0-based index) -> local 1-based id, 0 = not on this rank
    elem_map: np.ndarray               # global element -> local 1-based id, 0 = owned by another rank
    element_instance: np.ndarray       # instance of every global element
    owner: np.ndarray                  # global node id (1-based index, [0] unused) -> owning rank
    n_held: int                        # local ids 1..n_held are nodes of own elements, the rest are ghosts


@dataclass
class ContactLists:
    """Per-rank lists of the contact-surface all-gather (local 1-based node ids)."""
    export_nodes: np.ndarray           # surface nodes this rank owns (ascending global id)
    import_nodes: np.ndarray           # ghost copies of surface nodes owned elsewhere
    import_src: np.ndarray             # index of each ghost's record in the gathered buffer (rank*maxlen + position)
    surface_nodes: np.ndarray          # ALL surface nodes in global order (same order on every rank)
    maxlen: int                        # padded length of one rank's export block
    erosion: ErosionMaps = None        # set when deletions can expose new faces


def build_contact_lists(surf, owner, g2l, n_held, rank, n_ranks, erosion=None) -> ContactLists:
    """`surf`: ascending global ids of all current contact-surface nodes; `owner[g]`: rank whose copy is
    authoritative; `g2l`: global -> local 1-based (0 = absent).  Identical `surf` on every rank."""
    own = owner[surf]
    export_lists = [surf[own == r] for r in range(n_ranks)]
    maxlen = max(1, max(len(x) for x in export_lists))
    loc = g2l[surf]
    if np.any(loc == 0):
        raise ValueError("contact-surface node missing on rank %d (ghost set too small)" % rank)
    ghost = surf[loc > n_held]                                     # copies no local element updates
    g_owner = owner[ghost]
    src = np.zeros(len(ghost), np.int64)
    for o in range(n_ranks):
        sel = g_owner == o
        src[sel] = o * maxlen + np.searchsorted(export_lists[o], ghost[sel])
    return ContactLists(g2l[export_lists[rank]], g2l[ghost], src, loc, maxlen, erosion)


def _contact_candidates(setup: Setup):
    """(surface nodes now, nodes that may ever join the surface, erosion possible?) — global 1-based ids."""
    m = setup.model
    surf = np.zeros(0, np.int64)
    if not m.contact_flag:
        return surf, surf, False
    surf = np.unique(np.concatenate([np.concatenate([ct.c_nodes_i, ct.c_nodes_j, ct.c_triangles.reshape(-1)])
                                     for ct in setup.CT]))
    candidates = surf
    erosion = (any(mat.ductile.shape[0] > 0 for mat in m.MATERIAL)
               and any(len(ins.surfaces) > 0 for ins in m.INSTANCE))
    if erosion:
        inst = sorted({i for ct in setup.CT for i in (ct.i_instance, ct.j_instance)})
        candidates = np.unique(np.concatenate(
            [surf] + [np.arange(m.INSTANCE[i - 1].node_offset + 1,
                                m.INSTANCE[i - 1].node_offset + m.INSTANCE[i - 1].nNode + 1) for i in inst]))
    return surf, candidates, erosion


def partition_model(setup: Setup, n_ranks: int, only_rank: int = None) -> List[LocalDomain]:
    """Splits a (small) global model into contiguous element blocks.  Used by the tests and for general
    decks; the 16 M/GPU bench builds each slab directly (slab_deck) without materialising the global mesh.

    With contact, every rank receives the GLOBAL contact node lists (nodes it does not hold are appended as ghost
    nodes that no element references) and the master triangles of its own elements.  If a material can fail, the
    ghost set is every node of the instances in contact (any of them may become exposed) and the rank gets the
    global instance face tables + ErosionMaps.  only_rank: build just that rank's LocalDomain (the other list
    entries are None) — what a rank of a distributed run needs."""
    m = setup.model
    nE = m.nElement
    bounds = [(nE * r) // n_ranks for r in range(n_ranks + 1)]
    # holds[r, n] = rank r holds node n (bit matrix: n_ranks x nNode+1); owner = lowest rank holding it
    holds = np.zeros((n_ranks, m.nNode + 1), bool)
    doms = []
    locals_ = []
    for r in range(n_ranks):
        el = np.arange(bounds[r], bounds[r + 1])
        nodes = np.unique(m.elementmat[:, el])                    # ascending global ids
        holds[r, nodes] = True
        locals_.append((el, nodes))
    first_holder = np.argmax(holds, axis=0)
    surf, candidates, erosion = _contact_candidates(setup)
    for r in range(n_ranks):
        if only_rank is not None and r != only_rank:
            doms.append(None)
            continue
        el, nodes_own = locals_[r]
        ghosts = np.setdiff1d(candidates, nodes_own) if m.contact_flag else np.zeros(0, np.int64)
        nodes = np.concatenate([nodes_own, ghosts])              # local numbering: own nodes, then ghosts
        g2l = np.zeros(m.nNode + 1, np.int64)
        g2l[nodes] = np.arange(1, len(nodes) + 1)
        em = g2l[m.elementmat[:, el]]
        coord = m.coordmat[:, nodes - 1]
        lm = copy.copy(m)
        lm.nNode, lm.nElement = len(nodes), len(el)
        lm.coordmat, lm.elementmat = np.ascontiguousarray(coord), np.ascontiguousarray(em)
        lm.element_material = m.element_material[el]
        lm.element_instance = np.ones(len(el), np.int64)
        lm.BC, lm.IC = [], []
        for bc in m.BC:
            nb = I.BC(Nset_name=bc.Nset_name, amp_name=bc.amp_name, amplitude=bc.amplitude)
            nb.dof, nb.value = _restrict_dofs(bc.dof, bc.value, g2l)
            lm.BC.append(nb)
        for ic in m.IC:
            ni = I.IC(Nset_name=ic.Nset_name, type=ic.type)
            ni.dof, ni.value = _restrict_dofs(ic.dof, ic.value, g2l)
            lm.IC.append(ni)
        lst = Setup(lm, setup.d_time, setup.time_num, (None if setup.elementVolume is None else setup.elementVolume[el]),
                    np.repeat(setup.diag_M.reshape(-1, 3)[nodes - 1, 0], 3),     # global (summed) mass
                    setup.elementMinSize, setup.elementMaxSize)
        dom = LocalDomain(r, lst, nodes, el + 1)
        if m.contact_flag:
            from .model_setup import ContactTriangle
            e_g2l = np.zeros(nE + 1, np.int64)
            e_g2l[el + 1] = np.arange(1, len(el) + 1)
            maps = None
            if erosion:                                           # GLOBAL face tables (copied: the engine mutates its own)
                lm.INSTANCE = [copy.copy(ins) for ins in m.INSTANCE]
                maps = ErosionMaps(g2l[1:].copy(), e_g2l[1:].copy(), np.asarray(m.element_instance, np.int64),
                                   first_holder.astype(np.int64), len(nodes_own))
            else:
                lm.INSTANCE = []
            for ct in setup.CT:
                mine = e_g2l[ct.c_triangles_eleid] > 0
                lst.CT.append(ContactTriangle(ct.i_instance, ct.j_instance, g2l[ct.c_nodes_i], g2l[ct.c_nodes_j],
                                              g2l[ct.c_triangles[mine]], e_g2l[ct.c_triangles_eleid[mine]], ct.young))
            dom.contact = build_contact_lists(surf, first_holder, g2l, len(nodes_own), r, n_ranks, maps)
            if erosion:
                dom.contact_all = build_contact_lists(candidates, first_holder, g2l, len(nodes_own), r, n_ranks, maps)
        for q in range(n_ranks):
            if q == r:
                continue
            shared = nodes_own[holds[q, nodes_own]]
            if len(shared):
                dom.neighbors.append(q)
                dom.halo_nodes.append(g2l[shared])
        doms.append(dom)
    return doms


def slab_deck(deck, rank: int, world: int):
    """Rank-local slab of the weak-scaling deck W: `deck` describes ONE GPU's block (nx,ny,nz); the global
    mesh is nx x ny x (nz*world), split in z.  Returns the local StretchDeck (global loading) and the
    interface layers (local 1-based node ids) shared with rank-1 / rank+1."""
    local = copy.copy(deck)
    local.layer_offset = rank * deck.nz
    local.global_nz = deck.nz * world
    local.jitter_by_layer = True          # noise is a function of the GLOBAL node layer: the slabs tile one global mesh
    per = (deck.nx + 1) * (deck.ny + 1)
    nbrs, halos = [], []
    if rank > 0:
        nbrs.append(rank - 1)
        halos.append(np.arange(1, per + 1, dtype=np.int64))
    if rank < world - 1:
        nbrs.append(rank + 1)
        halos.append(np.arange(deck.nz * per + 1, (deck.nz + 1) * per + 1, dtype=np.int64))
    return local, nbrs, halos


def slab_parity_check(make_engine, torch_device, rank: int, world: int, n_steps: int = 32, nx: int = 24, ny: int = 24,
                      nz_per_rank: int = 8, engine_comm: bool = False, **params):
    """Checks the z-slab / force-halo path against an unpartitioned run of the SAME mesh and the same kernels.

    A small GLOBAL deck (nx x ny x nz_per_rank*world ductile block, jitter seeded per global node layer, stretched fast
    enough that elements delete within `n_steps`) is stepped by `world` ranks through slab_deck + SlabRunner (pack ->
    send/recv -> split step); rank 0 also runs the whole mesh on one engine.  Returns on rank 0 (None elsewhere):
    max relative field error, whether the deleted-element sets and the element flags are identical, and whether the
    node layer shared by two ranks is bit-identical on both.  bench.py runs it at N > 1 before anything is timed."""
    import torch.distributed as dist
    from .mesh import StretchDeck, steel
    from .model_setup import prepare
    rows = [[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]
    mk = lambda nz: StretchDeck(nx, ny, nz, material=steel("steel_Ductile", ductile=rows), jitter=0.05,
                                strain_per_step=1.0e-3, jitter_by_layer=True)
    fields = ("disp", "integ_eq_plastic_strain", "integ_stress", "element_flag")
    deck, nbrs, halos = slab_deck(mk(nz_per_rank), rank, world)
    runner = SlabRunner(make_engine, prepare(deck.build_model()), nbrs, halos, torch_device, sum_mass=True, rank=rank,
                        engine_comm=engine_comm, **params)
    runner.run(1, n_steps)
    d = runner.engine.download(fields=fields)
    mine = dict(disp=d["disp"], eps=d["integ_eq_plastic_strain"], stress=np.ascontiguousarray(d["integ_stress"]),
                flag=d["element_flag"], deleted=np.sort(runner.engine.deleted_ids()))
    runner.engine.close()
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    if rank != 0:
        return None
    g = configure_engine(make_engine, prepare(mk(nz_per_rank * world).build_model()), **params)
    g.step(1, n_steps)
    ref = g.download(fields=fields)
    ref_del = np.sort(g.deleted_ids())
    g.close()
    per, nEl = (nx + 1) * (ny + 1), nx * ny * nz_per_rank
    err, flags_equal, bitwise, got_del = 0.0, True, True, []
    rs = np.asarray(ref["integ_stress"])
    su, se, ss = np.abs(ref["disp"]).max(), max(ref["integ_eq_plastic_strain"].max(), 1e-300), max(np.abs(rs).max(), 1e-300)
    for r, p in enumerate(parts):
        n0, e0 = r * nz_per_rank * per, r * nEl
        err = max(err, np.abs(p["disp"] - ref["disp"][3 * n0:3 * (n0 + (nz_per_rank + 1) * per)]).max() / su)
        err = max(err, np.abs(p["eps"] - ref["integ_eq_plastic_strain"][8 * e0:8 * (e0 + nEl)]).max() / se)
        err = max(err, np.abs(p["stress"] - rs[:, 8 * e0:8 * (e0 + nEl)]).max() / ss)
        flags_equal = flags_equal and np.array_equal(p["flag"], ref["element_flag"][e0:e0 + nEl])
        got_del.append(p["deleted"] + e0)
        if r + 1 < world:        # last node layer of rank r == first node layer of rank r+1
            bitwise = bitwise and np.array_equal(p["disp"][3 * nz_per_rank * per:], parts[r + 1]["disp"][:3 * per])
    del_equal = bool(np.array_equal(np.sort(np.concatenate(got_del)), ref_del)) and flags_equal
    return {"n_ranks": world, "deck": f"{nx}x{ny}x{nz_per_rank * world} ductile block, {n_steps} steps, {world} slabs over "
                                      f"the halo exchange vs the same mesh unpartitioned on one engine",
            "exchange": "engine-owned NCCL communicator (hk_comm_init), all steps enqueued in one call" if engine_comm
                        else "host-driven (torch.distributed P2P per step)",
            "max_rel_err": float(err), "deleted_equal": del_equal, "n_deleted": int(len(ref_del)),
            "interface_bitwise": bool(bitwise),
            "ok": bool(err <= 1e-10 and del_equal and bitwise and 0 < len(ref_del))}


class HaloExchanger:
    """send/recv buffers (torch tensors on the engine's device) + one batch of P2P ops per step.

    engine_comm=True: the ENGINE owns an NCCL communicator (hk_comm_init; the 128-byte id is created by rank 0 through
    the library and broadcast here) and its own exchange blocks; `hk_step_enqueue(t, n)` then packs, sends, receives
    and steps n times without Python in the loop."""

    def __init__(self, engine, neighbors, halo_nodes, device, rank=None, engine_comm=False):
        import torch
        self.torch = torch
        self.engine = engine
        self.neighbors = list(neighbors)
        self.engine_comm = False
        self._bytes = sum(3 * len(h) * 8 for h in halo_nodes) * 2
        if rank is not None and self.neighbors:
            engine.set_halo_ranks(rank, self.neighbors)          # rank-ordered sums: any number of holders per node
        if engine_comm:
            import torch.distributed as dist
            world = dist.get_world_size()
            me = dist.get_rank() if rank is None else rank
            idt = torch.zeros(128, dtype=torch.uint8, device=device)
            if me == 0:
                idt.copy_(torch.frombuffer(bytearray(engine.comm_unique_id()), dtype=torch.uint8))
            dist.broadcast(idt, 0)
            engine.comm_init(bytes(idt.cpu().numpy().tobytes()), me, world)
            self.engine_comm = True
            self.send = self.recv = []
            return
        self.send = [torch.zeros(3 * len(h), dtype=torch.float64, device=device) for h in halo_nodes]
        self.recv = [torch.zeros(3 * len(h), dtype=torch.float64, device=device) for h in halo_nodes]
        for i in range(len(self.neighbors)):
            engine.halo_bind(i, self.send[i].data_ptr(), self.recv[i].data_ptr())

    def start(self):
        """pack + post the sends/receives; returns the requests to wait on (stream-ordered for NCCL)."""
        import torch.distributed as dist
        if not self.neighbors:
            return []
        self.engine.halo_pack()
        ops = []
        for i, nb in enumerate(self.neighbors):
            ops.append(dist.P2POp(dist.isend, self.send[i], nb))
            ops.append(dist.P2POp(dist.irecv, self.recv[i], nb))
        return dist.batch_isend_irecv(ops)

    @staticmethod
    def wait(reqs):
        for req in reqs:
            req.wait()

    def exchange(self):
        self.wait(self.start())

    @property
    def bytes_per_step(self):
        return self._bytes


def exchange_sum(values_per_nbr, neighbors, device):
    """Sum of per-neighbour arrays with the neighbours' counterparts (used once, for the lumped mass)."""
    import torch
    import torch.distributed as dist
    send = [torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64, device=device) for v in values_per_nbr]
    recv = [torch.zeros_like(s) for s in send]
    ops = []
    for i, nb in enumerate(neighbors):
        ops.append(dist.P2POp(dist.isend, send[i], nb))
        ops.append(dist.P2POp(dist.irecv, recv[i], nb))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return [r.cpu().numpy() for r in recv]


class ContactExchanger:
    """All-gather of contact-surface node {position, velocity} and of the fixed-point force accumulators."""

    def __init__(self, engine, lists: ContactLists, world: int, device, rank=None, node_l2g=None, n_pairs=0,
                 force_exchange="allreduce", engine_comm=False, static_lists: ContactLists = None,
                 max_deleted_per_step: int = 4096):
        """force_exchange: "allgather" (6 x u64 per surface node from every rank, summed on import) or "allreduce"
        (three 43-bit limbs per accumulator in int64 lanes, one SUM all-reduce: world/1.5 times fewer bytes).
        static_lists (surfaces that erode): exchange lists over EVERY candidate surface node, never rebuilt — the engine
        then replays the deletions of all ranks on the device (hk_comm_erosion); with engine_comm it also gathers them,
        so run() needs no host between steps."""
        if force_exchange not in ("allgather", "allreduce"):
            raise ValueError("force_exchange: allgather | allreduce")
        self.force_exchange = force_exchange
        self.engine_comm = engine_comm       # the engine's own communicator runs the exchange inside hk_step_enqueue
        self.engine, self.world, self.device = engine, world, device
        self.rank, self.node_l2g, self.n_pairs = rank, node_l2g, n_pairs
        self.erosion = lists.erosion
        self.device_erosion = self.erosion is not None and static_lists is not None
        if self.erosion is not None:
            if rank is None or node_l2g is None:
                raise ValueError("erosion across ranks needs rank and node_l2g")
            er = self.erosion
            engine.set_global_maps(er.node_map, er.elem_map, er.element_instance)
            self._g2l = np.concatenate([[0], er.node_map])
            self._surf0 = np.asarray(node_l2g)[lists.surface_nodes - 1]
            if self.device_erosion:
                engine.comm_erosion(max_deleted_per_step)
                lists = static_lists
        self.set_lists(lists)

    def set_lists(self, lists: ContactLists):
        import torch
        engine, world, device = self.engine, self.world, self.device
        self.lists = lists
        if self.engine_comm:
            engine.set_node_list(0, lists.export_nodes)
            engine.set_node_list(1, lists.import_nodes)
            engine.set_node_list(2, lists.surface_nodes)
            engine.comm_contact(lists.maxlen, lists.import_src)
            return
        self.send_nodes = torch.zeros(lists.maxlen * 6, dtype=torch.float64, device=device)
        self.all_nodes = torch.zeros(world * lists.maxlen * 6, dtype=torch.float64, device=device)
        n_surf = len(lists.surface_nodes)
        if self.force_exchange == "allreduce":
            self.limbs = torch.zeros(n_surf * 9, dtype=torch.int64, device=device)
        else:
            self.send_acc = torch.zeros(n_surf * 6, dtype=torch.int64, device=device)
            self.all_acc = torch.zeros(world * n_surf * 6, dtype=torch.int64, device=device)
        engine.set_node_list(0, lists.export_nodes)
        engine.set_node_list(1, lists.import_nodes)
        engine.set_node_list(2, lists.surface_nodes)
        self._first = True

    def exchange_deleted(self, fresh_global):
        """All ranks apply the same ascending list of GLOBAL element ids deleted in the step that just ended
        (the reference visits them in element order, J2:767-804), then rebuild the surface lists if they grew."""
        import torch
        import torch.distributed as dist
        n = torch.tensor([len(fresh_global)], dtype=torch.int64, device=self.device)
        counts = [torch.zeros_like(n) for _ in range(self.world)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        if sum(counts) == 0:
            return 0
        cap = max(counts)
        mine = torch.zeros(cap, dtype=torch.int64, device=self.device)
        mine[:len(fresh_global)] = torch.as_tensor(np.asarray(fresh_global, np.int64), device=self.device)
        got = [torch.zeros_like(mine) for _ in range(self.world)]
        dist.all_gather(got, mine)
        ids = np.sort(np.concatenate([g[:c].cpu().numpy() for g, c in zip(got, counts)]))
        self.engine.apply_deleted(ids)
        if self.device_erosion:              # lists cover every candidate node: nothing to rebuild
            return len(ids)
        l2g = np.asarray(self.node_l2g)
        parts = [self._surf0]
        for c in range(self.n_pairs):
            info = self.engine.contact_pair(c)
            parts += [l2g[info["c_nodes_i"] - 1], l2g[info["c_nodes_j"] - 1]]
        surf = np.unique(np.concatenate(parts))
        if len(surf) != len(self.lists.surface_nodes):
            er = self.erosion
            self.set_lists(build_contact_lists(surf, er.owner, self._g2l, er.n_held, self.rank, self.world, er))
        return len(ids)

    def run(self):
        """positions -> ghosts, contact pass on the local triangles, exact sum of the forces over ranks."""
        import torch.distributed as dist
        eng = self.engine
        if self.engine_comm:
            return                          # done by the engine at the start of the step it is about to run
        eng.nodes_export(self.send_nodes.data_ptr())
        dist.all_gather(list(self.all_nodes.chunk(self.world)), self.send_nodes)
        eng.nodes_import(self.all_nodes.data_ptr(), self.lists.import_src if self._first else None)
        self._first = False
        eng.contact_enqueue()
        if self.force_exchange == "allreduce":
            eng.contact_export_limbs(self.limbs.data_ptr())
            dist.all_reduce(self.limbs, op=dist.ReduceOp.SUM)
            eng.contact_import_limbs(self.limbs.data_ptr())
        else:
            eng.contact_export(self.send_acc.data_ptr())
            dist.all_gather(list(self.all_acc.chunk(self.world)), self.send_acc)
            eng.contact_import(self.all_acc.data_ptr(), self.world)


class SlabRunner:
    """Engine + halo exchange of one rank.  `engine_cls` is Engine (CUDA) — the CPU tests pass the
    host-compiled kernel build, whose "device" pointers are host pointers."""

    def __init__(self, engine_cls, setup: Setup, neighbors, halo_nodes, torch_device, sum_mass=False, contact=None,
                 world=1, rank=None, node_l2g=None, elem_l2g=None, force_exchange="allreduce", engine_comm=False,
                 contact_all=None, **params):
        self.setup = setup
        if sum_mass and neighbors:
            # interface nodes: add the neighbour's partial lumped mass (J2:201-215 summed over ALL elements)
            m = setup.diag_M.reshape(-1, 3)
            parts = [m[h - 1, 0].copy() for h in halo_nodes]
            got = exchange_sum(parts, neighbors, torch_device)
            for h, g in zip(halo_nodes, got):
                m[h - 1, :] += g[:, None]
        model = setup.model

        def with_halo(**p):
            eng = engine_cls(**p)
            if neighbors:
                eng.set_halo(halo_nodes)
            return eng
        self.engine = configure_engine(with_halo, setup, **params)
        self.halo = HaloExchanger(self.engine, neighbors, halo_nodes, torch_device, rank=rank, engine_comm=engine_comm)
        self.contact = (ContactExchanger(self.engine, contact, world, torch_device, rank, node_l2g, len(setup.CT),
                                         force_exchange, engine_comm=engine_comm, static_lists=contact_all)
                        if contact is not None else None)
        self.erosion = self.contact is not None and self.contact.erosion is not None
        # the engine gathers and replays the deletions of all ranks itself: no host between steps
        self.erosion_on_device = self.erosion and engine_comm and self.contact.device_erosion
        if self.erosion and elem_l2g is None:
            raise ValueError("erosion across ranks needs elem_l2g")
        self.elem_l2g = None if elem_l2g is None else np.asarray(elem_l2g, np.int64)
        self.nElement = model.nElement

    @classmethod
    def from_domain(cls, engine_cls, dom: LocalDomain, torch_device, world, force_exchange="allreduce",
                    device_erosion=False, **params):
        """device_erosion: surfaces that erode across ranks are kept current on the device (static exchange lists over all
        candidate nodes, hk_comm_erosion) instead of by a host replay + list rebuild after every step."""
        return cls(engine_cls, dom.setup, dom.neighbors, dom.halo_nodes, torch_device, contact=dom.contact,
                   world=world, rank=dom.rank, node_l2g=dom.node_l2g, elem_l2g=dom.elem_l2g,
                   force_exchange=force_exchange, contact_all=dom.contact_all if device_erosion else None, **params)

    def _after_step(self) -> int:
        """Erosion across ranks: one host sync per step (the deleted set decides the next step's contact surface)."""
        n = self.engine.sync()
        fresh = self.engine.deleted_ids()[-n:] if n else np.zeros(0, np.int64)
        self.contact.exchange_deleted(self.elem_l2g[np.asarray(fresh, np.int64) - 1])
        return n

    def step(self, t: int) -> int:
        if self.halo.engine_comm:
            return self.run(t, 1)
        if self.contact is not None:
            self.contact.run()
        self.halo.exchange()
        n = self.engine.step(t, 1)
        if self.erosion:
            self.contact.exchange_deleted(self.elem_l2g[self.engine.deleted_ids()[-n:] - 1] if n
                                          else np.zeros(0, np.int64))
        return n

    def run(self, t_first: int, n_steps: int, frame_at_end: bool = False, sync: bool = True) -> int:
        """Enqueues pack -> exchange -> step for every step without blocking the host, then synchronises once.
        frame_at_end: an output frame follows (the last step stores integ_triax_stress, hk_mark_frame).
        sync=False (no contact erosion across ranks): only enqueue; the caller calls engine.sync() itself."""
        import time as _time
        _t0 = _time.perf_counter()
        n_del = 0
        if self.halo.engine_comm and (not self.erosion or self.erosion_on_device):
            # one call: the engine packs, exchanges (its own NCCL), steps and keeps eroding surfaces current
            if frame_at_end:
                self.engine.mark_frame()
            self.engine.step_enqueue(t_first, n_steps)
            self.last_enqueue_s = _time.perf_counter() - _t0
            return self.engine.sync() if sync else 0
        if self.halo.engine_comm:                        # eroding contact surfaces: the host replays deletions every step
            for t in range(t_first, t_first + n_steps):
                if frame_at_end and t == t_first + n_steps - 1:
                    self.engine.mark_frame()
                self.engine.step_enqueue(t, 1)
                n_del += self._after_step()
            return n_del
        for t in range(t_first, t_first + n_steps):
            if frame_at_end and t == t_first + n_steps - 1:
                self.engine.mark_frame()
            if self.contact is not None:
                self.contact.run()
            if self.halo.neighbors:
                reqs = self.halo.start()             # partial forces on their way ...
                self.engine.step_begin(t)            # ... while all non-interface nodes are updated
                self.halo.wait(reqs)
                self.engine.step_finish(t)           # interface nodes, element kernel
            else:
                self.engine.step_enqueue(t, 1)
            if self.erosion:
                n_del += self._after_step()
        self.last_enqueue_s = _time.perf_counter() - _t0      # host time to enqueue (diagnostic)
        return n_del + (self.engine.sync() if sync else 0)


# ------------------------------------------------------------------ ghost-element partitions (partition-independent bits)
@dataclass
class GhostDomain:
    """Element block of a rank plus one layer of ghost elements (every element sharing a node with the block)."""
    rank: int
    setup: Setup                       # local model: own + ghost elements in ascending global order, global lumped mass
    node_l2g: np.ndarray               # local -> global node id (1-based values, ascending)
    elem_l2g: np.ndarray               # local -> global element id (1-based values, ascending)
    own_elem: np.ndarray               # bool per local element: in this rank's block (ghost copies are recomputed)
    own_node: np.ndarray               # bool per local node: node of an own element (complete force sum locally)
    neighbors: List[int] = field(default_factory=list)
    send_nodes: List[np.ndarray] = field(default_factory=list)     # local 1-based ids per neighbour (ascending global)
    recv_nodes: List[np.ndarray] = field(default_factory=list)
    contact: object = None             # ContactLists when the model has contact (master triangles of OWN elements only)


def partition_model_ghost(setup: Setup, n_ranks: int, only_rank: int = None) -> List[GhostDomain]:
    """SURVEY §8e's optional mode.  Every node of a rank's own elements has ALL its incident elements on the rank
    (own or ghost), listed in ascending global order, so its internal-force sum is the single-domain sum bit for bit;
    the outer nodes of the ghost layer are overwritten each step with the state computed by a rank that owns them.
    Results do not depend on the number of ranks.
    Contact works as in partition_model (global slave lists, ghost copies of remote surface nodes appended after the
    element-layer nodes, exact integer force sums over ranks — order independent, hence also partition independent);
    a master triangle is processed by the rank that OWNS its element, never by a rank holding it as a ghost."""
    m = setup.model
    surf, candidates, erosion = _contact_candidates(setup)
    nE = m.nElement
    bounds = [(nE * r) // n_ranks for r in range(n_ranks + 1)]
    holds = np.zeros((n_ranks, m.nNode + 1), bool)              # node of an OWN element of rank r
    for r in range(n_ranks):
        holds[r, np.unique(m.elementmat[:, bounds[r]:bounds[r + 1]])] = True
    owner = np.argmax(holds, axis=0)                            # lowest rank that computes the node completely
    local_el, ghost_nodes = [], []
    for r in range(n_ranks):
        el = np.flatnonzero(holds[r][m.elementmat].any(axis=0))                # own block + ghost layer, ascending
        nodes = np.unique(m.elementmat[:, el])
        local_el.append((el, nodes))
        ghost_nodes.append(nodes[~holds[r, nodes]])
    doms = []
    for r in range(n_ranks):
        if only_rank is not None and r != only_rank:
            doms.append(None)
            continue
        el, layer_nodes = local_el[r]
        n_layer = len(layer_nodes)                              # nodes of local (own + ghost) elements come first ...
        nodes = np.concatenate([layer_nodes, np.setdiff1d(candidates, layer_nodes)])    # ... then contact-only ghosts
        g2l = np.zeros(m.nNode + 1, np.int64)
        g2l[nodes] = np.arange(1, len(nodes) + 1)
        lm = copy.copy(m)
        lm.nNode, lm.nElement = len(nodes), len(el)
        lm.coordmat = np.ascontiguousarray(m.coordmat[:, nodes - 1])
        lm.elementmat = np.ascontiguousarray(g2l[m.elementmat[:, el]])
        lm.element_material = m.element_material[el]
        lm.element_instance = np.ones(len(el), np.int64)
        lm.BC, lm.IC, lm.INSTANCE = [], [], []
        for bc in m.BC:
            nb = I.BC(Nset_name=bc.Nset_name, amp_name=bc.amp_name, amplitude=bc.amplitude)
            nb.dof, nb.value = _restrict_dofs(bc.dof, bc.value, g2l)
            lm.BC.append(nb)
        for ic in m.IC:
            ni = I.IC(Nset_name=ic.Nset_name, type=ic.type)
            ni.dof, ni.value = _restrict_dofs(ic.dof, ic.value, g2l)
            lm.IC.append(ni)
        lst = Setup(lm, setup.d_time, setup.time_num, (None if setup.elementVolume is None else setup.elementVolume[el]),
                    np.repeat(setup.diag_M.reshape(-1, 3)[nodes - 1, 0], 3), setup.elementMinSize, setup.elementMaxSize)
        own_elem = (el >= bounds[r]) & (el < bounds[r + 1])
        dom = GhostDomain(r, lst, nodes, el + 1, own_elem, holds[r, nodes])
        if m.contact_flag:
            from .model_setup import ContactTriangle
            e_g2l = np.zeros(nE + 1, np.int64)                  # global element -> local id, OWN elements only
            e_g2l[el[own_elem] + 1] = np.flatnonzero(own_elem) + 1
            maps = None
            if erosion:
                lm.INSTANCE = [copy.copy(ins) for ins in m.INSTANCE]
                maps = ErosionMaps(g2l[1:].copy(), e_g2l[1:].copy(), np.asarray(m.element_instance, np.int64),
                                   owner.astype(np.int64), n_layer)
            for ct in setup.CT:
                mine = e_g2l[ct.c_triangles_eleid] > 0
                lst.CT.append(ContactTriangle(ct.i_instance, ct.j_instance, g2l[ct.c_nodes_i], g2l[ct.c_nodes_j],
                                              g2l[ct.c_triangles[mine]], e_g2l[ct.c_triangles_eleid[mine]], ct.young))
            dom.contact = build_contact_lists(surf, owner, g2l, n_layer, r, n_ranks, maps)
        for q in range(n_ranks):
            if q == r:
                continue
            recv = ghost_nodes[r][owner[ghost_nodes[r]] == q]               # q computes them, I copy them
            send = ghost_nodes[q][owner[ghost_nodes[q]] == r]               # I compute them, q copies them
            if len(recv) or len(send):
                dom.neighbors.append(q)
                dom.recv_nodes.append(g2l[recv])
                dom.send_nodes.append(g2l[send])
        doms.append(dom)
    return doms


class GhostRunner:
    """Engine + state exchange of one rank of a ghost-element partition: per step
    hk_step_begin (nodal update) -> hk_state_export -> send/recv -> hk_state_import -> hk_step_finish (elements)."""

    def __init__(self, engine_cls, dom: GhostDomain, torch_device, world=None, force_exchange="allreduce", **params):
        import torch
        self.dom = dom
        self.engine = configure_engine(engine_cls, dom.setup, **params)
        self.contact = None
        if dom.contact is not None:
            if world is None:
                raise ValueError("contact decks: pass world")
            self.contact = ContactExchanger(self.engine, dom.contact, world, torch_device, dom.rank, dom.node_l2g,
                                            len(dom.setup.CT), force_exchange)
        self.erosion = self.contact is not None and self.contact.erosion is not None
        cat = lambda lists: np.concatenate(lists) if lists else np.zeros(0, np.int64)
        self.engine.set_node_list(3, cat(dom.send_nodes))
        self.engine.set_node_list(4, cat(dom.recv_nodes))
        self.send = torch.zeros(6 * sum(len(x) for x in dom.send_nodes), dtype=torch.float64, device=torch_device)
        self.recv = torch.zeros(6 * sum(len(x) for x in dom.recv_nodes), dtype=torch.float64, device=torch_device)
        so = np.cumsum([0] + [6 * len(x) for x in dom.send_nodes])
        ro = np.cumsum([0] + [6 * len(x) for x in dom.recv_nodes])
        self.send_parts = [self.send[so[i]:so[i + 1]] for i in range(len(dom.neighbors))]
        self.recv_parts = [self.recv[ro[i]:ro[i + 1]] for i in range(len(dom.neighbors))]
        self.n_reported = 0
        self._seen_erosion = 0

    def run(self, t_first: int, n_steps: int, frame_at_end: bool = False) -> int:
        """Returns the number of OWN elements deleted in these steps (ghost copies delete in step but are not counted)."""
        import torch.distributed as dist
        eng = self.engine
        for t in range(t_first, t_first + n_steps):
            if frame_at_end and t == t_first + n_steps - 1:
                eng.mark_frame()
            if self.contact is not None:
                self.contact.run()          # ghost-layer nodes are current (state_import of the previous step)
            eng.step_begin(t)
            if self.dom.neighbors:
                eng.state_export(self.send.data_ptr())
                ops = []
                for i, nb in enumerate(self.dom.neighbors):
                    if len(self.send_parts[i]):
                        ops.append(dist.P2POp(dist.isend, self.send_parts[i], nb))
                    if len(self.recv_parts[i]):
                        ops.append(dist.P2POp(dist.irecv, self.recv_parts[i], nb))
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
                eng.state_import(self.recv.data_ptr())
            eng.step_finish(t)
            if self.erosion:                # every rank replays the same ascending list of freshly deleted GLOBAL ids
                eng.sync()
                ids = eng.deleted_ids()[self._seen_erosion:]
                self._seen_erosion += len(ids)
                self.contact.exchange_deleted(self.dom.elem_l2g[ids[self.dom.own_elem[ids - 1]] - 1])
        eng.sync()
        ids = eng.deleted_ids()
        fresh = ids[self.n_reported:]
        self.n_reported = len(ids)
        return int(self.dom.own_elem[fresh - 1].sum())

    def deleted_global_ids(self) -> np.ndarray:
        """Global ids of the OWN elements deleted so far, in deletion order."""
        ids = self.engine.deleted_ids()
        return self.dom.elem_l2g[ids[self.dom.own_elem[ids - 1]] - 1]
