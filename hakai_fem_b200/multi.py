"""Multi-GPU driver: element-block (slab) partition, shared-node force halos over torch.distributed.

One process per GPU.  Every rank owns a contiguous block of elements and ALL nodes those
elements touch; nodes on a partition interface are duplicated.  Per time step
    engine.halo_pack()                      partial internal force of the interface nodes -> send buffers
    batch_isend_irecv with the neighbours   NCCL send/recv over NVLink (gloo in the CPU tests)
    engine.step(t, 1)                       adds the received partials, then the usual nodal + element kernels
Both sides of an interface then update the shared nodes redundantly from bit-identical inputs
(a + b == b + a), so positions never need to be exchanged.  Contact across ranks is not implemented yet
(DESIGN.md §7); fracture is rank-local and works unchanged.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import inp as I
from .model_setup import Setup, prepare, configure_engine, lumped_mass, element_volumes, element_sizes


@dataclass
class LocalDomain:
    rank: int
    setup: Setup                       # local model + local (already halo-summed) diag_M
    node_l2g: np.ndarray               # local -> global node id (1-based values)
    elem_l2g: np.ndarray
    neighbors: List[int] = field(default_factory=list)
    halo_nodes: List[np.ndarray] = field(default_factory=list)     # local 1-based ids per neighbour
    contact: object = None             # ContactLists when the model has contact


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
class ContactLists:
    """Per-rank lists of the contact-surface all-gather (local 1-based node ids)."""
    export_nodes: np.ndarray           # surface nodes this rank owns (ascending global id)
    import_nodes: np.ndarray           # ghost copies of surface nodes owned elsewhere
    import_src: np.ndarray             # index of each ghost's record in the gathered buffer (rank*maxlen + position)
    surface_nodes: np.ndarray          # ALL surface nodes in global order (same order on every rank)
    maxlen: int                        # padded length of one rank's export block


def partition_model(setup: Setup, n_ranks: int) -> List[LocalDomain]:
    """Splits a (small) global model into contiguous element blocks.  Used by the tests and for general
    decks; the 16 M/GPU bench builds each slab directly (slab_deck) without materialising the global mesh.

    With contact, every rank receives the GLOBAL contact node lists (nodes it does not hold are appended as ghost
    nodes that no element references) and the master triangles of its own elements; exposed-face updates after
    element deletion are not propagated across ranks yet."""
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
    surf = np.zeros(0, np.int64)
    if m.contact_flag:
        surf = np.unique(np.concatenate([np.concatenate([ct.c_nodes_i, ct.c_nodes_j, ct.c_triangles.reshape(-1)])
                                         for ct in setup.CT]))
        owner = first_holder[surf]
        export_lists = [surf[owner == r] for r in range(n_ranks)]
        maxlen = max(len(x) for x in export_lists)
    for r in range(n_ranks):
        el, nodes_own = locals_[r]
        ghosts = np.setdiff1d(surf, nodes_own) if m.contact_flag else np.zeros(0, np.int64)
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
        lst = Setup(lm, setup.d_time, setup.time_num, setup.elementVolume[el],
                    np.repeat(setup.diag_M.reshape(-1, 3)[nodes - 1, 0], 3),     # global (summed) mass
                    setup.elementMinSize, setup.elementMaxSize)
        dom = LocalDomain(r, lst, nodes, el + 1)
        if m.contact_flag:
            from .model_setup import ContactTriangle
            e_g2l = np.zeros(nE + 1, np.int64)
            e_g2l[el + 1] = np.arange(1, len(el) + 1)
            lm.INSTANCE = []                                      # no cross-rank exposed-face update yet
            for ct in setup.CT:
                mine = e_g2l[ct.c_triangles_eleid] > 0
                lst.CT.append(ContactTriangle(ct.i_instance, ct.j_instance, g2l[ct.c_nodes_i], g2l[ct.c_nodes_j],
                                              g2l[ct.c_triangles[mine]], e_g2l[ct.c_triangles_eleid[mine]], ct.young))
            g_owner = first_holder[ghosts]
            src = np.zeros(len(ghosts), np.int64)
            for o in range(n_ranks):
                sel = g_owner == o
                src[sel] = o * maxlen + np.searchsorted(export_lists[o], ghosts[sel])
            dom.contact = ContactLists(g2l[export_lists[r]], g2l[ghosts], src, g2l[surf], maxlen)
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
    local.seed = deck.seed + rank
    per = (deck.nx + 1) * (deck.ny + 1)
    nbrs, halos = [], []
    if rank > 0:
        nbrs.append(rank - 1)
        halos.append(np.arange(1, per + 1, dtype=np.int64))
    if rank < world - 1:
        nbrs.append(rank + 1)
        halos.append(np.arange(deck.nz * per + 1, (deck.nz + 1) * per + 1, dtype=np.int64))
    return local, nbrs, halos


class HaloExchanger:
    """send/recv buffers (torch tensors on the engine's device) + one batch of P2P ops per step."""

    def __init__(self, engine, neighbors, halo_nodes, device):
        import torch
        self.torch = torch
        self.engine = engine
        self.neighbors = list(neighbors)
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
        return sum(t.numel() * 8 for t in self.send) * 2


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

    def __init__(self, engine, lists: ContactLists, world: int, device):
        import torch
        self.engine, self.lists, self.world = engine, lists, world
        self.send_nodes = torch.zeros(lists.maxlen * 6, dtype=torch.float64, device=device)
        self.all_nodes = torch.zeros(world * lists.maxlen * 6, dtype=torch.float64, device=device)
        n_surf = len(lists.surface_nodes)
        self.send_acc = torch.zeros(n_surf * 6, dtype=torch.int64, device=device)
        self.all_acc = torch.zeros(world * n_surf * 6, dtype=torch.int64, device=device)
        engine.set_node_list(0, lists.export_nodes)
        engine.set_node_list(1, lists.import_nodes)
        engine.set_node_list(2, lists.surface_nodes)
        self._first = True

    def run(self):
        """positions -> ghosts, contact pass on the local triangles, exact sum of the forces over ranks."""
        import torch.distributed as dist
        eng = self.engine
        eng.nodes_export(self.send_nodes.data_ptr())
        dist.all_gather(list(self.all_nodes.chunk(self.world)), self.send_nodes)
        eng.nodes_import(self.all_nodes.data_ptr(), self.lists.import_src if self._first else None)
        self._first = False
        eng.contact_enqueue()
        eng.contact_export(self.send_acc.data_ptr())
        dist.all_gather(list(self.all_acc.chunk(self.world)), self.send_acc)
        eng.contact_import(self.all_acc.data_ptr(), self.world)


class SlabRunner:
    """Engine + halo exchange of one rank.  `engine_cls` is Engine (CUDA) — the CPU tests pass the
    host-compiled kernel build, whose "device" pointers are host pointers."""

    def __init__(self, engine_cls, setup: Setup, neighbors, halo_nodes, torch_device, sum_mass=False, contact=None,
                 world=1, **params):
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
        self.halo = HaloExchanger(self.engine, neighbors, halo_nodes, torch_device)
        self.contact = ContactExchanger(self.engine, contact, world, torch_device) if contact is not None else None
        self.nElement = model.nElement

    def step(self, t: int) -> int:
        if self.contact is not None:
            self.contact.run()
        self.halo.exchange()
        return self.engine.step(t, 1)

    def run(self, t_first: int, n_steps: int) -> int:
        """Enqueues pack -> exchange -> step for every step without blocking the host, then synchronises once."""
        import time as _time
        _t0 = _time.perf_counter()
        for t in range(t_first, t_first + n_steps):
            if self.contact is not None:
                self.contact.run()
            if self.halo.neighbors:
                reqs = self.halo.start()             # partial forces on their way ...
                self.engine.step_begin(t)            # ... while all non-interface nodes are updated
                self.halo.wait(reqs)
                self.engine.step_finish(t)           # interface nodes, element kernel
            else:
                self.engine.step_enqueue(t, 1)
        self.last_enqueue_s = _time.perf_counter() - _t0      # host time to enqueue (diagnostic)
        return self.engine.sync()
