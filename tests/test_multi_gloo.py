"""world_size-2 (and 3) CPU tests of the multi-GPU path: slab partition + force halos over gloo.

Each rank drives the host-compiled kernel build (tests/emu) on its element block and exchanges interface
partial forces with torch.distributed; rank 0 gathers the fields and compares them with ONE oracle run on
the unpartitioned mesh.  Covers partition_model (general decks) and slab_deck (the bench's direct slab build)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, mode, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.model_setup import prepare
        from hakai_fem_b200.multi import partition_model, slab_deck, SlabRunner
        from hakai_fem_b200.mesh import StretchDeck, steel
        from tests.emu.emu_engine import EmuEngine
        n_steps = 90
        if mode == "partition":
            deck = StretchDeck(4, 3, 9, jitter=0.1, strain_per_step=8e-4,
                               material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]))
            gsetup = prepare(deck.build_model())
            dom = partition_model(gsetup, world)[rank]
            run = SlabRunner(EmuEngine, dom.setup, dom.neighbors, dom.halo_nodes, "cpu")
            node_l2g, elem_l2g = dom.node_l2g, dom.elem_l2g
        else:
            deck = StretchDeck(3, 4, 3, jitter=0.08, strain_per_step=3e-4)
            local, nbrs, halos = slab_deck(deck, rank, world)
            lsetup = prepare(local.build_model())
            run = SlabRunner(EmuEngine, lsetup, nbrs, halos, "cpu", sum_mass=True)
            per = (deck.nx + 1) * (deck.ny + 1)
            nloc = lsetup.model.nNode
            node_l2g = np.arange(1, nloc + 1) + rank * deck.nz * per
            elem_l2g = np.arange(1, lsetup.model.nElement + 1) + rank * lsetup.model.nElement
            q.put(("coords", rank, lsetup.model.coordmat, node_l2g))
        nd = run.run(1, n_steps)
        d = run.engine.download()
        q.put(("result", rank, dict(disp=d["disp"], eps=d["integ_eq_plastic_strain"], stress=np.asarray(d["integ_stress"]),
                                    flag=d["element_flag"], node_l2g=node_l2g, elem_l2g=elem_l2g, nd=nd)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _run(world, mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + (7 if mode == "slab" else 0) + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    msgs = []
    need = world * (2 if mode == "slab" else 1)
    while len(msgs) < need:
        msgs.append(q.get(timeout=300))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return msgs


def _compare(msgs, gsetup, n_steps=90):
    from hakai_fem_b200.model_setup import configure_engine
    from oracle.oracle_engine import OracleEngine
    o = configure_engine(OracleEngine, gsetup)
    nd_ref = o.step(1, n_steps)
    ref = o.download()
    nd_tot = 0
    for kind, rank, r in [m for m in msgs if m[0] == "result"]:
        n = r["node_l2g"] - 1
        e = r["elem_l2g"] - 1
        gd = ref["disp"].reshape(-1, 3)[n].reshape(-1)
        scale = np.abs(ref["disp"]).max()
        assert np.abs(gd - r["disp"]).max() <= 1e-11 * scale, f"rank {rank} disp"
        ip = (e[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        assert np.abs(ref["integ_eq_plastic_strain"][ip] - r["eps"]).max() <= 1e-11 * max(ref["integ_eq_plastic_strain"].max(), 1e-30)
        assert np.abs(np.asarray(ref["integ_stress"])[:, ip] - r["stress"]).max() <= 1e-10 * np.abs(ref["integ_stress"]).max()
        assert np.array_equal(ref["element_flag"][e], r["flag"]), f"rank {rank} flags"
        nd_tot += r["nd"]
    assert nd_tot == nd_ref
    return nd_ref


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_run_matches_single_domain(world):
    from hakai_fem_b200.model_setup import prepare
    from hakai_fem_b200.mesh import StretchDeck, steel
    msgs = _run(world, "partition")
    deck = StretchDeck(4, 3, 9, jitter=0.1, strain_per_step=8e-4,
                       material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]))
    nd = _compare(msgs, prepare(deck.build_model()))
    assert nd > 0, "deck should delete elements inside the window (fracture across ranks)"


def test_slab_deck_matches_global_deck():
    """The bench's per-rank slab build (no global mesh in memory) is the same problem as one global deck."""
    from hakai_fem_b200.model_setup import prepare
    from hakai_fem_b200.mesh import StretchDeck
    world = 2
    msgs = _run(world, "slab")
    deck = StretchDeck(3, 4, 3 * world, jitter=0.0, strain_per_step=3e-4)
    gm = deck.build_model()
    for kind, rank, coord, l2g in [m for m in msgs if m[0] == "coords"]:      # the slabs carry their own jitter
        gm.coordmat[:, l2g - 1] = coord
    gm.PART[0].coordmat = gm.coordmat
    _compare(msgs, prepare(gm))


def _worker_upload(rank, world, port, q):
    """Split step after hk_upload_state(Q=...): the uploaded local partial force + the neighbour's halo partial."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.model_setup import prepare
        from hakai_fem_b200.multi import slab_deck, SlabRunner
        from hakai_fem_b200.mesh import StretchDeck
        from tests.emu.emu_engine import EmuEngine
        deck = StretchDeck(3, 3, 3, jitter=0.05, strain_per_step=5e-4)
        outs = []
        for mode in ("plain", "reupload"):
            local, nbrs, halos = slab_deck(deck, rank, world)
            run = SlabRunner(EmuEngine, prepare(local.build_model()), nbrs, halos, "cpu", sum_mass=True)
            run.run(1, 10)
            if mode == "reupload":                       # download everything, upload it again, continue
                d = run.engine.download()
                ex = run.engine.download_ex(fields=("disp_pre", "Q", "integ_yield_stress"))
                run.engine.upload_state(disp=d["disp"], disp_pre=ex["disp_pre"], velo=d["velo"], Q=ex["Q"],
                                        integ_stress=d["integ_stress"], integ_strain=d["integ_strain"],
                                        integ_eq_plastic_strain=d["integ_eq_plastic_strain"],
                                        integ_yield_stress=ex["integ_yield_stress"])
            run.run(11, 10)
            outs.append(run.engine.download())
        same = all(np.array_equal(np.asarray(outs[0][k]), np.asarray(outs[1][k])) for k in ("disp", "integ_stress", "integ_eq_plastic_strain"))
        q.put((rank, same))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_upload_state_then_split_step_is_transparent():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_upload, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def _worker_contact(rank, world, port, q):
    """Two-body impact split over ranks: contact-surface all-gather + exact force exchange."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.model_setup import prepare
        from hakai_fem_b200.multi import partition_model, SlabRunner
        from hakai_fem_b200.mesh import ImpactDeck
        from tests.emu.emu_engine import EmuEngine
        gsetup = prepare(ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3)).build_model())
        dom = partition_model(gsetup, world)[rank]
        run = SlabRunner.from_domain(EmuEngine, dom, "cpu", world, contact_myu=0.25)
        run.run(1, 40)
        d = run.engine.download()
        n_own = len(np.unique(dom.setup.model.elementmat))          # held nodes first, ghosts after
        q.put((rank, dict(disp=d["disp"][:3 * n_own], eps=d["integ_eq_plastic_strain"], node_l2g=dom.node_l2g[:n_own],
                          elem_l2g=dom.elem_l2g, hits=int(run.engine.counters()[1]), n_ghost=len(dom.contact.import_nodes))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_contact_across_ranks_matches_single_domain(world):
    from hakai_fem_b200.model_setup import prepare, configure_engine
    from hakai_fem_b200.mesh import ImpactDeck
    from oracle.oracle_engine import OracleEngine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_contact, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    gsetup = prepare(ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3)).build_model())
    o = configure_engine(OracleEngine, gsetup, contact_myu=0.25)
    o.step(1, 40)
    ref = o.download()
    assert o.counters()[1] > 0
    assert sum(r["hits"] for _, r in res) == o.counters()[1], "every hit is found by exactly one rank"
    assert any(r["n_ghost"] > 0 for _, r in res)
    scale = np.abs(ref["disp"]).max()
    for rank, r in res:
        n = r["node_l2g"] - 1
        assert np.abs(ref["disp"].reshape(-1, 3)[n].reshape(-1) - r["disp"]).max() <= 1e-10 * scale, f"rank {rank}"
        ip = ((r["elem_l2g"] - 1)[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        assert np.abs(ref["integ_eq_plastic_strain"][ip] - r["eps"]).max() <= 1e-10 * max(ref["integ_eq_plastic_strain"].max(), 1e-30)


def _erosion_setup():
    from hakai_fem_b200.model_setup import prepare
    from hakai_fem_b200.mesh import ImpactDeck
    model = ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0).build_model()
    model.MATERIAL[0].ductile = np.array([[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]])
    return prepare(model)


def _worker_erosion(rank, world, port, q, device_erosion=False):
    """Brittle plate split over ranks: faces exposed by deletions on one rank join the contact surface on all.
    device_erosion: the gathered ids are replayed on the "device" (hk_comm_erosion, static exchange lists)."""
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.multi import partition_model, SlabRunner
        from tests.emu.emu_engine import EmuEngine
        dom = partition_model(_erosion_setup(), world)[rank]
        assert dom.contact.erosion is not None
        run = SlabRunner.from_domain(EmuEngine, dom, "cpu", world, device_erosion=device_erosion)
        assert run.contact.device_erosion == device_erosion
        n_del = run.run(1, 400)
        d = run.engine.download()
        n_own = dom.contact.erosion.n_held
        pairs = []
        for c in range(len(dom.setup.CT)):
            info = run.engine.contact_pair(c)
            pairs.append(dict(nodes_i=dom.node_l2g[info["c_nodes_i"] - 1], nodes_j=dom.node_l2g[info["c_nodes_j"] - 1],
                              tri=dom.node_l2g[info["c_triangles"] - 1], tele=dom.elem_l2g[info["c_triangles_eleid"] - 1]))
        q.put((rank, dict(disp=d["disp"][:3 * n_own], eps=d["integ_eq_plastic_strain"], flag=d["element_flag"],
                          node_l2g=dom.node_l2g[:n_own], elem_l2g=dom.elem_l2g, n_del=n_del,
                          deleted=dom.elem_l2g[run.engine.deleted_ids() - 1], pairs=pairs,
                          n_surf=len(run.contact.lists.surface_nodes), n_surf0=len(dom.contact.surface_nodes))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,device_erosion", [(2, False), (3, False), (2, True), (3, True)])
def test_erosion_across_ranks_matches_single_domain(world, device_erosion):
    from hakai_fem_b200.model_setup import configure_engine
    from oracle.oracle_engine import OracleEngine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + world + (10 if device_erosion else 0)
    procs = [ctx.Process(target=_worker_erosion, args=(r, world, port, q, device_erosion)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    o = configure_engine(OracleEngine, _erosion_setup())
    o.step(1, 400)
    ref = o.download()
    ids = o.deleted_ids()
    assert len(ids) > 0, "nothing eroded: test is vacuous"
    got = np.concatenate([res[r]["deleted"] for r in range(world)])
    assert np.array_equal(np.sort(got), np.sort(ids)), "deleted-element set differs"
    assert sum(res[r]["n_del"] for r in range(world)) == len(ids)
    if not device_erosion:                                        # (static candidate lists never change)
        assert any(res[r]["n_surf"] > res[r]["n_surf0"] for r in range(world)), "surface never grew"
    for c in range(2):
        po = o.contact_pair(c)
        for r in range(world):                                   # node lists: identical, same order, on every rank
            assert np.array_equal(res[r]["pairs"][c]["nodes_i"], po["c_nodes_i"]), (c, r)
            assert np.array_equal(res[r]["pairs"][c]["nodes_j"], po["c_nodes_j"]), (c, r)
        tri = np.concatenate([np.column_stack([res[r]["pairs"][c]["tri"], res[r]["pairs"][c]["tele"]])
                              for r in range(world)])
        want = np.column_stack([po["c_triangles"], po["c_triangles_eleid"]])
        assert len(tri) == len(want)                             # every triangle lives on exactly one rank
        assert np.array_equal(tri[np.lexsort(tri.T[::-1])], want[np.lexsort(want.T[::-1])])
    scale = np.abs(ref["disp"]).max()
    for r in range(world):
        n = res[r]["node_l2g"] - 1
        e = res[r]["elem_l2g"] - 1
        want = ref["disp"].reshape(-1, 3)[n].reshape(-1)
        assert np.abs(res[r]["disp"] - want).max() <= 1e-7 * scale
        assert np.array_equal(res[r]["flag"], ref["element_flag"][e])
        ep = ref["integ_eq_plastic_strain"].reshape(-1, 8)[e].reshape(-1)
        assert np.abs(res[r]["eps"] - ep).max() <= 1e-7 * max(np.abs(ep).max(), 1e-30)


def _vtk_sections(path):
    """ASCII legacy VTK -> {section name: flat array}"""
    out, name = {}, None
    for line in open(path):
        w = line.split()
        if not w:
            continue
        if w[0] in ("POINTS", "CELLS", "CELL_TYPES"):
            name = w[0]
            out[name] = []
        elif w[0] in ("SCALARS", "VECTORS"):
            name = w[1]
            out[name] = []
        elif w[0] in ("LOOKUP_TABLE", "POINT_DATA", "#", "Test", "ASCII", "DATASET"):
            continue
        elif name is not None:
            out[name].extend(float(v) for v in w)
    return {k: np.array(v) for k, v in out.items()}


def _worker_host(rank, world, port, deck_path, outdir, q, partition="halo"):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.host import hakai_distributed
        from tests.emu.emu_engine import EmuEngine
        runner, frames = hakai_distributed(deck_path, outdir, engine_cls=EmuEngine, torch_device="cpu", output_num=6,
                                           verbose=False, partition=partition)
        q.put((rank, len(frames)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["fracture", "impact_erosion"])
def test_distributed_host_driver_writes_the_single_domain_frames(kind, tmp_path):
    """`torchrun -m hakai_fem_b200.host deck.inp` path: partitioned run + rank-0 assembly of the frames (nodal sums of
    interface nodes added across ranks) against the single-domain driver on the same deck."""
    from hakai_fem_b200.host import hakai
    from hakai_fem_b200.mesh import ImpactDeck, StretchDeck, steel
    from tests.emu.emu_engine import EmuEngine
    deck_path = str(tmp_path / "deck.inp")
    if kind == "fracture":
        StretchDeck(4, 3, 8, jitter=0.05, strain_per_step=8e-4, n_steps=120,
                    material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]])).write_inp(deck_path)
    else:
        ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0, n_steps=120,
                   plate_ductile=[[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]]).write_inp(deck_path)
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 2000) + (0 if kind == "fracture" else 1)
    procs = [ctx.Process(target=_worker_host, args=(r, world, port, deck_path, str(tmp_path / "multi"), q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eng, ref = hakai(deck_path, str(tmp_path / "single"), engine_cls=EmuEngine, output_num=6, verbose=False)
    assert len(eng.deleted_ids()) > 0, "nothing deleted: CELLS would not be exercised"
    assert res[0] == len(ref) == 7 and res[1] == 0
    for f in ref:
        a = _vtk_sections(f)
        b = _vtk_sections(os.path.join(str(tmp_path / "multi"), os.path.basename(f)))
        assert list(a) == list(b) and len(a) == 22
        for k in a:
            assert a[k].shape == b[k].shape, (os.path.basename(f), k)          # same live-cell list
            if k in ("CELLS", "CELL_TYPES", "POINTS"):
                assert np.array_equal(a[k], b[k]), (os.path.basename(f), k)
            else:                                                   # %1.6e text of values equal to ~1e-9 relative
                scale = max(np.abs(a[k]).max(), 1e-300)
                assert np.allclose(a[k], b[k], rtol=0, atol=3e-6 * scale), (os.path.basename(f), k, scale)


def _worker_force_exchange(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.model_setup import prepare
        from hakai_fem_b200.multi import partition_model, SlabRunner
        from hakai_fem_b200.mesh import ImpactDeck
        from tests.emu.emu_engine import EmuEngine
        out = {}
        for mode in ("allgather", "allreduce"):
            gsetup = prepare(ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3)).build_model())
            dom = partition_model(gsetup, world, only_rank=rank)[rank]
            run = SlabRunner.from_domain(EmuEngine, dom, "cpu", world, force_exchange=mode, contact_myu=0.25)
            run.run(1, 60)
            d = run.engine.download()
            x = run.engine.download_ex(fields=("external_force",))
            out[mode] = dict(disp=d["disp"], velo=d["velo"], eps=d["integ_eq_plastic_strain"], ext=x["external_force"],
                             hits=int(run.engine.counters()[1]))
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_contact_force_exchange_by_limb_allreduce_is_bit_identical():
    """The 128-bit accumulators as three 43-bit limbs through ONE integer all-reduce give exactly the state the
    all-gather + 128-bit sum gives (positive and negative forces, carries across limbs)."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_force_exchange, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sum(res[r]["allgather"]["hits"] for r in range(world)) > 0
    for r in range(world):
        a, b = res[r]["allgather"], res[r]["allreduce"]
        assert np.abs(a["ext"]).max() > 0 and (a["ext"] < 0).any() and (a["ext"] > 0).any()
        for k in ("disp", "velo", "eps", "ext"):
            assert np.array_equal(a[k], b[k]), (r, k)
        assert a["hits"] == b["hits"]


def test_limb_split_and_recombine_formulas_are_exact_mod_2_128():
    """The bit manipulation of hk_launch_cacc_export_limbs / _import_limbs (hk_exact.cu), restated with Python integers
    masked to 64 bits, over random 128-bit two's-complement values and up to 4096 ranks."""
    import random
    rnd = random.Random(5)
    M64, M43 = (1 << 64) - 1, (1 << 43) - 1

    def export(x):
        lo, hi = x & M64, (x >> 64) & M64
        return [lo & M43, ((lo >> 43) | ((hi << 21) & M64)) & M43, hi >> 22]

    def recombine(s0, s1, s2):
        lo, hi = s0 & M64, 0
        a_lo, a_hi = (s1 << 43) & M64, s1 >> 21
        nlo = (lo + a_lo) & M64
        hi = (hi + a_hi + (1 if nlo < lo else 0)) & M64
        lo = nlo
        hi = (hi + ((s2 << 22) & M64)) & M64
        return lo | (hi << 64)
    for ranks in (1, 2, 3, 8, 4096):
        for _ in range(200):
            vals = [rnd.choice([rnd.getrandbits(128), (1 << 128) - rnd.getrandbits(70), rnd.getrandbits(50), (1 << 128) - 1, 0])
                    for _ in range(ranks)]
            limbs = [export(v) for v in vals]
            sums = [sum(l[i] for l in limbs) for i in range(3)]
            assert all(0 <= s < (1 << 63) for s in sums)               # fits the int64 lanes of the all-reduce
            assert recombine(*sums) == sum(vals) % (1 << 128)


def _worker_ghost(rank, world, port, exact, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.model_setup import prepare
        from hakai_fem_b200.multi import partition_model_ghost, GhostRunner
        from hakai_fem_b200.mesh import StretchDeck, steel
        from tests.emu.emu_engine import EmuEngine
        deck = StretchDeck(5, 4, 9, jitter=0.1, strain_per_step=8e-4,
                           material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]))
        dom = partition_model_ghost(prepare(deck.build_model()), world, only_rank=rank)[rank]
        run = GhostRunner(EmuEngine, dom, "cpu", element_mode=1 if exact else 0)
        nd = run.run(1, 60) + run.run(61, 30)
        d = run.engine.download()
        x = run.engine.download_ex(fields=("Q", "integ_yield_stress"))
        own_n = np.flatnonzero(dom.own_node)
        own_e = np.flatnonzero(dom.own_elem)
        ip = (own_e[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        q.put((rank, dict(nodes=dom.node_l2g[own_n] - 1, elems=dom.elem_l2g[own_e] - 1, nd=nd,
                          disp=d["disp"].reshape(-1, 3)[own_n], Q=x["Q"].reshape(-1, 3)[own_n],
                          stress=d["integ_stress"][:, ip], eps=d["integ_eq_plastic_strain"][ip],
                          flag=d["element_flag"][own_e], deleted=run.deleted_global_ids(),
                          n_ghost_el=int((~dom.own_elem).sum()), n_ghost_nodes=int((~dom.own_node).sum()))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,exact", [(2, False), (3, False), (3, True)])
def test_ghost_element_partition_is_bit_identical_to_single_domain(world, exact):
    """SURVEY 8e optional mode: with a ghost-element layer the partitioned run reproduces the single-domain run of the
    same kernels BIT FOR BIT for any number of ranks; with element_mode=1 that single-domain run is the oracle's."""
    from hakai_fem_b200.model_setup import prepare, configure_engine
    from hakai_fem_b200.mesh import StretchDeck, steel
    from oracle.oracle_engine import OracleEngine
    from tests.emu.emu_engine import EmuEngine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 2000) + 2 * world + int(exact)
    procs = [ctx.Process(target=_worker_ghost, args=(r, world, port, exact, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    deck = StretchDeck(5, 4, 9, jitter=0.1, strain_per_step=8e-4,
                       material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]))
    st = prepare(deck.build_model())
    ref_eng = configure_engine(OracleEngine, st) if exact else configure_engine(EmuEngine, st)
    nd_ref = ref_eng.step(1, 90)
    ref = ref_eng.download()
    refx = ref_eng.download_ex(fields=("Q",))
    assert nd_ref > 0, "nothing deleted: the fracture path would not be exercised"
    assert sum(res[r]["nd"] for r in range(world)) == nd_ref
    assert np.array_equal(np.sort(np.concatenate([res[r]["deleted"] for r in range(world)])), np.sort(ref_eng.deleted_ids()))
    for r in range(world):
        a = res[r]
        assert a["n_ghost_el"] > 0 and a["n_ghost_nodes"] > 0
        ip = (a["elems"][:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        assert np.array_equal(a["disp"], ref["disp"].reshape(-1, 3)[a["nodes"]]), r
        assert np.array_equal(a["Q"], refx["Q"].reshape(-1, 3)[a["nodes"]]), r
        assert np.array_equal(a["stress"], ref["integ_stress"][:, ip]), r
        assert np.array_equal(a["eps"], ref["integ_eq_plastic_strain"][ip]), r
        assert np.array_equal(a["flag"], ref["element_flag"][a["elems"]]), r


def test_distributed_host_driver_with_ghost_partitions_writes_identical_files(tmp_path):
    """partition="ghost": the frames of a 3-rank run are byte-for-byte the single-domain frames."""
    from hakai_fem_b200.host import hakai
    from hakai_fem_b200.mesh import StretchDeck, steel
    from tests.emu.emu_engine import EmuEngine
    deck_path = str(tmp_path / "deck.inp")
    StretchDeck(4, 3, 8, jitter=0.05, strain_per_step=8e-4, n_steps=120,
                material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]])).write_inp(deck_path)
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30300 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_host, args=(r, world, port, deck_path, str(tmp_path / "multi"), q, "ghost"))
             for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    eng, ref = hakai(deck_path, str(tmp_path / "single"), engine_cls=EmuEngine, output_num=6, verbose=False)
    assert len(eng.deleted_ids()) > 0 and res[0] == len(ref) == 7
    for f in ref:
        assert open(f).read() == open(os.path.join(str(tmp_path / "multi"), os.path.basename(f))).read(), f


def _ghost_contact_setup(kind):
    from hakai_fem_b200.model_setup import prepare
    from hakai_fem_b200.mesh import ImpactDeck
    if kind == "crash_tube":                                     # the reference's self-contact deck (tie-sensitive)
        from tests import util
        return util.deck_setup("crash_tube"), 400
    if kind == "erosion":
        return prepare(ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0,
                                  plate_ductile=[[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]]).build_model()), 120
    return prepare(ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3)).build_model()), 60


def _worker_ghost_contact(rank, world, port, kind, exact, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.multi import partition_model_ghost, GhostRunner
        from tests.emu.emu_engine import EmuEngine
        st, n_steps = _ghost_contact_setup(kind)
        dom = partition_model_ghost(st, world, only_rank=rank)[rank]
        run = GhostRunner(EmuEngine, dom, "cpu", world=world, force_exchange="allreduce" if kind == "impact" else "allgather",
                          element_mode=1 if exact else 0)
        nd = run.run(1, n_steps)
        d = run.engine.download()
        x = run.engine.download_ex(fields=("external_force",))
        own_n, own_e = np.flatnonzero(dom.own_node), np.flatnonzero(dom.own_elem)
        ip = (own_e[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        q.put((rank, dict(nodes=dom.node_l2g[own_n] - 1, elems=dom.elem_l2g[own_e] - 1, nd=nd,
                          disp=d["disp"].reshape(-1, 3)[own_n], velo=d["velo"].reshape(-1, 3)[own_n],
                          ext=x["external_force"].reshape(-1, 3)[own_n], eps=d["integ_eq_plastic_strain"][ip],
                          stress=d["integ_stress"][:, ip], flag=d["element_flag"][own_e],
                          deleted=run.deleted_global_ids(), hits=int(run.engine.counters()[1]),
                          n_ghost_el=int((~dom.own_elem).sum()))))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind,world,exact", [("impact", 2, False), ("erosion", 3, False), ("crash_tube", 2, True)])
def test_ghost_partition_with_contact_is_bit_identical(kind, world, exact):
    """Contact decks on ghost-element partitions: exact (order-independent) contact sums + complete local element sums
    => bit-identical to the single-domain run for any rank count; with element_mode=1 bit-identical to the ORACLE, even
    on the reference's tie-sensitive self-contact deck."""
    from hakai_fem_b200.model_setup import configure_engine
    from oracle.oracle_engine import OracleEngine
    from tests.emu.emu_engine import EmuEngine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30500 + (os.getpid() % 2000) + ("impact", "erosion", "crash_tube").index(kind)
    procs = [ctx.Process(target=_worker_ghost_contact, args=(r, world, port, kind, exact, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=400) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    st, n_steps = _ghost_contact_setup(kind)
    ref_eng = configure_engine(OracleEngine if exact else EmuEngine, st)
    nd_ref = ref_eng.step(1, n_steps)
    ref = ref_eng.download()
    refx = ref_eng.download_ex(fields=("external_force",))
    assert ref_eng.counters()[1] > 0, "no contact: vacuous"
    assert sum(res[r]["hits"] for r in range(world)) == ref_eng.counters()[1], "every hit found by exactly one rank"
    assert sum(res[r]["nd"] for r in range(world)) == nd_ref
    if kind == "erosion":
        assert nd_ref > 10
    assert np.array_equal(np.sort(np.concatenate([res[r]["deleted"] for r in range(world)])), np.sort(ref_eng.deleted_ids()))
    for r in range(world):
        a = res[r]
        ip = (a["elems"][:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        for k, want in (("disp", ref["disp"].reshape(-1, 3)[a["nodes"]]), ("velo", ref["velo"].reshape(-1, 3)[a["nodes"]]),
                        ("ext", refx["external_force"].reshape(-1, 3)[a["nodes"]]),
                        ("eps", ref["integ_eq_plastic_strain"][ip]), ("stress", ref["integ_stress"][:, ip]),
                        ("flag", ref["element_flag"][a["elems"]])):
            assert np.array_equal(a[k], want), (kind, r, k)


# ---- bench.py's N > 1 parity check (slab_parity_check) on CPU ranks ------------------------------------------------
def _worker_parity_check(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.multi import slab_parity_check
        from tests.emu.emu_engine import EmuEngine
        res = slab_parity_check(EmuEngine, "cpu", rank, world, n_steps=30, nx=6, ny=5, nz_per_rank=3)
        q.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_parity_check_green_on_cpu_ranks(world):
    """The check bench.py emits as `parity_check` at N > 1: slabs of ONE global jittered mesh (global-layer seeds) over
    the halo exchange vs the same mesh unpartitioned — deleted set identical, fields <= 1e-10, shared layers bitwise."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_parity_check, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    r0 = res[0]
    assert all(res[r] is None for r in range(1, world))
    assert r0["ok"], r0
    assert r0["n_ranks"] == world and r0["n_deleted"] > 0 and r0["interface_bitwise"] and r0["max_rel_err"] <= 1e-10


def test_slab_jitter_is_a_function_of_the_global_layer():
    """slab_deck: the slabs of an N-rank run tile exactly the mesh a single domain would build (per-rank
    seeds and unjittered interface planes made the weak-scaling mesh one no single-domain run computes)."""
    from hakai_fem_b200.mesh import StretchDeck
    from hakai_fem_b200.multi import slab_deck
    deck = StretchDeck(5, 4, 3, jitter=0.07, jitter_by_layer=True)
    world = 3
    whole, _ = StretchDeck(5, 4, 3 * world, jitter=0.07, jitter_by_layer=True).coord_elem()
    per = 6 * 5
    for r in range(world):
        local, nbrs, halos = slab_deck(deck, r, world)
        c, _ = local.coord_elem()
        assert np.array_equal(c, whole[:, r * 3 * per:(r * 3 + 4) * per])
    inner = whole[:, 3 * per:4 * per]                         # an interface plane: interior nodes ARE jittered
    base, _ = StretchDeck(5, 4, 3 * world).coord_elem()
    assert np.any(inner != base[:, 3 * per:4 * per])


# ---- interface nodes held by three or more ranks: rank-ordered sums keep every copy bit-identical ----------
def _worker_threeway(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from hakai_fem_b200.model_setup import prepare
        from hakai_fem_b200.multi import partition_model, SlabRunner
        from hakai_fem_b200.mesh import StretchDeck
        from tests.emu.emu_engine import EmuEngine
        gsetup = prepare(StretchDeck(4, 3, 2, jitter=0.1, strain_per_step=3e-4).build_model())
        dom = partition_model(gsetup, world)[rank]
        run = SlabRunner.from_domain(EmuEngine, dom, "cpu", world)
        run.run(1, 200)
        d = run.engine.download(fields=("disp",))
        n_own = len(np.unique(dom.setup.model.elementmat))
        q.put((rank, dict(disp=d["disp"][:3 * n_own].reshape(-1, 3), node_l2g=dom.node_l2g[:n_own])))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_nodes_shared_by_three_ranks_stay_bit_identical():
    """Contiguous element blocks of a 4 x 3 x 2 mesh on 4 ranks give nodes held by 3 ranks.  With hk_set_halo_ranks the
    partial forces are added in ascending global-rank order on every holder, so all copies of a node carry the same
    bits after 200 steps (own-first order let them drift apart by ~1e-15 per step)."""
    world = 4
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 32500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_threeway, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    copies = {}
    for r in range(world):
        for g, u in zip(res[r]["node_l2g"], res[r]["disp"]):
            copies.setdefault(int(g), []).append(u)
    multi = {g: c for g, c in copies.items() if len(c) >= 3}
    assert len(multi) >= 4, "mesh has no node held by three ranks: test is vacuous"
    for g, c in copies.items():
        for u in c[1:]:
            assert np.array_equal(c[0], u), f"node {g} ({len(c)} holders) differs across ranks"
    assert np.abs(np.concatenate([c[0] for c in copies.values()])).max() > 0


# ---- the reference's eroding example decks, partitioned, surfaces kept current on the device (hk_comm_erosion) --------
def reference_deck_rank(name, n_steps, rank, world, make_engine, device, engine_comm):
    """One rank of a reference deck split into element blocks with device-side erosion; returns what the checker needs
    (global ids).  Shared with tests/test_gpu_multi.py (CUDA engine, NCCL, engine_comm=True: ONE hk_step_enqueue)."""
    from hakai_fem_b200.multi import partition_model, SlabRunner
    from tests import util
    dom = partition_model(util.deck_setup(name), world, only_rank=rank)[rank]
    run = SlabRunner.from_domain(make_engine, dom, device, world, engine_comm=engine_comm, device_erosion=True)
    assert run.contact.device_erosion and run.erosion_on_device == engine_comm
    n_del = run.run(1, n_steps)
    d = run.engine.download()
    n_own = dom.contact.erosion.n_held
    pairs = []
    for c in range(len(dom.setup.CT)):
        info = run.engine.contact_pair(c)
        pairs.append(dict(nodes_i=dom.node_l2g[info["c_nodes_i"] - 1], nodes_j=dom.node_l2g[info["c_nodes_j"] - 1],
                          tri=dom.node_l2g[info["c_triangles"] - 1], tele=dom.elem_l2g[info["c_triangles_eleid"] - 1]))
    return dict(disp=d["disp"][:3 * n_own], flag=d["element_flag"], node_l2g=dom.node_l2g[:n_own], elem_l2g=dom.elem_l2g,
                n_del=n_del, deleted=dom.elem_l2g[run.engine.deleted_ids() - 1], pairs=pairs,
                hits=int(run.engine.counters()[1]))


def check_reference_deck_ranks(name, n_steps, expect_deleted, res, tol=1e-7):
    """res[rank] = reference_deck_rank(...) of every rank, against ONE oracle run of the unpartitioned deck."""
    from hakai_fem_b200.model_setup import configure_engine
    from oracle.oracle_engine import OracleEngine
    from tests import util
    world = len(res)
    o = configure_engine(OracleEngine, util.deck_setup(name))
    o.step(1, n_steps)
    ref = o.download()
    ids = o.deleted_ids()
    assert len(ids) == expect_deleted
    got = np.concatenate([res[r]["deleted"] for r in range(world)])
    assert np.array_equal(np.sort(got), np.sort(ids)), "deleted-element set differs"
    assert sum(res[r]["n_del"] for r in range(world)) == len(ids)
    assert sum(res[r]["hits"] for r in range(world)) == o.counters()[1] > 0, "contact hit counts differ"
    for c in range(len(res[0]["pairs"])):
        po = o.contact_pair(c)
        for r in range(world):                                   # node lists: identical, same order, on every rank
            assert np.array_equal(res[r]["pairs"][c]["nodes_i"], po["c_nodes_i"]), (c, r)
            assert np.array_equal(res[r]["pairs"][c]["nodes_j"], po["c_nodes_j"]), (c, r)
        tri = np.concatenate([np.column_stack([res[r]["pairs"][c]["tri"], res[r]["pairs"][c]["tele"]]) for r in range(world)])
        want = np.column_stack([po["c_triangles"], po["c_triangles_eleid"]])
        assert len(tri) == len(want)                             # every triangle lives on exactly one rank
        assert np.array_equal(tri[np.lexsort(tri.T[::-1])], want[np.lexsort(want.T[::-1])])
    scale = np.abs(ref["disp"]).max()
    for r in range(world):
        n, e = res[r]["node_l2g"] - 1, res[r]["elem_l2g"] - 1
        assert np.abs(res[r]["disp"] - ref["disp"].reshape(-1, 3)[n].reshape(-1)).max() <= tol * scale, r
        assert np.array_equal(res[r]["flag"], ref["element_flag"][e])


def _worker_reference_deck(rank, world, port, name, n_steps, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tests.emu.emu_engine import EmuEngine
        q.put((rank, reference_deck_rank(name, n_steps, rank, world, EmuEngine, "cpu", False)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,n_steps,expect_deleted", [("bullet_impact", 3000, 13)])
def test_reference_deck_erodes_across_ranks_with_device_side_lists(name, n_steps, expect_deleted):
    """bullet-impact.inp on 2 ranks: the 13 deletions expose faces on both; the pair lists every rank holds on the
    "device" equal the oracle's, in its order (host gathers the ids here; over NCCL the engine does: test_gpu_multi)."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_reference_deck, args=(r, world, port, name, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=1500) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    check_reference_deck_ranks(name, n_steps, expect_deleted, res)
