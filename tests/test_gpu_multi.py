"""-m gpu, needs >= 2 devices (skipped on single-GPU boxes): the multi-GPU paths over NCCL — force halos with
fracture, contact across ranks (surface all-gather + exact force exchange), and erosion of the contact surface
across ranks — against one oracle run."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _global_setup(mode):
    from hakai_fem_b200.model_setup import prepare
    from hakai_fem_b200.mesh import ImpactDeck, StretchDeck, steel
    if mode == "contact":
        return prepare(ImpactDeck(plate=(40, 40, 6), proj=(9, 9, 9)).build_model()), dict(contact_myu=0.25), 40
    if mode == "erosion":                                  # brittle plate: deletions expose new contact faces
        model = ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0).build_model()
        model.MATERIAL[0].ductile = np.array([[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]])
        return prepare(model), {}, 400
    deck = StretchDeck(12, 10, 24, jitter=0.1, strain_per_step=8e-4,
                       material=steel("steel_Ductile", ductile=[[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]))
    return prepare(deck.build_model()), {}, 90


def _worker(rank, world, port, mode, q, engine_comm=False, device_erosion=False):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hakai_fem_b200.engine import Engine
        from hakai_fem_b200.multi import partition_model, SlabRunner
        gsetup, prm, n_steps = _global_setup(mode)
        dom = partition_model(gsetup, world)[rank]
        stream = torch.cuda.current_stream()

        def make(**p):
            e = Engine(**p)
            e.set_stream(stream.cuda_stream)
            return e
        run = SlabRunner.from_domain(make, dom, torch.device("cuda", rank), world, device=rank, engine_comm=engine_comm,
                                     device_erosion=device_erosion, **prm)
        if device_erosion:
            assert run.erosion_on_device == engine_comm
        nd = run.run(1, n_steps)
        d = run.engine.download()
        n_own = len(np.unique(dom.setup.model.elementmat))          # held nodes come first, ghosts after
        q.put((rank, dict(disp=d["disp"][:3 * n_own], eps=d["integ_eq_plastic_strain"], flag=d["element_flag"],
                          node_l2g=dom.node_l2g[:n_own], elem_l2g=dom.elem_l2g, hits=int(run.engine.counters()[1]), nd=nd,
                          deleted=dom.elem_l2g[run.engine.deleted_ids() - 1],
                          n_surf=len(run.contact.lists.surface_nodes) if run.contact else 0,
                          n_surf0=len(dom.contact.surface_nodes) if dom.contact else 0)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode,world,engine_comm,device_erosion",
                         [("fracture", 2, False, False), ("contact", 2, False, False), ("erosion", 2, False, False),
                          ("erosion", 1, False, False), ("fracture", 2, True, False), ("contact", 2, True, False),
                          ("erosion", 2, True, False), ("erosion", 1, True, False),
                          ("erosion", 2, True, True), ("erosion", 1, True, True), ("erosion", 2, False, True)])
def test_ranks_match_single_domain_oracle(mode, world, engine_comm, device_erosion):
    """world == 1 runs the whole domain-runner path (global maps, hk_apply_deleted, list rebuild, NCCL calls) on a
    single-GPU box, where the 2-rank cases are skipped.  engine_comm: the ENGINE's communicator runs every exchange
    (hk_comm_init / hk_comm_contact: halo send/recv, surface all-gather, limb all-reduce) inside hk_step_enqueue.
    device_erosion: exposed faces of elements deleted on ANY rank join the surfaces on the device (hk_comm_erosion); with
    engine_comm the whole run is ONE hk_step_enqueue(1, n_steps) per rank — no host between steps."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from hakai_fem_b200.model_setup import configure_engine
    from oracle.oracle_engine import OracleEngine
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = (29200 + (os.getpid() % 2000) + ("fracture", "contact", "erosion").index(mode) + 10 * world + (40 if engine_comm else 0)
            + (80 if device_erosion else 0))
    procs = [ctx.Process(target=_worker, args=(r, world, port, mode, q, engine_comm, device_erosion)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gsetup, prm, n_steps = _global_setup(mode)
    o = configure_engine(OracleEngine, gsetup, **prm)
    nd_ref = o.step(1, n_steps)
    ref = o.download()
    scale = np.abs(ref["disp"]).max()
    tol = 1e-7 if mode == "erosion" else 1e-9               # same bound as the single-GPU erosion case
    for rank, r in res:
        n = r["node_l2g"] - 1
        assert np.abs(ref["disp"].reshape(-1, 3)[n].reshape(-1) - r["disp"]).max() <= tol * scale, f"rank {rank} disp"
        e = r["elem_l2g"] - 1
        ip = (e[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        assert np.abs(ref["integ_eq_plastic_strain"][ip] - r["eps"]).max() <= tol * max(ref["integ_eq_plastic_strain"].max(), 1e-30)
        assert np.array_equal(ref["element_flag"][e], r["flag"])
    if mode == "contact":
        assert sum(r["hits"] for _, r in res) == o.counters()[1] > 0
    elif mode == "erosion":
        got = np.sort(np.concatenate([r["deleted"] for _, r in res]))
        assert len(got) > 0 and np.array_equal(got, np.sort(o.deleted_ids())), "deleted-element set differs"
        if not device_erosion:                                  # (static candidate lists never change)
            assert any(r["n_surf"] > r["n_surf0"] for _, r in res), "contact surface never grew"
    else:
        assert sum(r["nd"] for _, r in res) == nd_ref > 0


def _worker_host(rank, world, port, deck_path, outdir, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hakai_fem_b200.host import hakai_distributed
        _, frames = hakai_distributed(deck_path, outdir, output_num=6, verbose=False)      # CUDA engine, NCCL
        q.put((rank, len(frames)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [1, 2])
def test_distributed_host_driver_frames_match_oracle_frames(world, tmp_path):
    """The `torchrun -m hakai_fem_b200.host deck.inp` path (partition, NCCL exchanges, rank-0 frame assembly from the
    device-side nodal sums) against the single-domain driver running the CPU oracle on the same deck."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from hakai_fem_b200.host import hakai
    from hakai_fem_b200.mesh import ImpactDeck
    from oracle.oracle_engine import OracleEngine
    from .test_multi_gloo import _vtk_sections
    deck_path = str(tmp_path / "deck.inp")
    ImpactDeck(plate=(8, 8, 2), proj=(3, 3, 3), v0=-900.0, n_steps=120,
               plate_ductile=[[0.02, 0.0, 30.0], [0.015, 0.4, 30.0]]).write_inp(deck_path)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 2000) + world
    procs = [ctx.Process(target=_worker_host, args=(r, world, port, deck_path, str(tmp_path / "multi"), q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    eng, ref = hakai(deck_path, str(tmp_path / "single"), engine_cls=OracleEngine, output_num=6, verbose=False)
    assert len(eng.deleted_ids()) > 0 and res[0] == len(ref) == 7
    for f in ref:
        a = _vtk_sections(f)
        b = _vtk_sections(os.path.join(str(tmp_path / "multi"), os.path.basename(f)))
        assert list(a) == list(b)
        for k in a:
            assert a[k].shape == b[k].shape, (os.path.basename(f), k)          # same live cells: same deleted set
            if k in ("CELLS", "CELL_TYPES", "POINTS"):
                assert np.array_equal(a[k], b[k]), (os.path.basename(f), k)
            elif k != "TRIAX_STRESS":                            # a ratio that is rounding noise while the plate is at rest
                scale = max(np.abs(a[k]).max(), 1e-300)
                assert np.allclose(a[k], b[k], rtol=0, atol=1e-5 * scale), (os.path.basename(f), k, scale)


# ---- ghost-element partitions (first run over NCCL in round 2: profiles/r2_multi_nccl_n2.log)

def _worker_ghost(rank, world, port, kind, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hakai_fem_b200.engine import Engine
        from hakai_fem_b200.multi import partition_model_ghost, GhostRunner
        gsetup, prm, n_steps = _global_setup(kind)
        dom = partition_model_ghost(gsetup, world, only_rank=rank)[rank]
        stream = torch.cuda.current_stream()

        def make(**p):
            e = Engine(**p)
            e.set_stream(stream.cuda_stream)
            return e
        run = GhostRunner(make, dom, torch.device("cuda", rank), world=world, device=rank,
                          force_exchange="allreduce" if kind == "contact" else "allgather", **prm)
        nd = run.run(1, n_steps)
        d = run.engine.download()
        own_n, own_e = np.flatnonzero(dom.own_node), np.flatnonzero(dom.own_elem)
        ip = (own_e[:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        q.put((rank, dict(nodes=dom.node_l2g[own_n] - 1, elems=dom.elem_l2g[own_e] - 1, nd=nd,
                          disp=d["disp"].reshape(-1, 3)[own_n], eps=d["integ_eq_plastic_strain"][ip],
                          flag=d["element_flag"][own_e], deleted=run.deleted_global_ids())))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["fracture", "contact", "erosion"])
def test_ghost_partitions_are_bit_identical_to_one_gpu(kind):
    """Two GPUs with ghost-element partitions vs ONE GPU running the same CUDA kernels: array_equal on every field."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from hakai_fem_b200.engine import Engine
    from hakai_fem_b200.model_setup import configure_engine
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 2000) + ("fracture", "contact", "erosion").index(kind)
    procs = [ctx.Process(target=_worker_ghost, args=(r, world, port, kind, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    gsetup, prm, n_steps = _global_setup(kind)
    one = configure_engine(Engine, gsetup, **prm)
    nd_ref = one.step(1, n_steps)
    ref = one.download()
    assert sum(res[r]["nd"] for r in range(world)) == nd_ref
    assert np.array_equal(np.sort(np.concatenate([res[r]["deleted"] for r in range(world)])), np.sort(one.deleted_ids()))
    for r in range(world):
        a = res[r]
        ip = (a["elems"][:, None] * 8 + np.arange(8)[None, :]).reshape(-1)
        assert np.array_equal(a["disp"], ref["disp"].reshape(-1, 3)[a["nodes"]]), (kind, r)
        assert np.array_equal(a["eps"], ref["integ_eq_plastic_strain"][ip]), (kind, r)
        assert np.array_equal(a["flag"], ref["element_flag"][a["elems"]]), (kind, r)


# ---- the engine's own NCCL communicator (hk_comm_init): all steps of a call enqueued by the library ------------------
def _worker_engine_comm(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hakai_fem_b200.engine import Engine
        from hakai_fem_b200.multi import slab_parity_check
        stream = torch.cuda.current_stream()

        def make(**p):
            e = Engine(**p)
            e.set_stream(stream.cuda_stream)
            return e
        res = {}
        for mode in (True, False):        # engine-owned communicator vs host-driven torch.distributed P2P
            res[mode] = slab_parity_check(make, torch.device("cuda", rank), rank, world, engine_comm=mode, device=rank)
        q.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_engine_owned_communicator_matches_unpartitioned_run():
    """hk_comm_init + hk_step_enqueue(t, n): n multi-GPU steps (pack -> ncclSend/ncclRecv on a side stream -> split
    step) enqueued by the library in ONE call, against the same mesh unpartitioned on one GPU: identical deleted set,
    fields <= 1e-10, shared node layer bit-identical on both ranks; and the same verdict through the host-driven path."""
    world = 2
    if torch.cuda.device_count() < world:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30100 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_engine_comm, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for mode in (True, False):
        r0 = res[0][mode]
        assert r0["ok"] and r0["interface_bitwise"] and r0["deleted_equal"] and r0["n_deleted"] > 0, (mode, r0)


# ---- SURVEY 8f.1 across ranks: the reference's eroding decks, ONE hk_step_enqueue(1, n) per rank, no host in between
def _worker_reference_deck(rank, world, port, name, n_steps, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        from hakai_fem_b200.engine import Engine
        from tests.test_multi_gloo import reference_deck_rank
        stream = torch.cuda.current_stream()

        def make(**p):
            e = Engine(device=rank, **p)
            e.set_stream(stream.cuda_stream)
            return e
        q.put((rank, reference_deck_rank(name, n_steps, rank, world, make, torch.device("cuda", rank), True)))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,n_steps,expect_deleted,world", [("bullet_impact", 3000, 13, 1), ("bullet_impact", 3000, 13, 2),
                                                               ("metal_cutting", 3000, 30, 2)])
def test_reference_deck_erodes_across_ranks_on_the_device(name, n_steps, expect_deleted, world):
    """bullet-impact.inp / metal-cutting.inp in element blocks over `world` GPUs: every rank logs its deletions, the
    engines all-gather them (NCCL, hk_comm_erosion) and replay them on the device — surfaces, deleted set, hit count and
    fields equal the oracle's unpartitioned run, and the whole run is one hk_step_enqueue(1, n_steps) per rank."""
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    from tests.test_multi_gloo import check_reference_deck_ranks
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 30300 + (os.getpid() % 2000) + 3 * world + (1 if name == "metal_cutting" else 0)
    procs = [ctx.Process(target=_worker_reference_deck, args=(r, world, port, name, n_steps, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=900) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    check_reference_deck_ranks(name, n_steps, expect_deleted, res)
