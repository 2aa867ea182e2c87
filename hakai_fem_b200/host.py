"""Host driver: the Python twin of `hakai(fname)` (HAKAI-v0.0.2/Julia/HAKAI_j.jl:81-978) with the loop
body replaced by the engine's C ABI.

    python -m hakai_fem_b200.host path/to/deck.inp [outdir]

Same entry semantics as `julia HAKAI_j.jl deck.inp`: reads the Abaqus deck, runs floor(end_time/d_time) steps
of fixed size, writes `file000.vtk` before the loop and one ASCII legacy VTK frame every
d_out = floor(time_num/output_num) steps (output_num = 100, J2:471-472), in the reference's field order and
`%1.6e` format.  Between frames the engine advances `d_out` steps on the GPU with no host round trip.  At a frame
the nodal averages come from the device (`hk_node_output`, 15 doubles per node) instead of downloading the
Gauss-point state (112 doubles per element) and averaging on the host; `node_output="host"` keeps the reference's
route through `cal_node_stress_strain`.  `vtk_format="binary"` writes the same sections as legacy-VTK BINARY
(big-endian float32/int32 — the ASCII file's `%1.6e` carries the same 7 digits) at about a fifth of the size.
(The reference's frame-buffer overflow for step counts that are not a multiple of output_num, SURVEY §3.1, is not
reproduced: frames beyond 100 are simply written.)
"""
from __future__ import annotations

import math
import os
import sys
import time

import numpy as np

from .inp import read_inp_file
from .model_setup import prepare, configure_engine


def cal_node_stress_strain(nNode, elementmat, integ_num, integ):
    """J2:3408-3486: Gauss points -> element mean -> nodal mean over incident elements -> von Mises."""
    nE = elementmat.shape[1]
    st = np.asarray(integ["integ_stress"]).T.reshape(nE, integ_num, 6).sum(axis=1) / integ_num
    sn = np.asarray(integ["integ_strain"]).T.reshape(nE, integ_num, 6).sum(axis=1) / integ_num
    ep = integ["integ_eq_plastic_strain"].reshape(nE, integ_num).sum(axis=1) / integ_num
    tx = integ["integ_triax_stress"].reshape(nE, integ_num).sum(axis=1) / integ_num
    node_stress = np.zeros((nNode, 6))
    node_strain = np.zeros((nNode, 6))
    node_ep = np.zeros(nNode)
    node_tx = np.zeros(nNode)
    inc = np.zeros(nNode)
    for k in range(8):
        idx = elementmat[k] - 1
        np.add.at(node_stress, idx, st)
        np.add.at(node_strain, idx, sn)
        np.add.at(node_ep, idx, ep)
        np.add.at(node_tx, idx, tx)
        np.add.at(inc, idx, 1.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        node_stress /= inc[:, None]
        node_strain /= inc[:, None]
        node_ep /= inc
        node_tx /= inc
    ox, oy, oz, txy, tyz, txz = (node_stress[:, i] for i in range(6))
    mises = np.sqrt(0.5 * ((ox - oy) ** 2 + (oy - oz) ** 2 + (ox - oz) ** 2 + 6 * (txy ** 2 + tyz ** 2 + txz ** 2)))
    return dict(node_stress=node_stress, node_strain=node_strain, node_eq_plastic_strain=node_ep,
                node_mises_stress=mises, node_triax_stress=node_tx)


def _flush(a):
    a = np.array(a, dtype=np.float64, copy=True)
    a[np.abs(a) < 1e-16] = 0.0                                   # J2:3531-3559
    return a


def write_vtk(outdir, index, coordmat, elementmat, element_flag, disp, velo, nd, binary=False):
    """J2:3517-3717 — ASCII legacy VTK, same sections, order and `%1.6e` format (binary=True: same sections as
    legacy BINARY)."""
    if binary:
        return _write_vtk_binary(outdir, index, coordmat, elementmat, element_flag, disp, velo, nd)
    nNode = coordmat.shape[1]
    disp3 = _flush(disp.reshape(nNode, 3))
    velo3 = _flush(velo.reshape(nNode, 3))
    ns, ne = _flush(nd["node_stress"]), _flush(nd["node_strain"])
    mises, eps, triax = (_flush(nd[k]) for k in ("node_mises_stress", "node_eq_plastic_strain", "node_triax_stress"))
    os.makedirs(outdir, exist_ok=True)
    fname = os.path.join(outdir, "file%03d.vtk" % index)
    live = np.flatnonzero(np.asarray(element_flag) == 1)
    with open(fname, "w") as f:
        f.write("# vtk DataFile Version 2.0\nTest\nASCII\nDATASET UNSTRUCTURED_GRID\n")
        f.write("POINTS %d float\n" % nNode)
        np.savetxt(f, coordmat.T, fmt="%1.6e")
        f.write("CELLS %d %d\n" % (len(live), len(live) * 9))
        cells = np.concatenate([np.full((len(live), 1), 8, np.int64), (elementmat[:, live] - 1).T], axis=1)
        np.savetxt(f, cells, fmt="%d")
        f.write("CELL_TYPES %d\n" % len(live))
        f.write("12\n" * len(live))
        f.write("POINT_DATA %d\n" % nNode)
        f.write("VECTORS DISPLACEMENT float\n")
        np.savetxt(f, disp3, fmt="%1.6e")

        def scalar(name, v):
            f.write("SCALARS %s float 1\nLOOKUP_TABLE default\n" % name)
            np.savetxt(f, v, fmt="%1.6e")
        for name, v in (("Vx", velo3[:, 0]), ("Vy", velo3[:, 1]), ("Vz", velo3[:, 2]),
                        ("E11", ne[:, 0]), ("E22", ne[:, 1]), ("E33", ne[:, 2]), ("E12", ne[:, 3]), ("E23", ne[:, 4]),
                        ("E13", ne[:, 5]), ("EQ_PSTRAIN", eps),
                        ("S11", ns[:, 0]), ("S22", ns[:, 1]), ("S33", ns[:, 2]), ("S12", ns[:, 3]), ("S23", ns[:, 4]),
                        ("S13", ns[:, 5]), ("MISES_STRESS", mises), ("TRIAX_STRESS", triax)):
            scalar(name, v)
    return fname


_SCALARS = (("E11", "node_strain", 0), ("E22", "node_strain", 1), ("E33", "node_strain", 2), ("E12", "node_strain", 3),
            ("E23", "node_strain", 4), ("E13", "node_strain", 5), ("EQ_PSTRAIN", "node_eq_plastic_strain", None),
            ("S11", "node_stress", 0), ("S22", "node_stress", 1), ("S33", "node_stress", 2), ("S12", "node_stress", 3),
            ("S23", "node_stress", 4), ("S13", "node_stress", 5), ("MISES_STRESS", "node_mises_stress", None),
            ("TRIAX_STRESS", "node_triax_stress", None))


def _write_vtk_binary(outdir, index, coordmat, elementmat, element_flag, disp, velo, nd):
    nNode = coordmat.shape[1]
    os.makedirs(outdir, exist_ok=True)
    fname = os.path.join(outdir, "file%03d.vtk" % index)
    live = np.flatnonzero(np.asarray(element_flag) == 1)
    disp3, velo3 = _flush(disp.reshape(nNode, 3)), _flush(velo.reshape(nNode, 3))

    def f32(a):
        return np.ascontiguousarray(a, dtype=">f4").tobytes()
    with open(fname, "wb") as f:
        f.write(b"# vtk DataFile Version 2.0\nTest\nBINARY\nDATASET UNSTRUCTURED_GRID\n")
        f.write(b"POINTS %d float\n" % nNode + f32(coordmat.T) + b"\n")
        cells = np.concatenate([np.full((len(live), 1), 8, np.int64), (elementmat[:, live] - 1).T], axis=1)
        f.write(b"CELLS %d %d\n" % (len(live), len(live) * 9) + cells.astype(">i4").tobytes() + b"\n")
        f.write(b"CELL_TYPES %d\n" % len(live) + np.full(len(live), 12, ">i4").tobytes() + b"\n")
        f.write(b"POINT_DATA %d\n" % nNode)
        f.write(b"VECTORS DISPLACEMENT float\n" + f32(disp3) + b"\n")
        for c, name in enumerate(("Vx", "Vy", "Vz")):
            f.write(b"SCALARS %s float 1\nLOOKUP_TABLE default\n" % name.encode() + f32(velo3[:, c]) + b"\n")
        for name, key, col in _SCALARS:
            v = _flush(nd[key] if col is None else nd[key][:, col])
            f.write(b"SCALARS %s float 1\nLOOKUP_TABLE default\n" % name.encode() + f32(v) + b"\n")
    return fname


def hakai(fname, outdir="temp", engine_cls=None, output_num=100, write_frames=True, verbose=True,
          node_output="device", vtk_format="ascii", checkpoint=None, checkpoint_frames=10, resume=None,
          contact_setup=None, **params):
    """hakai(fname), J2:81.  Returns the engine (state on the GPU) and the list of frame files.
    checkpoint: file rewritten every `checkpoint_frames` frames (checkpoint.py); resume: checkpoint to continue
    from (frames already written are kept; numbering continues).
    contact_setup: "device" — faces, exterior faces, pair lists and exposed-face twins are built by hk_build_contact
    (radix sort on the GPU; the reference's all-pairs face match, J2:1996-2084, is what keeps it from large decks) —
    or "host" (the NumPy mirror in model_setup.py; same tables, tests/parity_cases.py::case_build_contact).  Default:
    "device" when the engine class can (`builds_contact`: the CUDA engine), else "host" (the CPU oracle)."""
    if node_output not in ("device", "host") or vtk_format not in ("ascii", "binary"):
        raise ValueError("node_output: device|host, vtk_format: ascii|binary")
    if engine_cls is None:
        from .engine import Engine as engine_cls               # the CUDA engine; raises without a GPU
    if contact_setup is None:
        contact_setup = "device" if getattr(engine_cls, "builds_contact", False) else "host"
    if contact_setup not in ("device", "host"):
        raise ValueError("contact_setup: device|host")
    log = print if verbose else (lambda *a, **k: None)
    model = read_inp_file(fname)
    log("nNode:", model.nNode)
    log("nElement:", model.nElement)
    log("contact_flag:", model.contact_flag)
    setup = prepare(model, contact=contact_setup)
    log("mass_scaling:", model.mass_scaling)
    log("time_num:", setup.time_num)
    log("elementMinSize:", setup.elementMinSize)
    log("elementMaxSize:", setup.elementMaxSize)
    eng = configure_engine(engine_cls, setup, **params)
    n_total = int(math.floor(setup.time_num))                  # `for t = 1 : time_num` with Float64 time_num
    d_out = int(math.floor(setup.time_num / output_num))       # J2:472
    frames = []

    def frame(index):
        if node_output == "device":
            d = eng.download(fields=("disp", "velo", "element_flag"))
            nd = eng.node_output()
        else:
            d = eng.download()
            nd = cal_node_stress_strain(model.nNode, model.elementmat, 8, d)
        frames.append(write_vtk(outdir, index, model.coordmat, model.elementmat, d["element_flag"], d["disp"],
                                d["velo"], nd, binary=(vtk_format == "binary")))
    t, i_out = 0, 1
    if resume is not None:
        from .checkpoint import load_checkpoint
        t = load_checkpoint(eng, resume)
        i_out = t // d_out + 1 if d_out > 0 else 1
        log("resumed at step", t)
    elif write_frames:
        frame(0)                                                # J2:478-480
    t0 = time.perf_counter()
    while t < n_total:
        n = min(d_out - t % d_out, n_total - t) if d_out > 0 else n_total - t
        ndel = eng.step(t + 1, n)
        t += n
        if ndel:
            flags = eng.download(fields=("element_flag",))["element_flag"]
            log("Element deleted:", int(flags.sum()), "/", model.nElement)       # J2:736
        if d_out > 0 and t % d_out == 0:                         # rem(t, d_out) == 0, J2:932
            if write_frames:
                frame(i_out)
            if checkpoint is not None and i_out % checkpoint_frames == 0:     # also when no frames are written
                from .checkpoint import save_checkpoint
                save_checkpoint(eng, checkpoint, t)
            i_out += 1
        log("\r%.4e / %.4e     " % (t * setup.d_time, model.end_time), end="")
    log("\n%.3f seconds for %d steps" % (time.perf_counter() - t0, n_total))
    return eng, frames


def hakai_distributed(fname, outdir="temp", engine_cls=None, torch_device=None, output_num=100, write_frames=True,
                      verbose=True, vtk_format="ascii", partition="halo", **params):
    """hakai(fname) on several GPUs: one process per GPU inside an initialised torch.distributed group
    (`torchrun --nproc-per-node N -m hakai_fem_b200.host deck.inp`).  Every rank reads the deck, keeps its element
    block (multi.partition_model) and steps it with force halos / contact exchange (multi.SlabRunner); at a frame
    the ranks send disp, velo, flags and their undivided nodal sums (`hk_node_output`, raw) to rank 0, which adds
    the shares of interface nodes, divides by the incidence count (J2:3456-3469) and writes the same VTK file as
    the single-GPU driver.  partition="ghost" uses ghost-element partitions instead (multi.partition_model_ghost: no
    cross-rank summation order anywhere): every rank then holds complete nodal sums for the nodes of its own elements, the run and the
    frames are bit-identical to the single-GPU run for any number of ranks.
    Returns (runner, frames); frames is empty on ranks > 0."""
    if partition not in ("halo", "ghost"):
        raise ValueError("partition: halo | ghost")
    import torch
    import torch.distributed as dist
    from .multi import partition_model, SlabRunner, partition_model_ghost, GhostRunner
    rank, world = dist.get_rank(), dist.get_world_size()
    # frames travel as pickled host arrays: keep them off the GPU / NCCL (a gloo side group when the main one is NCCL)
    obj_group = dist.new_group(backend="gloo") if dist.get_backend() == "nccl" else None
    if engine_cls is None:
        from .engine import Engine
        torch_device = torch.device("cuda", torch.cuda.current_device())
        stream = torch.cuda.current_stream()

        def engine_cls(**p):                                    # engine on torch's stream: ordered with the NCCL calls
            e = Engine(**p)
            e.set_stream(stream.cuda_stream)
            return e
        params.setdefault("device", torch_device.index)
    log = print if (verbose and rank == 0) else (lambda *a, **k: None)
    model = read_inp_file(fname)
    setup = prepare(model)
    log("nNode:", model.nNode, " nElement:", model.nElement, " contact_flag:", model.contact_flag, " ranks:", world)
    ghost = partition == "ghost"
    if ghost:
        dom = partition_model_ghost(setup, world, only_rank=rank)[rank]
        runner = GhostRunner(engine_cls, dom, torch_device, world=world, **params)
        sel_n, sel_e = np.flatnonzero(dom.own_node), np.flatnonzero(dom.own_elem)
    else:
        dom = partition_model(setup, world, only_rank=rank)[rank]
        # over NCCL the engine runs every exchange with its own communicator (hk_comm_init / hk_comm_contact); over gloo
        # (CPU ranks with the host-compiled kernels) the host drives them through torch.distributed
        on_nccl = dist.get_backend() == "nccl"      # the engine exchanges by itself ...
        # ... and keeps eroding surfaces current on the device, unless that would make every step exchange far more
        # nodes than the surfaces hold: the static lists of hk_comm_erosion cover ALL nodes of the instances in contact
        dev_er = on_nccl and dom.contact_all is not None and (
            len(dom.contact_all.surface_nodes) <= max(4 * len(dom.contact.surface_nodes), 200_000))
        runner = SlabRunner.from_domain(engine_cls, dom, torch_device, world, engine_comm=on_nccl,
                                        device_erosion=dev_er, **params)
        n_held = len(np.unique(dom.setup.model.elementmat))     # local ids 1..n_held are nodes of own elements
        sel_n, sel_e = np.arange(n_held), np.arange(dom.setup.model.nElement)
    eng = runner.engine
    g_nodes = dom.node_l2g[sel_n] - 1
    n_total = int(math.floor(setup.time_num))
    d_out = int(math.floor(setup.time_num / output_num))
    frames = []

    def frame(index):
        d = eng.download(fields=("disp", "velo", "element_flag"))
        nd = eng.node_output(raw=True)      # ghost partitions: the sums of own nodes are already complete (and in
        part = dict(nodes=g_nodes, elems=dom.elem_l2g[sel_e] - 1, flag=d["element_flag"][sel_e],       # global order)
                    disp=d["disp"].reshape(-1, 3)[sel_n], velo=d["velo"].reshape(-1, 3)[sel_n],
                    **{k: nd[k][sel_n] for k in ("node_stress", "node_strain", "node_eq_plastic_strain",
                                                  "node_triax_stress", "inc_num")})
        parts = [None] * world if rank == 0 else None
        dist.gather_object(part, parts, dst=0, group=obj_group)
        if rank != 0:
            return
        nN = model.nNode
        disp, velo = np.zeros((nN, 3)), np.zeros((nN, 3))
        flag = np.zeros(model.nElement, np.int64)
        acc = dict(node_stress=np.zeros((nN, 6)), node_strain=np.zeros((nN, 6)), node_eq_plastic_strain=np.zeros(nN),
                   node_triax_stress=np.zeros(nN), inc_num=np.zeros(nN))
        for p in reversed(parts):                               # lowest rank last: the owner's copy of shared nodes wins
            disp[p["nodes"]] = p["disp"]
            velo[p["nodes"]] = p["velo"]
            flag[p["elems"]] = p["flag"]
        for p in (reversed(parts) if ghost else parts):         # halo: ascending rank = ascending element blocks
            for k in acc:
                if ghost:
                    acc[k][p["nodes"]] = p[k]                   # complete sums: copy (lowest rank last, as for disp)
                else:
                    acc[k][p["nodes"]] += p[k]
        inc = acc.pop("inc_num")
        with np.errstate(divide="ignore", invalid="ignore"):
            ns, ne = acc["node_stress"] / inc[:, None], acc["node_strain"] / inc[:, None]
            ep, tx = acc["node_eq_plastic_strain"] / inc, acc["node_triax_stress"] / inc
        ox, oy, oz, txy, tyz, txz = (ns[:, i] for i in range(6))
        mises = np.sqrt(0.5 * ((ox - oy) ** 2 + (oy - oz) ** 2 + (ox - oz) ** 2 + 6 * (txy ** 2 + tyz ** 2 + txz ** 2)))
        nodal = dict(node_stress=ns, node_strain=ne, node_eq_plastic_strain=ep, node_mises_stress=mises,
                     node_triax_stress=tx)
        frames.append(write_vtk(outdir, index, model.coordmat, model.elementmat, flag, disp.reshape(-1),
                                velo.reshape(-1), nodal, binary=(vtk_format == "binary")))
    if write_frames:
        frame(0)
    t0 = time.perf_counter()
    t, i_out = 0, 1
    n_deleted = 0
    while t < n_total:
        n = min(d_out - t % d_out, n_total - t) if d_out > 0 else n_total - t
        nd_local = runner.run(t + 1, n, frame_at_end=write_frames and d_out > 0 and (t + n) % d_out == 0)
        t += n
        tot = torch.tensor([nd_local], dtype=torch.int64, device=torch_device)
        dist.all_reduce(tot)
        if int(tot.item()):
            n_deleted += int(tot.item())
            log("Element deleted:", model.nElement - n_deleted, "/", model.nElement)
        if write_frames and d_out > 0 and t % d_out == 0:
            frame(i_out)
            i_out += 1
        log("\r%.4e / %.4e     " % (t * setup.d_time, model.end_time), end="")
    log("\n%.3f seconds for %d steps on %d ranks" % (time.perf_counter() - t0, n_total, world))
    return runner, frames


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if not argv:
        raise SystemExit("usage: [torchrun --nproc-per-node N -m | python -m] hakai_fem_b200.host deck.inp [outdir]")
    outdir = argv[1] if len(argv) > 1 else "temp"
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:              # launched by torchrun: one rank per GPU
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        try:
            hakai_distributed(argv[0], outdir)
        finally:
            dist.destroy_process_group()
    else:
        hakai(argv[0], outdir)


if __name__ == "__main__":
    main()
