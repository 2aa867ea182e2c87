#!/usr/bin/env python
"""bench.py — element-steps/s of the HAKAI time-step engine on B200 (BASELINE.json's metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W16|F16D|I8|B1|...] [--impl reference]

A "step" is one explicit time step of the hot path (contact kernels when the deck has contact, nodal
gather/update/BC/kinematics kernel, hex8 elastoplastic element kernel, deletion flush) on a synthetic deck of
SURVEY §8(d):
  W16 (default): 252x252x252 = 16 003 008 hex per GPU, elastoplastic steel, uniform stretch, jittered interior nodes —
      the mesh the north star's roofline target is quoted on and the per-GPU shard of the weak-scaling config.
  F16D: the same block with the ductile-damage table [0.03 0 30; 0.02 0.3 30] (BASELINE configs[2]): the timed window
      starts just before the first deletion and the live-element count falls inside it.
  I8:   two-instance impact, 400x400x48 plate + 68^3 projectile, frictionless penalty contact (BASELINE configs[3]).
  B1:   the 50x50x400 1 M-hex bar (BASELINE configs[1]).
Whatever --warmup says, the engine is advanced UNTIMED until the deck is in the regime the metric is quoted on (every
Gauss point yielding; F16D: deletion imminent; I8: bodies in contact) — `config.regime` records it.
State (>= 7 GB) is far larger than L2, so no L2 flush is needed between steps.
Prints ONE JSON line (see DESIGN.md §6 for every key).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_ELEMENT = 1904       # element kernel, per element-step
ALG_BYTES_NODAL = 224          # nodal update, per node-step
ALG_BYTES_STEP = 2128          # total per element-step (nN/nE -> 1)
DUCTILE_ROWS = [[0.03, 0.0, 30.0], [0.02, 0.3, 30.0]]


def make_deck(workload, strain_per_step=None):
    """-> (deck, kind) with kind in stretch | ductile | impact."""
    from hakai_fem_b200.mesh import StretchDeck, ImpactDeck, steel
    kw = {} if strain_per_step is None else dict(strain_per_step=strain_per_step)
    if workload in ("I8", "I1", "I0"):          # I1 / I0: 1 M / 125 k-element cuts of I8 (CPU samples, debugging)
        plate, proj = {"I8": ((400, 400, 48), (68, 68, 68)), "I1": ((200, 200, 24), (34, 34, 34)),
                       "I0": ((100, 100, 12), (17, 17, 17))}[workload]
        return ImpactDeck(plate=plate, proj=proj), "impact"
    ductile = workload.endswith("D")
    base = workload[:-1] if ductile else workload
    if base in ("W16", "F16"):
        n = (252, 252, 252)
    elif base == "B1":
        n = (50, 50, 400)
    elif base == "S1":          # 1 M-element slab of W16 (CPU sample)
        n = (252, 252, 16)
    elif base.startswith("N"):   # Nnx,ny,nz
        n = tuple(int(v) for v in base[1:].split(","))
    else:
        raise SystemExit(f"unknown workload {workload}")
    mat = steel("steel_Ductile", ductile=DUCTILE_ROWS) if ductile else steel()
    jitter = 0.0 if base == "B1" else 0.05
    return (StretchDeck(n[0], n[1], n[2], h=1.0, material=mat, jitter=jitter, n_steps=1.0e6, jitter_by_layer=True, **kw),
            "ductile" if ductile else "stretch")


def prepare_setup(deck, contact="host"):
    """contact="device": the contact tables are built by hk_build_contact on the GPU (CUDA engine only)."""
    from hakai_fem_b200.model_setup import prepare
    model = deck.build_model()
    vol = None
    if getattr(deck, "jitter", None) == 0.0:
        vol = np.full(model.nElement, deck.h ** 3)
    return prepare(model, elementVolume=vol, contact=contact)


def deck_text(deck, kind, myu=0.0):
    if kind == "impact":
        return (f"impact: plate {deck.plate[0]}x{deck.plate[1]}x{deck.plate[2]} (alum, elastoplastic + ductile) + projectile "
                f"{deck.proj[0]}x{deck.proj[1]}x{deck.proj[2]} (lead) at {deck.v0:g} m/s, "
                f"{'frictionless penalty contact' if myu == 0.0 else f'penalty contact with friction mu = {myu:g}'}")
    return (f"{deck.nx}x{deck.ny}x{deck.nz} hex8, steel {'elastoplastic + ductile damage' if kind == 'ductile' else 'elastoplastic'}"
            f", uniform stretch {deck.strain_per_step:g}/step, jitter {deck.jitter}")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks and throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_sample_for(workload):
    if workload.startswith("I"):
        return "I0"
    return "S1D" if workload.endswith("D") else "S1"


def cpu_oracle_rate(sample_workload, warmup, steps, threads=None, contact_myu=0.0):
    """Times the CPU oracle (C++ restatement of HAKAI_j.jl's loop, OpenMP) on a bounded sample."""
    from hakai_fem_b200.model_setup import configure_engine
    from oracle.oracle_engine import OracleEngine
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
    deck, kind = make_deck(sample_workload)
    st = prepare_setup(deck)
    prm = dict(contact_myu=contact_myu) if kind == "impact" else {}
    eng = configure_engine(OracleEngine, st, **prm)
    if threads:                 # torchrun exports OMP_NUM_THREADS=1 and libgomp may have read it already: set it directly
        try:
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(threads))
        except OSError:
            pass
    if kind == "impact":
        warmup = max(warmup, 12)        # the 0.1 h gap closes after 10 steps: time steps that do contact work
    eng.step(1, warmup)
    t0 = time.perf_counter()
    eng.step(warmup + 1, steps)
    dt = time.perf_counter() - t0
    nE = st.model.nElement
    eng.close()
    return nE * steps / dt, nE, dt


def run_reference(args, emit):
    """--impl reference: the reference's CPU algorithm (oracle port; Julia is not installed) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample or cpu_sample_for(args.workload)
    rate, nE, dt = cpu_oracle_rate(sample, args.warmup, args.steps, threads=cores, contact_myu=args.contact_myu)
    line = {
        "impl": "reference", "metric": "element-steps/sec (hex8 elastoplastic)", "value": rate,
        "unit": "element-steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "sample": sample, "elements_in_sample": nE},
        "cpu_baseline": {"value": rate, "unit": "element-steps/s", "cores": cores, "kind": "port",
                         "sample": f"{sample}: {nE} elements of the same deck recipe, {args.warmup} warm-up + "
                                   f"{args.steps} timed steps, OpenMP oracle (our C++ restatement of HAKAI_j.jl; Julia is "
                                   f"not installed), {cores} threads"},
        "e2e": {"value": rate, "unit": "element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def bind_to_gpu_numa_node(local_rank):
    """N > 1: run this rank on the cores next to its GPU (NVML's ideal CPU affinity), so that the pinned host arrays of
    the e2e leg are first-touched on that NUMA node and the GPU's DMA does not cross the socket link.  Round 1: eight ranks
    uploading from wherever the scheduler had put them took 0.68 s for what one rank does in 0.29 s.  Returns a
    description for the JSON line (or why nothing was done); never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        ideal = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus = sorted(ideal & allowed)
        if not cpus:
            return "NVML affinity has no CPU this process may use: unchanged"
        if len(cpus) == len(allowed):
            return f"NVML affinity = all {len(allowed)} allowed CPUs (one NUMA node): unchanged"
        os.sched_setaffinity(0, cpus)
        return f"{len(cpus)} of {len(allowed)} allowed CPUs (NVML ideal affinity of GPU {local_rank}: {cpus[0]}-{cpus[-1]})"
    except Exception as ex:          # no NVML, container without the permission, ...
        return f"unchanged ({type(ex).__name__}: {ex})"


def main():
    # libraries (NCCL, torchrun) print to stdout; the contract is ONE JSON line there, so everything else goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(os.dup(2), "w")

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--workload", default="W16")
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cpu-sample", default=None)
    ap.add_argument("--cpu-steps", type=int, default=40, help="timed steps of the CPU baseline sample (~10 s at 16 threads)")
    ap.add_argument("--strain-per-step", type=float, default=None,
                    help="override the deck's stretch rate (1e-6: purely elastic run, SURVEY §8d; disables the regime gate)")
    ap.add_argument("--contact-myu", type=float, default=0.0,
                    help="I8: friction coefficient (north star: frictionless; 0.25 = the reference's v0.0.2 default)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the partition parity check")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, emit)
        return

    import torch
    import torch.distributed as dist
    from hakai_fem_b200.engine import Engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    stream = torch.cuda.current_stream()

    pcheck = None
    if world > 1 and not args.no_parity:
        from hakai_fem_b200.multi import slab_parity_check

        def make_checked(**p):
            e_ = Engine(**p)
            e_.set_stream(stream.cuda_stream)
            return e_
        pcheck = slab_parity_check(make_checked, torch.device("cuda", local_rank), rank, world, engine_comm=True,
                                   device=local_rank)
        if rank == 0:
            print(f"[parity_check] {pcheck}", file=sys.stderr)

    from hakai_fem_b200.multi import slab_deck, SlabRunner
    deck, kind = make_deck(args.workload, strain_per_step=args.strain_per_step)
    if world > 1:
        if kind == "impact":
            raise SystemExit("workload I8 is a single-GPU bench line (multi-GPU contact: tests/test_gpu_multi.py)")
        # weak scaling (config W): the global mesh is nx x ny x (nz*world), split in z; every rank builds only its slab
        deck, nbrs, halos = slab_deck(deck, rank, world)
    else:
        nbrs, halos = [], []
    t_setup = time.perf_counter()
    st = prepare_setup(deck, contact="device" if kind == "impact" else "host")
    t_setup = time.perf_counter() - t_setup
    nE, nN = st.model.nElement, st.model.nNode
    prm = dict(contact_myu=args.contact_myu) if kind == "impact" else {}        # north star: frictionless

    def make_engine(**p):
        e_ = Engine(**p)
        e_.set_stream(stream.cuda_stream)
        return e_
    # N > 1: the engine owns the NCCL communicator and runs pack -> send/recv -> step for all steps of a call by itself
    t_setup0 = time.perf_counter()
    runner = SlabRunner(make_engine, st, nbrs, halos, torch.device("cuda", local_rank), sum_mass=True, rank=rank,
                        engine_comm=world > 1, device=local_rank, **prm)
    eng = runner.engine
    t_engine = time.perf_counter() - t_setup0

    def run_steps(t0, n, sync=True):
        """sync=False: only enqueue (the timed region ends with a CUDA event on the stream, BEFORE hk_sync's host work —
        copying and recording the deletion log of the steps)."""
        if world > 1:
            return runner.run(t0, n, sync=sync)   # hk_step_enqueue(t0, n): pack -> ncclSend/Recv -> split step, n times, in the engine
        eng.step_enqueue(t0, n)
        return eng.sync() if sync else 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def all_ranks(flag):
        if world == 1:
            return bool(flag)
        t_ = torch.tensor([1 if flag else 0], dtype=torch.int32, device="cuda")
        dist.all_reduce(t_, op=dist.ReduceOp.MIN)
        return bool(t_.item())

    # ---- untimed advance into the regime the metric is quoted on ----------------------------------------------------
    run_steps(1, args.warmup)
    t_next = args.warmup + 1
    regime, extra = "as given (--strain-per-step override)", 0
    if args.strain_per_step is None:
        if kind == "impact":
            regime = "in contact"
            ready = lambda: int(eng.counters()[1]) > 0
        elif kind == "ductile":
            regime = "plastic, first deletion imminent"
            eps_f = min(r[0] for r in DUCTILE_ROWS)
            ready = lambda: (lambda s: s["live_elements"] < nE or s["eps_max"] >= eps_f - 3.0 * deck.strain_per_step)(eng.state_summary())
        else:
            regime = "plastic (every Gauss point yielding)"
            ready = lambda: (lambda s: s["yielded_points"] == 8 * s["live_elements"])(eng.state_summary())
        chunk = 1 if kind == "ductile" else 4
        while not all_ranks(ready()) and extra < 400:
            run_steps(t_next, chunk)
            t_next += chunk
            extra += chunk
    s_start = eng.state_summary()

    # ---- device-resident throughput ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wait = time.time()
    while not sampler.rows and time.time() - t_wait < 3.0:      # first nvidia-smi sample before timing starts
        time.sleep(0.05)
    c0 = eng.counters()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # per-kernel times come from the SAME launches as `value`: hk_profile puts a CUDA-event pair on the engine's stream
    # around every launch of the timed region (2 x ~3 event records per 7 ms step, no synchronisation; a separate
    # profiling pass after the region read kernels up to 3 % faster or slower than the region itself — clocks drift)
    eng.profile(True)
    ev0.record(stream)
    run_steps(t_next, args.steps, sync=False)
    ev1.record(stream)
    eng.sync()
    kms, kn = eng.profile_read_ex()
    eng.profile(False)
    barrier()
    # live element-steps: elements alive at the start of each timed step.  The steps were enqueued in ONE call; the step
    # of every deletion comes from the engine's deletion log afterwards (hk_deleted_steps)
    dsteps = eng.deleted_steps()
    per_step_deleted = np.bincount(dsteps[dsteps >= t_next] - t_next, minlength=args.steps)[:args.steps]
    live_before = s_start["live_elements"] - np.concatenate([[0], np.cumsum(per_step_deleted)[:-1]])
    live_steps = int(live_before.sum())
    per_step_deleted = [int(v) for v in per_step_deleted] if kind == "ductile" else []
    t_next += args.steps
    ms = ev0.elapsed_time(ev1)
    print(f"[rank {rank}] host enqueue time of the timed steps: {getattr(runner, 'last_enqueue_s', 0.0) * 1e3:.1f} ms "
          f"of {ms:.1f} ms", file=sys.stderr)
    clocks = sampler.stop()
    c1 = eng.counters()
    launches = int(c1[3] - c0[3])
    s_end = eng.state_summary()
    t = torch.tensor([ms, float(live_steps)], dtype=torch.float64, device="cuda")
    if world > 1:
        tm = t.clone()
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        ms, live_steps = float(tm[0].item()), float(t[1].item())
    else:
        ms = float(t[0].item())
    value = live_steps / (ms * 1e-3)        # live element-steps of all ranks / max-over-ranks device time

    # ---- per-kernel timing (CUDA events on the engine's stream around every launch of the timed region) ----------
    n_prof = args.steps
    el_ms = kms[2] / max(kn[2], 1)
    nd_ms = kms[1] / max(n_prof, 1)          # per step (with halos the nodal update is two launches per step)
    ct_ms = kms[0] / max(n_prof, 1)
    ot_ms = kms[3] / max(n_prof, 1)
    ex_ms = kms[4] / max(n_prof, 1)          # ncclSend/ncclRecv group on the engine's side stream (overlaps the nodal update)
    dl_ms = kms[5] / max(n_prof, 1)
    gap_ms = kms[6] / max(n_prof, 1)         # idle time between consecutive launches of a step (device timestamps)
    per_rank = None
    if world > 1:          # kernel times of every rank: the exchange makes the slowest GPU set the pace
        tk = torch.tensor([el_ms, nd_ms, ot_ms, ex_ms], dtype=torch.float64, device="cuda")
        allk = [torch.zeros_like(tk) for _ in range(world)]
        dist.all_gather(allk, tk)
        per_rank = {"element_ms": [round(float(a[0]), 4) for a in allk], "nodal_ms_per_step": [round(float(a[1]), 4) for a in allk],
                    "halo_unpack_ms_per_step": [round(float(a[2]), 4) for a in allk],
                    "halo_exchange_ms_per_step": [round(float(a[3]), 4) for a in allk],
                    "note": "the exchange (pack -> ncclSend/ncclRecv on a side stream) overlaps the nodal update of the "
                            "non-interface nodes; what a step pays beyond the N = 1 kernels is the unpack and whatever of the "
                            "exchange outlasts that nodal update"}
    peaks = {}
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk_path):
        peaks = json.load(open(pk_path))
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    achieved = ALG_BYTES_ELEMENT * nE / (el_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        tr_path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tr_path):      # dram__bytes_read+write per element from the committed ncu --set full capture
            tj = json.load(open(tr_path))
            traffic = tj["element_kernel_dram_bytes_per_element"] * nE
            traffic_src = (f"profiles/{name}: ncu --set full on a {tj.get('mesh', '4 M-element')} mesh, per-element bytes "
                           f"scaled to this run's element count — a committed measurement, not taken in this run")
            break
    kname = "hk_element_simple_kernel" if os.environ.get("HK_ELEMENT_KERNEL") == "simple" else "hk_element_ring_kernel"
    step_ms = ms / args.steps
    roofline = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALG_BYTES_ELEMENT * nE, "avg_launch_ms": el_ms,
                "timing": "CUDA events on the engine's stream around every launch of the timed region (hk_profile); "
                          "same launches as `value`",
                "nodal_kernel": {"achieved": ALG_BYTES_NODAL * nN / (nd_ms * 1e-3) / 1e9, "ms_per_step": nd_ms},
                "deletion_pass_ms_per_step": dl_ms, "launch_gaps_ms_per_step": gap_ms,
                "whole_step": {"achieved": ALG_BYTES_STEP * nE / (step_ms * 1e-3) / 1e9,
                               "frac": ALG_BYTES_STEP * nE / (step_ms * 1e-3) / 1e9 / peak,
                               "note": "2128 B x ALL elements of the mesh (deleted ones still stream through the kernels)"}}
    fp_path = os.path.join(ROOT, "profiles", "r1_fp64_ops.json")
    if os.path.exists(fp_path):
        # secondary bound: FP64 issue.  Executed flops per element from the committed ncu instruction counts of the
        # plastic regime (the regime gate above guarantees it) x the live launch rate of this run
        fl = json.load(open(fp_path))["per_element"]["plastic"]["flops"]
        roofline["fp64"] = {"flops_per_element": fl, "achieved_tflops": fl * nE / (el_ms * 1e-3) / 1e12,
                            "source": "profiles/r1_fp64_ops.json (ncu instruction counts, same Gauss-point math)"}
    contact = None
    if kind == "impact":
        pinfo = [eng.contact_pair(c) for c in range(2)]
        contact = {"ms_per_step": ct_ms, "hits_per_step": float(c1[1] - c0[1]) / args.steps,
                   "tests_per_step": float(c1[2] - c0[2]) / args.steps,
                   "pairs": [dict(nn_i=len(p["c_nodes_i"]), nn_j=len(p["c_nodes_j"]), nTri=len(p["c_triangles_eleid"])) for p in pinfo],
                   "setup": "hk_build_contact: faces, orientation, exterior faces and exposed-face twins by radix sort on the GPU "
                            "(engine_setup_s includes it)",
                   "fixed_point_overflows": int(c1[5]),
                   "kernels": "hk_contact_{reset,bbox,cells,cull,narrow}_kernel per ordered pair + accumulator zeroing"}

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        fn, nip = 3 * nN, 8 * nE
        shapes = dict(disp=(fn,), velo=(fn,), disp_pre=(fn,), Q=(fn,), integ_stress=(nip, 6), integ_strain=(nip, 6),
                      integ_eq_plastic_strain=(nip,), integ_yield_stress=(nip,), integ_triax_stress=(nip,),
                      element_flag=(nE,), node_stress=(6, nN), node_strain=(6, nN), node_eq_plastic_strain=(nN,),
                      node_mises_stress=(nN,), node_triax_stress=(nN,), inc_num=(nN,))
        need = sum(int(np.prod(v)) * 8 for v in shapes.values())
        avail = psutil.virtual_memory().available / max(world, 1)
        if need * 1.3 > avail:
            e2e = {"value": None, "unit": "element-steps/s", "skipped": f"needs {need / 1e9:.1f} GB of pinned host memory "
                   f"per rank, {avail / 1e9:.1f} GB available per rank"}
        else:
            # ONE set of pinned host arrays per rank (the caller's arrays of the drop-in): filled by a download, then
            # uploaded, stepped and downloaded again inside the timed region
            pinned = {k: torch.empty(v, dtype=torch.int64 if k == "element_flag" else torch.float64).pin_memory()
                      for k, v in shapes.items()}
            hv = {k: v.numpy() for k, v in pinned.items()}
            frame = ("disp", "velo", "integ_stress", "integ_strain", "integ_eq_plastic_strain", "integ_triax_stress",
                     "element_flag")
            eng.download(fields=frame, out={k: hv[k] for k in frame})
            ex = eng.download_ex(fields=("disp_pre", "Q", "integ_yield_stress"))
            for k in ("disp_pre", "Q", "integ_yield_stress"):
                hv[k][...] = ex[k]
            del ex
            upl = ("disp", "disp_pre", "velo", "Q", "integ_stress", "integ_strain", "integ_eq_plastic_strain",
                   "integ_yield_stress")
            # the output frame the host consumer (write_vtk, J2:3517-3717) needs: disp, velo, element_flag and the
            # nodal averages, which hk_node_output computes on the device (cal_node_stress_strain, J2:3408-3486)
            raw = world > 1          # partitioned mesh: undivided sums + inc_num, the host adds the neighbours' shares
            nodal = ("node_stress", "node_strain", "node_eq_plastic_strain", "node_triax_stress", "inc_num") + \
                    (() if raw else ("node_mises_stress",))
            out_frame = ("disp", "velo", "element_flag")
            h2d = sum(pinned[k].numel() * 8 for k in upl)
            d2h = sum(pinned[k].numel() * 8 for k in out_frame + nodal)
            eng.node_output(raw=raw, out={k: hv[k] for k in nodal})     # untimed: allocates the engine's work buffers

            def frame_out():
                eng.download(fields=out_frame, out={k: hv[k] for k in out_frame})
                eng.node_output(raw=raw, out={k: hv[k] for k in nodal})
            barrier()
            w0 = time.perf_counter()
            eng.upload_state(disp=hv["disp"], disp_pre=hv["disp_pre"], velo=hv["velo"], Q=hv["Q"],
                             integ_stress=hv["integ_stress"].T, integ_strain=hv["integ_strain"].T,
                             integ_eq_plastic_strain=hv["integ_eq_plastic_strain"],
                             integ_yield_stress=hv["integ_yield_stress"])
            w_up = time.perf_counter() - w0
            run_steps(t_next, args.steps)
            t_next += args.steps
            w_st = time.perf_counter() - w0 - w_up
            frame_out()
            barrier()
            w = time.perf_counter() - w0
            # the drop-in loop itself (J2:487-951 with write_vtk every d_out steps): state stays on the device, the
            # host receives one frame per d_out = `steps` steps
            barrier()
            w1 = time.perf_counter()
            run_steps(t_next, args.steps)
            t_next += args.steps
            frame_out()
            barrier()
            wl = time.perf_counter() - w1
            tw = torch.tensor([w, wl], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tw, op=dist.ReduceOp.MAX)
            w, wl = float(tw[0].item()), float(tw[1].item())
            e2e = {"value": nE * world * args.steps / w, "unit": "element-steps/s",
                   "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
                   "seconds": {"upload": w_up, "steps": w_st, "frame": w - w_up - w_st},
                   "what": f"hk_upload_state (pinned host arrays, all loop state) + {args.steps} steps + one output "
                           f"frame (hk_download of disp, velo, element_flag + hk_node_output: the nodal averages "
                           f"write_vtk needs, computed on the device) per rank through the C ABI; wall clock, max "
                           f"over ranks",
                   "frame_loop": {"value": nE * world * args.steps / wl, "unit": "element-steps/s", "h2d_bytes_per_step": 0,
                                  "d2h_bytes_per_step": d2h / args.steps,
                                  "what": f"the drop-in loop: {args.steps} steps (d_out) + one output frame to host "
                                          f"buffers, state resident on the device; wall clock, max over ranks"}}
            del pinned, hv

    # ---- CPU baseline (oracle port on the host cores, bounded sample) ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or cpu_sample_for(args.workload)
        rate, nEs, dtc = cpu_oracle_rate(sample, 2, args.cpu_steps, threads=cores, contact_myu=args.contact_myu)
        rate1, _, dt1 = cpu_oracle_rate(sample, 1, 2, threads=1, contact_myu=args.contact_myu)      # also OMP_NUM_THREADS=1
        cpu = {"value": rate, "unit": "element-steps/s", "cores": cores, "kind": "port",
               "sample": f"{sample}: {nEs} elements of the same deck recipe, 2 warm-up + {args.cpu_steps} timed "
                         f"steps ({dtc:.1f} s), OpenMP C++ oracle (our restatement of HAKAI_j.jl, not Julia), {cores} threads",
               "single_thread": {"value": rate1, "cores": 1, "sample": f"same sample, 1 warm-up + 2 timed steps ({dt1:.1f} s)"}}

    if rank == 0:
        cfg = {"workload": args.workload, "elements_per_gpu": nE, "nodes_per_gpu": nN, "deck": deck_text(deck, kind, args.contact_myu),
               "regime": regime, "untimed_steps_before_timing": args.warmup + extra,
               "eps_at_start": [s_start["eps_min"], s_start["eps_max"]],
               "live_elements_start": s_start["live_elements"], "live_elements_end": s_end["live_elements"],
               "host_setup_s": round(t_setup, 2), "engine_setup_s": round(t_engine, 2),
               "l2": "state >> L2 (inputs larger than L2), no flush", "parallelism": f"z-slab x{world}",
               "cpu_affinity": numa,
               "halo_bytes_per_step_per_rank": runner.halo.bytes_per_step}
        if per_step_deleted:
            cfg["deleted_per_step"] = per_step_deleted
        line = {
            "metric": "element-steps/sec (hex8 elastoplastic)", "value": value, "unit": "element-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks,
        }
        if contact is not None:
            line["contact"] = contact
        if per_rank is not None:
            line["per_rank_kernel_ms"] = per_rank
        if pcheck is not None:
            line["parity_check"] = pcheck
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
