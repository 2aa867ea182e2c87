"""bench.py pieces that run without a GPU: the reference arm's JSON line (the driver launches it as
`bench.py --impl reference --gpus N --steps K --warmup W`), the workload factory and the rank-to-NUMA binding."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(*extra, env=None):
    e = dict(os.environ)
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                          "--cpu-sample", "N6,6,8", *extra], capture_output=True, text=True, timeout=600, env=e, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    return out.stdout.strip().splitlines()


def test_reference_arm_prints_one_contract_line():
    lines = _run()
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"].startswith("element-steps/sec") and j["unit"] == "element-steps/s"
    assert j["higher_is_better"] is True and j["steps"] == 2 and j["value"] > 0
    assert j["warmup"] == 3                       # the timing rules ask for at least 3 warm-up steps: a smaller W is raised
    assert j["config"]["workload"] == "W16" and j["config"]["elements_in_sample"] == 6 * 6 * 8
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and "N6,6,8" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "element-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_runs_on_rank_zero_only():
    assert _run("--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"}) == []
    j = json.loads(_run("--gpus", "2", env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0"})[0])
    assert j["n_gpus"] == 2


@pytest.mark.parametrize("workload,kind,n_el", [("N4,4,4", "stretch", 64), ("N4,4,4D", "ductile", 64), ("I0", "impact", 124913)])
def test_workload_factory(workload, kind, n_el):
    import bench
    deck, k = bench.make_deck(workload)
    assert k == kind
    if kind == "impact":
        assert deck.plate[0] * deck.plate[1] * deck.plate[2] + deck.proj[0] ** 3 == n_el
        assert "frictionless" in bench.deck_text(deck, k) and "mu = 0.25" in bench.deck_text(deck, k, 0.25)
    else:
        assert deck.nx * deck.ny * deck.nz == n_el and deck.jitter_by_layer
    assert bench.cpu_sample_for("I8") == "I0" and bench.cpu_sample_for("F16D") == "S1D" and bench.cpu_sample_for("W16") == "S1"
    with pytest.raises(SystemExit):
        bench.make_deck("nonsense")


def test_numa_binding_is_never_fatal():
    import bench
    before = os.sched_getaffinity(0)
    msg = bench.bind_to_gpu_numa_node(0)          # no NVML in the build container: reports why and changes nothing
    assert isinstance(msg, str) and msg
    if "unchanged" in msg:
        assert os.sched_getaffinity(0) == before
    os.sched_setaffinity(0, before)
