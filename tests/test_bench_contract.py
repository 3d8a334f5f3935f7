"""bench.py contract checks that need no GPU: the reference arm (the reference's own CPU object code, oracle/_ref) must
print exactly ONE JSON line with the keys the driver reads, alone and under a 2-rank torchrun launch (rank 0 prints,
the other rank exits 0 without work); the product arm must refuse to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REQUIRED = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(cmd, timeout=600):
    env = dict(os.environ, OMP_NUM_THREADS="4")
    return subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)


def check_line(stdout, n_gpus, steps, warmup):
    lines = [l for l in stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["metric"] == "hamming_comparisons_per_s" and d["unit"] == "Gcmp/s"
    assert d["n_gpus"] == n_gpus and d["steps"] == steps and d["warmup"] == warmup and d["value"] > 0
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]
    return d


def test_reference_arm_prints_one_contract_line():
    r = run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    check_line(r.stdout, 1, 2, 1)


def test_reference_arm_under_torchrun_rank0_only():
    r = run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
             "127.0.0.1", "--master-port", "29617", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "2",
             "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    check_line(r.stdout, 2, 2, 1)


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        return  # this check is for the CPU-only container
    r = run([sys.executable, "bench.py", "--steps", "1", "--warmup", "1"], timeout=300)
    assert r.returncode != 0
    assert not [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert "no CPU fallback" in (r.stderr + r.stdout)


import pytest


@pytest.mark.gpu
def test_product_arm_line_has_every_contract_key():
    """The default run (configs[1] on one B200): one JSON line with the base contract's keys plus roofline,
    cpu_baseline, e2e with host<->device bytes, clocks without thermal / hardware slowdown, and a launch count."""
    r = run([sys.executable, "bench.py", "--steps", "20", "--warmup", "3"], timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    need = (REQUIRED - {"impl"}) | {"roofline", "clocks", "gpu_launches"}
    assert need <= set(d), need - set(d)
    assert d["metric"] == "hamming_comparisons_per_s" and d["n_gpus"] == 1 and d["steps"] == 20 and d["warmup"] == 3
    assert d["scaling"] == "weak" and d["gpu_launches"] >= 20
    assert d["engine"].startswith("tensor") and d["dtype"].startswith("s8")  # a 10k x 10k pair goes to the tensor cores
    assert d["value"] > 100 and d["e2e"]["value"] > 50                       # Gcmp/s: far above any CPU path
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 10000 * 64 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(rf)
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-9 and rf["traffic"] > 1_000_000
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and 0 < cb["value"] < d["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert d["config"]["parity_full"] is True  # every query and every column of the timed pair against the oracle
    assert rf["bound"] == "tensor" and rf["unit"] == "TFLOP/s" and rf["popc_roofline_equivalent"]["frac"] > 1.0
    # the engine north_star specifies is measured in the same run, against the integer-pipe roofline
    ie = d["integer_pipe_engine"]
    assert ie["parity_full"] is True and 200 < ie["value"] < d["value"]
    irf = ie["roofline"]
    assert irf["profile"] and irf["profile"]["k1_variant"] == d["config"]["k1_variant"] and irf["mix_frac"] > 0.5
    assert abs(irf["frac"] - irf["achieved"] / irf["peak"]) < 1e-9 and irf["frac"] > 0.6  # north_star's target
    assert d["e2e"]["single_caller"]["value"] > 20
    # the sharded survey (configs[3]) is part of the driver-run line at every N, gather inside its timed region
    c4 = d["secondary"]["c4_survey"]
    assert c4["pairs"] == 9000 and c4["pairs_per_s"] > 2000 and c4["gather"]["records"] == c4["matches"] > 9000 * 500
    assert c4["parity_sample"]["equal"] is True and c4["parity_sample"]["pairs"] >= 8


def test_reference_arm_uses_the_product_arms_configuration():
    """Same workload, same config keys in both arms (the driver compares them): the full 10000 x 10000 pair per worker."""
    r = run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "1"])
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count("c2_config(") == 3  # one definition, one call per arm: the keys cannot drift apart
    assert {"workload", "n1", "n2", "comparisons_per_pair", "cross_check", "pairs_per_s", "l2", "k1_variant",
            "parity_full", "ratio_test_survivors"} == set(d["config"])
    assert d["config"]["n1"] == d["config"]["n2"] == 10000 and d["config"]["comparisons_per_pair"] == 10 ** 8
    assert "full 10000x10000" in d["cpu_baseline"]["sample"] and d["config"]["ratio_test_survivors"] == 5000
