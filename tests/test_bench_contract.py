"""The bench.py JSON contract (CPU): the committed lines of the last GPU runs carry every key the driver reads, the numbers
are consistent with each other, and the reference arm -- which needs no GPU -- prints a valid line here at a small size."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e"]


def _lines():
    out = []
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_h_bench_*gpu.json"))):
        with open(path) as fh:
            out.append((os.path.basename(path), json.loads(fh.read())))
    return out


@pytest.mark.parametrize("name,d", _lines())
def test_committed_bench_lines_keep_the_contract(name, d):
    for k in BASE_KEYS + ["roofline", "gpu_launches", "prove_ms"]:
        assert k in d, (name, k)
    assert d["metric"] == "lde_melem_per_s" and d["unit"] == "Melem/s" and d["higher_is_better"] is True
    assert d["dtype"] == "u64" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    assert d["steps"] >= 1 and d["warmup"] >= 3
    # value = L * C / time of ONE 2^22 x 32 problem (strong scaling: the same numerator at every N)
    L, C = (1 << 22) * 4, 32
    assert d["value"] == pytest.approx(L * C / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    r = d["roofline"]
    for k in ["bound", "achieved", "peak", "unit", "frac", "traffic"]:
        assert k in r, (name, k)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert r["achieved"] == pytest.approx(r["algorithmic_bytes_per_launch"] / (d["ms_per_step"] * 1e-3) / 1e9, rel=1e-6)
    assert r["kernel_sum_ms_per_step"] <= d["ms_per_step"] * 1.001  # the kernels' own time fits inside the step
    e = d["e2e"]
    for k in ["value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"]:
        assert k in e, (name, k)
    assert e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] < d["value"]
    assert d["gpu_launches"] > 0
    assert d["prove"]["matches_oracle_digest"] is True
    if d["n_gpus"] == 1:
        c = d["cpu_baseline"]
        for k in ["value", "unit", "cores", "kind", "sample"]:
            assert k in c, (name, k)
        assert c["kind"] == "port" and c["cores"] == 1
        assert r["traffic"] and r["traffic"] > r["algorithmic_bytes_per_launch"]  # measured (ncu), two-pass transform
        assert r["alu"]["frac"] and 0.5 < r["alu"]["frac"] < 1.5
        assert d["clocks"]["reasons"] == []
    else:
        assert d["scaling"] == "strong" and "lde_weak" in d
        five = d["baseline_configs"]["5a: 2^24 x 64, blowup 4, binary trees"]
        assert five["n_gpus"] == d["n_gpus"] and five["proof_bytes"] == 3221534464


def test_config_5a_is_one_proof_at_every_gpu_count():
    digests = set()
    for name, d in _lines():
        five = d.get("baseline_configs", {}).get("5a: 2^24 x 64, blowup 4, binary trees")
        if five and five.get("proof_sha256"):
            digests.add(five["proof_sha256"])
    assert len(digests) == 1, digests


def test_reference_arm_prints_a_valid_line_without_a_gpu():
    """bench.py --impl reference times the C restatement on the host: it runs here (small shape)."""
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--log-rows", "10", "--cols", "8"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in BASE_KEYS + ["impl", "cpu_baseline"]:
        assert k in d, k
    assert d["impl"] == "reference" and d["metric"] == "lde_melem_per_s" and d["unit"] == "Melem/s"
    assert d["e2e"] == {"value": d["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["prove"]["proof_bytes"] > 0 and d["prove_ms"] > 0
    # both arms describe the workload with the same `config` object (the driver compares them); what differs per arm is in `arm`
    assert set(d["config"]) == {"workload", "l2"} and "sample" in d["arm"]
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert src.count('"config": bench_config(args)') == 2
    # ranks other than 0 of a torchrun launch exit without work or output
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                         timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
