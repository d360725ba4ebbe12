"""bench.py's contract with the driver, checked on the CPU: the reference arm (`--impl reference`, the oracle port on the
host cores) prints ONE JSON line with the agreed keys, and its `config` is the object the GPU arm prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
        "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                          "--warmup", "0"], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.strip()]
    assert len(lines) == 1, lines                      # exactly one line on stdout
    d = json.loads(lines[0])
    assert KEYS <= set(d), KEYS - set(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["unit"] == "img/s" and d["value"] > 0 and d["dtype"] == "f32"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    sys.path.insert(0, ROOT)
    import bench
    wl = bench.get_workload("2")
    assert d["config"] == wl.config(1) and d["metric"] == wl.metric      # the same object in both arms
    assert "workload" in d["config"] and "model" not in d["config"]


def test_every_configuration_has_a_workload():
    sys.path.insert(0, ROOT)
    import bench
    for cid in ("2", "3a", "3b", "4", "5", "n3"):
        wl = bench.get_workload(cid)
        c = wl.config(2)
        assert c["config_id"] == cid and c["global_batch"] == 2 * c["batch_per_gpu"] and wl.metric and wl.bytes_per_px > 0
