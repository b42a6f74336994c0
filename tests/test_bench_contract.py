"""bench.py's reference arm runs on the host only: check the JSON contract of the line the driver parses (keys, units, the
e2e / cpu_baseline objects of the CPU arm) without a GPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "3"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert base["metric"].startswith(d["metric"].split(" at ")[0])           # BASELINE.json's metric, 640x480 configuration
    assert d["impl"] == "reference"
    assert d["metric"] == "frames/sec (track+depth-update) at 640x480" and d["unit"] == "frames/s"
    for k in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["value"] > 0 and abs(d["value"] - 1e3 / d["ms_per_step"]) < 1e-6 * d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 5 and cb["value"] == d["value"] and "sample" in cb
    e = d["e2e"]
    assert e["value"] == d["value"] and e["unit"] == d["unit"] and e["h2d_bytes_per_step"] == 0 and e["d2h_bytes_per_step"] == 0


def test_bench_refuses_to_run_the_gpu_arm_without_a_gpu():
    """no CPU fallback: without a CUDA device the product arm fails loudly instead of printing a number"""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert not any(ln.strip().startswith("{") for ln in r.stdout.splitlines())


def test_product_arm_line_contract_with_mocked_device_results(monkeypatch, capsys):
    """the JSON line of the product arm (everything after the device timing): all contract keys, roofline, cpu_baseline,
    e2e, clocks, gpu_launches -- exercised on the CPU by replacing only the device loop with canned timings"""
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench

    def fake_gpu_run(args, rank, world, local_rank):
        seq, frames = bench.render_frames(args.width, args.height, 1234, args.warmup + args.steps + 1)
        leg = dict(total_ms=2.0, launches=53, clocks={"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": [], "samples": 5}, wall=0.1,
                   kms=1.1, klaunch=10, kbytes=1.1e8, poses=np.zeros((args.steps, 7)), p50=0.2, p95=0.22)
        return seq, frames, {"resident": leg, "e2e": dict(leg, total_ms=2.2)}
    monkeypatch.setattr(bench, "gpu_run", fake_gpu_run)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "10", "--warmup", "3"])
    bench.main()
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 10 and d["gpu_launches"] == 53 and d["vs_baseline"] is None
    assert abs(d["value"] - 10 / 2.0e-3) < 1e-6 and abs(d["e2e"]["value"] - 10 / 2.2e-3) < 1e-6
    assert d["e2e"]["h2d_bytes_per_step"] == 640 * 480 and d["e2e"]["d2h_bytes_per_step"] > 0
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == 5 and cb["value"] > 0 and cb["single_core"]["value"] > 0
    assert 0.30 <= d["config"]["semi_dense_fraction"] <= 0.50 and "workload" in d["config"] and "l2" in d["config"]
