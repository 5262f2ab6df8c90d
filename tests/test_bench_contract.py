"""The reference arm of bench.py (`--impl reference`: the reference's own shaders compiled for the host, oracle/_ref -- or
the CPU oracle port where that library is absent -- timed on the host cores) runs without a GPU and prints ONE JSON line
with the keys the driver's contract names."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, ROOT)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3",
                          "--workload", "720p"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "impl", "cpu_baseline", "e2e"):
        assert key in j, key
    assert j["impl"] == "reference" and j["metric"] == "asvgf_denoise_throughput" and j["unit"] == "Mpixel/s"
    assert j["value"] > 0 and j["higher_is_better"] is True and j["vs_baseline"] is None
    from oracle import ref
    assert j["cpu_baseline"]["kind"] == ("reference" if ref.available() else "port")
    assert j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["value"] == j["value"]
    if ref.available():
        assert j["cpu_port"]["kind"] == "port" and j["cpu_port"]["value"] > j["value"] * 0.5   # the hand-written port, for context
    for key in ("width", "height", "atrous_iterations", "streams_per_gpu", "parallelism", "l2"):
        assert key in j["config"], key                        # the same config keys as the repo's own arm
    assert j["e2e"] == {"value": j["value"], "unit": j["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"]


def test_workloads_name_the_baseline_configs():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.WORKLOADS["4k"][:3] == (3840, 2160, 4) and bench.WORKLOADS["4k-n5"][2] == 5
    assert bench.WORKLOADS["8k-bands"][:2] == (7680, 4320) and bench.WORKLOADS["1080p"][:2] == (1920, 1080)
