"""The committed bench lines (profiles/r02_bench_config*_n1.json, written by `python bench.py` on a B200) carry every key of the
driver contract; `conv_traffic()` reads the fused-operand launch of the committed ncu summary.  Runs anywhere (no GPU)."""
import glob
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_committed_bench_lines_follow_the_contract():
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_config*_n1.json")))
    assert len(files) == 4
    for f in files:
        d = json.load(open(f))
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                  "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "clocks", "gpu_launches"):
            assert k in d, (f, k)
        assert d["n_gpus"] == 1 and d["higher_is_better"] is True and d["data"] == "synthetic" and d["vs_baseline"] is None
        assert "workload" in d["config"] and "model" not in d["config"]
        r = d["roofline"]
        for k in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
            assert k in r, (f, k)
        assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0.3 < r["frac"] < 1.0
        assert r["serialised"]["frac"] <= r["frac"] + 0.02          # the per-launch event intervals include the launch gaps
        c = d["cpu_baseline"]
        for k in ("value", "unit", "cores", "kind", "sample"):
            assert k in c, (f, k)
        assert c["kind"] in ("port", "reference") and c["value"] > 0
        e = d["e2e"]
        assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
        assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
        # a CPU arm four orders of magnitude slower says nothing about the kernels, but the ratio must at least be of that order
        assert d["value"] / c["value"] > 100


def test_conv_traffic_reads_the_fused_launch():
    t = _bench().conv_traffic()
    assert t is not None and "k_conv3x3_v3<3>" in t["kernel"]
    assert 0.8 < t["dram_bytes_per_launch"] / t["algorithmic_bytes_per_launch"] < 1.3
