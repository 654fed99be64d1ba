"""The bench line the driver parses: keys and types of the committed single-GPU run (profiles/r1e_bench_1M.json, produced by
`python bench.py` on a B200) and of the reference arm.  CPU-only; guards the JSON contract, not the numbers."""
import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(REPO, "profiles", name)) as f:
        return json.loads(f.read())


def test_b200_arm_line_has_the_contract_keys():
    d = _load("r1e_bench_1M.json")
    assert d["metric"] == "leaves_per_sec_encode_decode" and d["unit"] == "leaves/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] >= 1 and d["warmup"] >= 3 and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["vs_baseline"] is None                      # BASELINE.md publishes no number for this metric
    assert isinstance(d["value"], float) and d["value"] > 0 and abs(d["ms_per_step"] * d["value"] / 1e3 - d["config"]["leaves_per_gpu"]) < 1.0
    assert "workload" in d["config"] and "model" not in d["config"]
    r = d["roofline"]
    assert r["bound"] in ("hbm", "tensor") and r["unit"] in ("GB/s", "TFLOP/s")
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["traffic"] is None or r["traffic"] >= r["algorithmic_bytes"] * 0.9
    c = d["cpu_baseline"]
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1 and c["value"] > 0 and c["unit"] == "leaves/s" and c["sample"]
    e = d["e2e"]
    assert e["unit"] == "leaves/s" and e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] != d["value"]                      # measured separately, host buffers inside the timed region
    assert d["gpu_launches"] > 0
    k = d["clocks"]
    assert k["sm_mhz"] > 0 and k["sm_max_mhz"] >= k["sm_mhz"] and isinstance(k["reasons"], list)
    assert not set(k["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_reference_arm_line():
    d = _load("r1e_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "leaves_per_sec_encode_decode" and d["unit"] == "leaves/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_ncu_traffic_is_tied_to_the_kernel_source():
    # roofline.traffic comes from a committed ncu capture and must not outlive the kernel it was taken from
    import bench
    for k, e in bench.NCU_DRAM.items():
        assert os.path.exists(os.path.join(REPO, e["source"])) and len(e["blob"]) == 40
        traffic, why = bench.ncu_traffic(k, 1000)
        if bench.git_blob_hash(e["file"]) == e["blob"]:
            assert traffic == e["bytes_per_leaf"] * 1000 and why == e["source"]
        else:
            assert traffic is None and why.startswith("stale")
    saved = dict(bench.NCU_DRAM["encode"])
    try:
        bench.NCU_DRAM["encode"]["blob"] = "0" * 40
        assert bench.ncu_traffic("encode", 1000)[0] is None
    finally:
        bench.NCU_DRAM["encode"].update(saved)
