"""The bench line the driver parses: keys and types of the committed single-GPU run (profiles/r2_bench_1M.json, produced by
`python bench.py` on a B200), of the reference arm and of the 2-GPU runs.  CPU-only; guards the JSON contract, not the numbers."""
import json
import os

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load(name):
    with open(os.path.join(REPO, "profiles", name)) as f:
        return json.loads(f.read())


def test_b200_arm_line_has_the_contract_keys():
    d = _load("r2_bench_1M.json")
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
    # BASELINE's metric names "reconstruction PSNR vs reference": the line carries the comparison with the reference backend
    p = d["parity"]
    assert p["checker"].startswith("reference") and p["sample_leaves"] >= 1024 and p["index_total"] == 64 * p["sample_leaves"]
    assert p["index_mismatches"] <= 8 and p["max_margin_at_mismatch"] <= 1e-4 and abs(p["dpsnr_db"]) <= 0.1
    assert p["psnr_db_vs_reference_recon"] >= 55.0
    # the route the reference's callers use: pageable memory through the C++ virtual interface
    g = d["e2e_pageable"]
    assert g["value"] >= 0.8 * e["value"] and g["indices_equal_pinned_path"] is True
    assert r["traffic_source"] and (r["traffic"] is None or r["traffic_source"].startswith("profiles/"))


def test_multi_gpu_lines_verify_the_gathered_grid():
    for name, scaling, total in (("r2_bench_2gpu_weak.json", "weak", 2_000_000), ("r2_bench_2gpu_strong10M.json", "strong", 10_000_000),
                                 ("r2_bench_2gpu_weak_nccl.json", "weak", 2_000_000)):
        d = _load(name)
        assert d["n_gpus"] == 2 and d["scaling"] == scaling and d["config"]["leaves_total"] == total
        assert d["gather_verified"] is True
        assert abs(d["ms_per_step"] * d["value"] / 1e3 - total) < 2.0


def test_reference_arm_line():
    d = _load("r2_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["metric"] == "leaves_per_sec_encode_decode" and d["unit"] == "leaves/s"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    b = _load("r2_bench_1M.json")
    for key in ("workload", "leaves_per_gpu", "leaves_total", "weights"):      # same config as the B200 arm
        assert d["config"][key] == b["config"][key]
    assert d["metric"] == b["metric"] and d["higher_is_better"] == b["higher_is_better"] and d["scaling"] == b["scaling"]


def test_ncu_traffic_is_tied_to_the_kernel_source():
    # roofline.traffic comes from a committed ncu capture and must not outlive the kernel it was taken from
    import bench
    for k, e in bench.NCU_DRAM.items():
        files = e["file"] if isinstance(e["file"], list) else [e["file"]]      # a path made of two kernels lists both sources
        blobs = e["blob"] if isinstance(e["blob"], list) else [e["blob"]]
        assert all(os.path.exists(os.path.join(REPO, src)) for src in e["source"].split(" + "))
        assert len(files) == len(blobs) and all(len(b) == 40 for b in blobs)
        traffic, why = bench.ncu_traffic(k, 1000)
        if all(bench.git_blob_hash(f) == b for f, b in zip(files, blobs)):
            assert traffic == e["bytes_per_leaf"] * 1000 and why == e["source"]
        else:
            assert traffic is None and why.startswith("stale")
    saved = dict(bench.NCU_DRAM["encode"])
    try:
        bench.NCU_DRAM["encode"]["blob"] = "0" * 40
        assert bench.ncu_traffic("encode", 1000)[0] is None
    finally:
        bench.NCU_DRAM["encode"].update(saved)
