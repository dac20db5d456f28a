"""CPU: the pure parts of bench.py — algorithmic bytes per kernel (DESIGN.md §4 / SURVEY §8d) and the roofline /
stages objects of the JSON line — fed with the per-kernel timings of a committed B200 run, so that a typo in the
report code cannot surface for the first time on the GPU box."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _committed_line():
    return json.loads(open(os.path.join(ROOT, "profiles", "r01v_bench_n1.json")).read().strip().splitlines()[-1])


def test_algorithmic_bytes_follow_the_survey_figures():
    import bench
    P, R, HW = 1_000_000, 8_954_935, 640_000
    assert bench.algorithmic_bytes("preprocess", P, R, R, R, HW, P) == P * (236 + 83)
    assert bench.algorithmic_bytes("geom_backward", P, R, R, R, HW, P) == P * (303 + 256)
    assert bench.algorithmic_bytes("render_forward", P, R, 100, 50, HW, P) == 100 * 44 + HW * 24
    assert bench.algorithmic_bytes("render_backward", P, R, 100, 50, HW, P) == 50 * 112 + HW * 20
    assert bench.algorithmic_bytes("tile_sort.scatter", P, R, R, R, HW, P) == R * 16
    assert bench.algorithmic_bytes("depth_sort.hist", P, R, R, R, HW, P) == P * 4
    assert bench.algorithmic_bytes("no_such_kernel", P, R, R, R, HW, P) == 0


def test_roofline_report_reproduces_the_committed_line():
    import bench
    line = _committed_line()
    per_step = line["stages"]["ms_per_step_by_kernel"]
    by_k = line["stages"]["roofline_by_kernel"]
    nprof = 5
    acc = {}
    for k, ms in per_step.items():       # rebuild per-launch samples: ms per step / launches per step
        launches = max(1, round(ms / by_k[k]["ms_per_launch"])) if k in by_k else 1
        acc[k] = [ms / launches] * (launches * nprof)
    stats = {"R": line["stages"]["num_rendered"], "visible": line["stages"]["visible"],
             "mean_list": line["stages"]["mean_tile_list"], "max_list": line["stages"]["max_tile_list"],
             "R_fwd": 0, "R_bwd": 0}
    # swept-entry counts are not in the line: recover them from the committed algorithmic bytes of the render kernels
    dom_bytes = line["roofline"]["algorithmic_bytes_per_launch"]
    stats["R_bwd"] = (dom_bytes - 640_000 * 20) // 112
    roofline, stages = bench.roofline_report(acc, nprof, stats, 1_000_000, 640_000, line["ms_per_step"])
    assert roofline["kernel"] == line["roofline"]["kernel"] == "render_backward"
    assert roofline["bound"] == "hbm" and roofline["unit"] == "GB/s"
    assert roofline["algorithmic_bytes_per_launch"] == pytest.approx(dom_bytes, abs=112)
    assert roofline["achieved"] == pytest.approx(line["roofline"]["achieved"], rel=2e-3)
    assert roofline["frac"] == pytest.approx(roofline["achieved"] / roofline["peak"])
    assert roofline["traffic"] is not None and roofline["ncu"]["issue_active_pct"] > 0
    assert stages["roofline_by_kernel"]["geom_backward"]["achieved_GBps"] == pytest.approx(
        by_k["geom_backward"]["achieved_GBps"], rel=2e-3)
    assert stages["roofline_by_kernel"]["geom_backward"]["ncu_dram_pct_of_peak"] > 40
    assert 0.3 < stages["whole_path_frac_of_hbm_roofline"] < 0.7 or roofline["peak_source"].startswith("fallback")
    json.dumps({"roofline": roofline, "stages": stages})          # serialisable


def test_bench_line_contract_keys_of_the_committed_runs():
    """Every committed bench line carries the keys the driver's contract names."""
    need = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"}
    for f in ("r01v_bench_n1.json", "r01w_bench_n2.json", "r01w_bench_n8.json"):
        d = json.loads(open(os.path.join(ROOT, "profiles", f)).read().strip().splitlines()[-1])
        assert need <= set(d), (f, need - set(d))
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["config"]["workload"] == "lego_1m" and d["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
