"""bench.py's output contract: the reference arm runs here (CPU), the GPU arm is checked on its committed line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"}


def _reference_arm(*extra):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--rays", "4096",
                          "--steps", "1", "--warmup", "1", *extra], capture_output=True, text=True, timeout=600,
                         cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d) and d["impl"] == "reference"
    assert d["metric"] == "traced rays/s" and d["unit"] == "rays/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["gpu_launches"] == 0 and d["dtype"] == "f64"
    assert d["config"]["workload"] == "config4"
    assert (d["steps"], d["warmup"]) == (1, 1)  # what actually ran
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"] and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    return d


def test_reference_arm_port_only():
    """Without the staged NumPy reference the arm times the C restatement and says so."""
    d = _reference_arm("--no-numpy")
    assert d["cpu_baseline"]["kind"] == "port"


def test_reference_arm_times_the_numpy_reference_when_staged():
    from oracle import ref_shim

    if not ref_shim.available():
        import pytest

        pytest.skip("PyRayT reference not present")
    d = _reference_arm()
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and "UNMODIFIED NumPy reference" in cb["sample"]
    assert cb["port"]["kind"] == "port" and cb["port"]["value"] > cb["value"]  # the port is the harder baseline


def test_other_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--rays", "1024"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_committed_gpu_line_has_the_contract_keys():
    """profiles/bench_r2_final.json is the line `python bench.py` printed on a B200 (round-2 final build)."""
    d = json.loads(open(os.path.join(ROOT, "profiles", "bench_r2_final.json")).read().strip().splitlines()[-1])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["scaling"] == "weak" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["config"]["workload"] == "config4" and d["config"]["rays_per_gpu"] == 1 << 24
    roof = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(roof) and roof["bound"] == "hbm"
    assert abs(roof["frac"] - roof["achieved"] / roof["peak"]) < 1e-9 and roof["traffic"] > 0
    assert d["gpu_launches"] > 0 and d["warmup"] >= 3
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 13 * 8 * (1 << 24) and e2e["d2h_bytes_per_step"] > 0
    assert e2e["value"] < d["value"]  # host buffers in and out are slower than device-resident steps
    assert e2e["host_peak"]["d2h_gbs_aggregate"] > 0 and 0 < e2e["frac_of_host_peak"] <= 1.1
    # the metric's own baseline: the unmodified NumPy reference timed on the same box, the C port beside it
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] > 0
    assert cb["port"]["kind"] == "port" and cb["port"]["value"] > cb["value"]
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the counted exclusions and the optional fast mode
    nd = d["near_degenerate"]
    assert nd["rays"] > 0 and nd["grazing_rays"] >= 0 and nd["seam_rays"] >= 0
    assert d["argsort_mismatch"]["rows_differing"] >= 0
    f32 = d["fp32_mode"]
    assert f32["dtype"] == "f32" and f32["value"] > d["value"]
    ag = f32["agreement_with_fp64"]
    assert ag["id_columns_equal_on_compared_rows"] and ag["max_error_on_agreeing_rays"] <= 1e-5
    assert ag["rays_with_different_ids"] + ag["rays_beyond_tolerance"] <= 1e-4 * ag["rays"]
