"""bench.py's reference arm (the CPU oracle port timed on the host cores) prints the contract's JSON line;
runs without a GPU.  The GPU arm's line is produced on the B200 box (profiles/*.json)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, APB_REF_BUDGET_S="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "lm_iters_per_sec" and line["dtype"] == "f64"
    assert line["value"] > 0 and line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]


def test_workload_builders_lower_on_cpu():
    """Every bench workload's scale model builds and lowers (host logic only)."""
    sys.path.insert(0, ROOT)
    import bench
    import astrophot_b200 as ap
    from astrophot_b200.lowering import lower, tile_scene

    ap.AP_config.ap_device = "cpu"
    for wl, n_src in (("c3t", 15 + 78 + 1), ("c5t", 10 + 1)):
        model = bench.build_workload(ap, wl, 1, None)
        scene, _ = lower(model)
        assert len(scene.sources) == n_src
        tiled = tile_scene(scene, 2, 2)
        assert len(tiled.images) == 4 and tiled.owners is not None and len(tiled.owners) == n_src
        assert sum(im.H * im.W for im in tiled.images) == scene.images[0].H * scene.images[0].W
    model = bench.build_workload(ap, "c2b", 1, None)
    scene, _ = lower(model)
    assert [im.aux for im in scene.images] == [False, True] and scene.psfs[0].source == 0 and scene.n_par == 9
