"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
contract's keys, every workload spec is self-consistent, and rank > 0 of a reference run is silent."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "derivative grid-points/sec" and d["unit"] == "points/s"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["higher_is_better"] is True and d["dtype"] == "f64"
    assert d["config"]["workload"] == "channel_1024x257_b64"
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


@pytest.mark.parametrize("name", ["channel_1024x257_b64", "channel_1024x257_b64_fft", "polar_1024x512_b64", "polar_1024x512_b1",
                                  "polar_1024x512_b64_fft", "fd6_512cube", "fd6_1024cube", "cheb2048_b8192",
                                  "fd6_1024cube_slab", "cheb512cube_pencil"])
def test_workload_specs(name):
    import numpy as np

    import bench

    spec = bench.workload_spec(name)
    assert spec["name"] == name and spec["bound"] in ("tensor", "hbm")
    P = int(np.prod(spec["npts"]))
    assert spec["points_per_step"] % (P * spec["batch"]) == 0          # whole derivatives per point
    for label, kind, axis, order, alg in spec["ops"]:
        assert isinstance(label, str) and alg > 0 and 0 <= axis < len(spec["npts"]) and order >= 1
    with pytest.raises(SystemExit):
        bench.workload_spec("no-such-workload")
