"""Grid construction: pytristan_b200 (GPU expansion + readback, and device-resident) vs the NumPy
meshgrid+flatten formulation the reference uses (reference grid.py:83-88, restated here because the
reference is not on the GPU box).  Prints one JSON line per shape."""
import json, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import pytristan_b200 as pt

def reference_style(arrs):
    mg = np.meshgrid(*arrs, indexing="ij")
    return np.array([m.flatten(order="F") for m in mg])

def best(fn, reps):
    out = 1e30
    for _ in range(reps):
        t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); out = min(out, time.perf_counter() - t0); del r
    return out

for shape in [(1024, 257), (1024, 1024), (256, 256, 256), (512, 512, 512)]:
    arrs = [np.linspace(0.0, 1.0, n) for n in shape]
    pts = int(np.prod(shape))
    t_ref = best(lambda: reference_style(arrs), 2 if pts > 1e8 else 3)
    t_host = best(lambda: pt.Grid.from_arrs(*arrs), 3)
    t_dev = best(lambda: pt.DeviceGrid.from_arrs(*arrs), 3)
    print(json.dumps({"shape": shape, "numpy_meshgrid_s": round(t_ref, 5), "Grid_s": round(t_host, 5), "DeviceGrid_s": round(t_dev, 6),
                      "speedup_Grid": round(t_ref / t_host, 1), "speedup_DeviceGrid": round(t_ref / t_dev, 1),
                      "Mpts_s_numpy": round(pts / t_ref * 1e-6, 1), "Mpts_s_Grid": round(pts / t_host * 1e-6, 1), "Mpts_s_DeviceGrid": round(pts / t_dev * 1e-6, 1)}))
t = best(lambda: pt.DeviceGrid.from_arrs(*[np.linspace(0.0, 1.0, 1024)] * 3), 2)
print(json.dumps({"shape": (1024, 1024, 1024), "DeviceGrid_s": round(t, 5), "Mpts_s_DeviceGrid": round(1024 ** 3 / t * 1e-6, 1), "GB_written": 25.8,
                  "note": "reference: infeasible on the host (72 GiB peak)"}))
