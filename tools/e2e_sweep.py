import sys, json
import numpy as np, torch
sys.path.insert(0, ".")
import pytristan_b200 as pt
import bench
spec = bench.workload_spec("channel_1024x257_b64")
ops = bench.build_ops(spec, bench.axis_nodes(spec))
P, B = 1024 * 257, 64
h_in = torch.randn(B, P, dtype=torch.float64).pin_memory()
h_out = [torch.empty(B, P, dtype=torch.float64).pin_memory() for _ in ops]
for cf in (None, [1, 1, 2, 4, 8, 16, 16, 16], [1, 1, 2, 4, 8, 8, 8, 8, 8, 8, 8], [2, 2, 4, 8, 8, 8, 8, 8, 8, 8], [4] * 16, 32, 16, 8, 4, 2, 1):
    pt.apply_host(ops, h_in, h_out, chunk_fields=cf); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(5):
        pt.apply_host(ops, h_in, h_out, chunk_fields=cf)
    t1.record(); torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / 5
    print(json.dumps({"chunk_fields": str(cf), "ms": round(ms, 3), "gpts_s": round(spec["points_per_step"] / ms * 1e-6, 2)}))
