#!/bin/bash
# 1-GPU box: GPU test suite, bench lines of every single-GPU workload, ncu launch list of the default bench.
set -u
tag=${1:-r1e}
out=gpurun_out
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $out/pytest_gpu_$tag.log; cat $out/pytest_gpu_$tag.log
timeout 300 python bench.py --impl reference > $out/bench_${tag}_reference.json 2> $out/bench_${tag}_reference.err; cut -c1-200 $out/bench_${tag}_reference.json
timeout 400 python bench.py > $out/bench_$tag.json 2> $out/bench_$tag.err; cut -c1-300 $out/bench_$tag.json
for w in channel_1024x257_b64_fft polar_1024x512_b64 polar_1024x512_b64_fft polar_1024x512_b1 fd6_1024cube cheb2048_b8192; do
  timeout 400 python bench.py --workload $w --steps 30 --warmup 5 > $out/bench_${tag}_$w.json 2> $out/bench_${tag}_$w.err
  python - <<PY
import json
try:
    d = json.loads(open("$out/bench_${tag}_$w.json").read().strip().splitlines()[-1])
    print("$w", round(d["ms_per_step"], 4), "ms", "%.4g pts/s" % d["value"], "frac", round(d["roofline"]["frac"], 3), "e2e", "%.4g" % d["e2e"]["value"] if d.get("e2e") else None, "cpu", d["cpu_baseline"]["value"] if d.get("cpu_baseline") else None)
except Exception as e:
    print("$w FAILED", e); print(open("$out/bench_${tag}_$w.err").read()[-1500:])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu > $out/ncu_launch_$tag.log 2>&1; tail -2 $out/ncu_launch_$tag.log | cut -c1-200
