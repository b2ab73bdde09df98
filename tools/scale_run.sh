#!/bin/bash
# 8-GPU box: weak scaling of the default workload and strong scaling of the sharded workloads.
set -u
out=gpurun_out/scale_r1
mkdir -p $out
run() { # n workload steps
  local n=$1 w=$2 k=$3
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps $k --warmup 5 --workload $w --no-cpu > $out/${w}_n$n.json 2> $out/${w}_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps $k --warmup 5 --workload $w --no-cpu > $out/${w}_n$n.json 2> $out/${w}_n$n.err
  fi
  tail -1 $out/${w}_n$n.json | cut -c1-160
}
for n in 1 2 4 8; do run $n channel_1024x257_b64 50; done
for n in 1 2 4 8; do run $n fd6_1024cube_slab 10; done
for n in 1 2 4 8; do run $n cheb512cube_pencil 10; done
timeout 600 python -m pytest tests/test_dist.py -q -m gpu 2>&1 | tail -2
