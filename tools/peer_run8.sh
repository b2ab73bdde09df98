#!/bin/bash
# 8-GPU box: peer-memory paths of the sharded workloads at 4 and 8 GPUs (+ the multi-GPU peer test at 4 ranks).
set -u
out=gpurun_out/peer_r1
mkdir -p $out
run() { # n workload steps comm
  local n=$1 w=$2 k=$3 c=$4
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps $k --warmup 5 --workload $w --no-cpu --comm $c > $out/${w}_${c}_n$n.json 2> $out/${w}_${c}_n$n.err
  echo "$w $c n=$n: $(tail -1 $out/${w}_${c}_n$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"])' 2>&1 | tail -1)"
}
timeout 300 python -m pytest tests/test_peer.py -q -m gpu 2>&1 | tail -5
for n in 4 8; do
  run $n fd6_1024cube_slab 20 peer
  run $n cheb512cube_pencil 20 peer
done
run 8 fd6_1024cube_slab 20 nccl
run 8 cheb512cube_pencil 20 nccl
