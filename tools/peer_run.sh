#!/bin/bash
# N-GPU box: peer-memory paths vs the NCCL paths of the sharded workloads (+ the multi-GPU peer test).
# usage: tools/peer_run.sh "2 4 8"
set -u
out=gpurun_out/peer_r1
mkdir -p $out
run() { # n workload steps comm
  local n=$1 w=$2 k=$3 c=$4
  if [ "$n" = "1" ]; then
    timeout 300 python bench.py --gpus 1 --steps $k --warmup 5 --workload $w --no-cpu --comm $c > $out/${w}_${c}_n$n.json 2> $out/${w}_${c}_n$n.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n --steps $k --warmup 5 --workload $w --no-cpu --comm $c > $out/${w}_${c}_n$n.json 2> $out/${w}_${c}_n$n.err
  fi
  echo "$w $c n=$n: $(tail -1 $out/${w}_${c}_n$n.json | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["roofline"]["frac"])' 2>&1 | tail -1)"
  tail -3 $out/${w}_${c}_n$n.err | cut -c1-300
}
timeout 600 python -m pytest tests/test_peer.py tests/test_dist.py -q -m gpu 2>&1 | tail -15
for n in ${1:-2}; do
  for c in peer nccl; do
    run $n fd6_1024cube_slab 20 $c
    run $n cheb512cube_pencil 20 $c
  done
done
