#!/bin/bash
# compute-sanitizer passes over the small-shape GPU tests.
# usage (on the GPU box): bash tools/sanitize.sh <tag>   -> gpurun_out/san_<tag>/*.txt
tag=${1:-r02}
out=gpurun_out/san_$tag
mkdir -p $out
CS=/usr/local/cuda/bin/compute-sanitizer
T1="tests/test_dense_eo.py tests/test_gpu_random_shapes.py tests/test_fourier_fft.py"
run() {   # tool, budget seconds, name, pytest args...
  tool=$1; budget=$2; name=$3; shift 3
  timeout $budget $CS --tool $tool --print-limit 30 --error-exitcode 7 --log-file $out/${tool}_${name}.txt \
      python -m pytest "$@" -x -q -m gpu -p no:cacheprovider > $out/${tool}_${name}_pytest.log 2>&1
  echo "$tool $name exit=$?" | tee -a $out/summary.txt
  tail -n 3 $out/${tool}_${name}_pytest.log | tee -a $out/summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out/${tool}_${name}.txt | tail -n 2 | tee -a $out/summary.txt
}
run memcheck ${MEMCHECK_BUDGET:-420} tests $T1
run synccheck ${SYNC_BUDGET:-300} tests $T1
run racecheck ${RACE_BUDGET:-420} tests tests/test_dense_eo.py tests/test_gpu_random_shapes.py tests/test_fourier_fft.py \
    -k "${RACE_K:-not 512 and not 4096 and not 2048 and not 1024}"
# the two-process peer worker (IPC-mapped memory, device flag barrier) under memcheck
timeout ${PEER_BUDGET:-240} $CS --tool memcheck --target-processes all --print-limit 30 --log-file $out/memcheck_peer_%p.txt \
    python -m pytest tests/test_peer.py -x -q -m gpu -k two_processes -p no:cacheprovider > $out/memcheck_peer_pytest.log 2>&1
echo "memcheck peer exit=$?" | tee -a $out/summary.txt
tail -n 3 $out/memcheck_peer_pytest.log | tee -a $out/summary.txt
grep -h "ERROR SUMMARY" $out/memcheck_peer_*.txt | tee -a $out/summary.txt
