#!/bin/bash
# First GPU minutes of the next round: verify and measure the seeding paths that were only validated on the CPU.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_runs.sh'
SSQ_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_experimental.py -x -q -m gpu 2>&1 | tail -4
for cfg in "2 0" "3 0" "4 0" "4 8" "4 10" "4 11" "2 10"; do
  set -- $cfg
  echo "variant $1 kmer $2"
  if [ "$2" = "0" ]; then SSQ_SMEM_VARIANT=$1 timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline
  else SSQ_SMEM_VARIANT=$1 SSQ_KMER_K=$2 timeout 300 python tools/bench_brief.py --steps 2 --warmup 3 --no-cpu-baseline; fi
done
