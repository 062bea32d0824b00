SSQ_SMEM_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_cli.py tests/test_gpu_index.py -x -q -m gpu 2>&1 | tail -3
SSQ_SMEM_VARIANT=3 timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_split.json; tail -c 200 gpurun_out/bench_split.json
