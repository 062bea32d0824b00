timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
timeout 400 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline
echo "no p3 kernel"; SSQ_NO_P3_KERNEL=1 timeout 400 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline
echo "streams 5"; timeout 400 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline --streams 5
