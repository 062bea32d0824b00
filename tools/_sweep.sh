timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for cfg in "6 6" "7 4" "7 6" "8 4"; do set -- $cfg; echo "m32 blocks $1 listcap $2"; SSQ_SMEM_BLOCKS=$1 SSQ_LIST_CAP=$2 timeout 400 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline; done
