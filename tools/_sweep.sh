timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_index.py -x -q -m gpu 2>&1 | tail -2
for d in 8 4 16; do echo "sa dense $d"; SSQ_SA_DENSE=$d timeout 400 python tools/bench_brief.py --steps 3 --warmup 3 --no-cpu-baseline; done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_smem_m" -c 1 -o gpurun_out/smem_r3 python bench.py --reads 2000000 --steps 1 --warmup 0 --no-cpu-baseline --streams 1 > gpurun_out/smem_r3.log 2>&1; ls -la gpurun_out/smem_r3.ncu-rep
