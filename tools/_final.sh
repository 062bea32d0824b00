timeout 600 python bench.py 2>gpurun_out/final_bench.err > gpurun_out/final_bench.json; tail -c 300 gpurun_out/final_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --reads 2000000 --steps 1 --warmup 1 --no-cpu-baseline --streams 1 > gpurun_out/final_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_smem_m|k_smem_p3|k_chain_coop|k_chain$|k_sa$" -c 7 -o gpurun_out/final_full python bench.py --reads 2000000 --steps 1 --warmup 0 --no-cpu-baseline --streams 1 > gpurun_out/final_full.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"k_ext_run" -c 10 -o /tmp/final_ext python bench.py --reads 2000000 --steps 1 --warmup 0 --no-cpu-baseline --streams 1 > gpurun_out/final_ext.log 2>&1
ncu -i /tmp/final_ext.ncu-rep --page raw --csv > gpurun_out/final_ext_raw.csv 2>/dev/null
du -sh gpurun_out; ls -la gpurun_out/final_*
