set -x
K='regex:fill_rows|gather_rows|gene_gemm|heads_fwd|mid_|rmsprop|zinb_loss|expand_counts|flash_zinb'
timeout 300 python bench.py > gpurun_out/bench_c2_final.log 2>&1; tail -1 gpurun_out/bench_c2_final.log | cut -c1-200
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_final.log 2>&1; tail -1 gpurun_out/bench_c3_final.log | cut -c1-200
timeout 200 ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c2_final.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu -k "$K" --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_c3_final.csv python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch3.log 2>&1
