set -x
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_c2_final.log 2>&1; tail -1 gpurun_out/bench_c2_final.log | cut -c1-200
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_final.log 2>&1; tail -1 gpurun_out/bench_ref_final.log | cut -c1-300
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_final.log 2>&1; tail -1 gpurun_out/bench_c3_final.log | cut -c1-200
timeout 100 python bench.py --fused --no-cpu-baseline --no-e2e > gpurun_out/bench_c2_fused.log 2>&1
timeout 200 python bench.py --fused --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_fused.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_c2_final.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_c3_final.csv python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launch3.log 2>&1
