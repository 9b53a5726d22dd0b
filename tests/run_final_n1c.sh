timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -1 gpurun_out/pytest_gpu.log
timeout 100 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_c2_final.log 2>&1
timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_final.log 2>&1
for f in bench_c2_final bench_c3_final; do tail -1 gpurun_out/$f.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('$f value %.2fM ms %.4f frac %.3f e2e %.2fM ms %.4f passes %s'%(d['value']/1e6,d['ms_per_step'],d['roofline']['frac'],e['value']/1e6,e['ms_per_step'],e['ms_per_step_passes']))"; done
