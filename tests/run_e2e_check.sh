for sk in none profile standalone probe profile,standalone,probe; do
DCA_BENCH_SKIP=$sk timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('skip=$sk value %.2fM ms %.4f e2e %.2fM ms %.4f'%(d['value']/1e6,d['ms_per_step'],e['value']/1e6,e['ms_per_step']))"
done
