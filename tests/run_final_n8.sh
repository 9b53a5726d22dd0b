# 8-GPU box: scaling of the resident training step (weak scaling, 4096 cells per GPU per step)
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/scale_c2_n1.log 2>&1
timeout 200 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --no-cpu-baseline --no-e2e > gpurun_out/scale_c2_n2.log 2>&1
timeout 200 $TR --nproc-per-node 4 --master-port 29512 bench.py --gpus 4 --no-cpu-baseline --no-e2e > gpurun_out/scale_c2_n4.log 2>&1
timeout 200 $TR --nproc-per-node 8 --master-port 29513 bench.py --gpus 8 --no-cpu-baseline --no-e2e > gpurun_out/scale_c2_n8.log 2>&1
timeout 300 python bench.py --workload c5shard --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/scale_c5_n1.log 2>&1
timeout 300 $TR --nproc-per-node 8 --master-port 29514 bench.py --gpus 8 --workload c5shard --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/scale_c5_n8.log 2>&1
timeout 200 $TR --nproc-per-node 8 --master-port 29515 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/scale_c2_n8_e2e.log 2>&1
for f in scale_c2_n1 scale_c2_n2 scale_c2_n4 scale_c2_n8 scale_c5_n1 scale_c5_n8 scale_c2_n8_e2e; do tail -1 gpurun_out/$f.log | cut -c1-260; done
