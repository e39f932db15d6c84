set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -15
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tests/mgpu_worker.py --scene dam --steps 8 2>&1 | grep -E "slab\(|FAIL|ranks own" 
AKUA_SLAB_P2P=0 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 tests/mgpu_worker.py --scene dam --steps 20 --vx 1.5 2>&1 | grep -E "slab\(|FAIL|ranks own|migrated"
AKUA_SLAB_FUSED_PUSH=1 timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tests/mgpu_worker.py --scene dam --steps 20 --vx 1.5 2>&1 | grep -E "slab\(|FAIL|ranks own|migrated"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 20 --no-cpu-baseline > gpurun_out/bench_2gpu_r01f.json 2> gpurun_out/bench_2gpu_r01f.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_2gpu_r01f.json').read())
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['phases_ms'])
"
tail -3 gpurun_out/bench_2gpu_r01f.err
