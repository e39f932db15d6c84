# Round 2, GPU call 17 (EIGHT B200): BASELINE config 4 at N = 8 (64 M tank, 8 M per GPU) with asynchronous re-balancing; default-order worker at 8 ranks.
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c17
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-extra --no-cpu-baseline"
AKUA_SLAB_VERBOSE=1 timeout 300 $TR --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 $B --trace ${O}_trace_tank_n8 > ${O}_tank_n8.json 2> ${O}_tank_n8.err; tail -c 300 ${O}_tank_n8.json
( timeout 120 $TR --nproc-per-node 8 --master-port 29702 tests/mgpu_worker.py --scene tank --steps 20 --vx 1.5 --side 64; echo "exit $?" ) 2>&1 | grep -v Warn > ${O}_worker8_default_order.log; grep -E 'slab\(|FAIL|^exit|imbalance' ${O}_worker8_default_order.log | cut -c1-300
ls -la gpurun_out | grep c17
