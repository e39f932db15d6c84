# Round 2, GPU call 15 (EIGHT B200): multi-GPU parity at 8 and 4 ranks (canonical order: bit-identical to one GPU with migration,
# re-balancing and a skewed start), BASELINE config 4 at N = 8 (64 M tank, 8 M per GPU) and the 1 M-per-GPU dam break at N = 8.
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/r02_call9.sh'
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c15
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv | tee ${O}_smi.txt
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
( timeout 200 $TR --nproc-per-node 8 --master-port 29701 tests/mgpu_worker.py --scene tank --steps 30 --vx 1.0 --side 64 --canonical --rebalance-every 5 --skew 0.4; echo "exit $?" ) 2>&1 | grep -v Warn | tail -14 | tee ${O}_worker8_canonical_rebalance.log
( timeout 200 $TR --nproc-per-node 8 --master-port 29702 tests/mgpu_worker.py --scene tank --steps 20 --vx 1.5 --side 64; echo "exit $?" ) 2>&1 | grep -v Warn > ${O}_worker8_default_order.log; grep -E 'slab\(|FAIL|exit|imbalance' ${O}_worker8_default_order.log | cut -c1-300
( timeout 200 $TR --nproc-per-node 4 --master-port 29703 tests/mgpu_worker.py --scene tank --steps 24 --vx 1.5 --side 48 --canonical; echo "exit $?" ) 2>&1 | grep -v Warn | tail -14 | tee ${O}_worker4_canonical.log
B="--no-extra --no-cpu-baseline"
AKUA_SLAB_VERBOSE=1 timeout 400 $TR --nproc-per-node 8 --master-port 29704 bench.py --gpus 8 $B --trace ${O}_trace_tank_n8 > ${O}_tank_n8.json 2> ${O}_tank_n8.err; tail -c 500 ${O}_tank_n8.json; grep -v Warn ${O}_tank_n8.err | tail -4
timeout 300 $TR --nproc-per-node 8 --master-port 29705 bench.py --gpus 8 --workload dam --n-side 100 --no-selfcheck $B > ${O}_dam1m_n8.json 2> ${O}_dam1m_n8.err; tail -c 300 ${O}_dam1m_n8.json; grep -v Warn ${O}_dam1m_n8.err | tail -4
ls -la gpurun_out | grep c9
