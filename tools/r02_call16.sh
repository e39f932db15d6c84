# Round 2, GPU call 16 (two B200): asynchronous re-balancing (no host synchronisation in the timed region), wider windows.
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c16
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -5 | tee ${O}_pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-extra --no-cpu-baseline"
AKUA_SLAB_VERBOSE=1 timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 $B > ${O}_tank_n2.json 2> ${O}_tank_n2.err; tail -c 300 ${O}_tank_n2.json; grep "akua" ${O}_tank_n2.err | tail -6
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck $B > ${O}_dam1m_n2.json 2> ${O}_dam1m_n2.err; tail -c 300 ${O}_dam1m_n2.json
ls -la gpurun_out | grep c16
