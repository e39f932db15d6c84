# Round 2, GPU call 8 (two B200): 2048-particle migration tiles, publish folded into the plan kernel, plane check folded into the push.
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/r02_call8.sh'
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c8
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_abi.py -x -q -m gpu 2>&1 | tail -6 | tee ${O}_pytest_mgpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-extra --no-cpu-baseline"
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 $B --trace ${O}_trace_dam1m_n2 > ${O}_dam1m_n2.json 2> ${O}_dam1m_n2.err; tail -c 400 ${O}_dam1m_n2.json; grep -v Warn ${O}_dam1m_n2.err | tail -4
timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --no-selfcheck $B --trace ${O}_trace_tank_n2 > ${O}_tank_n2.json 2> ${O}_tank_n2.err; tail -c 400 ${O}_tank_n2.json; grep -v Warn ${O}_tank_n2.err | tail -4
ls -la gpurun_out | grep c8
