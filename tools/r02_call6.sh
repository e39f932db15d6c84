# Round 2, GPU call 6 (two B200): fused interior+boundary sweeps (one launch per sweep), canonical order, work-weighted re-balancing.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_call6.sh'
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c6
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee ${O}_pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="--no-extra --no-cpu-baseline"
timeout 300 python bench.py --workload dam --n-side 100 $B > ${O}_dam1m_n1.json 2> ${O}_dam1m_n1.err; tail -c 300 ${O}_dam1m_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 $B --trace ${O}_trace_dam1m_n2 > ${O}_dam1m_n2.json 2> ${O}_dam1m_n2.err; tail -c 600 ${O}_dam1m_n2.json; grep -v Warn ${O}_dam1m_n2.err | tail -4
AKUA_SLAB_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck $B > ${O}_dam1m_n2_nograph.json 2> ${O}_dam1m_n2_nograph.err; tail -c 300 ${O}_dam1m_n2_nograph.json
timeout 400 python bench.py $B > ${O}_tank_n1.json 2> ${O}_tank_n1.err; tail -c 300 ${O}_tank_n1.json
timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --no-selfcheck $B --trace ${O}_trace_tank_n2 > ${O}_tank_n2.json 2> ${O}_tank_n2.err; tail -c 600 ${O}_tank_n2.json; grep -v Warn ${O}_tank_n2.err | tail -4
ls -la gpurun_out | grep c6
