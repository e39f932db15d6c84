# Round 2, GPU call 5 (two B200): one-sweep sort on the GPU for the first time; launch timelines of the x-slab step.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_call5.sh'
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c5
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee ${O}_pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for cfg in "0 0" "1 0" "1 4" "1 16"; do set -- $cfg
  AKUA_SORT_MODE=$1 AKUA_SORT_ITEMS=$2 AKUA_TV_LAYOUTS=2 AKUA_TV_NSIDE=100,200 AKUA_TV_LIST_BUILD=1 timeout 300 python tools/time_variants.py 2>&1 | sed "s/^/sort_mode=$1 items=$2 /" | tee -a ${O}_sort_variants.txt
done
B="--no-extra --no-cpu-baseline"
timeout 300 python bench.py --workload dam --n-side 100 $B --trace ${O}_trace_dam1m_n1 > ${O}_dam1m_n1.json 2> ${O}_dam1m_n1.err; tail -c 300 ${O}_dam1m_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 $B --trace ${O}_trace_dam1m_n2 > ${O}_dam1m_n2.json 2> ${O}_dam1m_n2.err; tail -c 600 ${O}_dam1m_n2.json; grep -v Warn ${O}_dam1m_n2.err | tail -4
AKUA_SLAB_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck $B > ${O}_dam1m_n2_nograph.json 2> ${O}_dam1m_n2_nograph.err; tail -c 300 ${O}_dam1m_n2_nograph.json
AKUA_SLAB_BND_PRIORITY=0 timeout 300 $TR --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck $B > ${O}_dam1m_n2_noprio.json 2> ${O}_dam1m_n2_noprio.err; tail -c 300 ${O}_dam1m_n2_noprio.json
timeout 300 $TR --nproc-per-node 2 --master-port 29605 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck --rebalance-every 0 $B > ${O}_dam1m_n2_norebal.json 2> ${O}_dam1m_n2_norebal.err; tail -c 300 ${O}_dam1m_n2_norebal.json
timeout 400 python bench.py $B --trace ${O}_trace_tank_n1 > ${O}_tank_n1.json 2> ${O}_tank_n1.err; tail -c 300 ${O}_tank_n1.json
timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 --no-selfcheck $B --trace ${O}_trace_tank_n2 > ${O}_tank_n2.json 2> ${O}_tank_n2.err; tail -c 600 ${O}_tank_n2.json; grep -v Warn ${O}_tank_n2.err | tail -4
ls -la gpurun_out
