# Round 2, GPU call 4 (two B200): one-sweep sort vs three-kernel sort; slab step with explicit boundary priority + re-balancing.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_c4_pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for cfg in "0 0" "1 4" "1 8" "1 16"; do set -- $cfg
  AKUA_SORT_MODE=$1 AKUA_SORT_ITEMS=$2 AKUA_TV_LAYOUTS=2 AKUA_TV_NSIDE=100,160,252 AKUA_TV_LIST_BUILD=1 timeout 300 python tools/time_variants.py 2>&1 | sed "s/^/sort_mode=$1 items=$2 /" | tee -a gpurun_out/r02_c4_sort_variants.txt
done
timeout 300 python bench.py --workload dam --n-side 100 --no-extra --no-cpu-baseline > gpurun_out/r02_c4_dam1m_n1.json 2> gpurun_out/r02_c4_dam1m_n1.err; tail -c 300 gpurun_out/r02_c4_dam1m_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 > gpurun_out/r02_c4_dam1m_n2.json 2> gpurun_out/r02_c4_dam1m_n2.err; tail -c 600 gpurun_out/r02_c4_dam1m_n2.json; grep -v Warn gpurun_out/r02_c4_dam1m_n2.err | tail -4
AKUA_SLAB_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck > gpurun_out/r02_c4_dam1m_n2_nograph.json 2> gpurun_out/r02_c4_dam1m_n2_nograph.err; tail -c 300 gpurun_out/r02_c4_dam1m_n2_nograph.json
AKUA_SLAB_BND_PRIORITY=0 timeout 300 $TR --nproc-per-node 2 --master-port 29604 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck > gpurun_out/r02_c4_dam1m_n2_noprio.json 2> gpurun_out/r02_c4_dam1m_n2_noprio.err; tail -c 300 gpurun_out/r02_c4_dam1m_n2_noprio.json
timeout 400 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r02_c4_tank_n1.json 2> gpurun_out/r02_c4_tank_n1.err; tail -c 300 gpurun_out/r02_c4_tank_n1.json
timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 > gpurun_out/r02_c4_tank_n2.json 2> gpurun_out/r02_c4_tank_n2.err; tail -c 600 gpurun_out/r02_c4_tank_n2.json; grep -v Warn gpurun_out/r02_c4_tank_n2.err | tail -4
ls -la gpurun_out
