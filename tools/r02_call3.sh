# Round 2, GPU call 3 (two B200): graph-captured x-slab step, list-build variants with reachability culling.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_call3.sh'
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/r02_c3_pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# list build: scan / mask4 / mask8 (mask variants now cull unreachable cells), 1 M and 4 M dam break, one GPU
AKUA_TV_LAYOUTS=2 AKUA_TV_NSIDE=100,160 AKUA_TV_LIST_BUILD=0,1,2 timeout 400 python tools/time_variants.py 2>&1 | tee gpurun_out/r02_c3_list_build_variants.txt
# 1 M per GPU weak scaling, 1 vs 2 GPUs, graph vs eager
timeout 300 python bench.py --workload dam --n-side 100 --no-extra --no-cpu-baseline > gpurun_out/r02_c3_dam1m_n1.json 2> gpurun_out/r02_c3_dam1m_n1.err; tail -c 300 gpurun_out/r02_c3_dam1m_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 > gpurun_out/r02_c3_dam1m_n2.json 2> gpurun_out/r02_c3_dam1m_n2.err; tail -c 1800 gpurun_out/r02_c3_dam1m_n2.json; grep -v Warn gpurun_out/r02_c3_dam1m_n2.err | tail -5
AKUA_SLAB_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck > gpurun_out/r02_c3_dam1m_n2_nograph.json 2> gpurun_out/r02_c3_dam1m_n2_nograph.err; tail -c 300 gpurun_out/r02_c3_dam1m_n2_nograph.json
# default bench (config 4, tank 8 M per GPU): N = 1, N = 2
timeout 400 python bench.py --no-extra --no-cpu-baseline > gpurun_out/r02_c3_tank_n1.json 2> gpurun_out/r02_c3_tank_n1.err; tail -c 300 gpurun_out/r02_c3_tank_n1.json
timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 > gpurun_out/r02_c3_tank_n2.json 2> gpurun_out/r02_c3_tank_n2.err; tail -c 2500 gpurun_out/r02_c3_tank_n2.json; grep -v Warn gpurun_out/r02_c3_tank_n2.err | tail -5
ls -la gpurun_out
