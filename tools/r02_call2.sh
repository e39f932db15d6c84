# Round 2, GPU call 2 (two B200): the device-driven x-slab step on real GPUs.
#   gpurun --gpus 2 --timeout 1500 -- 'bash tools/r02_call2.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm --format=csv
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/r02_c2_pytest_gpu.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# 1 M per GPU weak scaling (the round-1 driver bench), 1 vs 2 GPUs
timeout 300 python bench.py --workload dam --n-side 100 --no-extra --no-cpu-baseline > gpurun_out/r02_c2_dam1m_n1.json 2> gpurun_out/r02_c2_dam1m_n1.err; tail -c 600 gpurun_out/r02_c2_dam1m_n1.json
timeout 300 $TR --nproc-per-node 2 --master-port 29601 bench.py --gpus 2 --workload dam --n-side 100 > gpurun_out/r02_c2_dam1m_n2.json 2> gpurun_out/r02_c2_dam1m_n2.err; tail -c 1500 gpurun_out/r02_c2_dam1m_n2.json; tail -5 gpurun_out/r02_c2_dam1m_n2.err
AKUA_SLAB_GRAPH=0 timeout 300 $TR --nproc-per-node 2 --master-port 29602 bench.py --gpus 2 --workload dam --n-side 100 --no-selfcheck > gpurun_out/r02_c2_dam1m_n2_nograph.json 2> gpurun_out/r02_c2_dam1m_n2_nograph.err; tail -c 400 gpurun_out/r02_c2_dam1m_n2_nograph.json
# default bench (config 4, tank 8 M per GPU): N = 1 with the reference arm, N = 2
timeout 400 python bench.py --impl reference > gpurun_out/r02_c2_ref_n1.json 2> gpurun_out/r02_c2_ref_n1.err; tail -c 700 gpurun_out/r02_c2_ref_n1.json
timeout 400 python bench.py > gpurun_out/r02_c2_tank_n1.json 2> gpurun_out/r02_c2_tank_n1.err; tail -c 2500 gpurun_out/r02_c2_tank_n1.json; tail -3 gpurun_out/r02_c2_tank_n1.err
timeout 400 $TR --nproc-per-node 2 --master-port 29603 bench.py --gpus 2 > gpurun_out/r02_c2_tank_n2.json 2> gpurun_out/r02_c2_tank_n2.err; tail -c 2500 gpurun_out/r02_c2_tank_n2.json; tail -5 gpurun_out/r02_c2_tank_n2.err
ls -la gpurun_out
