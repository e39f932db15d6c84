# Round 2, GPU call 1 (one B200): baseline of the round-1 code where round 1 never measured it.
#   gpurun --timeout 1200 -- 'bash tools/r02_call1.sh'
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv | tee gpurun_out/r02_c1_smi.txt
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_c1_pytest_gpu.log
# list-build variants (scan / mask4 / mask8), 1 M and 4 M dam break
AKUA_TV_LAYOUTS=2 AKUA_TV_NSIDE=100,160 AKUA_TV_LIST_BUILD=0,1,2 timeout 400 python tools/time_variants.py 2>&1 | tee gpurun_out/r02_c1_list_build_variants.txt
# config 4 per-GPU size (tank 200^3 = 8 M) and config 3 (16 M dam break) on one GPU, phase timings included
timeout 300 python bench.py --workload tank --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c1_bench_tank8m.json 2> gpurun_out/r02_c1_bench_tank8m.err
tail -c 1500 gpurun_out/r02_c1_bench_tank8m.json
timeout 300 python bench.py --n-side 252 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/r02_c1_bench_dam16m.json 2> gpurun_out/r02_c1_bench_dam16m.err
tail -c 1500 gpurun_out/r02_c1_bench_dam16m.json
# ncu --set full of every kernel of one step at 16 M (nothing fits in L2 there)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:'k_|rsort' -s 150 -c 40 \
  -o gpurun_out/r02_c1_ncu_16m python bench.py --n-side 252 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c1_ncu_16m.log 2>&1
tail -3 gpurun_out/r02_c1_ncu_16m.log
# L1 gather cost model probe
mkdir -p tools/_build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l1_gather_probe tools/l1_gather_probe.cu \
  && timeout 120 tools/_build/l1_gather_probe | tee gpurun_out/r02_c1_l1_gather_probe.jsonl | tail -5
ls -la gpurun_out
