# First GPU call of round 2 (one B200):  gpurun --timeout 1500 -- 'bash tools/r02_first_call.sh'
# Validates and times the opt-in list-build variants written without GPU access at the end of round 1
# (akuaengine_b200/csrc/list_build.cuh). Decision rule: if test_zz_list_build_gpu passes and mask4 or mask8 beats the scan
# kernel's "lists" time at both sizes, make it the default in akua_pbf_default_options and re-run bench.py + the launch list.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
AKUA_TV_LAYOUTS=2 AKUA_TV_NSIDE=100,160 AKUA_TV_LIST_BUILD=0,1,2 timeout 400 python tools/time_variants.py 2>&1 | tee gpurun_out/r02_list_build_variants.txt
# L1 gather cost model (lines vs sectors vs slot collisions): feeds tests/gather_locality_study.py
mkdir -p tools/_build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l1_gather_probe tools/l1_gather_probe.cu \
  && timeout 120 tools/_build/l1_gather_probe | tee gpurun_out/r02_l1_gather_probe.jsonl
timeout 200 ncu --metrics l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,gpu__time_duration.sum \
  --clock-control none --csv --log-file gpurun_out/r02_l1_gather_probe_ncu.csv tools/_build/l1_gather_probe > /dev/null 2>&1
for lb in 0 1 2; do
  AKUA_LIST_BUILD=$lb timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_build_neighbours -s 3 -c 1 \
    -o gpurun_out/r02_list_build_$lb python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_ncu_lb$lb.log 2>&1
done
for lb in 0 1 2; do
  AKUA_LIST_BUILD=$lb timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_lb$lb.json 2> gpurun_out/r02_bench_lb$lb.err
  tail -c 400 gpurun_out/r02_bench_lb$lb.json
done
# BASELINE config 5 at its largest size (round 1 stopped at 64 M), and the clustered scenes where the list build dominates,
# with the scan and the mask list build
for lb in 0 1; do
  AKUA_LIST_BUILD=$lb timeout 900 python tools/bench_neighbour_search.py --sizes 16,128 --reps 3 > gpurun_out/r02_neighbour_search_lb$lb.jsonl 2> gpurun_out/r02_neighbour_search_lb$lb.err
  tail -n 2 gpurun_out/r02_neighbour_search_lb$lb.jsonl | cut -c1-300
done
