# Round 2, GPU call 7 (one B200): ncu --set full at config-4 (8 M tank) and config-3 (16 M dam break) sizes, L1 gather probe,
# neighbour-order sensitivity, config-5 neighbour-search microbench 1 M - 128 M (uniform / clustered) with the mask4 list build.
#   gpurun --timeout 2400 -- 'bash tools/r02_call7.sh'
set -x
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/r02_c7
python -m akuaengine_b200.build >/dev/null 2>&1
timeout 120 python tools/order_sensitivity.py > ${O}_order_sensitivity.json 2> ${O}_order_sensitivity.err; tail -c 900 ${O}_order_sensitivity.json
mkdir -p tools/_build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l1_gather_probe tools/l1_gather_probe.cu \
  && timeout 120 tools/_build/l1_gather_probe > ${O}_l1_gather_probe.jsonl; wc -l ${O}_l1_gather_probe.jsonl
K='regex:k_density_lambda|k_delta_apply|k_vorticity|k_confinement|k_xsph|k_build_neighbours|k_onesweep|k_reorder_ranges|k_predict_key|k_hist'
# per step these match: 1 predict, 1 hist + 3 onesweep, 1 reorder, 1 build, 4 A, 4 B, 3 post = 18
timeout 600 ncu --set full --clock-control none -k "$K" -s $((18*100)) -c 18 -f -o /tmp/ncu/tank8m python tools/ncu_target.py tank200 100 3 > ${O}_ncu_tank8m.log 2>&1; tail -2 ${O}_ncu_tank8m.log
ncu -i /tmp/ncu/tank8m.ncu-rep --page raw --csv > ${O}_ncu_tank8m_raw.csv 2>/dev/null; ls -la /tmp/ncu
timeout 600 ncu --set full --clock-control none -k "$K" -s $((18*40)) -c 18 -f -o /tmp/ncu/dam16m python tools/ncu_target.py dam252 40 3 > ${O}_ncu_dam16m.log 2>&1; tail -2 ${O}_ncu_dam16m.log
ncu -i /tmp/ncu/dam16m.ncu-rep --page raw --csv > ${O}_ncu_dam16m_raw.csv 2>/dev/null
# the launch list of the bench command itself (per-launch durations, cold cache, serialised): kernel SHARES of the step
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|rsort' -s 2000 -c 120 --csv --log-file ${O}_launches_tank8m.csv python bench.py --settle 60 --steps 5 --warmup 3 --windows 1 --no-extra --no-cpu-baseline > ${O}_launches_tank8m.log 2>&1; tail -2 ${O}_launches_tank8m.csv
for f in tank8m dam16m; do s=$(stat -c %s /tmp/ncu/$f.ncu-rep 2>/dev/null || echo 0); if [ "$s" -gt 0 ] && [ "$s" -lt 22000000 ]; then cp /tmp/ncu/$f.ncu-rep ${O}_ncu_$f.ncu-rep; fi; done
# list-build variants after the culling was removed again, and the mask variants on the settled tank
AKUA_TV_LAYOUTS=2 AKUA_TV_NSIDE=100,200 AKUA_TV_LIST_BUILD=0,1,2 timeout 300 python tools/time_variants.py 2>&1 | tee ${O}_list_build_variants.txt
# config 5
timeout 900 python tools/bench_neighbour_search.py --sizes 1,16,64 --reps 3 > ${O}_neighbour_search.jsonl 2> ${O}_neighbour_search.err; cut -c1-330 ${O}_neighbour_search.jsonl
timeout 600 python tools/bench_neighbour_search.py --sizes 128 --reps 2 --linear-only >> ${O}_neighbour_search.jsonl 2>> ${O}_neighbour_search.err; tail -2 ${O}_neighbour_search.jsonl | cut -c1-330; tail -3 ${O}_neighbour_search.err
ls -la gpurun_out | grep c7
