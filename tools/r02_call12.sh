# Round 2, GPU call 12 (one B200): full GPU test suite, the default bench line as the driver runs it (wall clock), ncu launch list of the
# bench command, ncu --set full re-capture on the final kernel source.
#   gpurun --timeout 2400 -- 'bash tools/r02_call12.sh'
set -x
mkdir -p gpurun_out /tmp/ncu
O=gpurun_out/r02_c12
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee ${O}_pytest_gpu.log
/usr/bin/time -v -o ${O}_bench_default.time timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; grep -E "Elapsed|Maximum resident" ${O}_bench_default.time; tail -c 600 ${O}_bench_default.json
K='regex:k_density_lambda|k_delta_apply|k_vorticity|k_confinement|k_xsph|k_build_neighbours|k_onesweep|k_reorder_ranges|k_predict_key|k_hist'
timeout 600 ncu --set full --clock-control none -k "$K" -s $((18*100)) -c 18 -f -o /tmp/ncu/tank8m python tools/ncu_target.py tank200 100 3 > ${O}_ncu_tank8m.log 2>&1; tail -2 ${O}_ncu_tank8m.log
ncu -i /tmp/ncu/tank8m.ncu-rep --page raw --csv > ${O}_ncu_tank8m_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none -k "$K" -s $((18*40)) -c 18 -f -o /tmp/ncu/dam16m python tools/ncu_target.py dam252 40 3 > ${O}_ncu_dam16m.log 2>&1; tail -2 ${O}_ncu_dam16m.log
ncu -i /tmp/ncu/dam16m.ncu-rep --page raw --csv > ${O}_ncu_dam16m_raw.csv 2>/dev/null
# launch list of the bench command (graph-replayed steps; per-launch durations are cold-cache and serialised: SHARES of the step)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_|rsort' -s 900 -c 150 --csv --log-file ${O}_launches_tank8m.csv python bench.py --settle 40 --steps 5 --warmup 3 --windows 1 --no-extra --no-cpu-baseline > ${O}_launches_tank8m.log 2>&1; wc -l ${O}_launches_tank8m.csv
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > ${O}_bench_reference.json 2> ${O}_bench_reference.err; tail -c 400 ${O}_bench_reference.json
ls -la gpurun_out | grep c12
