# Round 2, GPU call 13 (one B200): the default bench line exactly as the driver runs it, with its wall clock.
set -x
mkdir -p gpurun_out
O=gpurun_out/r02_c13
S=$(date +%s.%N)
timeout 900 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err
E=$(date +%s.%N); echo "wall_s $(echo "$E - $S" | bc)" | tee ${O}_bench_default.wall
tail -c 400 ${O}_bench_default.json; tail -5 ${O}_bench_default.err
S=$(date +%s.%N)
timeout 900 python bench.py --steps 20 --warmup 5 > ${O}_bench_s20w5.json 2> ${O}_bench_s20w5.err
E=$(date +%s.%N); echo "wall_s $(echo "$E - $S" | bc)" | tee ${O}_bench_s20w5.wall
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee ${O}_smoke.log
ls -la gpurun_out | grep c13
