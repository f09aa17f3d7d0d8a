python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/gpu_tests_r2.log 2>&1; tail -3 gpurun_out/gpu_tests_r2.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for c in ensemble lm fc_lc ensemble1024 unet; do python bench.py --config $c --steps 10 --warmup 3 --cpu-frames 2 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; python - <<PY
import json
d=json.loads(open('gpurun_out/bench_r2_$c.json').read().strip().splitlines()[-1])
print('$c', round(d['value'],1), round(d['e2e']['value'],1), d['roofline']['frac'], {k:(v.get('ms_per_frame'),v.get('frac')) for k,v in d.get('per_network',{}).items()})
PY
done
for k in FC_LC VV LM; do python tools/bench_net.py $k 32 > gpurun_out/r2_net_${k}_final.log 2>&1; head -1 gpurun_out/r2_net_${k}_final.log | cut -c1-120; done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-rooflines > gpurun_out/plain_r2.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1800 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-rooflines > gpurun_out/ncu_r2_launches.log 2>&1; echo launches rc=$?
ncu --kernel-name-base demangled -k regex:"conv_tc_kernel<.*true>" -s 3 -c 1 --set full --clock-control none --import-source on -f -o gpurun_out/prof_conv_head_r2 python tools/bench_net.py FC_LC 32 > gpurun_out/ncu_head.log 2>&1; echo head rc=$?
python tools/bench_expand.py 32 > gpurun_out/expand_r2_final.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 30 -c 1 -f -o gpurun_out/prof_conv_expand48_r2 python tools/bench_expand.py 32 > gpurun_out/ncu_expand48.log 2>&1; echo expand rc=$?
