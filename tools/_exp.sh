timeout 400 python tests/gpu_diag.py pool_upsample module_ops step_golden infer_sequence engine step_vs_oracle_fullsize bi2 2>&1 | cut -c1-170 | tail -12
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1w.json').read().strip().splitlines()[-1])
r=d['roofline']
print(round(d['value']), 'fps', round(d['ms_per_step']*1e3), 'us/step | e2e', round(d['e2e']['value']), '| chain', round(r['us_per_launch'],1), round(r['frac'],3), '| launches', d['gpu_launches'], d['clocks'])
PY
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1w.csv python bench.py --steps 2 --warmup 3 --profile-only > gpurun_out/ncu1.log 2>&1; tail -1 gpurun_out/ncu1.log | cut -c1-100
