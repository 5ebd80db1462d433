timeout 200 python tests/gpu_diag.py warp step_golden_g15 bi2 2>&1 | cut -c1-150 | tail -8
for c in 1 2 4; do
TECOGAN_B200_TAIL_CHUNKS=$c python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1v_c$c.json 2> gpurun_out/bench_r1v_c$c.err
done
python - <<'PY'
import json
for c in (1,2,4):
    try:
        d=json.loads(open(f'gpurun_out/bench_r1v_c{c}.json').read().strip().splitlines()[-1])
        print('chunks',c, round(d['value']), 'fps', round(d['ms_per_step']*1e3), 'us/step | e2e', round(d['e2e']['value']), '| warp', {k: (round(v['us_per_launch'],1), round(v['frac'],3)) for k,v in d.items() if k.startswith('roofline_warp')}, d['gpu_launches'])
    except Exception as e:
        print(c, 'ERR', e); print(open(f'gpurun_out/bench_r1v_c{c}.err').read()[-500:])
PY
