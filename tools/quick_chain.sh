#!/bin/bash
# usage: tools/quick_chain.sh TAG [ablation combos]   (GPU box: chain checks, ablation table, event trace, short bench)
export TAG=$1
python -m pytest tests -m gpu -q -x -k "chain or step_golden" 2>&1 | tail -3
TG_ABLATE_COMBOS=${2:-0,128,2,130,1,16} python tools/conv_timers.py chain-ablate > gpurun_out/chain_ablate_$TAG.log 2>&1
python tools/conv_timers.py chain > gpurun_out/chain_trace_$TAG.log 2>&1
python bench.py --steps 20 --warmup 5 --no-eager > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<'PY'
import json, os
tag = os.environ['TAG']
d = json.load(open(f'gpurun_out/bench_{tag}.json'))
print('ms/step', d['ms_per_step'], 'fps', d['value'], 'chain us', d['roofline']['us_per_launch'], 'frac', d['roofline']['frac'],
      'single', d['roofline_conv_single']['us_per_launch'])
for line in open(f'gpurun_out/chain_ablate_{tag}.log'):
    if not line.startswith('{'):
        print(line.strip())
        continue
    d = json.loads(line)
    p = d['per_tile']
    print(f"abl={d['ablate']:3d} us={d['us']:6.1f} cyc/tile={d['kernel_cycles_per_tile']:5d} flags={p['prod_wait_flags']:5d} "
          f"mma_wait={p['mma_wait']:5d} i0={p['mma_issue0']:5d} i1={p['mma_issue1']:5d} bd={p['mma_boundary']:4d} "
          f"epi_tfull={p['epi_wait_tfull']:5d} | chk iters={p['chk_iters']} fence={p['chk_fence']} hits={p['chk_hits']} total={p['chk_total']}")
for line in open(f'gpurun_out/chain_trace_{tag}.log'):
    if 'trace_cta0' in line:
        for k, v in json.loads(line)['trace_cta0 p10/p50/p90 cycles'].items():
            print(v, k)
PY
