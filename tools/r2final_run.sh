# final consolidated run of round 2 (GPU box)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2final_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2final.json 2> gpurun_out/bench_r2final.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2final_ref.json 2> gpurun_out/bench_r2final_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/bench_r2final.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], 'chain', d['roofline']['us_per_launch'], d['roofline']['frac'], 'single', d['roofline_conv_single']['frac'], 'warp', d['roofline_warp']['frac'], d['roofline_warp_fused_lrflow']['frac'], 'eager', d.get('gpu_eager_baseline'))
"
