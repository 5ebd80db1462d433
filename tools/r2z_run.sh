# round-2 consolidated measurement run (GPU box): tests, benches, ncu launch list + full captures, sanitizer
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_pytest.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2z.json 2> gpurun_out/bench_r2z.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2z_ref.json 2> gpurun_out/bench_r2z_ref.err; echo "ref rc=$?"
python bench.py --workload bi2 --steps 20 --warmup 5 --no-eager > gpurun_out/bench_r2z_bi2.json 2> gpurun_out/bench_r2z_bi2.err; echo "bi2 rc=$?"
python -c "
import json
for f in ['bench_r2z','bench_r2z_bi2']:
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['value'], d['e2e']['value'], 'chain', d['roofline']['us_per_launch'], d['roofline']['frac'], 'single', d.get('roofline_conv_single',{}).get('frac'), 'warp', d.get('roofline_warp',{}).get('frac'))
print(open('gpurun_out/bench_r2z_ref.json').read()[:300])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2z.csv python bench.py --steps 2 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_chain -c 1 -f -o gpurun_out/prof_chain_r2z python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tail_tcgen05 -c 1 -f -o gpurun_out/prof_tail_r2z python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_s2d_concat_w -c 1 -f -o gpurun_out/prof_warp_r2z python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tcgen05_kernel -s 3 -c 1 -f -o gpurun_out/prof_conv_r2z python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ls -la gpurun_out/*r2z.ncu-rep
SAN_TIMEOUT=300 SAN_MEMCHECK="conv_chain_1tile conv_chain_ragged_repeat conv_chain_few_ctas conv_tc_halo_64 conv_tc_tap_convT dgrad_tc_conv_mask_res step_golden_g1" SAN_RACECHECK="conv_chain_ragged_repeat conv_tc_halo_64" bash tools/sanitize.sh
