mkdir -p gpurun_out
timeout 900 python tests/gpu_diag.py warp_ fused_tail engine_matches step_golden_g15 > gpurun_out/r2g_diag.log 2>&1; grep -E "^(PASS|FAIL)" gpurun_out/r2g_diag.log | cut -c1-160
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2g_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench rc=$?"
TG_WARP_KERNEL=cta python bench.py --steps 20 --warmup 5 --no-eager --sustain-s 0 > gpurun_out/bench_r2g_warpcta.json 2>/dev/null
python -c "
import json
for f in ['bench_r2g','bench_r2g_warpcta']:
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
    print(f, d['ms_per_step'], d['value'], d['e2e']['value'], 'warp', d['roofline_warp']['us_per_launch'], d['roofline_warp']['frac'], d['roofline_warp_fused_lrflow']['us_per_launch'], 'chain', d['roofline']['frac'])
"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2g.csv python bench.py --steps 2 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_chain -c 1 -f -o gpurun_out/prof_chain_r2g python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:tail_tcgen05 -c 1 -f -o gpurun_out/prof_tail_r2g python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:warp_s2d_concat_w -c 1 -f -o gpurun_out/prof_warp_r2g python bench.py --steps 1 --warmup 3 --profile-only > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:wgrad_tcgen05 -s 20 -c 1 -f -o gpurun_out/prof_wgrad_r2g python bench.py --workload train-frvsr --steps 1 --warmup 1 --batch 8 --no-eager > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
SAN_TIMEOUT=300 SAN_MEMCHECK="dgrad_tc_conv_mask_res dgrad_tc_convT wgrad_conv_ragged wgrad_convT_ragged backward_elementwise fused_tail_accumulate_bd4 st_discriminator_input warp_lrflow_bd4" SAN_RACECHECK="wgrad_conv_ragged fused_tail_accumulate_bd4 warp_lrflow_bd4" bash tools/sanitize.sh
