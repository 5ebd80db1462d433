mkdir -p gpurun_out
timeout 600 python tests/gpu_diag.py conv_pool step_golden infer_sequence_golden ragged_sizes engine_matches > gpurun_out/r2k_diag.log 2>&1; grep -E "^(PASS|FAIL)" gpurun_out/r2k_diag.log | cut -c1-200; grep -A12 "^FAIL" gpurun_out/r2k_diag.log | head -30
for k in 1 0; do TECOGAN_B200_POOL=$k python bench.py --steps 20 --warmup 5 --no-eager --sustain-s 0 > gpurun_out/bench_r2k_pool$k.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/bench_r2k_pool$k.json').read().strip().splitlines()[-1])
print('pool=$k', d['ms_per_step'], d['value'], d['e2e']['value'], d['launches_per_step'])
"; done
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2k_pytest.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 2400 --csv --log-file gpurun_out/launches_train_r2k.csv python bench.py --workload train-frvsr --steps 1 --warmup 1 --no-eager > /dev/null 2>&1; echo "ncu train rc=$?"
