mkdir -p gpurun_out
P=29511
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r2i_n2.json 2> gpurun_out/bench_r2i_n2.err; echo "infer n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --workload train-frvsr --steps 3 --warmup 1 > gpurun_out/bench_r2i_train_frvsr_n2.json 2> gpurun_out/bench_r2i_train_frvsr_n2.err; echo "train-frvsr n2 rc=$?"; tail -2 gpurun_out/bench_r2i_train_frvsr_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 2 --workload train --steps 2 --warmup 1 > gpurun_out/bench_r2i_train_n2.json 2> gpurun_out/bench_r2i_train_n2.err; echo "train n2 rc=$?"; tail -2 gpurun_out/bench_r2i_train_n2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+3)) bench.py --gpus 2 --workload bi2 --steps 20 --warmup 5 --no-eager > gpurun_out/bench_r2i_bi2_n2.json 2> gpurun_out/bench_r2i_bi2_n2.err; echo "bi2 n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+4)) bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2i_ref_n2.json 2>/dev/null; echo "ref n2 rc=$?"
python - <<'PY'
import json
for f in ['bench_r2i_n2','bench_r2i_train_frvsr_n2','bench_r2i_train_n2','bench_r2i_bi2_n2','bench_r2i_ref_n2']:
    try:
        d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1])
        print(f, d.get('n_gpus'), d.get('ms_per_step'), d.get('value'), d.get('e2e',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
