# GPU box (8 GPUs): the 2/4/8-GPU scaling lines, the NCCL variant, config 4 on 8 GPUs and a 1-GPU line on the same box -> gpurun_out/r02_*
B="--steps 20 --warmup 5"
run() { n=$1; tag=$2; shift 2; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $n $B "$@" > gpurun_out/r02_${tag}.json 2> gpurun_out/r02_${tag}.err; }
run 8 bench_8gpu
run 4 bench_4gpu
run 2 bench_2gpu
run 8 bench_8gpu_nccl --collective nccl
run 8 config4_8gpu --workload config4
python bench.py $B --no-cpu-baseline --no-quality --no-dropin > gpurun_out/r02_bench_1gpu_samebox.json 2>/dev/null
python - <<'PY'
import json
for t in ('bench_1gpu_samebox','bench_2gpu','bench_4gpu','bench_8gpu','bench_8gpu_nccl','config4_8gpu'):
    try:
        d=json.loads(open(f'gpurun_out/r02_{t}.json').read().strip().splitlines()[-1])
        print(t, 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value'],1), 'coll', d['config']['collective'], 'replicas', d.get('replicas_identical'), 'loss', d['final_loss'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()}, 'mem', d['device_memory'])
    except Exception as e:
        print(t, 'FAILED', e); print(open(f'gpurun_out/r02_{t}.err').read()[-800:])
PY
