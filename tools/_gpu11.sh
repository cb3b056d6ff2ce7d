python -m pytest tests/test_gpu_peer.py tests/test_gpu_fit_parity.py -q --timeout 900 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/peer_check.py 2>&1 | grep "^{" | tail -1
B="--steps 20 --warmup 5"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 $B > gpurun_out/r02k_2gpu_peer.json 2> gpurun_out/r02k_2gpu_peer.err; tail -3 gpurun_out/r02k_2gpu_peer.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 $B --collective nccl > gpurun_out/r02k_2gpu_nccl.json 2>/dev/null
python - <<'PY'
import json
for n in ('peer','nccl'):
    try:
        d=json.loads(open(f'gpurun_out/r02k_2gpu_{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'launches', d['launches_per_step'], d['config']['collective'], 'replicas', d.get('replicas_identical'), 'loss', d['final_loss'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()})
    except Exception as e: print(n, 'failed', e)
PY
