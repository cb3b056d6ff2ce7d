python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d['value'],1), 'launches', d['launches_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()}, 'loss', d['final_loss'])
PY
}
B="--steps 20 --no-cpu-baseline --no-quality --no-dropin"
for fr in 128 16; do python bench.py --frames $fr $B > gpurun_out/r02n_f${fr}.json 2>/dev/null; show gpurun_out/r02n_f${fr}.json "frames$fr"; done
