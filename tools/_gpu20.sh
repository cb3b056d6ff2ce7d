python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], 'it/s', round(d['value'],1), 'ms', round(d['ms_per_step'],4), 'b2b', round(d['back_to_back_iters_per_s'],1), 'e2e', round(d['e2e']['value'],1), 'loss', d['final_loss'])
PY
}
B="--steps 40 --no-cpu-baseline --no-quality --no-dropin"
for fr in 128 16; do
  python bench.py --frames $fr $B > gpurun_out/r02o_f${fr}_pdl.json 2>/dev/null; show gpurun_out/r02o_f${fr}_pdl.json "frames$fr pdl"
  SMALFIT_NO_PDL=1 python bench.py --frames $fr $B > gpurun_out/r02o_f${fr}_nopdl.json 2>/dev/null; show gpurun_out/r02o_f${fr}_nopdl.json "frames$fr nopdl"
done
