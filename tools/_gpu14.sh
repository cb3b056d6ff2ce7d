show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d['value'],1), 'fwd', round(d['roofline']['phase_ms']['raster_forward'],4), 'total', round(d['roofline']['phase_ms']['total'],4))
PY
}
B="--steps 20 --no-cpu-baseline --no-quality --no-dropin"
for fr in 16 32 64 128; do
 for mi in 4096 16000 32000 48000; do
  SMALFIT_RT_MINITEM=$mi python bench.py --frames $fr $B > gpurun_out/r02m_f${fr}_m${mi}.json 2>/dev/null; show gpurun_out/r02m_f${fr}_m${mi}.json "frames$fr minitem$mi"
 done
done
