python -m pytest tests/test_gpu_parity.py tests/test_gpu_fit_parity.py -q --timeout 900 > gpurun_out/r02j_tests.log 2>&1; tail -4 gpurun_out/r02j_tests.log
SMALFIT_LIB=build/variants/clocks.so python bench.py --frames 16 --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin 2>/dev/null | grep "cycles" | sort | uniq -c | sort -rn | awk '{$1="";print}' | awk '!seen[$1]++' | head -4
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d['value'],1), 'launches', d['launches_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()}, 'loss', d['final_loss'])
PY
}
B="--steps 20 --no-cpu-baseline --no-quality --no-dropin"
for fr in 128 16; do
  python bench.py --frames $fr $B > gpurun_out/r02j_f${fr}.json 2>/dev/null; show gpurun_out/r02j_f${fr}.json "frames$fr"
done
