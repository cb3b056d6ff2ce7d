set -x
python tools/ab_bench.py run --steps 20 2>&1 | tee gpurun_out/r02c_ab.log
for fair in 1 2 4 8 16; do
  SMALFIT_LIB=build/variants/bw4pack.so SMALFIT_RT_FAIR=$fair python bench.py --frames 16 --steps 20 --no-cpu-baseline --no-quality --no-dropin > gpurun_out/r02c_f16_fair$fair.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r02c_f16_fair$fair.json').read().strip().splitlines()[-1])
print('frames16 fair$fair', round(d['value'],1), {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()})
PY
done
# ncu: one launch of each raster kernel, after warm-up
for v in bw4pack bw4nopack; do
SMALFIT_LIB=build/variants/$v.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:raster_backward_kernel -s 6 -c 1 -o gpurun_out/r02c_bwd_$v -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2> gpurun_out/r02c_ncu_$v.err
done
SMALFIT_LIB=build/variants/bw4pack.so timeout 600 ncu --set full --import-source on --clock-control none -k regex:raster_tile_forward_kernel -s 6 -c 1 -o gpurun_out/r02c_fwd_pack -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2> gpurun_out/r02c_ncu_fwd.err
ls -la gpurun_out/*.ncu-rep
