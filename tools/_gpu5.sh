B="--steps 20 --no-cpu-baseline --no-quality --no-dropin"
SMALFIT_LIB=build/variants/clocks.so python bench.py --frames 16 --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin 2>/dev/null | grep cycles | sort | uniq -c | sort -rn | head -12
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d['value'],1), {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()})
PY
}
for fair in 2 4 8 16; do
  SMALFIT_RT_FAIR=$fair python bench.py $B > gpurun_out/r02e_f128_fair$fair.json 2>/dev/null; show gpurun_out/r02e_f128_fair$fair.json "frames128 fair$fair"
  SMALFIT_RT_FAIR=$fair timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:raster_tile_forward_kernel -s 6 -c 1 --csv --log-file gpurun_out/r02e_traffic_fair$fair.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2>&1
  grep -E "dram__bytes" gpurun_out/r02e_traffic_fair$fair.csv | awk -F'","' '{print "   ", $(NF-2), $NF}'
done
