set -x
python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/r02d_tests.log 2>&1; tail -8 gpurun_out/r02d_tests.log
B="--steps 20 --no-cpu-baseline --no-quality --no-dropin"
show() { python - "$1" "$2" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], round(d['value'],1), 'launches', d['launches_per_step'], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()}, 'loss', d['final_loss'])
PY
}
for fr in 128 16; do
  python bench.py --frames $fr $B > gpurun_out/r02d_f${fr}_cluster.json 2>/dev/null; show gpurun_out/r02d_f${fr}_cluster.json "frames$fr cluster"
  SMALFIT_SPLIT_FRONT=1 python bench.py --frames $fr $B > gpurun_out/r02d_f${fr}_split.json 2>/dev/null; show gpurun_out/r02d_f${fr}_split.json "frames$fr splitfront"
done
# DRAM traffic of the forward by L2-hint variant (and a small list scratch)
for v in base nohints nodiscard; do
  SMALFIT_LIB=build/variants/$v.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:raster_tile_forward_kernel -s 6 -c 1 --csv --log-file gpurun_out/r02d_traffic_$v.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2>&1
  echo "traffic $v"; grep -E "dram__bytes|gpu__time" gpurun_out/r02d_traffic_$v.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
done
SMALFIT_RT_LISTCAP=32768 SMALFIT_LIB=build/variants/base.so timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:raster_tile_forward_kernel -s 6 -c 1 --csv --log-file gpurun_out/r02d_traffic_cap32k.csv python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2>&1
echo "traffic cap32k"; grep -E "dram__bytes|gpu__time" gpurun_out/r02d_traffic_cap32k.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}'
# forward at 16 frames: occupancy / busy
timeout 300 ncu --set full --clock-control none -k regex:raster_tile_forward_kernel -s 6 -c 1 -o gpurun_out/r02d_fwd_f16 -f python bench.py --frames 16 --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2>&1
ls -la gpurun_out/r02d_fwd_f16.ncu-rep
