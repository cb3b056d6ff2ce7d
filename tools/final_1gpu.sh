# GPU box (1 GPU): the round-end sequence -- GPU test suite, ncu captures, default bench, reference arm, config 4 slice, configs 2 and 5 -> gpurun_out/
python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
bash tools/capture_profiles.sh r02 2>&1 | tail -2
cp gpurun_out/r02_raster_forward_traffic.json profiles/raster_forward_traffic.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; head -c 400 gpurun_out/r02_bench_1gpu.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_reference_arm.json 2>/dev/null; head -c 300 gpurun_out/r02_reference_arm.json; echo
python bench.py --workload config4 --steps 20 --warmup 5 > gpurun_out/r02_config4_1gpu.json 2>/dev/null; head -c 300 gpurun_out/r02_config4_1gpu.json; echo
python tools/run_configs.py config2 config5 > gpurun_out/r02_run_configs.log 2>&1; grep -c sweep gpurun_out/r02_run_configs.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_1gpu.json').read().strip().splitlines()[-1])
print(round(d['value'],1), d['ms_per_step'], 'e2e', d['e2e']['value'], 'traffic', d['roofline']['traffic'], d['roofline']['traffic_note'][:60], {k:round(v,4) for k,v in d['roofline']['phase_ms'].items()})
print('dropin', d.get('dropin_api',{}).get('value'), 'quality', d.get('quality',{}).get('d_kp_l2_vs_oracle'), d.get('quality',{}).get('d_iou_vs_oracle'))
PY
