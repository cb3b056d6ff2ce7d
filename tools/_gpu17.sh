python bench.py --workload config4 --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_config4_1gpu.json 2> gpurun_out/r02_config4_1gpu.err; tail -3 gpurun_out/r02_config4_1gpu.err; head -c 1500 gpurun_out/r02_config4_1gpu.json; echo
python tools/run_configs.py config2 config5 > gpurun_out/r02_run_configs.log 2>&1; tail -12 gpurun_out/r02_run_configs.log | cut -c1-600
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; head -c 600 gpurun_out/r02_bench_1gpu.json; echo
bash tools/capture_profiles.sh r02 2>&1 | tail -3
ls -la gpurun_out/r02_*ncu-rep
