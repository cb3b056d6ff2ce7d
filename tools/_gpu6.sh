SMALFIT_LIB=build/variants/clocks.so python bench.py --frames 16 --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin 2>/dev/null | grep cycles | sort | uniq -c | sort -rn | head -8
mv build/variants/clocks.so build/variants/clocks.so.skip
python tools/ab_bench.py run --steps 20 2>&1 | tee gpurun_out/r02f_ab.log
python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_sizes.py -q -x --timeout 900 2>&1 | tail -3
