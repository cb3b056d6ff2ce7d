SMALFIT_LIB=build/variants/clocks.so python bench.py --frames 16 --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin 2>/dev/null | grep "frame_backward" | sort | uniq -c | sort -rn | head -4
mv build/variants/clocks.so build/variants/clocks.so.skip
python tools/ab_bench.py run --steps 20 2>&1 | tee gpurun_out/r02h_ab.log
