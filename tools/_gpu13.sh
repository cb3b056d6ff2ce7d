python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_sizes.py -q -x --timeout 900 2>&1 | tail -3
for fr in 128 64 16; do python tools/ab_bench.py run --steps 20 -- --frames $fr 2>&1 | sed "s/^/frames$fr /"; done
for mi in 4096 24000; do SMALFIT_RT_MINITEM=$mi python tools/ab_bench.py run --steps 20 -- --frames 32 2>&1 | sed "s/^/frames32 min$mi /"; done
