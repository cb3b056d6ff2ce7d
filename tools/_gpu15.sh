python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
for fr in 128 64 32 16; do python tools/ab_bench.py run --steps 20 -- --frames $fr 2>&1 | sed "s/^/frames$fr /"; done
