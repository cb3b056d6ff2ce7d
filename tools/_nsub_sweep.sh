for fr in 128 64 16; do for fair in 2 3 4 6; do
  echo -n "frames $fr fair $fair: "; SMALFIT_RT_FAIR=$fair python bench.py --no-cpu-baseline --steps 20 --frames $fr 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['roofline']['phase_ms']; print(round(d['value'],1), round(p['raster_forward'],4), round(p['face_rects'],4), d['final_loss'])"
done; done
for S in 512 1024; do for fair in 2 3 6; do
  echo -n "S $S frames 32 fair $fair: "; SMALFIT_RT_FAIR=$fair python bench.py --no-cpu-baseline --steps 10 --frames 32 --size $S 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); p=d['roofline']['phase_ms']; print(round(d['value'],1), round(p['raster_forward'],4), d['final_loss'])"
done; done
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
