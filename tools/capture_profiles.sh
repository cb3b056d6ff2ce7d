#!/bin/bash
# GPU box: the ncu artefacts of one build (launch list, full captures of the two rasteriser kernels, DRAM traffic).
#   bash tools/capture_profiles.sh TAG          -> gpurun_out/TAG_*
# Timings printed by a run under ncu are never bench values (serialised, cold caches): they give the kernels' SHARES.
TAG=${1:-r02}
B="--steps 2 --warmup 1 --no-cpu-baseline --no-quality --no-dropin"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py $B > gpurun_out/${TAG}_launches_bench.log 2>&1
for k in raster_tile_forward_kernel raster_backward_kernel; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k -s 6 -c 1 -o gpurun_out/${TAG}_$k -f python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-quality --no-dropin > /dev/null 2> gpurun_out/${TAG}_$k.err
done
python - "$TAG" <<'PY'
import csv, json, subprocess, sys
sys.path.insert(0, ".")
from smalify_b200 import build as B
tag = sys.argv[1]
out = subprocess.run(["ncu", "-i", f"gpurun_out/{tag}_raster_tile_forward_kernel.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, vals, units = rows[0], rows[2], rows[1]
def get(name):
    i = hdr.index(name); v = float(vals[i]); u = units[i]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
sha = B.source_hash()
json.dump({"kernel": "raster_tile_forward_kernel", "frames_per_gpu": 128, "image_size": 256, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_launch": rd + wr, "src_sha16": sha, "capture": f"ncu --set full, {tag}"}, open(f"gpurun_out/{tag}_raster_forward_traffic.json", "w"), indent=1)
print("traffic", rd + wr, sha)
PY
