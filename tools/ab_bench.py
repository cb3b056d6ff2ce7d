#!/usr/bin/env python
"""A/B timing of compile-time variants of libsmalfit.so on the benchmark workload.

    python tools/ab_bench.py build  NAME:-DFLAG1,-DFLAG2 NAME2: ...     (here, no GPU needed)
    python tools/ab_bench.py run [--steps K]                            (on the GPU box)

`build` writes build/variants/NAME.so; `run` benches every variant found there through bench.py
(SMALFIT_LIB) and prints the per-phase device times side by side.
"""
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
VDIR = os.path.join(ROOT, "build", "variants")


def main():
    if sys.argv[1] == "build":
        from smalify_b200 import build as B
        os.makedirs(VDIR, exist_ok=True)
        for old in glob.glob(os.path.join(VDIR, "*.so")):
            os.remove(old)
        for spec in sys.argv[2:]:
            name, _, flags = spec.partition(":")
            defines = [f[2:] for f in flags.split(",") if f.startswith("-D")]
            print(B.build(force=True, defines=defines, out=os.path.join(VDIR, name + ".so")), defines)
    else:
        steps = sys.argv[sys.argv.index("--steps") + 1] if "--steps" in sys.argv else "20"
        extra = sys.argv[sys.argv.index("--") + 1:] if "--" in sys.argv else []
        for lib in sorted(glob.glob(os.path.join(VDIR, "*.so"))):
            env = dict(os.environ, SMALFIT_LIB=lib)
            res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu-baseline", "--no-quality", "--no-dropin", "--steps", steps] + extra,
                                 env=env, capture_output=True, text=True)
            try:
                d = json.loads(res.stdout.strip().splitlines()[-1])
                ph = d["roofline"]["phase_ms"]
                print(f"{os.path.basename(lib):28s} it/s {d['value']:7.1f}  fwd {ph['raster_forward']:.4f}  bwd {ph['raster_backward']:.4f}  "
                      f"front {ph['pose_forward']:.4f}  fback {ph['frame_backward']:.4f}  shape {ph['shape_backward']:.4f}  total {ph['total']:.4f}  "
                      f"e2e {d['e2e']['value']:.1f} (serial {d['e2e'].get('serial_value', 0):.1f})  loss {d['final_loss']}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(os.path.basename(lib), "FAILED", e, res.stderr[-400:], flush=True)


if __name__ == "__main__":
    main()
