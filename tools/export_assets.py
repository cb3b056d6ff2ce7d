#!/usr/bin/env python
"""Export the SMAL constants of a SMALify checkout into the compact npz asset
the B200 boxes use (they have no SMALify checkout).

    python tools/export_assets.py --smalify /root/reference --family 1

Also cross-checks the loader against the reference's own ``SMAL.__init__`` when
the checkout is importable (it is in the build container).
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from smalify_b200 import model_io  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--smalify", default="/root/reference")
    ap.add_argument("--family", type=int, default=1)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()

    c = model_io.load_from_smalify_data(os.path.join(args.smalify, "data"), args.family)
    out = args.out or model_io.default_asset_path(args.family)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    model_io.save_asset(c, out)
    back = model_io.load_asset(out)
    for k in ("v_template", "shapedirs", "j_regressor", "weights", "faces", "pose_prec", "unity_prec"):
        assert np.array_equal(getattr(c, k), getattr(back, k)), k
    print(f"wrote {out}: {os.path.getsize(out) / 1e6:.2f} MB; "
          f"V={c.v_template.shape[0]} F={c.faces.shape[0]} "
          f"nnz(weights)={int((c.weights != 0).sum())} nnz(Jreg)={int((c.j_regressor != 0).sum())}")


if __name__ == "__main__":
    main()
