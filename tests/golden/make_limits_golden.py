#!/usr/bin/env python
"""Golden vectors for the joint-limit term from the UNMODIFIED reference, imported in place.

Run in the build container only (needs /root/reference):

    python tests/golden/make_limits_golden.py

Loads ``LimitPrior`` from smal_fitter/priors/joint_limits_prior.py (pure numpy), stores its min / max
tables and its hinge ``LimitPrior.__call__(x, np)`` on a seeded batch of joint rotations in
``tests/golden/joint_limits_golden.npz``.  The reference evaluates 32 parts x 3 axes; the 35-joint model
has 34 pose joints (ears unbounded), so the batch covers the first 32 rows.
"""
import importlib.util
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("SMALIFY_REF", "/root/reference")


def main():
    spec = importlib.util.spec_from_file_location("joint_limits_prior", os.path.join(REF, "smal_fitter", "priors", "joint_limits_prior.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lp = mod.LimitPrior()
    rng = np.random.default_rng(0)
    x = rng.normal(scale=0.9, size=(5, 96))            # 5 frames of 32 x 3 rotations, many past the limits
    hinge = np.stack([lp(xi, np) for xi in x])
    np.savez(os.path.join(HERE, "joint_limits_golden.npz"), min_values=lp.min_values, max_values=lp.max_values, x=x, hinge=hinge)
    print("wrote joint_limits_golden.npz", hinge.shape, float(hinge.mean()))


if __name__ == "__main__":
    main()
