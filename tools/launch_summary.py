"""Summarises an ncu launch list (tools/capture_profiles.sh: TAG_launches.csv) per kernel: total, share of the step's own
kernels, launches, average.   python tools/launch_summary.py profiles/r02_launches.csv > profiles/r02_launch_summary.txt
Launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's phase_ms, not absolutes."""
import csv
import sys
from collections import defaultdict

NOT_STEP = ("fp32_peak_kernel", "region_tsum_kernel")   # roofline microbenchmark; one-time work of smalfit_set_targets


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    tot, cnt = defaultdict(float), defaultdict(int)
    for r in rows:
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        ns = float(r["Metric Value"]) * {"ns": 1, "us": 1e3, "ms": 1e6}.get(r["Metric Unit"], 1)
        tot[r["Kernel Name"]] += ns
        cnt[r["Kernel Name"]] += 1
    own = {k: v for k, v in tot.items() if k.startswith("smf::") and not any(s in k for s in NOT_STEP)}
    step = sum(own.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400, python bench.py --steps 2 --warmup 1 "
          "--no-cpu-baseline --no-quality --no-dropin (N=128, S=256)")
    print("# cold-cache serialised launches: compare SHARES, not absolutes.  share = of the step's own kernels (smf:: without the")
    print("# fp32_peak roofline microbenchmark and the one-time region_tsum of smalfit_set_targets)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        share = "%5.1f%%" % (100 * v / step) if k in own else "     -"
        print("%9.3f ms total  %s of step  n=%4d  avg %9.1f us  %s" % (v / 1e6, share, cnt[k], v / cnt[k] / 1e3, k[:110]))


if __name__ == "__main__":
    main(sys.argv[1])
