"""BASELINE config 2, three ways: profiles/r02_config2_comparison.json from the CPU oracle's fits
(tools/config2_oracle_fit.py -> profiles/r02_config2_oracle_fit.json) and the GPU fits of the same inputs
(tools/run_configs.py config2 -> profiles/r02_configs.json).   python tools/config2_comparison.py"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    c2 = json.load(open(os.path.join(ROOT, "profiles", "r02_configs.json")))["config2_full_fit"]
    gpu = c2["gpu"]

    def pick(d, keys=("kp_l2", "iou", "wall_s")):
        return {k: d[k] for k in keys if k in d}

    def diff(a, b):
        return {k: abs(a[k] - b[k]) for k in ("kp_l2", "iou")}
    f32, f64 = c2["oracle_f32"], c2["oracle_f64"]
    out = {
        "config": "BASELINE.json configs[1]: N = 10 frames, 256x256, the full 4-stage schedule (150 + 400 + 600 + 800 = 1950 Adam steps), "
                  "inputs tests/golden/config2_inputs.npz",
        "oracle_f64": pick(f64), "oracle_f32": pick(f32),
        "gpu_fused_graph": pick(gpu["fused_graph"], ("kp_l2", "iou", "wall_s", "iters_per_s", "final_stage_losses")),
        "gpu_dropin": pick(gpu["dropin"], ("kp_l2", "iou", "wall_s", "iters_per_s", "final_stage_losses")),
        "oracle_final_stage_losses": {"f32": f32["final_stage_losses"], "f64": f64["final_stage_losses"]},
        "abs_diff": {
            "gpu_fused_vs_oracle_f64": diff(gpu["fused_graph"], f64), "gpu_fused_vs_oracle_f32": diff(gpu["fused_graph"], f32),
            "gpu_dropin_vs_oracle_f64": diff(gpu["dropin"], f64), "gpu_dropin_vs_oracle_f32": diff(gpu["dropin"], f32),
            "oracle_f32_vs_f64": diff(f32, f64), "gpu_fused_vs_gpu_dropin": diff(gpu["fused_graph"], gpu["dropin"]),
        },
        "target": "north star: final kp-L2 (px) and IoU within 1e-3 of the reference's fit of the same inputs",
    }
    out["within_1e-3_of_both_oracle_precisions"] = all(v < 1e-3 for k in ("gpu_fused_vs_oracle_f64", "gpu_fused_vs_oracle_f32",
                                                                           "gpu_dropin_vs_oracle_f64", "gpu_dropin_vs_oracle_f32")
                                                        for v in out["abs_diff"][k].values())
    path = os.path.join(ROOT, "profiles", "r02_config2_comparison.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out["abs_diff"], indent=1), out["within_1e-3_of_both_oracle_precisions"])


if __name__ == "__main__":
    main()
