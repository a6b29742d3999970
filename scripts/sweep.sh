#!/bin/bash
# C5 sweep points on one GPU (bench.py with explicit sizes); one JSON line per point into gpurun_out/sweep.jsonl
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for cfg in "2048 8192" "4096 8192" "8192 4096" "8192 32768" "16384 8192"; do
  set -- $cfg
  python bench.py --train-points $1 --particles-per-gpu $2 --steps 2 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-real-shapes 2>/dev/null >> gpurun_out/sweep.jsonl
done
python - <<'PY'
import json
for l in open("gpurun_out/sweep.jsonl"):
    d = json.loads(l); c = d["config"]
    v = d.get("variants", {})
    print("N=%5d M=%6d  %.3e particle-steps/s  step %.0f ms  step_frac %.3f  gemm %.2f TF (%.3f)  share %.3f  precompute %.0f ms | ozaki8 %s ozaki7 %s" % (
        c["train_points"], c["particles_per_gpu"], d["value"], d["ms_per_step"], c["step_frac_of_fp64_peak"], d["roofline"]["achieved"],
        d["roofline"]["frac"], d["roofline"]["kernel_share_of_step"], c["precompute_ms"],
        "%.3e" % v["ozaki8"]["value"] if "ozaki8" in v else "-", "%.3e" % v["ozaki7"]["value"] if "ozaki7" in v else "-"))
PY
