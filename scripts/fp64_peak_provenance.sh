#!/bin/bash
# FP64 pipe peak with provenance (MEASURED_PEAKS.json has no FP64 entry): runs the DMMA / DFMA issue microbenchmark while sampling clocks,
# writes gpurun_out/fp64_peak.json.  usage (GPU box): bash scripts/fp64_peak_provenance.sh
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o profiles/microbench/fp64_peaks profiles/microbench/fp64_peaks.cu
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv,noheader,nounits -lms 200 > gpurun_out/fp64_peak_clocks.csv &
SMI=$!
./profiles/microbench/fp64_peaks > gpurun_out/fp64_peaks.txt
kill $SMI
python - <<'PY'
import json, re, statistics, subprocess, datetime
txt = open("gpurun_out/fp64_peaks.txt").read()
sus = float(re.search(r"DMMA884 sustained .*?: ([0-9.]+) TFLOP/s", txt).group(1))
burst = max(float(x) for x in re.findall(r"DMMA884 ilp\d+ blocks/SM \d+: [0-9.]+ ms\s+([0-9.]+) TFLOP/s", txt))
dfma = max(float(x) for x in re.findall(r"DFMA\s+ilp\d+ blocks/SM \d+: [0-9.]+ ms\s+([0-9.]+) TFLOP/s", txt))
rows = [l.split(",") for l in open("gpurun_out/fp64_peak_clocks.csv") if l.count(",") >= 6]
sm = [float(r[0]) for r in rows]; pw = [float(r[2]) for r in rows]
load = [s for s, p in zip(sm, pw) if p > 300]
reasons = sorted({n for r in rows for n, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]) if v.strip().lower().startswith("active")})
out = {"fp64_tflops_sustained": sus, "fp64_tflops_burst": burst, "dfma_tflops": dfma,
       "how": "profiles/microbench/fp64_peaks.cu: mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4) issue loop, 8 independent accumulators per warp, 1-8 blocks of 256 threads per SM; "
              "sustained = the same loop for 3 s; 2*8*8*4 flop per warp instruction; 128 flop/clk/SM x 148 SMs x 1.965 GHz = 37.2",
       "command": "bash scripts/fp64_peak_provenance.sh", "gpu_name": subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip(),
       "when": datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%M:%SZ"),
       "clocks_under_load": {"samples": len(sm), "samples_under_load": len(load), "sm_mhz_median": statistics.median(load) if load else None,
                             "sm_max_mhz": max(float(r[1]) for r in rows) if rows else None, "power_w_max": max(pw) if pw else None, "reasons": reasons}}
json.dump(out, open("gpurun_out/fp64_peak.json", "w"), indent=1)
print(json.dumps(out))
PY
