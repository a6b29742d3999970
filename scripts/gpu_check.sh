#!/bin/bash
# usual GPU-box sequence: smoke, GPU tests (full log kept in gpurun_out/pytest_gpu.log)
mkdir -p gpurun_out
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_gpu.log | tail -15
