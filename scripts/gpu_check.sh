#!/bin/bash
# usual GPU-box sequence: smoke, GPU tests
python __graft_entry__.py --smoke 2>&1 | tail -3
python -m pytest tests -m gpu -x -q 2>&1 | tail -25
