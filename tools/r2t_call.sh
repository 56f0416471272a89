#!/bin/bash
# Round 2, last GPU seconds: the whole GPU suite on the final sources (one GPU).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2t
mkdir -p $O
timeout 255 python -m pytest tests -x -q -m gpu -p no:cacheprovider --durations=8 > $O/pytest_gpu.log 2>&1
tail -15 $O/pytest_gpu.log
