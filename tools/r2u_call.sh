#!/bin/bash
# Round 2, the last GPU seconds: smoke() and the record-window test on the final sources.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2u
mkdir -p $O
timeout 40 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 45 python -m pytest tests/test_gpu_zz_run_to_file.py -x -q -m gpu -p no:cacheprovider > $O/pytest_run_to_file.log 2>&1; tail -3 $O/pytest_run_to_file.log
