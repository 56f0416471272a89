#!/bin/bash
# Round 2, two GPUs: tiles-over-ranks + strips tests, and what the strip record costs the end-to-end loop at N = 2
# (record copied inside the step by the library | record issued after the step from Python | no record).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2q
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_strips.py -x -q -m gpu > $O/pytest_2gpu.log 2>&1; tail -3 $O/pytest_2gpu.log
for v in instep norecord; do
  vv=$v; [ $v = instep ] && vv=""
  LM_RECORD_TIMING=1 LM_E2E_VARIANT=$vv timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 2 --steps 40 --warmup 5 --no-parity > $O/bench_n2_$v.json 2>$O/bench_n2_$v.err
  python -c "
import json,sys
d=json.loads(open('$O/bench_n2_$v.json').read().strip().splitlines()[-1]); print('$v', d['ms_per_step'], d['e2e'])"
done
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_shard_n1.json 2>$O/bench_n1.err; python -c "
import json
d=json.loads(open('$O/bench_shard_n1.json').read().strip().splitlines()[-1]); print(1, d['ms_per_step'], d['e2e'])"
ls -la $O
