#!/bin/bash
# Round 2, the very last GPU seconds: the N = 2 bench line on the final sources (parity twin with the larger capacity, record bytes counted from the strip).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2v
mkdir -p $O
timeout 34 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_shard_n2.json 2>$O/bench_n2.err
cut -c1-200 $O/bench_shard_n2.json; tail -3 $O/bench_n2.err
