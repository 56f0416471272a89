#!/bin/bash
# Round 2, eight GPUs (one call): strips over 4 real ranks (NCCL + peer), pinned PCIe bandwidth of 1 .. 8 GPUs at once,
# the bench at N = 4 and N = 8 (BASELINE config 4: 10^8 microbes) with the in-step strip record.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2r
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_strips.py -x -q -m gpu -k "over_nccl and 4" > $O/pytest_4gpu.log 2>&1; tail -3 $O/pytest_4gpu.log
timeout 300 python tools/pcie_bandwidth.py > $O/pcie_bandwidth.jsonl 2>$O/pcie.err; cat $O/pcie_bandwidth.jsonl
for n in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n --steps 50 --warmup 5 > $O/bench_shard_n$n.json 2>$O/bench_n$n.err
  python -c "
import json
d=json.loads(open('$O/bench_shard_n$n.json').read().strip().splitlines()[-1]); print($n, d['value'], d['ms_per_step'], d['e2e'], d.get('parity'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 \
    bench.py --gpus 8 --steps 50 --warmup 5 --transport nccl --no-e2e --no-parity > $O/bench_shard_n8_nccl.json 2>$O/bench_n8_nccl.err
python -c "
import json
d=json.loads(open('$O/bench_shard_n8_nccl.json').read().strip().splitlines()[-1]); print('8 nccl', d['value'], d['ms_per_step'])"
ls -la $O
