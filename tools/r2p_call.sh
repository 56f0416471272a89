#!/bin/bash
# Round 2, two GPUs: N_procs tiles over ranks, strips (NCCL + peer), pinned PCIe bandwidth of one and two GPUs at once,
# what each part of the in-step record costs the end-to-end loop, the bench at N = 1 / 2 with e2e.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2p
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
nvidia-smi topo -m > $O/topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_tiles_over_ranks.py tests/test_gpu_strips.py -x -q -m gpu > $O/pytest_2gpu.log 2>&1; tail -3 $O/pytest_2gpu.log
timeout 300 python tools/pcie_bandwidth.py > $O/pcie_bandwidth.jsonl 2>$O/pcie.err; cat $O/pcie_bandwidth.jsonl
timeout 600 python tools/scatter_probe.py 12500000 decompose > $O/record_parts.jsonl 2>$O/record_parts.err; cat $O/record_parts.jsonl; tail -2 $O/record_parts.err
timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_shard_n1.json 2>$O/bench_n1.err; cut -c1-300 $O/bench_shard_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > $O/bench_shard_n2.json 2>$O/bench_n2.err; cut -c1-300 $O/bench_shard_n2.json; tail -2 $O/bench_n2.err
ls -la $O
