#!/bin/bash
# Round 2, 2-GPU call: strips over NCCL against a single handle (the tests the 1-GPU driver box skips), the shard at N = 2 with
# the parity twin, BASELINE config 3 on 1 and on 2 GPUs (strong scaling).
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2h
mkdir -p $O
nvidia-smi -L > $O/gpus.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_strips.py -m gpu -q -rs -v > $O/pytest_strips_2gpu.log 2>&1; tail -12 $O/pytest_strips_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_shard_n2.json 2> $O/bench_shard_n2.err; tail -3 $O/bench_shard_n2.err
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_shard_n1.json 2> $O/bench_shard_n1.err
timeout 600 python bench.py --workload config3 --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_config3_n1.json 2> $O/bench_config3_n1.err
timeout 600 $TR bench.py --workload config3 --gpus 2 --steps 10 --warmup 3 > $O/bench_config3_n2.json 2> $O/bench_config3_n2.err; tail -3 $O/bench_config3_n2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2h/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "value %.3e" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "parity", (d.get("parity") or {}).get("match"), {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
ls $O
