#!/bin/bash
# Round 2, 2-GPU call: the peer-memory exchange (CUDA IPC, peer stores + flags) against NCCL send/recv -- strip tests over both,
# the shard at N = 2 and BASELINE config 3 at N = 2 with both transports.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2n
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_gpu_strips.py -m gpu -q -rs -v > $O/pytest_strips_2gpu.log 2>&1; tail -22 $O/pytest_strips_2gpu.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
for t in peer nccl peer nccl; do
  timeout 600 $TR bench.py --gpus 2 --steps 50 --warmup 3 --transport $t --no-e2e > $O/bench_shard_n2_$t.json 2> $O/bench_shard_n2_$t.err; tail -2 $O/bench_shard_n2_$t.err
  python -c "
import json; d=json.load(open('$O/bench_shard_n2_$t.json')); print('$t', 'value %.4e' % d['value'], 'ms %.4f' % d['ms_per_step'], 'parity', (d.get('parity') or {}).get('match'), {k: round(v,3) for k,v in d['phases_ms'].items()})"
done
timeout 600 python bench.py --gpus 1 --steps 50 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_shard_n1.json 2> $O/bench_shard_n1.err
for t in peer nccl; do
  timeout 600 $TR bench.py --workload config3 --gpus 2 --steps 10 --warmup 3 --transport $t --no-e2e > $O/bench_config3_n2_$t.json 2> $O/bench_config3_n2_$t.err; tail -2 $O/bench_config3_n2_$t.err
done
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 > $O/bench_shard_n2_e2e.json 2> $O/bench_shard_n2_e2e.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "value %.4e" % d["value"], "ms %.4f" % d["ms_per_step"], "e2e", d.get("e2e", {}).get("value"), "parity", (d.get("parity") or {}).get("match"), {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
PY
