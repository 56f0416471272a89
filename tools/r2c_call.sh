#!/bin/bash
# Round 2, third GPU call: the fused tile kernel with records in shared memory (F/R) -- parity suite, A/B, ncu.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2c
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rf -x > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
for w in shard config3 config2; do
  timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline --no-e2e --no-parity > $O/bench_${w}_im1.json 2> $O/bench_${w}_im1.err
done
timeout 300 python bench.py --workload shard --steps 20 --no-cpu-baseline --no-e2e --no-parity --interact-mode 0 > $O/bench_shard_im0.json 2> $O/bench_shard_im0.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2c/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()}, "rho %.2f" % d["rho"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_shard.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interact_tile|interact_cross" -c 8 -o $O/tile_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
timeout 600 python tools/config2_full.py --out $O/config2_full.jsonl > $O/config2_full.log 2>&1; tail -3 $O/config2_full.log
ls -la $O
