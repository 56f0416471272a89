#!/bin/bash
# Round 2, second GPU call: first hardware run of the fused tile kernel (LM_OPT_INTERACT_MODE = 1) and of the float32 RK4
# (LM_OPT_ADVECT_MODE = 1): the whole GPU suite, A/B bench lines against the round-1 pipeline, BASELINE config 2 to its
# full 7,670 steps with parity checkpoints, launch list + one ncu --set full capture of the tile kernel.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2b
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest_gpu.log 2>&1; tail -15 $O/pytest_gpu.log
for w in shard config3 config2; do
  for im in 1 0; do
    timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline --no-e2e --interact-mode $im --advect-mode 1 > $O/bench_${w}_im${im}.json 2> $O/bench_${w}_im${im}.err
  done
done
timeout 300 python bench.py --workload shard --steps 20 --no-cpu-baseline --no-e2e --interact-mode 1 --advect-mode 0 > $O/bench_shard_im1_am0.json 2> $O/bench_shard_im1_am0.err
for db in 8 12 16 24 28 32; do
  timeout 300 python bench.py --workload shard --steps 20 --no-cpu-baseline --no-e2e --draw-batch $db > $O/bench_shard_db${db}.json 2> $O/bench_shard_db${db}.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()}, "rho %.2f" % d["rho"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 600 python tools/config2_full.py --out $O/config2_full.jsonl > $O/config2_full.log 2>&1; tail -3 $O/config2_full.log
LM_INTERACT_MODE=1 LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config2 7670 590 > $O/config2_probe_im1.jsonl 2> $O/config2_probe_im1.err
LM_INTERACT_MODE=1 LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py shard 1000 250 > $O/shard_probe_im1.jsonl 2> $O/shard_probe_im1.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_shard.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interact_tile|advect_rk4_fast" -c 4 -o $O/tile_full \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
ls -la $O
