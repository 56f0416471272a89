#!/bin/bash
# Round 2, GPU call 5: the hybrid path (LM_OPT_INTERACT_MODE = 2: round-1 pipeline for the light units + device-wide queue of
# heavy units resolved in rounds of matchings) -- whole GPU suite, A/B against the round-1 pipeline alone, fresh and stirred.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2f
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rf > $O/pytest_gpu.log 2>&1; tail -8 $O/pytest_gpu.log
for im in 2; do
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py shard 1000 250 > $O/shard_probe_im${im}.jsonl 2> $O/shard_probe_im${im}.err
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config3 400 100 > $O/config3_probe_im${im}.jsonl 2> $O/config3_probe_im${im}.err
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config2 7670 590 > $O/config2_probe_im${im}.jsonl 2> $O/config2_probe_im${im}.err
done
for w in shard config3 config2; do
  for im in 2; do
    timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline --no-e2e --no-parity --interact-mode $im > $O/bench_${w}_im${im}.json 2> $O/bench_${w}_im${im}.err
  done
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2f/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()}, "rho %.2f" % d["rho"])
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("gpurun_out/r2f/*probe*.jsonl")):
    print(f.split("/")[-1])
    for l in open(f):
        d = json.loads(l)
        if "step" in d: print("   step", d["step"], "rho %.2f" % d["rho"], d["phases_ms"], "wall", d.get("wall_s"))
PY
timeout 600 python tools/config2_full.py --out $O/config2_full.jsonl > $O/config2_full.log 2>&1; tail -2 $O/config2_full.log
ls $O | head -50
