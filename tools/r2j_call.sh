#!/bin/bash
# Round 2: heavy units forked beside the light units of each phase -- config 2 / config 3 / shard, hybrid vs round-1 pipeline.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2k
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strips.py tests/test_gpu_config1.py -m gpu -q -x > $O/pytest.log 2>&1; tail -2 $O/pytest.log
for im in 2; do
  timeout 300 python bench.py --workload config2 --no-cpu-baseline --no-e2e --no-parity --interact-mode $im > $O/bench_config2_im${im}.json 2> $O/err.txt
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config2 7670 590 > $O/config2_probe_im${im}.jsonl 2>> $O/err.txt
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config3 400 100 > $O/config3_probe_im${im}.jsonl 2>> $O/err.txt
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py shard 1000 250 > $O/shard_probe_im${im}.jsonl 2>> $O/err.txt
done
timeout 600 python tools/config2_full.py --out $O/config2_full.jsonl > $O/config2_full.log 2>&1; tail -1 $O/config2_full.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2k/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()})
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("gpurun_out/r2k/*probe*.jsonl")):
    print(f.split("/")[-1])
    for l in open(f):
        d = json.loads(l)
        if "step" in d: print("   step", d["step"], "rho %.2f" % d["rho"], d["phases_ms"], "wall", d.get("wall_s"))
PY
