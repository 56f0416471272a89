#!/bin/bash
# Round 2, GPU call 4: fused tile kernel v3 (compact record walk) vs the round-1 pipeline, fresh and stirred states.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2d
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config1.py tests/test_gpu_strips.py tests/test_gpu_streamed_field.py tests/test_gpu_advect_fast.py -m gpu -q -rf -x > $O/pytest_gpu.log 2>&1; tail -5 $O/pytest_gpu.log
for im in 1 0; do
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py shard 1000 250 > $O/shard_probe_im${im}.jsonl 2> $O/shard_probe_im${im}.err
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config3 400 100 > $O/config3_probe_im${im}.jsonl 2> $O/config3_probe_im${im}.err
  LM_INTERACT_MODE=$im LM_ADVECT_MODE=1 timeout 300 python tools/long_run_probe.py config2 7670 590 > $O/config2_probe_im${im}.jsonl 2> $O/config2_probe_im${im}.err
done
for w in shard config3 config2; do
  timeout 300 python bench.py --workload $w --steps 20 --no-cpu-baseline --no-e2e --no-parity > $O/bench_${w}_im1.json 2> $O/bench_${w}_im1.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "ms %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["phases_ms"].items()}, "rho %.2f" % d["rho"])
    except Exception as e:
        print(f, "FAILED", e)
for f in sorted(glob.glob("gpurun_out/r2d/*probe*.jsonl")):
    print(f.split("/")[-1])
    for l in open(f):
        d = json.loads(l)
        if "step" in d: print("   step", d["step"], "rho %.2f" % d["rho"], d["phases_ms"])
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"interact_tile" -c 2 -o $O/tile_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
ls -la $O | head -40
