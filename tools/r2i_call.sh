#!/bin/bash
# Round 2: where does the end-to-end gap come from (tools/e2e_probe.py), the default bench line as the driver runs it, both arms of
# BASELINE config 1 (the same-size GPU / CPU comparison) and config 2.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2i
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 300 python tools/e2e_probe.py > $O/e2e_probe.txt 2>&1; cat $O/e2e_probe.txt
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err; tail -2 $O/bench_default.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $O/bench_default_reference.json 2> $O/bench_default_reference.err
timeout 600 python bench.py --workload config1 > $O/bench_config1.json 2> $O/bench_config1.err; tail -2 $O/bench_config1.err
timeout 600 python bench.py --workload config1 --impl reference > $O/bench_config1_reference.json 2> $O/bench_config1_reference.err
timeout 600 python bench.py --workload config2 --no-cpu-baseline > $O/bench_config2.json 2> $O/bench_config2.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i/bench_*.json")):
    try:
        d = json.load(open(f)); print(f.split("/")[-1], "value %.3e" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.3e" % d.get("e2e", {}).get("value", 0), "parity", (d.get("parity") or {}).get("match"), d.get("phases_ms"), (d.get("cpu_baseline") or {}).get("value"), (d.get("roofline") or {}).get("frac"))
    except Exception as e:
        print(f, "FAILED", e)
PY
