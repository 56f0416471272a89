#!/bin/bash
# Round 2, final state on one GPU: smoke, the whole GPU suite, the default bench line and the reference arm, configs 1-3,
# ncu launch list of the bench command.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2s
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1; tail -2 $O/smoke.log
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; tail -3 $O/pytest_gpu.log
timeout 900 python bench.py > $O/bench_default.json 2>$O/bench_default.err; cut -c1-250 $O/bench_default.json
timeout 900 python bench.py --impl reference > $O/bench_default_reference.json 2>$O/bench_default_reference.err; cut -c1-250 $O/bench_default_reference.json
for w in config1 config2 config3; do
  timeout 900 python bench.py --workload $w > $O/bench_$w.json 2>$O/bench_$w.err; cut -c1-200 $O/bench_$w.json
done
timeout 900 python bench.py --workload config1 --impl reference > $O/bench_config1_reference.json 2>$O/bench_config1_reference.err; cut -c1-250 $O/bench_config1_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_shard.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
ls -la $O
