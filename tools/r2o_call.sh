#!/bin/bash
# Round 2: profiles of the shipped default path (ncu launch list of the bench command + one ncu --set full capture of a whole step),
# the BASELINE config 5 sweep to 50 M microbes and its bench line.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2o
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_shard.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
# one whole step under --set full (launch-skip: set-up, spin-up and warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on --launch-skip 260 -c 40 -o $O/step_full \
    python bench.py --steps 4 --warmup 8 --no-cpu-baseline --no-e2e --no-parity > /dev/null 2>&1
ncu -i $O/step_full.ncu-rep --page raw --csv > $O/step_full_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/step_full_raw.csv shard 12500000 > $O/ncu_summary_shard.json 2>$O/ncu_summary.err; head -c 300 $O/ncu_summary_shard.json
ncu -i $O/step_full.ncu-rep --page source --csv --kernel-name regex:find_pairs --launch-skip 0 --launch-count 1 > $O/find_pairs_sass.csv 2>/dev/null
rm -f $O/step_full.ncu-rep          # gpurun brings back at most 64 MiB: keep the exports, not the report
timeout 900 python tools/sweep_interact.py --max-n 50000000 > $O/sweep_config5.jsonl 2> $O/sweep_config5.err; cut -c1-200 $O/sweep_config5.jsonl; tail -2 $O/sweep_config5.err
timeout 600 python bench.py --workload config5 > $O/bench_config5.json 2> $O/bench_config5.err; cut -c1-400 $O/bench_config5.json; tail -2 $O/bench_config5.err
timeout 900 python tools/scatter_probe.py > $O/scatter_probe.jsonl 2> $O/scatter_probe.err; cat $O/scatter_probe.jsonl; tail -3 $O/scatter_probe.err
ls -la $O
