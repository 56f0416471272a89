#!/bin/bash
# First GPU call of round 2: everything written at the end of round 1 after the GPU budget ran out, in ONE gpurun call
# (a box takes minutes to get; results come back under gpurun_out/).
#
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
#
# 1. the verified suite (must stay green), then the never-run kernels verbosely (xfail-allowed: XPASS = green)
# 2. A/B of the tiled resolver on fresh and stirred states, bench lines with both resolvers
# 3. timing of the analysis kernels against the reference's own benchmark size (10^4 points: 4.0 s on one Julia process)
# 4. launch list + one ncu --set full capture of the tiled resolver
# 5. pinned-memory PCIe bandwidth of the box (what bounds the end-to-end number)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out/r2a
mkdir -p $O
python -c "import __graft_entry__ as g; g.build(); g.smoke()" > $O/smoke.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_zz_analysis.py --deselect tests/test_gpu_zz_resolver_tiled.py --deselect tests/test_gpu_zz_run_to_file.py --deselect tests/test_gpu_zzz_record_delta.py > $O/pytest_verified.log 2>&1
timeout 600 python -m pytest tests/test_gpu_zz_resolver_tiled.py -m gpu -rxX -v --runxfail > $O/pytest_tiled.log 2>&1
timeout 600 python -m pytest tests/test_gpu_zz_analysis.py -m gpu -rxX -v --runxfail > $O/pytest_analysis.log 2>&1
timeout 600 python -m pytest tests/test_gpu_zz_run_to_file.py -m gpu -rxX -v --runxfail > $O/pytest_run_to_file.log 2>&1
timeout 300 python -m pytest tests/test_gpu_zzz_record_delta.py -m gpu -rxX -v --runxfail > $O/pytest_record_delta.log 2>&1
tail -3 $O/pytest_verified.log $O/pytest_tiled.log $O/pytest_analysis.log $O/pytest_run_to_file.log $O/pytest_record_delta.log
# pinned-memory PCIe bandwidth of the box (what bounds the end-to-end number)
timeout 120 python tools/pcie_bandwidth.py > $O/pcie_bandwidth.txt 2>&1
# BASELINE config 2 to its full length (7,670 hourly steps), nine-phase resolver: where does the step time go?
LM_RESOLVE_MODE=0 timeout 300 python tools/long_run_probe.py config2 7670 590 > $O/config2_full_mode0.jsonl 2> $O/config2_full_mode0.err
if grep -q "failed\|error" $O/pytest_tiled.log; then echo "tiled resolver NOT green: skipping its measurements"; else
  LM_RESOLVE_MODE=1 timeout 300 python tools/long_run_probe.py config2 7670 590 > $O/config2_full_mode1.jsonl 2> $O/config2_full_mode1.err
  timeout 900 python tools/tiled_sweep.py config2:1500:200 shard:0:20 shard:1000:20 config3:0:10 config3:400:10 > $O/tiled_sweep.jsonl 2> $O/tiled_sweep.err
  for w in shard config2 config3; do
    timeout 600 python bench.py --workload $w --no-cpu-baseline --resolve-mode 1 > $O/bench_${w}_tiled.json 2> $O/bench_${w}_tiled.err
    timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/bench_${w}_phases.json 2> $O/bench_${w}_phases.err
  done
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_shard_tiled.csv \
      python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --resolve-mode 1 > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:resolve_tiled -c 3 -o $O/tiled_full \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --resolve-mode 1 > /dev/null 2>&1
fi
if grep -q "failed\|error" $O/pytest_analysis.log; then echo "analysis kernels NOT green: skipping their timing"; else
  timeout 300 python tools/analysis_probe.py > $O/analysis_probe.jsonl 2> $O/analysis_probe.err
fi
if grep -q "failed\|error" $O/pytest_record_delta.log; then echo "delta record NOT green: skipping its probe"; else
  timeout 120 python tools/record_pack_probe.py > $O/record_pack_probe.json 2> $O/record_pack_probe.err
  timeout 600 python bench.py --no-cpu-baseline --packed-record > $O/bench_shard_packed_record.json 2> $O/bench_shard_packed_record.err
fi
ls -la $O
