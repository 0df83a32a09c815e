#!/bin/bash
# Round-2 final GPU run (1 GPU): the whole -m gpu suite, smoke(), every bench workload, the measured (no extrapolation) pair,
# launch list, compute-sanitizer on smoke().  Everything lands in gpurun_out/final/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final
mkdir -p $O
( timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_config2.json 2> $O/bench_config2.err
for w in config2-672 config2-nnr config2-nn config3 config1 config4 config5; do
  timeout 600 python bench.py --workload $w --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 600 python bench.py --n 1500 > $O/bench_config2_n1500_measured_cpu.json 2> $O/bench_config2_n1500.err
timeout 600 python bench.py --workload config3 --n 30000 --no-cpu > $O/bench_config3_n30000.json 2> $O/bench_config3_n30000.err
timeout 300 python tools/fd_time.py 441 672 > $O/fd_time.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l2.log 2>&1
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) > $O/sanitizer_memcheck_smoke.log 2>&1
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) > $O/sanitizer_racecheck_smoke.log 2>&1
echo done
