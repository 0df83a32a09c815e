#!/bin/bash
# Round-2 last GPU run (1 GPU): the whole -m gpu suite on the final code, bench lines of config 2 / 4 / 5.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final2
mkdir -p $O
( timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) > $O/smoke.log 2>&1
timeout 900 python bench.py > $O/bench_config2.json 2> $O/bench_config2.err
for w in config4 config5; do
  timeout 600 python bench.py --workload $w --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
done
echo done
