#!/bin/bash
# Round-2 GPU call 13: tail rounds with the active lists and bid slots in shared memory: KM / parity / extension tests, config 2 / 4 / 5.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c13
mkdir -p $O
( timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
for w in config2 config4 config5; do
  timeout 300 python bench.py --workload $w --no-cpu --steps 10 --warmup 4 > $O/bench_$w.json 2> $O/bench_$w.err
done
echo done
