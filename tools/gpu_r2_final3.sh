#!/bin/bash
# Round-2 closing GPU run (1 GPU): whole -m gpu suite + smoke() on the final commit, the GPU arm at 4000 x 4000 (its CPU counterpart,
# measured without extrapolation, is profiles/r02_bench_config2_n4000_reference_arm_measured_in_build_container.json), default bench line.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/final3
mkdir -p $O
( timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) >> $O/gpu_tests.log 2>&1
timeout 300 python bench.py --n 4000 --steps 6 --warmup 4 --no-cpu > $O/bench_config2_n4000.json 2> $O/bench_config2_n4000.err
timeout 900 python bench.py > $O/bench_config2.json 2> $O/bench_config2.err
echo done
