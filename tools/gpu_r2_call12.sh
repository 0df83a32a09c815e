#!/bin/bash
# Round-2 GPU call 12: auction state through L1 (plain loads; shipped) against L2-only loads (round-1 behaviour, variant library):
# KM tests on the shipped library, then both libraries on config 2 / 4 / 5.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c12
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_km_freerun.py tests/test_gpu_parity.py tests/test_zz_extensions.py tests/test_gpu_dropin.py -q -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
for lib in default l2; do
  if [ $lib = default ]; then unset GHICP_B200_LIB; else export GHICP_B200_LIB=$PWD/gh-icp_b200/variants/lib_auction_l2_loads.so; fi
  for w in config2 config4 config5; do
    timeout 300 python bench.py --workload $w --no-cpu --steps 10 --warmup 4 > $O/bench_${w}_$lib.json 2> $O/bench_${w}_$lib.err
  done
done
echo done
