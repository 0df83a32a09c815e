#!/bin/bash
# Round-2 GPU call 14: k_ff_sweep with two target columns per thread (shipped) against one (variant library): FPFH tests, config 3.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c14
mkdir -p $O
( timeout 600 python -m pytest tests/test_zz_extensions.py tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider -k "fpfh or FPFH"; echo "rc=$?" ) > $O/gpu_tests_fpfh.log 2>&1
( timeout 300 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) >> $O/gpu_tests_fpfh.log 2>&1
for lib in default cpt1; do
  if [ $lib = default ]; then unset GHICP_B200_LIB; else export GHICP_B200_LIB=$PWD/gh-icp_b200/variants/lib_ff_cpt1.so; fi
  timeout 300 python bench.py --workload config3 --no-cpu --steps 8 --warmup 4 > $O/bench_config3_$lib.json 2> $O/bench_config3_$lib.err
done
echo done
