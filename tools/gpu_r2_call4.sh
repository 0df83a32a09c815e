#!/bin/bash
# Round-2 GPU call 4: K-chunked tcgen05 FD build (672-bit) on hardware, k_stream tuning variants A/B, auction debug on config4.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c4
mkdir -p $O
( timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
timeout 600 python bench.py --workload config2-672 --no-cpu > $O/bench_config2-672.json 2> $O/bench_config2-672.err
GHICP_FD_POPC=1 timeout 600 python bench.py --workload config2-672 --no-cpu > $O/bench_config2-672_popc.json 2> $O/bench_config2-672_popc.err
for lib in default u2_s8_b3 u2_s6_b3 u2_s12_b3 u4_s6_b3 u4_s8_b2; do
  for w in config2 config2-nn config2-nnr; do
    if [ $lib = default ]; then unset GHICP_B200_LIB; else export GHICP_B200_LIB=$PWD/gh-icp_b200/variants/lib_$lib.so; fi
    timeout 300 python bench.py --workload $w --no-cpu --steps 20 --warmup 5 > $O/var_${lib}_$w.json 2> $O/var_${lib}_$w.err
  done
done
unset GHICP_B200_LIB
GHICP_AUCTION_DEBUG=1 timeout 600 python bench.py --workload config4 --steps 1 --warmup 3 --no-cpu > $O/auction_debug_config4.json 2> $O/auction_debug_config4.log
NCU="ncu --clock-control none"
timeout 900 $NCU --set full --import-source on -k regex:k_fd_bsc_tc_kc -c 1 -f -o /tmp/fdkc python bench.py --workload config2-672 --steps 1 --warmup 3 --no-cpu > $O/ncu_fdkc.log 2>&1
if [ -f /tmp/fdkc.ncu-rep ]; then ncu -i /tmp/fdkc.ncu-rep --page raw --csv > $O/k_fd_bsc_tc_kc.raw.csv; ncu -i /tmp/fdkc.ncu-rep --page details --csv > $O/k_fd_bsc_tc_kc.details.csv; cp /tmp/fdkc.ncu-rep $O/k_fd_bsc_tc_kc.ncu-rep; fi
echo done
