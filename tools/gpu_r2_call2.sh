#!/bin/bash
# Round-2 GPU call 2: A/B of the auction changes (D-budget stop, warps-per-bidder grid rounds), KM / prep tests, config4 after
# the pool allocator, ncu captures missed by call 1 (NNR main pass, k_bsc).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c2
mkdir -p $O
timeout 600 python tools/auction_ab.py 50000 "" "GHICP_AUCTION_NOCUT=1" "GHICP_AUCTION_EPSF=0.25" "GHICP_AUCTION_EPSF=0.1" "GHICP_AUCTION_SMALL=46,GHICP_AUCTION_NOCUT=1" "GHICP_AUCTION_SMALL=46" > $O/auction_ab.log 2>&1
GHICP_AUCTION_DEBUG=1 timeout 300 python tools/auction_ab.py 50000 "" > $O/auction_debug.log 2>&1
( timeout 900 python -m pytest tests/test_gpu_km_freerun.py tests/test_gpu_parity.py tests/test_zz_extensions.py tests/test_zz2_prep_gpu.py tests/test_zz3_bsc_gpu.py tests/test_gpu_dropin.py -q -p no:cacheprovider -s; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
timeout 900 python bench.py --workload config4 --no-cpu > $O/bench_config4.json 2> $O/bench_config4.err
timeout 600 python bench.py --no-cpu > $O/bench_config2.json 2> $O/bench_config2.err
NCU="ncu --clock-control none"
full() {  # name, kernel regex, launch-skip, count, command...
  local name=$1 k=$2 skip=$3 cnt=$4; shift 4
  timeout 900 $NCU --set full --import-source on --kernel-name-base demangled -k "regex:$k" --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > $O/ncu_$name.log 2>&1
  if [ -f /tmp/$name.ncu-rep ]; then
    ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
    ncu -i /tmp/$name.ncu-rep --page details --csv > $O/$name.details.csv 2>/dev/null
    sz=$(stat -c %s /tmp/$name.ncu-rep)
    if [ "$sz" -lt 9000000 ]; then cp /tmp/$name.ncu-rep $O/; fi
  fi
}
full k_stream_nnr_main 'k_stream<1, ' 6 1 python bench.py --workload config2-nnr --steps 2 --warmup 3 --no-cpu
full k_bsc 'k_bsc' 0 1 python tools/bench_prep.py --points 1000000 --reps 1 --no-cpu
echo done
