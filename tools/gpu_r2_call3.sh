#!/bin/bash
# Round-2 GPU call 3: settled KM route (sync-free iteration), NNR fixed column thresholds, eps split variants on the
# free-running KM parity tests, config5, NNR ncu capture.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c3
mkdir -p $O
( timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
timeout 600 python bench.py --no-cpu > $O/bench_config2.json 2> $O/bench_config2.err
GHICP_KM_GENERAL=1 timeout 600 python bench.py --no-cpu > $O/bench_config2_general_route.json 2> $O/bench_config2_general_route.err
timeout 600 python bench.py --workload config2-nnr --no-cpu > $O/bench_config2-nnr.json 2> $O/bench_config2-nnr.err
timeout 600 python bench.py --workload config2-nn --no-cpu > $O/bench_config2-nn.json 2> $O/bench_config2-nn.err
for f in 0.1 0.25; do
  ( GHICP_AUCTION_EPSF=$f timeout 600 python -m pytest tests/test_gpu_km_freerun.py tests/test_gpu_parity.py -q -p no:cacheprovider -s -k "km or KM"; echo "rc=$?" ) > $O/km_tests_epsf_$f.log 2>&1
done
GHICP_AUCTION_EPSF=0.1 timeout 600 python bench.py --no-cpu > $O/bench_config2_epsf0.1.json 2> $O/bench_config2_epsf0.1.err
timeout 900 python bench.py --workload config4 --no-cpu > $O/bench_config4.json 2> $O/bench_config4.err
timeout 1200 python bench.py --workload config5 --no-cpu > $O/bench_config5.json 2> $O/bench_config5.err
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
full k_stream_nnr_main 'k_stream<.int.1, .bool.1, .bool.1' 6 1 python bench.py --workload config2-nnr --steps 2 --warmup 3 --no-cpu
timeout 600 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l2.log 2>&1
echo done
