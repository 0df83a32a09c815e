#!/bin/bash
# Round-2 GPU call 1: the whole -m gpu suite (no -x), every bench workload, launch lists and ncu --set full captures of the
# kernels VERDICT r01 asked evidence for.  Run through gpurun from the repo root; everything lands in gpurun_out/c1/.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c1
mkdir -p $O
nvidia-smi -L > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt
( timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
timeout 900 python bench.py > $O/bench_config2.json 2> $O/bench_config2.err
for w in config2-672 config2-nnr config2-nn config3 config1; do
  timeout 500 python bench.py --workload $w --no-cpu > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 900 python bench.py --workload config4 --no-cpu > $O/bench_config4.json 2> $O/bench_config4.err

NCU="ncu --clock-control none"
# launch lists (cold-cache, serialised: compare shares)
timeout 600 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l2.log 2>&1
timeout 900 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $O/launches_config4.csv python bench.py --workload config4 --steps 2 --warmup 3 --no-cpu > $O/ncu_l4.log 2>&1

full() {  # name, kernel regex, launch-skip, count, command...
  local name=$1 k=$2 skip=$3 cnt=$4; shift 4
  timeout 900 $NCU --set full --import-source on -k "regex:$k" --launch-skip $skip -c $cnt -f -o /tmp/$name "$@" > $O/ncu_$name.log 2>&1
  if [ -f /tmp/$name.ncu-rep ]; then
    ncu -i /tmp/$name.ncu-rep --page raw --csv > $O/$name.raw.csv 2>/dev/null
    ncu -i /tmp/$name.ncu-rep --page details --csv > $O/$name.details.csv 2>/dev/null
    sz=$(stat -c %s /tmp/$name.ncu-rep)
    if [ "$sz" -lt 9000000 ]; then cp /tmp/$name.ncu-rep $O/; fi
  fi
}
# steady-state k_stream of the KM loop: 2 registration passes (6 launches each) + warm-up iterations 0-2 (5) come first
full k_stream_km k_stream 18 1 python bench.py --steps 2 --warmup 3 --no-cpu
full k_stream_nnr k_stream 20 1 python bench.py --workload config2-nnr --steps 2 --warmup 3 --no-cpu
# the reverse phase of iteration 0 = the 8th persistent-auction launch (7 forward phases first)
full k_auction_rev k_auction_persistent 7 1 python bench.py --steps 1 --warmup 3 --no-cpu
full k_ff_sweep k_ff_sweep 6 1 python bench.py --workload config3 --steps 1 --warmup 3 --no-cpu
full prep 'k_bsc|k_pca|k_nms_round' 0 8 python tools/bench_prep.py --points 1000000 --reps 1 --no-cpu
ls -la $O /tmp/*.ncu-rep > $O/listing.txt 2>&1
echo done
