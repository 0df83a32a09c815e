#!/bin/bash
# Round-2 GPU call 11: ncu --set full of the reverse auction phase AFTER this round's changes (iteration 0 of config 2), launch
# list of the config-4 pipeline.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c11
mkdir -p $O
NCU="ncu --clock-control none"
# forward phases of iteration 0: the schedule has 8 phases at f = 0.1 -> the reverse launch is the 9th persistent-auction launch
timeout 600 $NCU --set full --import-source on --kernel-name-base demangled -k 'regex:k_auction_persistent<.bool.1' -c 1 -f -o /tmp/aucrev python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_auction_rev_after.log 2>&1
if [ -f /tmp/aucrev.ncu-rep ]; then ncu -i /tmp/aucrev.ncu-rep --page raw --csv > $O/k_auction_reverse_after.raw.csv; ncu -i /tmp/aucrev.ncu-rep --page details --csv > $O/k_auction_reverse_after.details.csv; fi
timeout 600 $NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file $O/launches_config4.csv python bench.py --workload config4 --steps 1 --warmup 3 --no-cpu > $O/ncu_l4.log 2>&1
echo done
