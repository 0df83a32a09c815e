#!/bin/bash
# Round-2 GPU call 15: whole -m gpu suite + smoke() on the final commit; ncu --set full of the two-column k_ff_sweep (config 3).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c15
mkdir -p $O
( timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
( timeout 300 python -c "import __graft_entry__ as e; e.smoke()"; echo "rc=$?" ) >> $O/gpu_tests.log 2>&1
timeout 600 ncu --clock-control none --set full --import-source on --kernel-name-base demangled -k 'regex:k_ff_sweep<.bool.0, .bool.1' --launch-skip 2 -c 1 -f -o /tmp/ff2 python bench.py --workload config3 --steps 1 --warmup 3 --no-cpu > $O/ncu_k_ff_sweep.log 2>&1
if [ -f /tmp/ff2.ncu-rep ]; then ncu -i /tmp/ff2.ncu-rep --page raw --csv > $O/k_ff_sweep_two_columns.raw.csv; ncu -i /tmp/ff2.ncu-rep --page details --csv > $O/k_ff_sweep_two_columns.details.csv; fi
echo done
