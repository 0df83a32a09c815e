#!/bin/bash
# Round-2 GPU call 10: the eps split f (eps_last = f*KM_eps, D budget = (1-f)*n*KM_eps) on the small dense instances of the
# pipeline workloads.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c10
mkdir -p $O
for f in 0.1 0.5 0.75 0.9; do
  for w in config4 config5; do
    GHICP_AUCTION_EPSF=$f timeout 300 python bench.py --workload $w --no-cpu --steps 3 --warmup 3 > $O/bench_${w}_f$f.json 2> $O/bench_${w}_f$f.err
  done
done
echo done
