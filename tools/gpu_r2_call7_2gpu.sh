#!/bin/bash
# Round-2 GPU call 7 (2 GPUs): per-iteration trace of sharded NNR, scaling after the tail cuts (no CSC build in a settled loop,
# thread-per-bidder rounds, thread-per-pair FD lookup).
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c7
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
ALL_RANKS=1 timeout 300 $TR tools/iter_diag.py config2-nnr 14 > $O/diag_nnr_2.log 2>&1
timeout 300 python tools/iter_diag.py config2-nnr 10 > $O/diag_nnr_1.log 2>&1
timeout 300 python tools/iter_diag.py config2 10 > $O/diag_km_1.log 2>&1
timeout 300 $TR tools/iter_diag.py config2 10 > $O/diag_km_2.log 2>&1
timeout 600 python bench.py --gpus 1 --no-cpu > $O/scale_config2_1.json 2> $O/scale_config2_1.err
timeout 600 $TR bench.py --gpus 2 --no-cpu > $O/scale_config2_2.json 2> $O/scale_config2_2.err
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_km_freerun.py tests/test_gpu_multi.py -q -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_tests.log 2>&1
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 4000 --csv --log-file $O/launches_config2.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/ncu_l2.log 2>&1
echo done
