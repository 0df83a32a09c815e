#!/bin/bash
# Round-2 GPU call 5 (2 GPUs): multi-GPU parity tests, strong scaling 1 -> 2 on config 2 with the settled KM route (one
# all-gather per iteration) against the general route, FD build timing.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c5
mkdir -p $O
nvidia-smi -L > $O/gpu.txt
( timeout 1200 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_multi_tests.log 2>&1
run() {  # name, ngpu, extra env..., then bench args
  local name=$1 n=$2; shift 2
  if [ $n = 1 ]; then timeout 600 python bench.py --gpus 1 --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu "$@" > $O/$name.json 2> $O/$name.err; fi
}
run scale_config2_1 1
run scale_config2_2 2
GHICP_KM_GENERAL=1 run scale_config2_2_general_route 2
run scale_config2-nn_2 2 --workload config2-nn
run scale_config2-nnr_2 2 --workload config2-nnr
run scale_config4_2 2 --workload config4
timeout 300 python tools/fd_time.py 441 672 > $O/fd_time.log 2>&1
GHICP_FD_POPC=1 timeout 300 python tools/fd_time.py 672 >> $O/fd_time.log 2>&1
echo done
