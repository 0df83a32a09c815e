#!/bin/bash
# Round-2 GPU call 9 (4 GPUs): final multi-GPU tests and strong scaling 1 / 2 / 4 on the final code.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c9
mkdir -p $O
( timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_multi_tests.log 2>&1
run() {  # name, ngpu, bench args
  local name=$1 n=$2; shift 2
  if [ $n = 1 ]; then timeout 400 python bench.py --gpus 1 --no-cpu "$@" > $O/$name.json 2> $O/$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --no-cpu "$@" > $O/$name.json 2> $O/$name.err; fi
}
run scale_config2_1 1
run scale_config2_2 2
run scale_config2_4 4
run scale_config2-nnr_2 2 --workload config2-nnr
run scale_config2-nnr_4 4 --workload config2-nnr
run scale_config2-nn_2 2 --workload config2-nn
run scale_config2-nn_4 4 --workload config2-nn
echo done
