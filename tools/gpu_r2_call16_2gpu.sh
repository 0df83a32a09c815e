#!/bin/bash
# Round-2 GPU call 16 (2 GPUs): the multi-GPU parity tests on the final commit + one sharded bench line.
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c16
mkdir -p $O
( timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider; echo "rc=$?" ) > $O/gpu_multi_tests.log 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --no-cpu > $O/scale_config2_2.json 2> $O/scale_config2_2.err
echo done
