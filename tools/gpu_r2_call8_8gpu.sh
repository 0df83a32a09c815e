#!/bin/bash
# Round-2 GPU call 8 (8 GPUs): strong scaling at 8 on config 2 (KM, NN, NNR) and the two pipeline workloads of BASELINE.json
# (configs 4 / 5: "8 x B200 sharded").
cd "${GRAFT_REPO_ROOT:-.}"
O=gpurun_out/c8
mkdir -p $O
nvidia-smi -L > $O/gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR bench.py --gpus 8 --no-cpu > $O/scale_config2_8.json 2> $O/scale_config2_8.err
timeout 400 $TR bench.py --gpus 8 --no-cpu --workload config2-nn > $O/scale_config2-nn_8.json 2> $O/scale_config2-nn_8.err
timeout 400 $TR bench.py --gpus 8 --no-cpu --workload config2-nnr > $O/scale_config2-nnr_8.json 2> $O/scale_config2-nnr_8.err
timeout 500 $TR bench.py --gpus 8 --no-cpu --workload config4 > $O/scale_config4_8.json 2> $O/scale_config4_8.err
timeout 600 $TR bench.py --gpus 8 --no-cpu --workload config5 > $O/scale_config5_8.json 2> $O/scale_config5_8.err
timeout 300 python bench.py --gpus 1 --no-cpu > $O/scale_config2_1.json 2> $O/scale_config2_1.err
echo done
