#!/bin/bash
# tools/scatter_probe.sh <gpus>: the sharded leg under a few NCCL settings (one torchrun each)
N=$1
run() { env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/scatter_probe.py 2>&1 | grep -E "ms per rollout|Error|error" | head -3; }
run X=0
run NCCL_MIN_P2P_NCHANNELS=8 NCCL_MAX_P2P_NCHANNELS=32
run NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run NCCL_MIN_P2P_NCHANNELS=16 NCCL_MAX_P2P_NCHANNELS=32 NCCL_P2P_NVL_CHUNKSIZE=1048576
