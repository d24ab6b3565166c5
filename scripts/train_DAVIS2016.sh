#!/bin/bash
# Train on DAVIS 2016 with the reference's hyper-parameters (scripts/train_DAVIS2016.sh of antonilo/unsupervised_detection).
# --flow_ckpt / --recover_ckpt accept the authors' TF checkpoints (prefix, .index or .data-* file name) or native .pt files.
# Multi-GPU: NGPU=8 scripts/train_DAVIS2016.sh   (--batch_size stays the whole job's batch and is sharded over the ranks)
ROOT_DIR=${ROOT_DIR:-/path/to/DAVIS_2016/}
FLOW_CKPT=${FLOW_CKPT:-/path/to/PWCNet/pwcnet-lg-6-2-multisteps-chairsthingsmix/pwcnet.ckpt-595000}
RECOVER_CKPT=${RECOVER_CKPT:-/path/to/pretrained_recover/model-175}
NGPU=${NGPU:-1}
if [ "$NGPU" -gt 1 ]; then
  LAUNCH="python3 -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29500}"
else
  LAUNCH="python3"
fi
$LAUNCH train.py \
--flow_normalizer=80.0 \
--epsilon=75.0 \
--max_temporal_len=2 \
--train_crop=0.6 \
--test_crop=0.9 \
--iters_rec=1 \
--iters_gen=3 \
--dataset=DAVIS2016 \
--root_dir="$ROOT_DIR" \
--flow_ckpt="$FLOW_CKPT" \
--recover_ckpt="$RECOVER_CKPT" \
--test_temporal_shift=1 \
--checkpoint_dir=${CHECKPOINT_DIR:-/tmp/tests} "$@"
