#!/bin/bash
# Multi-crop / multi-shift prediction buffers for the reference's offline post-processing
# (scripts/generate_buffer_DAVIS2016.sh of the reference: test_generator_ensemble.py over temporal shifts -2..2).
CKPT_FILE=${CKPT_FILE:-/path/to/checkpoint}
PWC_CKPT_FILE=${PWC_CKPT_FILE:-/path/to/pwc_ckpt/}
DATASET_FILE=${DATASET_FILE:-/path/to/DAVIS_2016/}
RESULT_DIR=${RESULT_DIR:-./results/DAVIS2016_buffer}
NGPU=${NGPU:-1}      # NGPU=8 shards the frames over 8 GPUs
if [ "$NGPU" -gt 1 ]; then LAUNCH="python3 -m torch.distributed.run --nnodes=1 --nproc-per-node $NGPU --master-addr 127.0.0.1 --master-port ${MASTER_PORT:-29500}"; else LAUNCH=python3; fi
for SHIFT in -2 -1 1 2; do
  $LAUNCH test_generator_ensemble.py \
  --dataset=DAVIS2016 \
  --ckpt_file=$CKPT_FILE \
  --flow_ckpt=$PWC_CKPT_FILE \
  --test_temporal_shift=$SHIFT \
  --root_dir=$DATASET_FILE \
  --test_partition='val' \
  --generate_visualization=True \
  --test_save_dir=${RESULT_DIR}/shift_${SHIFT} "$@"
done
