#!/bin/bash
# Run the mask generator on your own video (role of the reference's scripts/test_video.sh, minus the downloads):
#   VIDEO_FILE=clip.mp4 CKPT_FILE=... PWC_CKPT_FILE=... scripts/test_video.sh
SCRIPT_DIR=$(cd "$(dirname "$0")" && pwd)
VIDEO_FILE=${VIDEO_FILE:?set VIDEO_FILE to a video file}
DATASET_DIR=${DATASET_DIR:-./download/video}
RESULT_DIR=${RESULT_DIR:-./results/video}
python3 ${SCRIPT_DIR}/create_data_frvideo.py "$VIDEO_FILE" --out "$DATASET_DIR" || exit 1
mkdir -p ${RESULT_DIR}
python3 test_generator.py \
--dataset=DAVIS2016 \
--ckpt_file=${CKPT_FILE:?set CKPT_FILE} \
--flow_ckpt=${PWC_CKPT_FILE:?set PWC_CKPT_FILE} \
--test_crop=0.9 \
--test_temporal_shift=1 \
--root_dir=$DATASET_DIR \
--generate_visualization=True \
--test_save_dir=${RESULT_DIR} "$@"
