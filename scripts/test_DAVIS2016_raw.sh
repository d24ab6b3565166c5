#!/bin/bash
# Raw (no post-processing) evaluation on DAVIS 2016 (scripts/test_DAVIS2016_raw.sh of the reference, minus the downloads: this
# environment has no network).  CKPT_FILE / PWC_CKPT_FILE may be TF checkpoint prefixes, their .index / .data-* files, or .pt files.
DOWNLOAD_DIR=${DOWNLOAD_DIR:-./download}
CKPT_FILE=${CKPT_FILE:-${DOWNLOAD_DIR}/unsupervised_detection_models/davis_best_model/model.best}
PWC_CKPT_FILE=${PWC_CKPT_FILE:-${DOWNLOAD_DIR}/pwcnet-lg-6-2-multisteps-chairsthingsmix/pwcnet.ckpt-595000.data-00000-of-00001}
DATASET_FILE=${DATASET_FILE:-${DOWNLOAD_DIR}/DAVIS}
RESULT_DIR=${RESULT_DIR:-./results/DAVIS2016_raw}
mkdir -p ${RESULT_DIR}
python3 test_generator.py \
--dataset=DAVIS2016 \
--ckpt_file=$CKPT_FILE \
--flow_ckpt=$PWC_CKPT_FILE \
--test_crop=0.9 \
--test_temporal_shift=1 \
--root_dir=$DATASET_FILE \
--generate_visualization=True \
--test_save_dir=${RESULT_DIR} "$@"
