"""CLI multi-crop evaluation, same flag surface as the reference's test_generator_ensemble.py (:20-135): crops 0.85/0.9/0.95/1.0 of
every test frame go through PWC-Net + generator (batched here instead of batch-1 graph copies), per-frame `.mat` buffers with the
keys post_processing/generate_soft_score_from_buffer.py:45-92 expects are written with --generate_visualization."""
import os
import sys

import numpy as np
from absl import flags as gflags

from unsupervised_detection_b200.common_flags import FLAGS
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
from unsupervised_detection_b200.models.utils.general_utils import compute_IoU, compute_mae, postprocess_image, postprocess_mask

des_width, des_height = 640, 384


def _test_masks():
    learner = AdversarialLearner()
    learner.setup_inference(FLAGS, aug_test=True)
    if not FLAGS.ckpt_file:
        raise IOError("Checkpoint file not found")
    learner.restore(FLAGS.ckpt_file)
    print("Resume model from checkpoint {}".format(FLAGS.ckpt_file))
    CategoryIou, CategoryMae = {}, {}
    test_crops = learner.test_crops
    i = 0
    for step in range(learner.test_samples):
        inference = learner.inference(None)
        outputs = inference['outs']
        fname = inference['img_fname']
        fname = fname.item() if hasattr(fname, 'item') else fname
        cropped_iou, cropped_mae = [], []
        for crop in test_crops:
            iou, out_mask = compute_IoU(gt_mask=outputs['gt_masks'][crop], pred_mask_f=outputs['pred_masks'][crop])
            outputs['pred_masks'][crop] = out_mask                      # test_generator_ensemble.py:66-67 "take the best one"
            cropped_iou.append(iou)
            cropped_mae.append(compute_mae(gt_mask=outputs['gt_masks'][crop], pred_mask_f=out_mask))
        category = fname.decode("utf-8").split('/')[-2]
        # the reference stores the LAST crop's numbers for the first frame of a category (:76-81); the crop mean is used for all frames here
        CategoryIou.setdefault(category, []).append(np.mean(cropped_iou))
        CategoryMae.setdefault(category, []).append(np.mean(cropped_mae))
        if FLAGS.generate_visualization:
            import cv2
            import scipy.io as sio
            save_dir = os.path.join(FLAGS.test_save_dir, category)
            os.makedirs(save_dir, exist_ok=True)
            k = len(CategoryIou[category])
            bgr = postprocess_image(outputs['img_1s'][test_crops[-1]])
            red = postprocess_mask(outputs['pred_masks'][test_crops[-1]].astype(np.float32))
            cv2.imwrite(os.path.join(save_dir, "frame_{:08d}.png".format(k)),
                        cv2.resize(cv2.addWeighted(bgr, 0.5, red, 0.4, 0), (des_width, des_height)))
            matlab_out = {}
            for crop in test_crops:
                matlab_out['img_1_{:03d}'.format(int(crop * 100))] = outputs['img_1s'][crop]
                matlab_out['pred_mask_{:03d}'.format(int(crop * 100))] = outputs['pred_masks'][crop]
                matlab_out['gt_mask_{:03d}'.format(int(crop * 100))] = outputs['gt_masks'][crop]
            sio.savemat(os.path.join(save_dir, 'result_{}.mat'.format(k)), matlab_out)
        i += 1
    tot_ious = tot_maes = 0
    for cat, list_iou in CategoryIou.items():
        print("Category {}: IoU is {} and MAE is {}".format(cat, np.mean(list_iou), np.mean(CategoryMae[cat])))
        tot_ious += np.sum(list_iou)
        tot_maes += np.sum(CategoryMae[cat])
    print("The Average over the dataset: IoU is {} and MAE is {}".format(tot_ious / float(i), tot_maes / float(i)))
    print("Success: Processed {} frames".format(i))


def _test_masks_dp():
    """Batch-sharded variant used under torchrun (BASELINE config 4): rank r evaluates frames r, r + world, ...; scores are merged
    with one all_gather_object and rank 0 prints the report.  Same per-frame work and file names as `_test_masks`."""
    from unsupervised_detection_b200 import eval_dp
    learner = AdversarialLearner()
    learner.setup_inference(FLAGS, aug_test=True)
    if not FLAGS.ckpt_file:
        raise IOError("Checkpoint file not found")
    learner.restore(FLAGS.ckpt_file)
    rank, world, total = learner.rank, learner.world, int(learner.test_samples)
    names = eval_dp.global_names(learner)
    counters = eval_dp.category_counters(names)
    test_crops = learner.test_crops
    local = []
    for step in range(eval_dp.steps_for(total, 1, world)):
        inference = learner.inference(None)                # every rank steps its iterator every time: the shards stay aligned
        gidx = eval_dp.owned_indices(step, 1, rank, world, total)[0]
        if gidx >= total:
            continue                                       # wrap-around duplicate of the endless iterator
        outputs = inference['outs']
        cropped_iou, cropped_mae = [], []
        for crop in test_crops:
            iou, out_mask = compute_IoU(gt_mask=outputs['gt_masks'][crop], pred_mask_f=outputs['pred_masks'][crop])
            outputs['pred_masks'][crop] = out_mask
            cropped_iou.append(iou)
            cropped_mae.append(compute_mae(gt_mask=outputs['gt_masks'][crop], pred_mask_f=out_mask))
        category = names[gidx].split('/')[-2]
        local.append((gidx, category, float(np.mean(cropped_iou)), float(np.mean(cropped_mae))))
        if FLAGS.generate_visualization:
            import cv2
            import scipy.io as sio
            save_dir = os.path.join(FLAGS.test_save_dir, category)
            os.makedirs(save_dir, exist_ok=True)
            k = counters[gidx]                             # the frame's running index inside its category in list order
            bgr = postprocess_image(outputs['img_1s'][test_crops[-1]])
            red = postprocess_mask(outputs['pred_masks'][test_crops[-1]].astype(np.float32))
            cv2.imwrite(os.path.join(save_dir, "frame_{:08d}.png".format(k)),
                        cv2.resize(cv2.addWeighted(bgr, 0.5, red, 0.4, 0), (des_width, des_height)))
            matlab_out = {}
            for crop in test_crops:
                matlab_out['img_1_{:03d}'.format(int(crop * 100))] = outputs['img_1s'][crop]
                matlab_out['pred_mask_{:03d}'.format(int(crop * 100))] = outputs['pred_masks'][crop]
                matlab_out['gt_mask_{:03d}'.format(int(crop * 100))] = outputs['gt_masks'][crop]
            sio.savemat(os.path.join(save_dir, 'result_{}.mat'.format(k)), matlab_out)
    scores = eval_dp.merge_scores(local)
    if rank == 0:
        eval_dp.report(scores)
    return scores


def main(argv):
    try:
        argv = FLAGS(argv)
    except gflags.Error:
        print('Usage: %s ARGS\n%s' % (sys.argv[0], FLAGS))
        sys.exit(1)
    from unsupervised_detection_b200 import eval_dp
    _test_masks_dp() if eval_dp.is_distributed_launch() else _test_masks()


if __name__ == "__main__":
    main(sys.argv)
