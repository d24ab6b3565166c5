"""CLI eval, same flag surface as the reference's test_generator.py (:42-144): per-category IoU / MAE of the generated
masks (threshold 0.1, border-score disambiguation); with --generate_visualization the PNG overlays and the .mat files of :93-118."""
import os
import sys

import numpy as np
from absl import flags as gflags

from unsupervised_detection_b200.common_flags import FLAGS
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner
from unsupervised_detection_b200.models.utils.general_utils import compute_IoU, compute_mae, postprocess_image, postprocess_mask

des_width, des_height = 640, 384      # test_generator.py:14-15


def _test_masks():
    learner = AdversarialLearner()
    learner.setup_inference(FLAGS, aug_test=False)
    if not FLAGS.ckpt_file:
        raise IOError("Checkpoint file not found")           # test_generator.py:58
    learner.restore(FLAGS.ckpt_file)
    print("Resume model from checkpoint {}".format(FLAGS.ckpt_file))
    CategoryIou, CategoryMae = {}, {}
    n_steps = int(np.ceil(learner.test_samples / float(FLAGS.batch_size)))
    i = 0
    for step in range(n_steps):
        inference = learner.inference(None)
        for b in range(inference['input_image'].shape[0]):
            generated_mask = inference['gen_masks'][b]
            gt_mask = inference['gt_masks'][b]
            category = inference['img_fname'][b].decode("utf-8").split('/')[-2]
            iou, out_mask = compute_IoU(gt_mask=gt_mask, pred_mask_f=generated_mask)
            mae = compute_mae(gt_mask=gt_mask, pred_mask_f=out_mask)
            CategoryIou.setdefault(category, []).append(iou)
            CategoryMae.setdefault(category, []).append(mae)
            if FLAGS.generate_visualization:                       # test_generator.py:93-118
                import cv2
                import scipy.io as sio
                save_dir = os.path.join(FLAGS.test_save_dir, category)
                os.makedirs(save_dir, exist_ok=True)
                k = len(CategoryIou[category])
                bgr = postprocess_image(inference['input_image'][b])
                red = postprocess_mask(out_mask.astype(np.float32))
                res = cv2.resize(cv2.addWeighted(bgr, 0.5, red, 0.4, 0), (des_width, des_height))
                cv2.imwrite(os.path.join(save_dir, "frame_{:08d}.png".format(k)), res)
                sio.savemat(os.path.join(save_dir, 'result_{}.mat'.format(k)),
                            {'flow': inference['gt_flow'][b], 'img1': cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), 'pred_mask': out_mask,
                             'gt_mask': inference['gt_masks'][b]})
            i += 1
    tot_ious = tot_maes = 0
    per_cat_iou = []
    for cat, list_iou in CategoryIou.items():
        print("Category {}: IoU is {} and MAE is {}".format(cat, np.mean(list_iou), np.mean(CategoryMae[cat])))
        tot_ious += np.sum(list_iou)
        tot_maes += np.sum(CategoryMae[cat])
        per_cat_iou.append(np.mean(list_iou))
    print("The Average over the dataset: IoU is {} and MAE is {}".format(tot_ious / float(i), tot_maes / float(i)))
    print("The Average over sequences IoU is {}".format(np.mean(per_cat_iou)))
    print("Success: Processed {} frames".format(i))


def _test_masks_dp():
    """Batch-sharded variant used under torchrun: every rank evaluates its slice of each global batch of `batch_size * world` frames;
    scores are merged with one all_gather_object and rank 0 prints the report.  Same per-frame work and file names as `_test_masks`."""
    from unsupervised_detection_b200 import eval_dp
    learner = AdversarialLearner()
    learner.setup_inference(FLAGS, aug_test=False)
    if not FLAGS.ckpt_file:
        raise IOError("Checkpoint file not found")
    learner.restore(FLAGS.ckpt_file)
    rank, world, total, b = learner.rank, learner.world, int(learner.test_samples), int(FLAGS.batch_size)
    names = eval_dp.global_names(learner)
    vt = eval_dp.virtual_total(total, b)      # the single-process loop scores ceil(total/b)*b frames (its last batch wraps around)
    names = [names[g % total] for g in range(vt)]
    counters = eval_dp.category_counters(names)
    local = []
    for step in range(eval_dp.steps_for(total, b, world)):
        inference = learner.inference(None)
        for j, gidx in enumerate(eval_dp.owned_indices(step, b, rank, world, total)):
            if gidx >= vt or j >= inference['input_image'].shape[0]:
                continue
            gt_mask = inference['gt_masks'][j]
            iou, out_mask = compute_IoU(gt_mask=gt_mask, pred_mask_f=inference['gen_masks'][j])
            mae = compute_mae(gt_mask=gt_mask, pred_mask_f=out_mask)
            category = names[gidx].split('/')[-2]
            local.append((gidx, category, float(iou), float(mae)))
            if FLAGS.generate_visualization:
                import cv2
                import scipy.io as sio
                save_dir = os.path.join(FLAGS.test_save_dir, category)
                os.makedirs(save_dir, exist_ok=True)
                k = counters[gidx]
                bgr = postprocess_image(inference['input_image'][j])
                red = postprocess_mask(out_mask.astype(np.float32))
                cv2.imwrite(os.path.join(save_dir, "frame_{:08d}.png".format(k)),
                            cv2.resize(cv2.addWeighted(bgr, 0.5, red, 0.4, 0), (des_width, des_height)))
                sio.savemat(os.path.join(save_dir, 'result_{}.mat'.format(k)),
                            {'flow': inference['gt_flow'][j], 'img1': cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB), 'pred_mask': out_mask,
                             'gt_mask': inference['gt_masks'][j]})
    scores = eval_dp.merge_scores(local)
    if rank == 0:
        eval_dp.report(scores, sequence_average=True)
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
