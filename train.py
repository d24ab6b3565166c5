"""CLI entry, same flag surface as the reference's train.py (:16-43): seeds, prints flags, AdversarialLearner().train(FLAGS)."""
import os
import pprint
import random
import sys

import numpy as np
import torch
from absl import flags as gflags

from unsupervised_detection_b200.common_flags import FLAGS
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner


def _main():
    seed = 8964                                     # train.py:18
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    pp = pprint.PrettyPrinter()
    from unsupervised_detection_b200.common_flags import FLAG_NAMES
    pp.pprint({k: getattr(FLAGS, k) for k in FLAG_NAMES})
    if FLAGS.checkpoint_dir and not os.path.exists(FLAGS.checkpoint_dir):
        os.makedirs(FLAGS.checkpoint_dir)
    trl = AdversarialLearner()
    trl.train(FLAGS)


def main(argv):
    try:
        argv = FLAGS(argv)  # parse flags
    except gflags.Error:
        print('Usage: %s ARGS\n%s' % (sys.argv[0], FLAGS))
        sys.exit(1)
    _main()


if __name__ == "__main__":
    main(sys.argv)
