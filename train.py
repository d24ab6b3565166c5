"""Training CLI with the reference's flag surface (its train.py:16-43): fixed seed 8964, flag dump, then
`AdversarialLearner().train(FLAGS)`.  Under torchrun every rank runs this file; only rank 0 prints."""
import os
import pprint
import random
import sys

import numpy as np
import torch
from absl import flags as absl_flags

from unsupervised_detection_b200.common_flags import FLAGS, FLAG_NAMES
from unsupervised_detection_b200.models.adversarial_learner import AdversarialLearner

SEED = 8964


def seed_everything(seed=SEED):
    for fn in (torch.manual_seed, np.random.seed, random.seed):
        fn(seed)


def run(config):
    seed_everything()
    if int(os.environ.get('RANK', '0')) == 0:
        pprint.pprint({name: getattr(config, name) for name in FLAG_NAMES})
    if config.checkpoint_dir:
        os.makedirs(config.checkpoint_dir, exist_ok=True)
    AdversarialLearner().train(config)


def main(argv):
    try:
        FLAGS(argv)
    except absl_flags.Error as err:
        sys.exit('%s\nUsage: %s ARGS\n%s' % (err, argv[0], FLAGS))
    run(FLAGS)


if __name__ == "__main__":
    main(sys.argv)
