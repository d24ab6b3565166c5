"""Reference module path models/utils/loss_utils.py: `charbonnier_loss` (:34-51) and `train_op` (:12-32) of the function-level API
(implemented in ..functional on libcis_b200 kernels)."""
from ..functional import charbonnier_loss, train_op  # noqa: F401
