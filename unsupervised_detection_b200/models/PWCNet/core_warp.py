"""Reference module path models/PWCNet/core_warp.py: `dense_image_warp` (:153-202), implemented in ...functional on cis_dense_image_warp."""
from ..functional import dense_image_warp  # noqa: F401
