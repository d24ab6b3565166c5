"""Reference module path models/PWCNet/core_costvol.py: `cost_volume` (:20-40), implemented in ...functional on cis_warp_costvol."""
from ..functional import cost_volume  # noqa: F401
