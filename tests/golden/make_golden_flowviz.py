"""Golden vectors for the flow colour code: runs the REFERENCE's numpy functions (models/utils/flow_utils.py:14-100) in this
container -- `import tensorflow` is stubbed out, only the numpy part is executed -- and stores input + output.
Usage: python tests/golden/make_golden_flowviz.py   (needs /root/reference; the .npz is committed)."""
import os
import types

import numpy as np

src = open('/root/reference/models/utils/flow_utils.py').read().replace('import tensorflow as tf', 'tf = None')
ref = types.ModuleType('ref_flow_utils')
exec(compile(src, 'ref_flow_utils', 'exec'), ref.__dict__)

rng = np.random.RandomState(2024)
flow = (rng.randn(3, 20, 28, 2) * np.array([1.0, 6.0, 0.3])[:, None, None, None]).astype(np.float32)
flow[0, 0, 0] = 0.0
flow[1, 3, 4, 0] = 1e9            # "unknown" flow marker
flow[2, 5, 5] = (0.0, -0.25)
img = ref.flow_to_image(flow.copy())
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'flowviz.npz'), flow=flow, image=img.astype(np.uint8),
                    wheel=ref.make_color_wheel().astype(np.uint8))
print('flowviz.npz', img.shape, img.dtype, float(img.mean()))
