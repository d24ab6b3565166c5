"""Generates tests/golden/*.npz with the fp64 oracle (run here, in the build container; /root/reference cannot execute:
TF1.13 is not installable).  Fixtures are small seeded (input, weights-seed, output) triples."""
import os
import sys
import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import params as OP, nets as ON, pwcnet as PW, losses as OL  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def smooth(g, B, H, W, C, amp):
    lo = torch.randn(B, C, max(H // 8, 2), max(W // 8, 2), generator=g, dtype=torch.float64)
    return (F.interpolate(lo, size=(H, W), mode='bicubic', align_corners=False) * amp).permute(0, 2, 3, 1).contiguous()


def main():
    g = torch.Generator().manual_seed(1234)
    p = {k: v.double() for k, v in OP.make_params(seed=77, jitter=0.1).items()}   # same fp32 values, fp64 arithmetic
    B, H, W = 1, 32, 48
    image = torch.rand(B, H, W, 3, generator=g, dtype=torch.float64) - 0.5
    flow = smooth(g, B, H, W, 2, 0.3)
    L = OL.adversarial_losses(image, flow, p)
    np.savez_compressed(os.path.join(HERE, 'cis_losses_32x48.npz'), image=image.numpy(), flow=flow.numpy(), mask=L['masks'].numpy(),
                        pred=L['pred'].numpy(), pred_c=L['pred_c'].numpy(), pred_i=L['pred_i'].numpy(),
                        generator=float(L['generator']), recover=float(L['recover']), seed=77, jitter=0.1)
    img1 = smooth(g, 1, 64, 64, 3, 0.25).clamp(-0.5, 0.5)
    img2 = torch.roll(img1, shifts=(1, 2), dims=(1, 2))
    fl, pyr, c1, c2 = PW.predict_from_img_pairs(img1, img2, p, return_pyr=True)
    np.savez_compressed(os.path.join(HERE, 'pwc_64x64.npz'), img1=img1.numpy(), img2=img2.numpy(), flow=fl.numpy(),
                        flow6=pyr[0].numpy(), flow2=pyr[-1].numpy(), c1_3=c1[3].numpy(), seed=77, jitter=0.1)
    c1 = torch.randn(1, 6, 7, 8, generator=g, dtype=torch.float64)
    c2 = torch.randn(1, 6, 7, 8, generator=g, dtype=torch.float64)
    fw = smooth(g, 1, 6, 7, 2, 1.5)
    wr = PW.dense_image_warp(c2, fw)
    cv = PW.cost_volume(c1, wr)
    np.savez_compressed(os.path.join(HERE, 'warp_costvol_6x7.npz'), c1=c1.numpy(), c2=c2.numpy(), flow=fw.numpy(), warp=wr.numpy(), cv=cv.numpy())
    print('golden fixtures written to', HERE)


if __name__ == '__main__':
    main()
