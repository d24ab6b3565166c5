"""Scratch probe (GPU): full step graph vs the CPU oracle.  Writes gpurun_out/probe_graph.txt."""
import os, sys, json, time, traceback
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))   # repo root
from oracle import params as OP, losses as OL, nets as ON, pwcnet as OW
from unsupervised_detection_b200.step_graph import CISGraph

os.makedirs('gpurun_out', exist_ok=True)
out = open('gpurun_out/probe_graph.txt', 'w')


def log(*a):
    s = ' '.join(str(x) for x in a)
    print(s)
    out.write(s + '\n')
    out.flush()


def err(a, b):
    a = a.detach().float().cpu()
    b = b.detach().float().cpu()
    return dict(max=float((a - b).abs().max()), mean=float((a - b).abs().mean()), ref=float(b.abs().max()))


def smooth(B, H, W, C, amp, gen):
    lo = torch.randn(B, C, max(H // 16, 2), max(W // 16, 2), generator=gen)
    return (F.interpolate(lo, size=(H, W), mode='bicubic', align_corners=False) * amp).permute(0, 2, 3, 1).contiguous()


def main():
    gen = torch.Generator().manual_seed(0)
    B, H, W = 2, 64, 96
    p = OP.make_params(seed=1, jitter=0.1)
    g = CISGraph(H, W, B, with_pwc=False)
    g.load_params(p)
    image = torch.rand(B, H, W, 3, generator=gen) - 0.5
    flow = smooth(B, H, W, 2, 0.3, gen)
    g.image.copy_(image)
    g.flow.copy_(flow)
    g.forward()
    torch.cuda.synchronize()
    pr = {k: v.clone().requires_grad_(not k.startswith('pwcnet')) for k, v in p.items()}
    L = OL.adversarial_losses(image, flow, pr)
    log('mask', err(g.mask, L['masks']))
    log('mask thr mismatches', int(((g.mask.cpu() > 0.1) != (L['masks'] > 0.1)).sum()), 'of', L['masks'].numel())
    for j, k in enumerate(('pred', 'pred_c', 'pred_i')):
        log(k, err(g.pred[j * B:(j + 1) * B], L[k]))
    log('losses gpu', g.losses(), 'oracle', float(L['generator']), float(L['recover']))
    for mode, key, scope in (('R', 'recover', 'FlownetS/'), ('G', 'generator', 'MaskNet/')):
        names = [n for n in pr if n.startswith(scope)]
        grads = torch.autograd.grad(L[key], [pr[n] for n in names], retain_graph=True)
        g.bwd[mode].run()
        torch.cuda.synchronize()
        store = g.rec_store if mode == 'R' else g.gen_store
        worst = []
        fa = torch.cat([store.view(n, 'grad').reshape(-1).cpu() for n in names])
        fb = torch.cat([gr.reshape(-1) for gr in grads])
        log('grads', mode, 'flat cosine', float(torch.dot(fa, fb) / (fa.norm() * fb.norm())), 'norm ratio', float(fa.norm() / fb.norm()))
        for n, gr in zip(names, grads):
            e = err(store.view(n, 'grad'), gr)
            worst.append((e['max'] / (e['ref'] + 1e-12), n, e))
        worst.sort(reverse=True)
        log('grads', mode, 'worst5', json.dumps([(round(w[0], 4), w[1], w[2]) for w in worst[:5]]))
        log('grads', mode, 'median rel', sorted(w[0] for w in worst)[len(worst) // 2])
    # ---- training steps vs oracle
    # oracle train_step needs (img, flow) directly: inline variant
    pt = {k: v.clone() for k, v in p.items()}
    opt = OL.TFAdam()
    for step in range(1, 6):
        rec_step = OL.is_recover_step(step)
        scope = 'FlownetS/' if rec_step else 'MaskNet/'
        names = [n for n in pt if n.startswith(scope)]
        for n in names:
            pt[n].requires_grad_(True)
        Ls = OL.adversarial_losses(image, flow, pt)
        loss = Ls['recover'] if rec_step else Ls['generator']
        grads = torch.autograd.grad(loss, [pt[n] for n in names])
        for n in names:
            pt[n].requires_grad_(False)
        clipped, _ = OL.clip_or_noise(list(grads), 0.2, can_change=not rec_step)
        opt.apply(pt, names, clipped)
        g.train_step('R' if rec_step else 'G')
        torch.cuda.synchronize()
        ex = g.export_params()
        worst = max(((float((ex[n].cpu() - pt[n]).abs().max()), n) for n in names))
        log('step', step, 'R' if rec_step else 'G', 'loss gpu', g.losses(), 'oracle', float(Ls['generator']), float(Ls['recover']),
            'max param diff', worst)
    # ---- PWC-Net
    Bp, ph, pw = 1, 128, 192
    g2 = CISGraph(64, 96, Bp, with_pwc=True, pwc_hw=(ph, pw), train=False)
    g2.load_params(p)
    img1 = smooth(Bp, ph, pw, 3, 0.25, gen).clamp(-0.5, 0.5)
    img2 = torch.roll(img1, shifts=(1, 2), dims=(1, 2)) + 0.01 * torch.randn(Bp, ph, pw, 3, generator=gen)
    g2.img1.copy_(img1)
    g2.img2.copy_(img2)
    g2.forward()
    torch.cuda.synchronize()
    fo, pyr, c1, c2 = OW.predict_from_img_pairs(img1, img2, p, return_pyr=True)
    for l in range(1, 7):
        log('c1 lvl', l, err(g2.pwc.c1[l].float(), c1[l]), 'c2', err(g2.pwc.c2[l].float(), c2[l]))
    for i, l in enumerate(range(6, 1, -1)):
        log('flow lvl', l, err(g2.pwc.flows[l], pyr[i]))
    log('flow_pred', err(g2.flow_full, fo))
    im, fl = OL.resize_inputs(img1, fo, 64, 96)
    log('image_s', err(g2.image, im), 'flow_s', err(g2.flow, fl))


try:
    main()
except Exception:
    log('EXC', traceback.format_exc())
