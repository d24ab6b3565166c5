"""Data-parallel contract of SURVEY.md section 8e on CPU (gloo, world_size 2): with every rank's loss divided by the GLOBAL batch,
one SUM all-reduce of the active network's gradients reproduces the single-process global-batch gradient, after which clip /
TF-Adam are identical on every rank.  (The same contract is what AdversarialLearner._allreduce implements with NCCL.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import params as OP, losses as OL

H, W, GB = 32, 48, 2


def _inputs():
    g = torch.Generator().manual_seed(11)
    image = torch.rand(GB, H, W, 3, generator=g) - 0.5
    lo = torch.randn(GB, 2, 4, 6, generator=g)
    flow = (torch.nn.functional.interpolate(lo, size=(H, W), mode='bicubic') * 0.3).permute(0, 2, 3, 1).contiguous()
    return image, flow


def _grads(p, image, flow, scope, key, gb):
    names = [n for n in p if n.startswith(scope)]
    pr = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    L = OL.adversarial_losses(image, flow, pr, global_batch=gb)
    return names, torch.autograd.grad(L[key], [pr[n] for n in names]), float(L[key])


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    torch.set_num_threads(2)
    p = OP.make_params(seed=21, jitter=0.05, nets=('MaskNet', 'FlownetS'))
    image, flow = _inputs()
    lb = GB // world
    sl = slice(rank * lb, (rank + 1) * lb)          # rank r takes samples [r*b, (r+1)*b)
    out = {}
    for scope, key in (('FlownetS/', 'recover'), ('MaskNet/', 'generator')):
        names, gr, loss = _grads(p, image[sl], flow[sl], scope, key, GB)
        flat = torch.cat([g.reshape(-1) for g in gr])
        dist.all_reduce(flat)                          # SUM, no division by world
        lt = torch.tensor([loss])
        dist.all_reduce(lt)
        out[key] = (flat, float(lt))
    if rank == 0:
        q.put({k: (v[0].numpy(), v[1]) for k, v in out.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_sum_allreduce_reproduces_global_batch_gradient():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    p = OP.make_params(seed=21, jitter=0.05, nets=('MaskNet', 'FlownetS'))
    image, flow = _inputs()
    for scope, key in (('FlownetS/', 'recover'), ('MaskNet/', 'generator')):
        _, gr, loss = _grads(p, image, flow, scope, key, GB)
        flat = torch.cat([g.reshape(-1) for g in gr])
        got = torch.from_numpy(res[key][0])
        assert abs(res[key][1] - loss) < 1e-5 * max(1.0, abs(loss))
        assert float((got - flat).abs().max()) <= 1e-5 * float(flat.abs().max()) + 1e-9, key
