"""Device-side step graph: everything adversarial_learner.py:72-258 puts in the TF graph, as static launch lists.

One `CISGraph` owns the parameters (flat fp32 master copies per scope), all activation buffers for a fixed
(batch, 384x640 -> HxW) geometry, and the launch lists: forward (PWC-Net -> resize -> generator -> mask (x) flow ->
3x recover -> Charbonnier losses), backward for the recover step, backward for the generator step, and clip + TF-Adam.
"""
import torch

from . import _lib
from .engine import Builder, ParamStore, Plan, Act
from .models.nets import GeneratorNet, RecoverNet
from .models.PWCNet.model_pwcnet import ModelPWCNet

PWC_H, PWC_W = 384, 640   # data/davis2016_data_utils.py:87-88: frames are resized to 384x640 before PWC-Net


class CISGraph(object):
    def __init__(self, img_height, img_width, batch, device='cuda', global_batch=None, flow_normalizer=80.0, cbn=0.5, epsilon=75.0,
                 beta1=0.9, with_pwc=True, train=True, pwc_hw=(PWC_H, PWC_W), seed=8964):
        _lib.load()
        self.H, self.W, self.B = img_height, img_width, batch
        self.GB = global_batch or batch
        self.dev = device
        self.cbn, self.eps_rr, self.flow_norm, self.beta1 = cbn, epsilon, flow_normalizer, beta1
        self.with_pwc, self.train = with_pwc, train
        self.seed = seed
        f32 = lambda *s: torch.zeros(*s, dtype=torch.float32, device=device)
        B, H, W = batch, img_height, img_width
        # ---- parameter stores (scopes of adversarial_learner.py:211-214 and model_pwcnet.py)
        self.gen_store, self.rec_store, self.pwc_store = ParamStore(device), ParamStore(device), ParamStore(device)
        self.gen = GeneratorNet(self.gen_store)
        self.rec = RecoverNet(self.rec_store)
        self.gen_store.finalize(True)
        self.rec_store.finalize(True)
        if with_pwc:
            self.pwc = ModelPWCNet(self.pwc_store)
            self.pwc_store.finalize(False)
        self.step_state = torch.zeros(1, dtype=torch.int64, device=device)   # shared Adam beta-power step (App. A.14)
        self.avg_abs = f32(1)
        # ---- static buffers
        bld = Builder(device)
        self.bld = bld
        P = bld.fwd
        if with_pwc:
            ph, pw = pwc_hw
            self.img1, self.img2 = f32(B, ph, pw, 3), f32(B, ph, pw, 3)
            self.flow_full = f32(B, ph, pw, 2)
            i1 = Act(B, ph, pw, 3, device, name='img1_8')
            i2 = Act(B, ph, pw, 3, device, name='img2_8')
            npx = B * ph * pw
            P.add('cis_pack_f32_to_bf16', self.img1.data_ptr(), npx, 3, 0.5, i1.ptr, 8, 0)     # adapt_x: img + 0.5
            P.add('cis_pack_f32_to_bf16', self.img2.data_ptr(), npx, 3, 0.5, i2.ptr, 8, 0)
            self.pwc.build(bld, i1, i2, self.flow_full)
        self.image = f32(B, H, W, 3)
        self.flow = f32(B, H, W, 2)
        self._pwc_ops = 0
        if with_pwc:
            # adversarial_learner.py:87-97: legacy bilinear to (H,W); flow / flow_normalizer.  The frozen flow network writes its
            # results to STAGE buffers; the two small device copies below hand them to the trainable part.  That split is what lets
            # train_step(pipeline=True) run PWC-Net for the NEXT batch on a second stream while this batch trains (the flow network
            # has no dependency on the parameters being trained).
            self.image_st, self.flow_st = f32(B, H, W, 3), f32(B, H, W, 2)
            P.add('cis_resize_bilinear_f32', self.img1.data_ptr(), B, ph, pw, 3, self.image_st.data_ptr(), H, W, 1.0)
            P.add('cis_resize_bilinear_f32', self.flow_full.data_ptr(), B, ph, pw, 2, self.flow_st.data_ptr(), H, W, 1.0 / flow_normalizer)
            self._pwc_ops = len(P.ops)                  # fwd.ops[:_pwc_ops] = everything that depends only on the frame pair
            P.add_py(self._take_stage, 'take_stage')
        self.stats = torch.zeros(B, 4, dtype=torch.float64, device=device)
        self.gen_in = Act(B, H, W, 5, device, name='gen_in')
        self.img8 = Act(B, H, W, 3, device, name='img8')
        hw = H * W
        P.zero(self.stats)
        P.add('cis_flow_stats', self.flow.data_ptr(), B, hw, self.stats.data_ptr())                       # flow_utils.py:10
        P.add('cis_pack_generator_input', self.image.data_ptr(), self.flow.data_ptr(), self.stats.data_ptr(), B, hw, self.gen_in.ptr)
        P.add('cis_pack_f32_to_bf16', self.image.data_ptr(), B * hw, 3, 0.0, self.img8.ptr, 8, 0)
        self.mask = f32(B, H, W, 1)
        bld.lane = 1
        i0 = len(P.ops)
        self.rec.build_a_encoder(bld, self.img8)        # side stream, overlaps the generator
        i1 = len(P.ops)
        bld.lane = 0
        self.gen.build(bld, self.gen_in, self.mask)
        self._mask_ops = (i0, i1, len(P.ops))           # forward ops [0,i0) + [i1,end) produce the masks (no recover net, no losses)
        # mask (x) flow -> recover inputs for the 3 calls (adversarial_learner.py:107-131)
        self.rec_in = Act(3 * B, H, W, 4, device, name='rec_in', dep={'G'})
        self.rec_in.gen_rows = 2 * B
        P.add('cis_mask_apply', self.flow.data_ptr(), self.mask.data_ptr(), B, hw, self.rec_in.ptr)
        self.dmask = f32(B, H, W)
        logits = self.gen.logits

        def mask_bwd(bp, mode):
            if mode != 'G':
                return
            g = self.rec_in.get_grad()
            assert self.rec_in.grad_written.get('G'), 'recover backward did not reach the mask inputs'
            lg = logits.get_grad()
            bp.add('cis_mask_bwd', self.flow.data_ptr(), self.mask.data_ptr(), self.dmask.data_ptr(), g.ptr, B, hw, lg.ptr)
            logits.grad_written['G'] = True
        bld.tape.append(mask_bwd)
        h1, w1 = -(-H // 2), -(-W // 2)
        self.h1, self.w1 = h1, w1
        self.flow1 = f32(3 * B, h1, w1, 2)
        P.join()
        self.rec.build(bld, self.img8, self.rec_in, self.flow1)
        # ---- losses (adversarial_learner.py:141-204)
        self.sums = torch.zeros(B, 5, dtype=torch.float64, device=device)
        self.scalars = f32(8)
        self.coef = f32(B, 4)
        self.pred = f32(3 * B, H, W, 2)
        P.zero(self.sums)
        P.add('cis_cis_loss_fwd', self.flow.data_ptr(), self.mask.data_ptr(), self.flow1.data_ptr(), B, H, W, h1, w1, cbn,
              self.sums.data_ptr(), self.pred.data_ptr())
        P.add('cis_cis_loss_reduce', self.sums.data_ptr(), B, self.GB, hw, epsilon, self.scalars.data_ptr(), self.coef.data_ptr())
        self.fwd = P
        # ---- weight packing plans
        self.pack_pwc = Plan('pack_pwc')
        if with_pwc:
            for L in self.pwc.all_layers():
                L.plan_pack(self.pack_pwc)
        self.bwd = {}
        self.adam = {}
        if train:
            self.dpred = f32(3 * B, H, W, 2)
            for mode, which in (('R', 0), ('G', 1)):
                nb = 3 * B if mode == 'R' else 2 * B
                head = Plan('loss_bwd_' + mode)
                head.add('cis_cis_loss_bwd', self.flow.data_ptr(), self.mask.data_ptr(), self.flow1.data_ptr(), self.coef.data_ptr(),
                         self.scalars.data_ptr(), B, H, W, h1, w1, cbn, which, self.dpred.data_ptr(), self.dmask.data_ptr())
                fg = self.rec.flow1.get_grad()
                head.add('cis_resize_f32_bwd_to_bf16', self.dpred.data_ptr(), nb, H, W, 2, h1, w1, fg.ptr, fg.pitch)
                body = bld.build_backward(mode, [self.rec.flow1])
                store = self.rec_store if mode == 'R' else self.gen_store
                layers = self.rec.all_layers() if mode == 'R' else self.gen.all_layers()
                # nothing to zero: every real entry of the flat gradient buffer is overwritten by cis_unpack_wgrad / cis_bn_chain each
                # step, and its padding slots are never written (they stay at their initial 0, also through the all-reduce)
                pre = Plan('zero_' + mode)
                fin = Plan('fin_' + mode)
                for L in layers:
                    L.plan_finalize(fin, mode)
                full = Plan('bwd_' + mode)
                for pl in (pre, head, body):
                    full.extend(pl)
                full.join()            # weight-gradient lane -> main lane before the packed gradients are unpacked
                full.extend(fin.batch_param_ops(device))
                self.bwd[mode] = full
                ad = Plan('adam_' + mode)
                if mode == 'G':
                    # can_change branch of train_op (loss_utils.py:18-26)
                    self.seg = torch.tensor([v for pr in store.seg_pairs for v in pr], dtype=torch.int64, device=device)
                    ad.zero(self.avg_abs)
                    ad.add('cis_grad_avg_abs', store.grad.data_ptr(), self.seg.data_ptr(), len(store.seg_pairs), self.avg_abs.data_ptr())
                ad.add('cis_clip_adam', store.flat.data_ptr(), store.m.data_ptr(), store.v.data_ptr(), store.grad.data_ptr(), store.size, 1.0,
                       0.2, 1e-4, beta1, 0.999, 1e-8, self.step_state.data_ptr(), self.avg_abs.data_ptr(), 1 if mode == 'G' else 0, seed)
                self.adam[mode] = ad
        # packing of the trainable nets (forward + data-gradient orientation), after backward planning decided what is needed
        self.pack_gen, self.pack_rec = Plan('pack_gen'), Plan('pack_rec')
        for L in self.gen.all_layers():
            L.plan_pack(self.pack_gen, dgrad=train)
        for L in self.rec.all_layers():
            L.plan_pack(self.pack_rec, dgrad=train)
        self.pack_gen, self.pack_rec, self.pack_pwc = (pl.batch_param_ops(device) for pl in (self.pack_gen, self.pack_rec, self.pack_pwc))
        self._pwc_packed = False
        self._dirty = True      # packed bf16 operands out of date w.r.t. the fp32 master weights
        self.graphs = {}

    # ------------------------------------------------------------------------------------------------ parameters
    def load_params(self, params):
        """params: dict name -> tensor with the reference's variable layout (see oracle/params.py for the names)."""
        self.gen_store.load(params)
        self.rec_store.load(params)
        self._dirty = True
        if self.with_pwc:
            self.pwc_store.load(params)
            self._pwc_packed = False

    def export_params(self):
        out = {}
        out.update(self.gen_store.export())
        out.update(self.rec_store.export())
        if self.with_pwc:
            out.update(self.pwc_store.export())
        return out

    def param_count(self):
        return self.gen_store.real_count() + self.rec_store.real_count() + (self.pwc_store.real_count() if self.with_pwc else 0)

    # ------------------------------------------------------------------------------------------------ execution
    def _ensure_pwc(self):
        if self.with_pwc and not self._pwc_packed:
            self.pack_pwc.run()
            self._pwc_packed = True

    def _ensure_packed(self):
        self._ensure_pwc()
        if self._dirty:
            self.pack_gen.run()
            self.pack_rec.run()
            self._dirty = False

    def _pack_of(self, mode):
        return self.pack_rec if mode == 'R' else self.pack_gen

    def forward(self):
        self._ensure_packed()
        self.pipeline_drain()
        self._stage_valid = False
        self.fwd.run()

    def forward_masks(self, use_graph=False):
        """Mask path only: [PWC-Net -> resize ->] flow normalisation -> generator -> self.mask (what test_generator*.py consume;
        the reference's multi-crop test graph, adversarial_learner.py:525-592, builds nothing else).  use_graph replays it as one
        CUDA graph."""
        self._ensure_packed()
        self.pipeline_drain()
        self._stage_valid = False
        if getattr(self, '_mask_plan', None) is None:
            i0, i1, i2 = self._mask_ops
            mp = Plan('fwd_masks')
            mp.ops = self.fwd.ops[:i0] + self.fwd.ops[i1:i2]
            mp.keep = self.fwd.keep
            self._mask_plan = mp
        if not use_graph:
            self._mask_plan.run()
            return
        g = self.graphs.get('masks')
        if g is None:
            torch.cuda.synchronize()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._mask_plan.run()          # warm-up outside capture (function attributes, lazy allocations)
            torch.cuda.current_stream().wait_stream(s)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._mask_plan.run()
            self.graphs['masks'] = g
        g.replay()

    def _take_stage(self):
        self.image.copy_(self.image_st, non_blocking=True)
        self.flow.copy_(self.flow_st, non_blocking=True)

    def launches_per_step(self, mode):
        return self.fwd.count() + self.bwd[mode].count() + self.adam[mode].count() + self._pack_of(mode).count()

    def train_step(self, mode, allreduce=None, use_graph=False, pipeline=False, inputs_ready=None):
        """One alternating step body (adversarial_learner.py:380-397): mode 'R' = train_recover_op, 'G' = train_generator_op.

        pipeline=True (CUDA graphs only): software pipeline over steps.  The step trains on the (image, flow) pair that the PREVIOUS
        call's side branch left in the stage buffers and, concurrently on a second stream, runs the frozen PWC-Net on the frame
        pair currently in self.img1 / self.img2 -- the NEXT batch -- for the next call.  `inputs_ready`: an event after which
        img1 / img2 hold that next batch (the host-to-device copy).  The first pipelined call primes the stage from img1 / img2.
        Results are identical to the sequential order; only the schedule changes."""
        self._ensure_packed()
        if use_graph and pipeline and self.with_pwc:
            return self._train_step_pipelined(mode, allreduce, inputs_ready)
        if inputs_ready is not None:
            torch.cuda.current_stream().wait_event(inputs_ready)
        self.pipeline_drain()
        self._stage_valid = False
        if use_graph:
            g = self.graphs.get(mode)
            if g is None:
                g = self._capture(mode)
            g[0].replay()
            if allreduce is not None:
                allreduce((self.rec_store if mode == 'R' else self.gen_store).grad)
            g[1].replay()
            return
        self.fwd.run()
        self.bwd[mode].run()
        if allreduce is not None:
            allreduce((self.rec_store if mode == 'R' else self.gen_store).grad)
        self.adam[mode].run()
        self._pack_of(mode).run()      # only the updated network's bf16 operands are re-packed

    def _capture(self, mode):
        g1 = self._capture_plans('seq_fwd_bwd_' + mode, [self.fwd, self.bwd[mode]])
        g2 = self._capture_plans('seq_adam_' + mode, [self.adam[mode], self._pack_of(mode)], warm=False)
        self.graphs[mode] = (g1, g2)
        return self.graphs[mode]

    def _sub_plan(self, lo, hi, name):
        sp = Plan(name)
        sp.ops = self.fwd.ops[lo:hi]
        sp.keep = self.fwd.keep
        return sp

    def _capture_plans(self, key, plans, lane_key=0, warm=True):
        """One CUDA graph of `plans`.  warm: run them once outside capture first (sets function attributes, lazy allocations) -- never
        for plans that change state (the optimiser step)."""
        g = self.graphs.get(key)
        if g is None:
            torch.cuda.synchronize()
            if warm:
                s = torch.cuda.Stream()
                s.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(s):
                    for pl in plans:
                        pl.run(lane_key=lane_key)
                torch.cuda.current_stream().wait_stream(s)
                torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for pl in plans:
                    pl.run(lane_key=lane_key)
            self.graphs[key] = g
        return g

    def _pipe_state(self):
        if getattr(self, '_pipe', None) is None:
            k = self._pwc_ops
            self._pipe = dict(stream=torch.cuda.Stream(), done=None, free=None, pwc=self._sub_plan(0, k, 'fwd_pwc'),
                              rest=self._sub_plan(k + 1, len(self.fwd.ops), 'fwd_rest'))
        return self._pipe

    def prime_pipeline(self):
        """Run the frozen flow network (on the current stream) for the frame pair now in img1 / img2 so that the next
        train_step(pipeline=True) trains on it.  Needed before the first pipelined step and whenever the stream of batches restarts."""
        self._ensure_packed()
        pp = self._pipe_state()
        self.pipeline_drain()
        self._capture_plans('pipe_pwc', [pp['pwc']], lane_key=1).replay()
        pp['done'] = None
        pp['free'] = torch.cuda.Event()
        pp['free'].record(torch.cuda.current_stream())
        self._stage_valid = True

    def pipeline_inputs_free(self):
        """Event after which img1 / img2 may be overwritten with the next frame pair (the flow network last reading them is done)."""
        pp = self._pipe_state()
        return pp['done'] if pp['done'] is not None else pp['free']

    def _train_step_pipelined(self, mode, allreduce, inputs_ready):
        pp = self._pipe_state()
        main, side = torch.cuda.current_stream(), pp['stream']
        g_pwc = self._capture_plans('pipe_pwc', [pp['pwc']], lane_key=1)
        g_rest = self._capture_plans('pipe_rest_' + mode, [pp['rest'], self.bwd[mode]])
        g_adam = self._capture_plans('pipe_adam_' + mode, [self.adam[mode], self._pack_of(mode)], warm=False)
        if not getattr(self, '_stage_valid', False):
            # prime: the stage must hold the flow of the batch this call trains on (= what img1 / img2 hold right now)
            if inputs_ready is not None:
                main.wait_event(inputs_ready)
            self.prime_pipeline()
        if pp['done'] is not None:
            main.wait_event(pp['done'])              # the previous call's side branch filled the stage
        self._take_stage()
        taken = torch.cuda.Event()
        taken.record(main)
        side.wait_event(taken)                       # stage and level buffers are free again
        if inputs_ready is not None:
            side.wait_event(inputs_ready)
        with torch.cuda.stream(side):
            g_pwc.replay()                           # PWC-Net on the NEXT batch, concurrent with everything below
            pp['done'] = torch.cuda.Event()
            pp['done'].record(side)
        g_rest.replay()
        if allreduce is not None:
            allreduce((self.rec_store if mode == 'R' else self.gen_store).grad)
        g_adam.replay()

    def pipeline_drain(self):
        """Join the side branch (call before reading PWC-Net outputs or re-feeding img1 / img2 outside train_step)."""
        pp = getattr(self, '_pipe', None)
        if pp is not None and pp['done'] is not None:
            torch.cuda.current_stream().wait_event(pp['done'])

    def losses(self, full=False, reduce=None):
        """The `losses` dict of adversarial_learner.py:196-204 (device -> host read).  full=True adds the four first-sample
        diagnostics (:201-204) taken from the per-sample Charbonnier sums {rec, rec_c, prior, den, den_c}.  `reduce`: the SUM
        all-reduce of a data-parallel job -- the local scalars are this rank's share of the global-batch losses."""
        if reduce is not None:
            t = self.scalars[:4].clone()
            reduce(t)
            s = t.tolist()
        else:
            s = self.scalars.tolist()
        out = dict(generator=s[0], recover=s[1], red_rate=s[2], red_rate_compl=s[3])
        if full:
            r = self.sums[0].tolist()
            out.update(reconstruction_loss=r[0], reconstruction_compl_loss=r[1], denominator_red_rate=r[3] + self.eps_rr,
                       denominator_red_rate_compl=r[4] + self.eps_rr)
        return out
