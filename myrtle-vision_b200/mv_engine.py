"""Fused sm_100a execution of the quantised ViT encoder (patch embedding + transformer blocks).

Host-side orchestration only: every arithmetic step is a kernel of libmv_b200.so called through
the C ABI (include/mv_b200.h).  Replaces the module-by-module execution of
src/myrtle_vision/models/vit.py:267-314 (ViT.forward up to the decoder) and its autograd
backward, with the fake-quantisers of SURVEY.md Appendix A folded into the neighbouring kernels:

  forward, per block (H = the input/weight format, o = Linear/LN output format, f = FloatFunctional
  format; o and f are identity except in FP16_16):
    xn1  = H(LN(H(x)))                                   mv_layernorm_q_fwd      -> fp16
    qkv  = o(xn1 . H(Wqkv)^T + b)                        mv_gemm                 -> fp16
    att  = H(softmax(q k^T / 8) v)                       mv_attention_fwd        -> fp16
    x1   = f(o(att . H(Wo)^T + b) + x)                   mv_gemm (+residual)     -> fp32
    xn2  = H(LN(H(x1)))                                  mv_layernorm_q_fwd      -> fp16
    u,h  = o(xn2 . H(W1)^T + b),  H(gelu(u))             mv_gemm (GELU epilogue) -> fp16, fp16
    x2   = f(o(h . H(W2)^T + b) + x1)                    mv_gemm (+residual)     -> fp32
  backward: straight-through quantisers (utils/quantize.py:87-89); dgrad / wgrad on the same GEMM
  kernel (wgrad reads dY and X in place as MN-major operands, split-K red.add into a flat fp32
  gradient buffer), GELU' in the dgrad epilogue, LN backward fused with the residual-gradient add,
  the fp16 operand copy and the bias-gradient column sums.

Gradients travel as fp16 tensor-core operands scaled by a power of two S chosen on the device from
the incoming gradient (the reference itself trains under GradScaler(65536), classification/
train.py:167); parameter gradients are un-scaled in fp32 before they are returned.
"""
import contextlib
import os

import torch

import mv_native as mv


class _Range:
    """`profile=True` of the reference's ViT (models/vit.py:115-124, 205-213: autograd-profiler record_function
    contexts): the same five labels, as a record_function (torch profiler) AND an NVTX range (nsys / ncu)."""

    def __init__(self, name):
        self.name = name
        self.rf = None

    def __enter__(self):
        torch.cuda.nvtx.range_push(self.name)
        self.rf = torch.autograd.profiler.record_function(self.name)
        self.rf.__enter__()
        return self

    def __exit__(self, *exc):
        self.rf.__exit__(*exc)
        torch.cuda.nvtx.range_pop()
        return False


class EngineConfig:
    def __init__(self, dim, heads, mlp_dim, depth, patch, plan):
        self.dim, self.heads, self.mlp_dim, self.depth, self.patch, self.plan = (
            dim, heads, mlp_dim, depth, patch, plan)


PER_LAYER = 12   # ln1.w ln1.b qkv.w qkv.b out.w out.b ln2.w ln2.b fc1.w fc1.b fc2.w fc2.b


class EncoderEngine:
    def __init__(self, cfg, params):
        """params: [pe.w, pe.b] + PER_LAYER tensors per block, all fp32 CUDA nn.Parameters."""
        self.cfg = cfg
        self.params = list(params)
        assert len(self.params) == 2 + PER_LAYER * cfg.depth
        # debugging aids of the reference's model, off the training path: profile ranges and, per block, the module
        # the reference passes the softmax probabilities through (models/vit.py:80-82, 94 — a hook point)
        self.profile = False
        self.probes = None
        plan = cfg.plan
        if plan.inp not in ((5, 10), (8, 10), None):
            raise NotImplementedError("no tensor-core operand container for input format %r" % (plan.inp,))
        # 32-bit formats (TF32: inputs/weights on the (8,10) grid; FP32: no quantisers) run their
        # contractions as kind::tf32 on fp32 containers — what the reference's torch 1.11 does with its
        # default allow_tf32 matmuls; (8,10) values are exact tf32 operands.
        self.wide = plan.inp != (5, 10)
        # FP32 (no quantisers at all): fp32-grade arithmetic.  Linears run as 3xTF32 — operands split into
        # tf32 hi + lo parts concatenated along K (mv_split_tf32), so the kind::tf32 kernel accumulates
        # hi*hi + lo*hi + hi*lo — and attention / GELU, which have no quantiser to fuse with in this format,
        # use torch's fp32 ops.  This is the reference's "plumbing" configuration, not a performance path.
        self.exact = plan.inp is None
        # gradient quantiser of the input / weight quantisers (QPyTorch backward_number; QuantPlan.grad)
        self.gfmt = getattr(plan, "grad", None)
        if self.gfmt is not None and self.wide:
            raise NotImplementedError("a gradient format needs the 16-bit operand path (q_format=FP16_32)")
        if cfg.dim != cfg.heads * 64:
            raise ValueError("dim must equal heads * 64 (dim_head is fixed at 64, models/vit.py:178)")
        self.fmt = plan.inp
        self.dev = self.params[0].device
        self._wq = None
        self._wq_versions = None
        self._wq_capture_done = None
        self.capture_generation = 0
        sizes = [p.numel() for p in self.params]
        # gflat[0:4] is a slot that travels through the last gradient all-reduce: under data parallelism it carries
        # this rank's overflow flag to every rank (no extra collective)
        self.offsets = [4]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + ((s + 3) // 4) * 4)
        self.gflat = torch.zeros(self.offsets[-1], dtype=torch.float32, device=self.dev)
        self.gviews = [self.gflat[self.offsets[i]:self.offsets[i] + sizes[i]].view_as(p)
                       for i, p in enumerate(self.params)]
        # Overflow handling in place of the reference's GradScaler (classification/train.py:167, 259-277), all on the
        # device: `overflow` is raised by the backward kernels (mv_set_overflow_flag), `scaler_state` =
        # {found_inf of the last backward, scale_target, good_steps}; an optimizer that understands found_inf
        # (utils.fused_adamw.FusedAdamW.step(found_inf=engine.found_inf)) skips the update, as GradScaler.step does
        self.overflow = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.scaler_state = torch.tensor([0.0, 1024.0, 0.0], dtype=torch.float32, device=self.dev)
        self.reducer = None      # set by parallel.DataParallel: called with flat gradient slices
        self.bucket_blocks = max(1, int(os.environ.get("MV_DP_BUCKET_BLOCKS", "3")))   # encoder blocks per all-reduce
        # The persistent kernels launched right behind a bucket's all-reduce are sized for the SMs NCCL leaves free
        # (mv_set_option "sm_limit" / "sm_limit_launches", csrc/common.cuh persistent_sms): on a full-size grid their
        # last CTAs wait for an SM and run as a second wave — the kernel takes twice as long.
        self.dp_sm_limit = int(os.environ.get("MV_DP_SM_LIMIT", "148"))      # off: measured no gain at N = 2 and N = 8
        self.dp_limit_launches = int(os.environ.get("MV_DP_LIMIT_LAUNCHES", "4"))
        # True once an optimizer (utils/fused_adamw.FusedAdamW) emits q(W) / q(W)^T itself after every
        # update: a CUDA-graph capture then leaves the re-quantisation out of the graph
        self.external_requant = False

    # ------------------------------------------------------------------ weights
    def _weight_indices(self):
        idx = [0]
        for l in range(self.cfg.depth):
            base = 2 + PER_LAYER * l
            idx += [base + 2, base + 4, base + 8, base + 10]
        return idx

    def quantised_weights(self):
        """weight_fake_quant of every Linear (torch.nn.qat.Linear.forward re-quantises each step):
        fp16 q(W) [out,in] for forward and q(W)^T [in,out] for dgrad; cached on parameter versions."""
        idx = self._weight_indices()
        versions = tuple(self.params[i]._version for i in idx) + tuple(
            self.params[i].data_ptr() for i in idx)
        capturing = torch.cuda.is_current_stream_capturing()
        if self._wq is not None and versions == self._wq_versions and (
                not capturing or self.external_requant):
            return self._wq
        if capturing and self._wq is not None and self._wq_capture_done == self._capture_key():
            return self._wq          # already re-quantised once inside this capture (forward)
        if self.exact:
            wq = {i: (mv.split_tf32(self.params[i].detach(), 1),
                      mv.split_tf32(self.params[i].detach().t().contiguous(), 1)) for i in idx}
        else:
            wq = self._requantise_all(idx)
        self._wq, self._wq_versions = wq, versions
        # inside a CUDA-graph capture the re-quantisation must be part of the graph (weights change
        # between replays) but only once per step: backward reuses the forward's operands
        self._wq_capture_done = self._capture_key() if capturing else None
        return wq

    def _capture_key(self):
        """Identifies the running capture: the capturing stream's handle and the capture generation
        (utils.graph.GraphedTrainStep bumps it around every capture, so a flag left by an earlier capture on the
        same stream never matches a later one)."""
        return (torch.cuda.current_stream().cuda_stream, self.capture_generation)

    def _requantise_all(self, idx):
        """q(W) and q(W)^T of every Linear in ONE launch: the multi-tensor optimizer kernel (mv_adamw_step)
        in its skip mode leaves parameters and moments untouched and only emits the operand tiles, so the
        per-step weight_fake_quant of 4 * depth + 1 Linears costs one kernel instead of one each."""
        import ctypes
        from myrtle_vision.utils.fused_adamw import AdamwTensor
        dt = torch.float32 if self.wide else torch.float16
        e, m = self.fmt if self.fmt else (0, 0)
        key = tuple(self.params[i].data_ptr() for i in idx)
        tab = getattr(self, "_rq_table", None)
        if tab is None or tab[0] != key:
            wq, rows, chunk = {}, [], 0
            for i in idx:
                w = self.params[i].detach()
                r, c = w.shape
                q = torch.empty(r, c, dtype=dt, device=self.dev)
                qt = torch.empty(c, r, dtype=dt, device=self.dev)
                wq[i] = (q, qt)
                t = AdamwTensor()
                t.param, t.grad, t.exp_avg, t.exp_avg_sq = w.data_ptr(), w.data_ptr(), w.data_ptr(), w.data_ptr()
                t.wq, t.wq_t, t.n, t.rows, t.cols = q.data_ptr(), qt.data_ptr(), w.numel(), r, c
                t.wq_dtype, t.q_exp, t.q_man, t.lr, t.weight_decay, t.chunk0 = mv._DT[dt], e, m, 0.0, 0.0, chunk
                chunk += ((r + 31) // 32) * ((c + 31) // 32)
                rows.append(t)
            arr = (AdamwTensor * len(rows))(*rows)
            table = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.dev)
            hyper = torch.tensor([0.9, 0.999, 1e-8, 1.0, 1.0, 1.0], dtype=torch.float32, device=self.dev)  # found_inf = 1: skip
            tab = (key, wq, table, hyper, len(rows), chunk)
            self._rq_table = tab
        _, wq, table, hyper, n, chunks = tab
        mv._check(mv.lib().mv_adamw_step(ctypes.c_void_p(table.data_ptr()), n, chunks, mv._ptr(hyper),
                                         mv._stream()), "mv_adamw_step (re-quantise)")
        return wq

    def invalidate_weights(self):
        """The parameters changed behind the version counters: rebuild the operands on the next use."""
        self._wq, self._wq_versions = None, None

    def mark_weights_fresh(self):
        """The operand buffers were just rewritten from the current parameter values (fused optimizer)."""
        idx = self._weight_indices()
        self._wq_versions = tuple(self.params[i]._version for i in idx) + tuple(
            self.params[i].data_ptr() for i in idx)

    @property
    def found_inf(self):
        """fp32 [1] on the device: 1.0 when the last backward overflowed (identical on every data-parallel rank)."""
        return self.scaler_state[0:1]

    def _gradient_scale(self, gx):
        """Power-of-two operand scale S chosen on the device: max |gx| * S lands in [target / 2, target]."""
        mv.set_overflow_flag(self.overflow)        # (cleared by the training forward this backward belongs to)
        if self.gfmt is not None:
            # the gradient quantisers must see the gradient as autograd presents it (what QPyTorch's would see under
            # the reference's GradScaler): no rescaling here, loss scaling is the caller's
            return self._one()
        amax = torch.linalg.vector_norm(gx, float("inf")).clamp_min(1e-30)          # one reduction, no |gx| temporary
        return torch.exp2(torch.floor(torch.log2(self.scaler_state[1] / amax))).clamp(2.0 ** -60, 2.0 ** 60)

    # ------------------------------------------------------------------ forward
    def _range(self, name):
        return _Range(name) if self.profile else contextlib.nullcontext()

    def _probe_attention(self, l, qkv, B, H, N):
        """Debug only: if forward hooks are registered on block l's `attn_output` module, hand it the softmax
        probabilities [B, H, N, N] (computed unfused from the same qkv; the fused kernels never materialise them).
        The training arithmetic is not touched; Python hooks do not replay inside a captured CUDA graph."""
        m = None if self.probes is None else self.probes[l]
        if m is None or not (m._forward_hooks or m._forward_pre_hooks):
            return
        q, k = qkv.view(B, N, 3, H, 64).float().permute(2, 0, 3, 1, 4)[:2]
        m(torch.softmax(torch.matmul(q, k.transpose(-2, -1)) * 0.125, dim=-1))

    def forward(self, img, pos_full, cls_token, save):
        # a training forward opens an overflow window that its backward closes (mv_overflow_update); inference
        # forwards report nothing
        mv.set_overflow_flag(self.overflow if save else None)
        if save:
            self.overflow.zero_()
        if self.wide:
            return self._forward_wide(img, pos_full, cls_token, save)
        cfg, plan, fmt = self.cfg, self.cfg.plan, self.fmt
        B, C, Hh, Ww = img.shape
        P = cfg.patch
        N = (Hh // P) * (Ww // P) + 1
        M, D, Mm, H = B * N, cfg.dim, cfg.mlp_dim, cfg.heads
        prm = [p.detach() for p in self.params]
        wq = self.quantised_weights()
        f16 = torch.float16
        dev = img.device

        with self._range("patch_to_embedding"):
            patches = mv.patchify_q(img, P, q_in=fmt, cls_slot=True)            # [M, P*P*C], zero cls rows
            x = torch.empty(M, D, dtype=torch.float32, device=dev)
            pos32 = pos_full.detach().reshape(N, D).contiguous()
            mv.gemm(patches, wq[0][0], x, bias=prm[1], q_out=plan.out, residual=pos32, q_res=plan.ff,
                    rows_per_img=N)
            mv.cls_rows(cls_token.detach().reshape(D).contiguous(), pos32, x, B, N, D, q_ff=plan.ff)

        saved = {"B": B, "N": N, "patches": patches, "layers": []} if save else None
        with self._range("transformer"):
            for l in range(cfg.depth):
                b0 = 2 + PER_LAYER * l
                with self._range("transformer:attention"):
                    xn1, mean1, rstd1 = mv.layernorm_q_fwd(x, prm[b0], prm[b0 + 1], q_in=fmt, q_post=fmt)
                    qkv = torch.empty(M, 3 * D, dtype=f16, device=dev)
                    mv.gemm(xn1, wq[b0 + 2][0], qkv, bias=prm[b0 + 3], q_out=plan.out)
                    self._probe_attention(l, qkv, B, H, N)
                    att, lse = mv.attention_fwd(qkv, B, H, N, scale=0.125, q_out=fmt)
                    x1 = torch.empty(M, D, dtype=torch.float32, device=dev)
                    mv.gemm(att, wq[b0 + 4][0], x1, bias=prm[b0 + 5], q_out=plan.out, residual=x,
                            q_res=plan.ff)
                with self._range("transformer:feedforward"):
                    xn2, mean2, rstd2 = mv.layernorm_q_fwd(x1, prm[b0 + 6], prm[b0 + 7], q_in=fmt, q_post=fmt)
                    u = torch.empty(M, Mm, dtype=f16, device=dev)
                    h = torch.empty(M, Mm, dtype=f16, device=dev)
                    mv.gemm(xn2, wq[b0 + 8][0], h, bias=prm[b0 + 9], q_out=plan.out, aux=u,
                            epilogue=mv.EPI_GELU, q_res=fmt)
                    x2 = torch.empty(M, D, dtype=torch.float32, device=dev)
                    mv.gemm(h, wq[b0 + 10][0], x2, bias=prm[b0 + 11], q_out=plan.out, residual=x1,
                            q_res=plan.ff)
                if save:
                    saved["layers"].append((x, xn1, mean1, rstd1, qkv, att, lse, x1, xn2, mean2, rstd2, u, h))
                x = x2
        return x.view(B, N, D), saved

    # ----------------------------------------------------------------- backward
    def backward(self, saved, gx):
        if self.wide:
            return self._backward_wide(saved, gx)
        cfg, fmt = self.cfg, self.fmt
        B, N = saved["B"], saved["N"]
        M, D, Mm, H = B * N, cfg.dim, cfg.mlp_dim, cfg.heads
        prm = [p.detach() for p in self.params]
        wq = self.quantised_weights()
        g = self.gviews
        dev = gx.device
        f16 = torch.float16

        # power-of-two gradient scale chosen on the device (no host sync)
        gx = gx.reshape(M, D)
        S = self._gradient_scale(gx)
        dx, dx_h = mv.scale_f32(gx, S, want_f16=True)                             # fp32 master and fp16 operand in one pass
        self.gflat.zero_()
        last = cfg.depth - 1
        mv.colsum(dx, g[2 + PER_LAYER * last + 11].view(-1))               # fc2 bias of the last block

        # Gradient quantiser G (QuantPlan.grad, normally None): the backward of every QuantStub in front of a Linear is
        # the q_out of that Linear's dgrad GEMM (before the gelu' multiply for fc2's), of every stub in front of a
        # LayerNorm a rounding inside mv_layernorm_q_bwd (before the residual gradient is added), of every weight
        # quantiser a pass over the accumulated weight gradient (before the data-parallel reduction, as autograd
        # would apply it to the local gradient)
        gq = self.gfmt
        mv.set_grad_format(gq)

        def wq_grad(i):
            if gq is not None:
                mv.float_quantize(g[i], gq[0], gq[1], out=g[i])

        for l in range(last, -1, -1):
            b0 = 2 + PER_LAYER * l
            x, xn1, mean1, rstd1, qkv, att, lse, x1, xn2, mean2, rstd2, u, h = saved["layers"][l]
            # ---- FeedForward
            du = torch.empty(M, Mm, dtype=f16, device=dev)
            # fc1's bias gradient (column sums of du) rides in the epilogue
            mv.gemm(dx_h, wq[b0 + 10][1], du, aux=u, epilogue=mv.EPI_DGELU, colsum=g[b0 + 9].view(-1), q_out=gq)
            # fc2's [D, 4D] gradient as its transpose h^T dx: 256 x 384 tiles over 4D rows instead of 1.5 padded 256-row tiles
            mv.gemm(h, dx_h, g[b0 + 10], a_major=1, b_major=1, accumulate=True, transpose_out=True)
            wq_grad(b0 + 10)
            dxn2 = torch.empty(M, D, dtype=f16, device=dev)
            mv.gemm(du, wq[b0 + 8][1], dxn2, tag="dgrad", q_out=gq)
            mv.gemm(du, xn2, g[b0 + 8], a_major=1, b_major=1, accumulate=True)
            wq_grad(b0 + 8)
            del du
            dx1, dx1_h = mv.layernorm_q_bwd(dxn2, x1, prm[b0 + 6], mean2, rstd2, dres=dx, q_in=fmt,
                                            dgamma=g[b0 + 6], dbeta=g[b0 + 7], dbias_prev=g[b0 + 5])
            # ---- Attention
            datt = torch.empty(M, D, dtype=f16, device=dev)
            mv.gemm(dx1_h, wq[b0 + 4][1], datt, tag="dgrad", q_out=gq)
            mv.gemm(dx1_h, att, g[b0 + 4], a_major=1, b_major=1, accumulate=True)
            wq_grad(b0 + 4)
            dqkv = mv.attention_bwd(qkv, att, datt, lse, B, H, N, scale=0.125,
                                    dbias=g[b0 + 3].view(-1))          # to_qkv's bias gradient fused
            dxn1 = dxn2      # reuse
            mv.gemm(dqkv, wq[b0 + 2][1], dxn1, tag="dgrad", q_out=gq)
            mv.gemm(dqkv, xn1, g[b0 + 2], a_major=1, b_major=1, accumulate=True)
            wq_grad(b0 + 2)
            prev_bias = g[b0 - 1] if l > 0 else None                       # fc2 bias of block l-1
            dx, dx_h = mv.layernorm_q_bwd(dxn1, x, prm[b0], mean1, rstd1, dres=dx1, q_in=fmt,
                                          dgamma=g[b0], dbeta=g[b0 + 1], dbias_prev=prev_bias,
                                          dx=dx, dx_f16=dx_h)
            if self.reducer is not None and l % self.bucket_blocks == 0:
                # one all-reduce per bucket of `bucket_blocks` encoder blocks (their gradient slices are contiguous),
                # issued as soon as the bucket's last wgrad is enqueued
                top = min(cfg.depth, l + self.bucket_blocks)
                lo, hi = self.offsets[b0], self.offsets[2 + PER_LAYER * top]
                self.reducer.reduce_slice(self.gflat[lo:hi], 1.0 / S)
                if self.reducer.world > 1 and self.dp_sm_limit < 148 and l > 0:
                    mv.set_option("sm_limit", self.dp_sm_limit)
                    mv.set_option("sm_limit_launches", self.dp_limit_launches)
        # ---- patch embedding: dW = dx^T patches (cls rows of `patches` are zero)
        mv.gemm(dx_h, saved["patches"], g[0], a_major=1, b_major=1, accumulate=True)
        wq_grad(0)
        mv.set_grad_format(None)
        return self._finish_backward(dx, S, B, N, D)

    def _finish_backward(self, dx, S, B, N, D):
        g, dev = self.gviews, dx.device
        dpos = torch.zeros(N * D, dtype=torch.float32, device=dev)
        mv.colsum(dx.view(B, N * D), dpos)
        dpos = dpos.view(1, N, D)
        g[1].copy_(dpos[0, 1:].sum(0))                                     # bias: patch rows only
        inv = 1.0 / S
        slot = None
        if self.reducer is not None:
            # this rank's overflow flag rides in the slot in front of the last bucket (any positive value survives
            # the 1 / (S * world) factor: S is clamped to 2^+-60)
            self.gflat[0:1].copy_(self.overflow)
            self.reducer.reduce_slice(self.gflat[:self.offsets[2]], inv)
            self.reducer.wait()
            slot = self.gflat[0:1]
        dpos = dpos * inv
        dcls = dpos[:, 0:1, :].clone()
        # hand autograd its own copy: self.gflat is reused (zeroed) by the next backward.  The same pass un-scales
        # (single GPU) and tests every parameter gradient for inf / NaN (the overflow sink)
        if self.reducer is not None:
            out, _ = mv.scale_f32(self.gflat, self._one(), invert=False)
        else:
            out, _ = mv.scale_f32(self.gflat, S, invert=True)
        mv.overflow_update(self.overflow, self.scaler_state, shared_slot=slot)
        grads = [out[self.offsets[i]:self.offsets[i] + p.numel()].view_as(p)
                 for i, p in enumerate(self.params)]
        return dpos, dcls, grads

    def _one(self):
        if getattr(self, "_one_t", None) is None:
            self._one_t = torch.ones(1, dtype=torch.float32, device=self.dev)
        return self._one_t


    # ------------------------------------------------- 32-bit formats (TF32 / FP32)
    # Same step as above on fp32 containers.  Every Linear is a kind::tf32 GEMM on K-major operands;
    # wgrad therefore takes explicit transposes (mv_widen_transpose) instead of the in-place MN-major
    # reads of the 16-bit path.  Attention keeps the fp16 tcgen05 kernels: qkv / attention output /
    # their gradients carry 11 significant bits, the same as a tf32 operand.
    def _forward_wide(self, img, pos_full, cls_token, save):
        cfg, plan, fmt = self.cfg, self.cfg.plan, self.fmt
        B, C, Hh, Ww = img.shape
        P = cfg.patch
        N = (Hh // P) * (Ww // P) + 1
        M, D, Mm, H = B * N, cfg.dim, cfg.mlp_dim, cfg.heads
        prm = [p.detach() for p in self.params]
        wq = self.quantised_weights()
        f16, f32 = torch.float16, torch.float32
        dev = img.device
        ex = self.exact
        A_ = (lambda t: mv.split_tf32(t, 0)) if ex else (lambda t: t)        # A operand of a Linear

        patches = mv.patchify_q(img, P, q_in=fmt, out_dtype=f32, cls_slot=True)
        x = torch.empty(M, D, dtype=f32, device=dev)
        pos32 = pos_full.detach().reshape(N, D).contiguous()
        mv.gemm(A_(patches), wq[0][0], x, bias=prm[1], q_out=plan.out, residual=pos32, q_res=plan.ff,
                rows_per_img=N)
        mv.cls_rows(cls_token.detach().reshape(D).contiguous(), pos32, x, B, N, D, q_ff=plan.ff)

        saved = {"B": B, "N": N, "patches": patches, "layers": []} if save else None
        for l in range(cfg.depth):
            b0 = 2 + PER_LAYER * l
            xn1, mean1, rstd1 = mv.layernorm_q_fwd(x, prm[b0], prm[b0 + 1], q_in=fmt, q_post=fmt,
                                                   out_dtype=f32)
            qkv = torch.empty(M, 3 * D, dtype=f32 if ex else f16, device=dev)
            mv.gemm(A_(xn1), wq[b0 + 2][0], qkv, bias=prm[b0 + 3], q_out=plan.out)
            self._probe_attention(l, qkv, B, H, N)
            if ex:
                att, lse = self._attention_exact_fwd(qkv, B, H, N)           # lse slot holds the probabilities
                att16 = att
            else:
                att16, lse = mv.attention_fwd(qkv, B, H, N, scale=0.125, q_out=fmt)
                att, _ = mv.widen_transpose(att16, want_copy=True, want_t=False)
            x1 = torch.empty(M, D, dtype=f32, device=dev)
            mv.gemm(A_(att), wq[b0 + 4][0], x1, bias=prm[b0 + 5], q_out=plan.out, residual=x, q_res=plan.ff)
            xn2, mean2, rstd2 = mv.layernorm_q_fwd(x1, prm[b0 + 6], prm[b0 + 7], q_in=fmt, q_post=fmt,
                                                   out_dtype=f32)
            h = torch.empty(M, Mm, dtype=f32, device=dev)
            if ex:
                gd = torch.empty(M, Mm, dtype=f32, device=dev)               # u (pre-GELU), fp32
                mv.gemm(A_(xn2), wq[b0 + 8][0], gd, bias=prm[b0 + 9])
                h = torch.nn.functional.gelu(gd)
            else:
                gd = torch.empty(M, Mm, dtype=f16, device=dev)               # gelu'(u)
                mv.gemm(xn2, wq[b0 + 8][0], h, bias=prm[b0 + 9], q_out=plan.out, aux=gd,
                        epilogue=mv.EPI_GELU, q_res=fmt)
            x2 = torch.empty(M, D, dtype=f32, device=dev)
            mv.gemm(A_(h), wq[b0 + 10][0], x2, bias=prm[b0 + 11], q_out=plan.out, residual=x1, q_res=plan.ff)
            if save:
                saved["layers"].append((x, xn1, mean1, rstd1, qkv, att16, lse, x1, xn2, mean2, rstd2, gd, h))
            x = x2
        return x.view(B, N, D), saved

    @staticmethod
    def _attention_exact_fwd(qkv, B, H, N):
        q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        prob = torch.softmax(torch.matmul(q, k.transpose(-2, -1)) * 0.125, dim=-1)
        return torch.matmul(prob, v).transpose(1, 2).reshape(B * N, H * 64), prob

    @staticmethod
    def _attention_exact_bwd(qkv, prob, d_att, B, H, N):
        q, k, v = qkv.view(B, N, 3, H, 64).permute(2, 0, 3, 1, 4)
        do = d_att.view(B, N, H, 64).transpose(1, 2)
        dv = torch.matmul(prob.transpose(-2, -1), do)
        dp = torch.matmul(do, v.transpose(-2, -1))
        ds = prob * (dp - (dp * prob).sum(-1, keepdim=True)) * 0.125
        dq, dk = torch.matmul(ds, k), torch.matmul(ds.transpose(-2, -1), q)
        return torch.stack((dq, dk, dv), dim=0).permute(1, 3, 0, 2, 4).reshape(B * N, 3 * H * 64)

    def _backward_wide(self, saved, gx):
        cfg, fmt = self.cfg, self.fmt
        B, N = saved["B"], saved["N"]
        M, D, Mm, H = B * N, cfg.dim, cfg.mlp_dim, cfg.heads
        prm = [p.detach() for p in self.params]
        wq = self.quantised_weights()
        g = self.gviews
        dev = gx.device
        f16, f32 = torch.float16, torch.float32

        ex = self.exact
        A_ = (lambda t: mv.split_tf32(t, 0)) if ex else (lambda t: t)

        def wgrad(dy, act, out):                       # out[o, i] += sum_m dy[m, o] act[m, i]
            _, dy_t = mv.widen_transpose(dy, padded=ex)
            _, act_t = mv.widen_transpose(act, padded=ex)
            if ex:
                dy_t, act_t = mv.split_tf32(dy_t, 0), mv.split_tf32(act_t, 1)
            mv.gemm(dy_t, act_t, out, accumulate=True)

        # the fp16 attention backward wants its gradient operand inside fp16's range: same
        # power-of-two scale as the 16-bit path, removed from the parameter gradients at the end
        gx = gx.reshape(M, D)
        S = self._gradient_scale(gx)
        dx, _ = mv.scale_f32(gx, S)
        self.gflat.zero_()
        last = cfg.depth - 1
        mv.colsum(dx, g[2 + PER_LAYER * last + 11].view(-1))

        for l in range(last, -1, -1):
            b0 = 2 + PER_LAYER * l
            x, xn1, mean1, rstd1, qkv, att16, lse, x1, xn2, mean2, rstd2, gd, h = saved["layers"][l]
            # ---- FeedForward
            du = torch.empty(M, Mm, dtype=f32, device=dev)
            if ex:
                mv.gemm(A_(dx), wq[b0 + 10][1], du)
                u = gd                                                       # gelu'(u) = Phi(u) + u phi(u)
                du.mul_(0.5 * (1.0 + torch.erf(u * 0.7071067811865476))
                        + u * torch.exp(-0.5 * u * u) * 0.3989422804014327)
            else:
                mv.gemm(dx, wq[b0 + 10][1], du, aux=gd, epilogue=mv.EPI_DGELU)
            wgrad(dx, h, g[b0 + 10])
            dxn2 = torch.empty(M, D, dtype=f32, device=dev)
            mv.gemm(A_(du), wq[b0 + 8][1], dxn2, tag="dgrad")
            wgrad(du, xn2, g[b0 + 8])
            mv.colsum(du, g[b0 + 9].view(-1))
            del du
            dx1, _ = mv.layernorm_q_bwd(dxn2, x1, prm[b0 + 6], mean2, rstd2, dres=dx, q_in=fmt,
                                        dgamma=g[b0 + 6], dbeta=g[b0 + 7], dbias_prev=g[b0 + 5],
                                        want_f16=False)
            # ---- Attention
            datt = torch.empty(M, D, dtype=f32 if ex else f16, device=dev)
            mv.gemm(A_(dx1), wq[b0 + 4][1], datt, tag="dgrad")
            wgrad(dx1, att16, g[b0 + 4])
            if ex:
                dqkv16 = dqkv = self._attention_exact_bwd(qkv, lse, datt, B, H, N)
                _, dqkv_t = mv.widen_transpose(dqkv, padded=True)
            else:
                dqkv16 = mv.attention_bwd(qkv, att16, datt, lse, B, H, N, scale=0.125, dbias=g[b0 + 3].view(-1))
                dqkv, dqkv_t = mv.widen_transpose(dqkv16, want_copy=True)
            dxn1 = dxn2
            mv.gemm(A_(dqkv), wq[b0 + 2][1], dxn1, tag="dgrad")
            _, xn1_t = mv.widen_transpose(xn1, padded=ex)
            if ex:
                dqkv_t, xn1_t = mv.split_tf32(dqkv_t, 0), mv.split_tf32(xn1_t, 1)
            mv.gemm(dqkv_t, xn1_t, g[b0 + 2], accumulate=True)
            if ex:
                mv.colsum(dqkv16, g[b0 + 3].view(-1))
            prev_bias = g[b0 - 1] if l > 0 else None
            dx, _ = mv.layernorm_q_bwd(dxn1, x, prm[b0], mean1, rstd1, dres=dx1, q_in=fmt,
                                       dgamma=g[b0], dbeta=g[b0 + 1], dbias_prev=prev_bias,
                                       want_f16=False, dx=dx)
            if self.reducer is not None and l % self.bucket_blocks == 0:
                # one all-reduce per bucket of `bucket_blocks` encoder blocks (their gradient slices are contiguous),
                # issued as soon as the bucket's last wgrad is enqueued
                top = min(cfg.depth, l + self.bucket_blocks)
                lo, hi = self.offsets[b0], self.offsets[2 + PER_LAYER * top]
                self.reducer.reduce_slice(self.gflat[lo:hi], 1.0 / S)
        wgrad(dx, saved["patches"], g[0])
        return self._finish_backward(dx, S, B, N, D)


class EncoderFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, img, pos_full, cls_token, *params):
        need = any(ctx.needs_input_grad)      # False under torch.no_grad() / eval without grads
        with torch.no_grad():
            x, saved = engine.forward(img, pos_full, cls_token, save=need)
        ctx.engine, ctx.saved = engine, saved
        return x

    @staticmethod
    def backward(ctx, gx):
        engine, saved = ctx.engine, ctx.saved
        if saved is None:
            raise RuntimeError("backward through an encoder forward that ran without grad")
        if engine.reducer is not None:
            engine.reducer.queue_finalize()
        with torch.no_grad():
            dpos, dcls, g = engine.backward(saved, gx.contiguous())
        ctx.saved = None
        return (None, None, dpos, dcls) + tuple(g)
