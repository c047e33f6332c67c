"""Data-parallel gradient reduction for the fused ViT (one process per GPU, NCCL over NVLink).

Replaces torch.nn.parallel.DistributedDataParallel as used by the reference
(classification/train.py:155-158): gradients are averaged over ranks, the reduction of each
transformer block's gradient bucket is issued on a communication stream as soon as that block's
wgrad kernels have been enqueued (so NCCL overlaps the rest of backward), and parameters that
received no gradient (det_tokens, pos_embedding_det — SURVEY.md fact 7, which makes stock DDP
raise on the second iteration) are simply skipped.
"""
import torch
import torch.distributed as dist
from torch import nn


class GradReducer:
    def __init__(self, process_group=None):
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.stream = None
        self.rest_params = []
        self._queued = False
        self.buckets_reduced = 0

    def _comm_stream(self, device):
        if device.type != "cuda":
            return None
        if self.stream is None:
            self.stream = torch.cuda.Stream(device=device)
        return self.stream

    def reduce_slice(self, flat, inv_scale=1.0):
        """flat <- mean over ranks of (flat * inv_scale); asynchronous on the comm stream."""
        if flat.is_cuda and torch.is_tensor(inv_scale):
            # one pass: un-scale, divide by the world size and test for inf / NaN (the library's overflow sink)
            import mv_native as mv
            mv.scale_f32(flat, (inv_scale / self.world).reshape(1).float(), out=flat)
        else:
            flat.mul_(inv_scale / self.world)
        self.buckets_reduced += 1
        if self.world == 1:
            return
        cs = self._comm_stream(flat.device)
        if cs is None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return
        cs.wait_stream(torch.cuda.current_stream(flat.device))
        with torch.cuda.stream(cs):
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        flat.record_stream(cs)

    def wait(self):
        if self.stream is not None:
            torch.cuda.current_stream().wait_stream(self.stream)

    def reduce_params(self, params):
        """Average the .grad of the given parameters, skipping those without a gradient."""
        grads = [p.grad for p in params if p.grad is not None]
        if not grads or self.world == 1:
            return len(grads)
        flat = torch.cat([g.reshape(-1) for g in grads])
        flat.div_(self.world)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        return len(grads)

    def queue_finalize(self):
        """Called from inside backward: reduce the non-engine parameters once backward ends."""
        if self._queued:
            return
        self._queued = True

        def _cb():
            self._queued = False
            self.reduce_params(self.rest_params)

        torch.autograd.Variable._execution_engine.queue_callback(_cb)


class DataParallel(nn.Module):
    """model = DataParallel(vit)  — then use it exactly like the reference uses DDP."""

    def __init__(self, module, process_group=None, broadcast=True):
        super().__init__()
        self.module = module
        self.reducer = GradReducer(process_group)
        if broadcast and self.reducer.world > 1:
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        self._attach()

    def _attach(self):
        engine = self.module.engine()
        engine.reducer = self.reducer
        owned = {id(p) for p in engine.params}
        self.reducer.rest_params = [p for p in self.module.parameters() if id(p) not in owned]

    def forward(self, *args, **kwargs):
        if self.module._engine is None or self.module._engine.reducer is not self.reducer:
            self._attach()
        return self.module(*args, **kwargs)
