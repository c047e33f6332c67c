"""CUDA-graph capture of the training step (fixed shapes).

The fused step launches ~280 small kernels; enqueueing them from Python costs about as much host
time as the GPU needs to run them.  `GraphedTrainStep` captures zero_grad -> forward -> loss ->
backward (including the NCCL bucket all-reduces of `DataParallel`) once and replays it, so a
step costs one graph launch.  Gradients live in the graph's static memory and are re-attached
to the parameters after every replay; the optimizer step stays outside (e.g. timm AdamW as in the
reference's classification/train.py:161-166, 274-277).
"""
import torch


class GraphedTrainStep:
    def __init__(self, net, loss_fn, img, target, warmup=3):
        """net: ViT or DataParallel(ViT); img/target: example batch on the device (shapes are
        frozen).  loss_fn(output, target) -> scalar."""
        self.net, self.loss_fn = net, loss_fn
        self.model = getattr(net, "module", net)
        self.static_img = img.clone()
        self.static_target = ({k: v.clone() for k, v in target.items()} if isinstance(target, dict)
                              else target.clone())
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.model.zero_grad(set_to_none=True)
                self.loss_fn(self.net(self.static_img), self.static_target).backward()
        torch.cuda.current_stream().wait_stream(side)
        self.model.zero_grad(set_to_none=True)
        self.graph = torch.cuda.CUDAGraph()
        engine = self.model.engine() if hasattr(self.model, "engine") else None
        if engine is not None:
            engine.capture_generation += 1       # operands re-quantised by an earlier capture do not count for this one
        with torch.cuda.graph(self.graph):
            self.output = self.net(self.static_img)
            self.loss = self.loss_fn(self.output, self.static_target)
            self.loss.backward()
        if engine is not None:
            engine.capture_generation += 1
        self.params = [p for p in self.model.parameters() if p.grad is not None]
        self.grads = [p.grad for p in self.params]

    def __call__(self, img=None, target=None):
        """Copy the batch into the static buffers (skip with None), replay, return the loss tensor."""
        if img is not None:
            self.static_img.copy_(img, non_blocking=True)
        if target is not None:
            if isinstance(target, dict):
                for k, v in target.items():
                    self.static_target[k].copy_(v, non_blocking=True)
            else:
                self.static_target.copy_(target, non_blocking=True)
        self.graph.replay()
        for p, g in zip(self.params, self.grads):
            if p.grad is not g:
                p.grad = g
        return self.loss
