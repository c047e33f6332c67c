"""Throughput of BASELINE.json configs 4 and 5 (segmentation 256^2 / 17 classes, detection 800^2 / 20
classes, N = 2501 tokens) on one B200: images/s of zero_grad -> forward -> loss -> backward, with the
CUDA-event kernel breakdown.  usage: python tools/bench_configs.py [segmentation|detection] [--batch B]
[--q-format F] [--arch small|tiny]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "myrtle-vision_b200")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
import mv_native  # noqa: E402
from myrtle_vision.datasets.synthetic import SyntheticVision, detection_collate  # noqa: E402
from myrtle_vision.models.vit import ViT  # noqa: E402
from myrtle_vision.utils.graph import GraphedTrainStep  # noqa: E402
from myrtle_vision.utils.trainer import build_criterion, to_device  # noqa: E402

ARCHS = {"small": dict(dim=384, depth=12, heads=6, mlp_dim=1536), "tiny": dict(dim=192, depth=12, heads=3, mlp_dim=768)}
ap = argparse.ArgumentParser()
ap.add_argument("task", choices=["classification", "segmentation", "detection"])
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--q-format", default="FP16_32")
ap.add_argument("--arch", default="small")
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--unfused-seg-loss", action="store_true", help="segmentation: upsample + CrossEntropyLoss in PyTorch")
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--profile", action="store_true", help="print the 30 largest CUDA kernels of three eager steps (torch profiler) and exit")
ap.add_argument("--host-matcher", action="store_true",
                help="detection: SciPy assignment on the host between two captured graphs (the round-1e path)")
args = ap.parse_args()
size, classes, batch = {"classification": (256, 45, 256), "segmentation": (256, 17, 256), "detection": (800, 20, 8)}[args.task]
batch = args.batch or batch
dev = torch.device("cuda")
torch.manual_seed(1234)
model = ViT(decoder=args.task, image_size=size, patch_size=16, num_classes=classes, q_format=args.q_format,
            **ARCHS[args.arch]).to(dev).train()
ds = SyntheticVision(args.task, 2 * batch, size, classes, seed=1234)
items = [ds[i] for i in range(2 * batch)]
if args.task == "detection":
    batches = [detection_collate(items[:batch]), detection_collate(items[batch:])]
else:
    batches = [(torch.stack([a for a, _ in part]), torch.stack([torch.as_tensor(b) for _, b in part])) for part in (items[:batch], items[batch:])]
if args.task == "detection" and not args.host_matcher:
    from myrtle_vision.models.matcher import pad_targets
    cap = -(-max(int(t["boxes"].shape[0]) for _, ts in batches for t in ts) // 16) * 16
    batches = [(img, pad_targets(ts, capacity=cap)) for img, ts in batches]      # fixed [B, cap] target block
batches = [(img.to(dev), to_device(t, dev)) for img, t in batches]
train_cfg = {"loss_ce": 1.0, "class_error": 0.0, "loss_bbox": 5.0, "loss_giou": 2.0, "cardinality_error": 0.0, "eos_coef": 0.1}
train_cfg["fused_seg_loss"] = not args.unfused_seg_loss
criterion = build_criterion(args.task, train_cfg, classes, dev)
if args.task == "segmentation" and not args.unfused_seg_loss:
    model.decoder.fused_loss = True


def eager(img, tgt):
    model.zero_grad(set_to_none=True)
    loss = criterion(model(img), tgt)
    loss.backward()
    return loss


for i in range(3):
    eager(*batches[i % 2])
torch.cuda.synchronize()
if args.profile:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for i in range(3):
            eager(*batches[i % 2])
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    print("total device time per step %.3f ms" % (tot / 3e3))
    for e in rows[:30]:
        print("%8.3f ms/step  n=%4d  %s" % (e.device_time_total / 3e3, e.count // 3, e.key[:110]))
    sys.exit(0)
mv_native.enable_timing(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(args.steps):
    eager(*batches[i % 2])
e1.record(); torch.cuda.synchronize()
ms_eager = e0.elapsed_time(e1) / args.steps
kern = {k: round(v[1] / args.steps, 3) for k, v in sorted(mv_native.timing_summary().items(), key=lambda kv: -kv[1][1])}
mv_native.enable_timing(False)
ms = ms_eager
graph = not args.no_graph
if graph and (args.task == "segmentation" or not args.host_matcher):
    gs = GraphedTrainStep(model, criterion, *batches[0])
    run = lambda img, tgt: gs(img, tgt)
elif graph:
    # detection: the Hungarian matcher runs on the host between forward and backward, so the model's
    # forward and backward are captured as two separate graphs around it
    gm = torch.cuda.make_graphed_callables(model, (batches[0][0],), allow_unused_input=True)

    def run(img, tgt):
        model.zero_grad(set_to_none=True)
        loss = criterion(gm(img), tgt)
        loss.backward()
        return loss
if graph:
    for i in range(3):
        run(*batches[i % 2])
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        run(*batches[i % 2])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
n_tok = (size // 16) ** 2 + 1
print(json.dumps({"task": args.task, "arch": args.arch, "q_format": args.q_format, "batch": batch, "tokens": n_tok,
                  "images_per_s": batch / ms * 1e3, "ms_per_step": ms, "eager_ms_per_step": ms_eager,
                  "cuda_graph": graph, "matcher": ("host" if args.host_matcher else "device") if args.task == "detection" else None,
                  "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
                  "top_kernels_ms_per_step": dict(list(kern.items())[:8])}))
