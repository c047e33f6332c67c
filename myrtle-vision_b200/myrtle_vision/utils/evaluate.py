"""Evaluation entry points: `test_deit(config)` of the reference's classification/test.py:16-79,
segmentation/test.py:18-86 and detection/test.py:18-70 behind one function (the three CLIs in
{classification,segmentation,detection}/test.py call it with their task).

Same flow as the reference: drop-out off, `get_models`, `prepare_model_and_load_ckpt` (a checkpoint path is
required), eval-mode forward over the test split on the fused sm_100a path, then the task's report —
sklearn's classification report, mIoU with the per-class table, or box AP.  The reference's detection CLI
hands `PostProcess` results to pycocotools' COCOeval; that package is a dataset-side dependency outside the
hot path, so the same `PostProcess` results go to `box_average_precision` below (COCO's definition: AP
averaged over IoU 0.50:0.05:0.95 with 101-point interpolation, per class, then the mean).
Data: the synthetic sets or `data_config["dataset_factory"]`, as in utils/trainer.py."""
import warnings

import numpy as np
import torch
from torch.utils.data import DataLoader

from myrtle_vision.models.matcher import cxcywh_to_xyxy
from myrtle_vision.utils.models import get_models, prepare_model_and_load_ckpt
from myrtle_vision.utils.trainer import build_datasets, to_device
from myrtle_vision.utils.utils import parse_config


def pairwise_iou(a, b):
    """IoU of every xyxy box in a [n, 4] with every box in b [m, 4] -> [n, m]."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.maximum(a[:, None, :2], b[None, :, :2])
    rb = torch.minimum(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area_a[:, None] + area_b[None, :] - inter).clamp(min=1e-12)


def box_average_precision(detections, ground_truth, num_classes, iou_thresholds=None, max_dets=100):
    """COCO-style box AP.  detections: per image {'scores' [q], 'labels' [q], 'boxes' [q, 4] xyxy};
    ground_truth: per image {'labels' [k], 'boxes' [k, 4] xyxy} (same coordinate frame).
    -> {'AP': mean over thresholds and classes, 'AP50': at IoU 0.5, 'per_class': AP per class (nan if absent)}."""
    thr = np.arange(0.5, 0.96, 0.05) if iou_thresholds is None else np.asarray(iou_thresholds, dtype=np.float64)
    recall_grid = np.linspace(0.0, 1.0, 101)
    ap = np.full((len(thr), num_classes), np.nan)
    for c in range(num_classes):
        scores, hits, n_gt = [], [], 0
        for det, gt in zip(detections, ground_truth):
            g = gt["boxes"][gt["labels"] == c].float().cpu()
            n_gt += len(g)
            keep = det["labels"].cpu() == c
            s, b = det["scores"].cpu()[keep], det["boxes"].float().cpu()[keep]
            order = torch.argsort(s, descending=True)[:max_dets]
            s, b = s[order], b[order]
            hit = np.zeros((len(thr), len(s)), dtype=bool)
            if len(g) and len(s):
                iou = pairwise_iou(b, g).numpy()
                for t, th in enumerate(thr):
                    taken = np.zeros(len(g), dtype=bool)
                    for d in range(len(s)):             # greedy, best remaining ground truth first (COCOeval)
                        cand = np.where(~taken, iou[d], -1.0)
                        j = int(cand.argmax())
                        if cand[j] >= th:
                            taken[j] = True
                            hit[t, d] = True
            scores.append(s.numpy())
            hits.append(hit)
        if n_gt == 0:
            continue
        scores = np.concatenate(scores) if scores else np.zeros(0)
        hits = np.concatenate(hits, axis=1) if hits else np.zeros((len(thr), 0), dtype=bool)
        order = np.argsort(-scores, kind="mergesort")
        for t in range(len(thr)):
            tp = np.cumsum(hits[t, order])
            fp = np.cumsum(~hits[t, order])
            recall = tp / n_gt
            precision = tp / np.maximum(tp + fp, 1)
            for i in range(len(precision) - 1, 0, -1):   # precision envelope
                precision[i - 1] = max(precision[i - 1], precision[i])
            idx = np.searchsorted(recall, recall_grid, side="left")
            ap[t, c] = np.where(idx < len(precision), precision[np.minimum(idx, len(precision) - 1)], 0.0).mean() \
                if len(precision) else 0.0
    with np.errstate(invalid="ignore"), warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)            # classes without ground truth stay nan
        per_class = np.nanmean(ap, axis=0) if np.isfinite(ap).any() else np.full(num_classes, np.nan)
        i50 = int(np.argmin(np.abs(thr - 0.5)))
        return {"AP": float(np.nanmean(ap)) if np.isfinite(ap).any() else float("nan"),
                "AP50": float(np.nanmean(ap[i50])) if np.isfinite(ap[i50]).any() else float("nan"),
                "per_class": per_class}


@torch.no_grad()
def test_deit(config, task=None):
    train_config, vit_config = config["train_config"], config["vit_config"]
    task = task or vit_config["decoder"]
    data_config = config.get("data_config") or parse_config(config["data_config_path"])
    config["data_config"] = data_config
    num_classes = data_config["number_of_classes"]
    device = torch.device("cuda")
    _, testset, collate = build_datasets(task, data_config, vit_config)
    loader = DataLoader(testset, num_workers=0, batch_size=train_config["local_batch_size"], pin_memory=False,
                        drop_last=train_config["drop_last_batch"], collate_fn=collate)
    vit_config["dropout"] = 0.0
    vit_config["emb_dropout"] = 0.0
    vit, _ = get_models(config)
    vit = vit.to(device)
    assert train_config["checkpoint_path"] != "", "Must provide a checkpoint path in the config file"
    prepare_model_and_load_ckpt(train_config=train_config, model=vit)
    vit.eval()
    if task == "classification":
        truth, pred = [], []
        for imgs, labels in loader:
            pred.extend(vit(imgs.to(device)).argmax(dim=1).cpu().numpy())
            truth.extend(np.asarray(labels))
        accuracy = float(np.mean(np.asarray(truth) == np.asarray(pred))) if truth else float("nan")
        try:
            from sklearn.metrics import classification_report
            print(classification_report(truth, pred, labels=np.arange(num_classes), zero_division=0))
        except ImportError:
            print(f"accuracy: {accuracy:.4f}")
        return {"accuracy": accuracy}
    if task == "segmentation":
        from myrtle_vision.utils.miou import MIoU
        miou = MIoU(num_classes, device)
        for imgs, labels in loader:
            miou.add_img(vit(imgs.to(device)).argmax(dim=1), labels.to(device))
        per_class = miou.get_per_class_iou()
        print(f"mIoU is: {100 * miou.get_miou():.2f}%")
        print("IoU per class:")
        for ix in range(num_classes):
            print(f"  {data_config.get('class_names', {}).get(str(ix), ix)!s:<11} - {100 * float(per_class[ix]):.2f}%")
        return {"miou": miou.get_miou(), "per_class_iou": per_class.cpu()}
    from myrtle_vision.models.detector import PostProcess
    post = PostProcess().eval()
    detections, truth = [], []
    for imgs, targets in loader:
        out = vit(imgs.to(device))
        targets = to_device(targets, device)
        sizes = torch.stack([t.get("orig_size", torch.tensor(imgs.shape[-2:], device=device)) for t in targets])
        detections.extend(post(out, sizes))
        for t, (h, w) in zip(targets, sizes.tolist()):
            scale = torch.tensor([w, h, w, h], dtype=torch.float32, device=device)
            truth.append({"labels": t["labels"], "boxes": cxcywh_to_xyxy(t["boxes"]) * scale})
    res = box_average_precision(detections, truth, num_classes)
    print(f"Average Precision  (AP) @[ IoU=0.50:0.95 | maxDets=100 ] = {res['AP']:.3f}")
    print(f"Average Precision  (AP) @[ IoU=0.50      | maxDets=100 ] = {res['AP50']:.3f}")
    return res
