"""Synthetic stand-ins for the reference datasets (RESISC45 / DLRSD / DIOR are not redistributable and
their loaders — src/myrtle_vision/datasets/*.py — are outside the hot path).

Each dataset yields tensors of exactly the shape and dtype the reference loaders hand to
`train_deit`: classification `(img[3,S,S], label)`, segmentation `(img[3,S,S], mask[S,S] int64)`,
detection `(img[3,S,S], {"labels": int64[k], "boxes": float[k,4] cxcywh in 0..1})`.  Images are
normalised like `data_configs/data_config.json:11-14` (mean 0.5 / std 0.5 => roughly [-1, 1]).
Samples are generated on the fly from (seed, index), so a dataset of any length costs no memory.
"""
import torch
from torch.utils.data import Dataset


class SyntheticVision(Dataset):
    def __init__(self, task, length, image_size, num_classes, seed=1234, max_boxes=20):
        assert task in ("classification", "segmentation", "detection")
        self.task, self.length, self.size = task, length, image_size
        self.num_classes, self.seed, self.max_boxes = num_classes, seed, max_boxes

    def __len__(self):
        return self.length

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        img = torch.randn(3, self.size, self.size, generator=g).clamp(-1, 1)
        if self.task == "classification":
            return img, int(torch.randint(0, self.num_classes, (1,), generator=g))
        if self.task == "segmentation":
            # blocky masks: class regions of 32x32 pixels, like land-cover tiles
            s = max(1, self.size // 32)
            coarse = torch.randint(0, self.num_classes, (s, s), generator=g)
            mask = coarse.repeat_interleave(32, 0).repeat_interleave(32, 1)[:self.size, :self.size]
            return img, mask.contiguous()
        k = int(torch.randint(1, self.max_boxes + 1, (1,), generator=g))
        cxcy = torch.rand(k, 2, generator=g) * 0.8 + 0.1
        wh = torch.rand(k, 2, generator=g) * 0.25 + 0.05
        return img, {"labels": torch.randint(0, self.num_classes, (k,), generator=g),
                     "boxes": torch.cat([cxcy, wh], dim=1)}


def detection_collate(batch):
    """Images have one size here, so the batch is a plain tensor (the reference pads into a
    NestedTensor, transforms/detection.py) and the targets stay a list of dicts."""
    return torch.stack([b[0] for b in batch]), [b[1] for b in batch]
