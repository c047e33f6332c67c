"""Mean intersection-over-union for segmentation validation (reference src/myrtle_vision/utils/miou.py).

Same `MIoU(num_classes, device)` accumulator (`add_img(prediction, ground_truth)`, `get_per_class_iou()`,
`get_miou()`); the per-image statistics come from one `bincount` of `label * C + prediction` (a confusion
matrix) instead of three `histc` passes, so an image costs one kernel and no host synchronisation."""
import torch


def intersect_and_union(pred_label, label, num_classes):
    """-> (intersection, union, prediction histogram, label histogram), one entry per class."""
    pred = pred_label.reshape(-1).long()
    lab = label.reshape(-1).long()
    keep = (lab >= 0) & (lab < num_classes) & (pred >= 0) & (pred < num_classes)
    conf = torch.bincount(lab[keep] * num_classes + pred[keep], minlength=num_classes * num_classes)
    conf = conf.view(num_classes, num_classes).double()
    area_intersect = conf.diagonal()
    area_pred_label = torch.bincount(pred[(pred >= 0) & (pred < num_classes)], minlength=num_classes).double()
    area_label = torch.bincount(lab[(lab >= 0) & (lab < num_classes)], minlength=num_classes).double()
    return area_intersect, area_pred_label + area_label - area_intersect, area_pred_label, area_label


class MIoU:
    def __init__(self, num_classes, device):
        self.num_classes = num_classes
        self.total_area_intersect = torch.zeros(num_classes, dtype=torch.float64, device=device)
        self.total_area_union = torch.zeros(num_classes, dtype=torch.float64, device=device)

    def add_img(self, prediction_img, ground_truth_img):
        inter, union, _, _ = intersect_and_union(prediction_img, ground_truth_img, self.num_classes)
        self.total_area_intersect += inter.to(self.total_area_intersect.device)
        self.total_area_union += union.to(self.total_area_union.device)

    def get_per_class_iou(self):
        return self.total_area_intersect / self.total_area_union

    def get_miou(self):
        return torch.mean(self.get_per_class_iou()).item()
