"""detectron2.modeling.roi_heads.ROIHeads._sample_proposals (restated from detectron2 0.5
modeling/roi_heads/roi_heads.py); the reference's OpenVocabularyRes5ROIHeads inherits it."""
import torch
from torch import nn

from detectron2.modeling.sampling import subsample_labels


class ROIHeads(nn.Module):
    def _sample_proposals(self, matched_idxs, matched_labels, gt_classes):
        has_gt = gt_classes.numel() > 0
        if has_gt:
            gt_classes = gt_classes[matched_idxs]
            gt_classes[matched_labels == 0] = self.num_classes
            gt_classes[matched_labels == -1] = -1
        else:
            gt_classes = torch.zeros_like(matched_idxs) + self.num_classes
        sampled_fg_idxs, sampled_bg_idxs = subsample_labels(gt_classes, self.batch_size_per_image,
                                                            self.positive_fraction, self.num_classes)
        sampled_idxs = torch.cat([sampled_fg_idxs, sampled_bg_idxs], dim=0)
        return sampled_idxs, gt_classes[sampled_idxs]
