"""B200-native PAMR + pseudo-label epilogue (drop-in for the hot path of EnchanterXiao/1-stage-wseg).

The directory name `1-stage-wseg_b200` is not a Python identifier; import it as `wseg_b200`
(the alias module at the repository root registers this package under that name).
"""
from . import _lib
from .pamr import PAMR, LocalAffinity, LocalAffinityAbs, LocalAffinityCopy, LocalStDev
from .pamr import local_affinity, local_std, propagate, resize_bilinear
from .stage import (IGNORE_INDEX, VOC_MEAN, VOC_STD, HostPipeline, denorm_resize, labels_from_pseudo_gt, pseudo_gtmask, pseudo_labels, refine_and_label,
                    rescale_and_clean, run_pamr)
from .dist import OverlappedLabelGather, ShardedPseudoLabeler, gather_labels, shard_batch, shard_range
from .inference import merge_and_predict, merge_masks
from .loss import balanced_mask_loss_ce, balanced_mask_loss_ce_from_labels, labels_from_onehot

__all__ = [
    "PAMR", "LocalAffinity", "LocalAffinityAbs", "LocalAffinityCopy", "LocalStDev", "local_affinity", "local_std", "propagate",
    "resize_bilinear", "run_pamr", "rescale_and_clean", "pseudo_gtmask", "pseudo_labels", "labels_from_pseudo_gt",
    "refine_and_label", "HostPipeline", "IGNORE_INDEX", "ShardedPseudoLabeler", "OverlappedLabelGather", "gather_labels", "shard_batch", "shard_range",
    "denorm_resize", "VOC_MEAN", "VOC_STD", "merge_masks", "merge_and_predict", "balanced_mask_loss_ce", "balanced_mask_loss_ce_from_labels", "labels_from_onehot",
]
