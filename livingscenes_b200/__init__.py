"""livingscenes_b200 -- B200-native (sm_100a) implementation of the LivingScenes per-instance
inference hot path behind the reference's own Python API.

    from livingscenes_b200 import Shape_Prior, VecDGCNN_att, sequential_matcher, nn_matcher, \
        kabsch_transformation_estimation, More_Solver

Every op runs in the hand-written CUDA library ``_ls_b200.so`` (C ABI: include/livingscenes_b200.h).
There is no CPU path and no fallback: importing works anywhere, calling an op without the built
extension or with CPU tensors raises.
"""
from . import _lib  # noqa: F401
from .decoder import DeepSDF_Decoder, FieldWrapper
from .encoder import VecDGCNN_att
from .matcher_new import (eq_seq_matcher, nn_matcher, nn_matcher_batched, sequential_matcher, sequential_matcher_batched,
                          sim3_seq_matcher, sinkhorn_matcher)
from .model_utils import Shape_Prior, extract_checkpoint, slice_code_dict
from .more_solver import More_Solver
from .ops import farthest_point_sample, knn_points, sample_farthest_points, vn_linear
from .pose_estimation import kabsch_from_codes, kabsch_transformation_estimation, rotation_error, translation_error

__all__ = [
    "Shape_Prior", "VecDGCNN_att", "DeepSDF_Decoder", "FieldWrapper", "More_Solver",
    "sequential_matcher", "sequential_matcher_batched", "nn_matcher", "nn_matcher_batched",
    "kabsch_transformation_estimation", "kabsch_from_codes", "rotation_error", "translation_error",
    "knn_points", "sample_farthest_points", "vn_linear", "farthest_point_sample", "extract_checkpoint", "slice_code_dict",
]
