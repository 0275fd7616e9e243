"""diffuncertainty_b200 -- the ValUES per-pixel uncertainty hot path on B200.

Hand-written sm_100a CUDA (libvalunc.so, C ABI in include/valunc.h) behind the
Python call surface of JakobLC/DiffUncertainty's
``uncertainty_modeling/unc_mod_utils/test_utils.py`` (C2 measures),
``evaluation/uncertainty_aggregation`` (C3 aggregation) and
``evaluation/metrics`` (ECE/ACE, NCC, AURC inputs).  There is no CPU fallback.
"""
from . import _lib  # noqa: F401
from .uncertainty import (FusedResult, GroundTruth, Groups, calculate_one_minus_msr, calculate_uncertainty,  # noqa: F401
                          calculate_uncertainty_from_logits, fused_pass, group_members, map_stats, mean_argmax_labels)

from .members import MemberScoreBuffers, fused_pass_with_member_scores  # noqa: F401,E402

__version__ = "0.2.0"
