"""mi-seg_b200: B200-native `instance_cond` (modality-conditioned instance norm) for MI-Seg.

One hot path only: hand-written sm_100a CUDA forward/backward kernels behind
the C ABI of include/micn.h (csrc/ -> libmicn.so), and the Python mirror of the reference's module
interface that calls them.  The directory name carries a hyphen, so import it with
`importlib.import_module("mi-seg_b200")` or through the `mi_seg_b200` alias module at the repo root.
"""
from . import _lib
from .blocks import fuse_blocks
from .functional import binding_in_use, instance_cond, reset_workspaces, set_binding, set_channels_last_native
from .inference import sliding_window_inference, window_slices
from .integration import convert_module, convert_plain, install, install_plain, uninstall
from .parallel import PeerExchange, allreduce_style_grads, shard_range, style_parameters
from .norms import (FastConditionalInstanceNorm1d, FastConditionalInstanceNorm2d, FastConditionalInstanceNorm3d,
                    FastInstanceNorm1d, FastInstanceNorm2d, FastInstanceNorm3d, check_status, fast_instance_norm,
                    make_dropin_classes, set_sync_free_styles)

__all__ = ["instance_cond", "reset_workspaces", "install", "install_plain", "uninstall", "convert_module",
           "convert_plain", "fuse_blocks", "FastInstanceNorm1d", "FastInstanceNorm2d", "FastInstanceNorm3d",
           "fast_instance_norm", "set_channels_last_native", "sliding_window_inference", "window_slices",
           "FastConditionalInstanceNorm1d", "FastConditionalInstanceNorm2d", "FastConditionalInstanceNorm3d",
           "make_dropin_classes", "set_sync_free_styles", "check_status", "set_binding", "binding_in_use", "PeerExchange", "allreduce_style_grads",
           "shard_range", "style_parameters", "_lib"]
