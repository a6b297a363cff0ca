"""Registration of the fast classes behind MI-Seg's own layer factory (the drop-in boundary).

    import sys; sys.path.insert(0, "/path/to/MI-Seg")
    import importlib; importlib.import_module("mi-seg_b200").install()
    # ... from here on `--encoder_norm_name=instance_cond` builds FastConditionalInstanceNorm*d

`Norm.add_factory_callable` (networks/layers/factories.py:97-102) overwrites the "instance_cond"
entry that `instance_cond_factory` (:227-232) registered; `get_norm_layer` (networks/layers/utils.py:
43-50) then instantiates our class with the same kwargs (num_styles, affine, num_features).  Nothing in
`networks/` is edited.
"""
from __future__ import annotations

import importlib

from . import norms

_installed = None
_installed_plain = None


def install(factories_module: str = "networks.layers.factories",
            norm_module: str = "networks.norms.conditional_instance_norm"):
    """Register the fast classes under "instance_cond".  Returns the (1d, 2d, 3d) classes.
    MI-Seg's repository root must already be importable (sys.path)."""
    global _installed
    factories = importlib.import_module(factories_module)
    ref = importlib.import_module(norm_module)
    classes = norms.make_dropin_classes(ref)

    def fast_instance_cond_factory(dim: int):
        return classes[dim - 1]

    factories.Norm.add_factory_callable("instance_cond", fast_instance_cond_factory)
    _installed = (factories, ref, classes)
    return classes


def uninstall():
    """Restore MI-Seg's own classes under "instance_cond" (and torch's under "instance")."""
    global _installed, _installed_plain
    if _installed is not None:
        factories, ref, _ = _installed
        types = (ref.ConditionalInstanceNorm1d, ref.ConditionalInstanceNorm2d, ref.ConditionalInstanceNorm3d)
        factories.Norm.add_factory_callable("instance_cond", lambda dim: types[dim - 1])
        _installed = None
    if _installed_plain is not None:
        import torch.nn as nn

        plain = (nn.InstanceNorm1d, nn.InstanceNorm2d, nn.InstanceNorm3d)
        _installed_plain.Norm.add_factory_callable("instance", lambda dim: plain[dim - 1])
        _installed_plain = None
    while _patched_functional:
        mod, real = _patched_functional.pop()
        mod.F = real


class _FunctionalProxy:
    """`torch.nn.functional` as seen by ONE reference module (networks/nets/swin_transformer.py:6 `import
    torch.nn.functional as F`): everything forwards to torch, except `instance_norm` with batch statistics, which runs
    on the sm_100a kernels (SwinTransformer.proj_out, swin_transformer.py:135-136: `F.instance_norm(x)` on the five
    hidden states of every forward)."""

    def __init__(self, real):
        self._real = real

    def __getattr__(self, name):
        return getattr(self._real, name)

    def instance_norm(self, input, running_mean=None, running_var=None, weight=None, bias=None,
                      use_input_stats=True, momentum=0.1, eps=1e-5):
        if running_mean is not None or running_var is not None or not use_input_stats or input.dim() < 3:
            return self._real.instance_norm(input, running_mean, running_var, weight, bias, use_input_stats, momentum, eps)
        return norms.fast_instance_norm(input, weight, bias, eps)  # (CUDA only: raises for CPU tensors)


_patched_functional = []


def install_plain(factories_module: str = "networks.layers.factories",
                  functional_users=("networks.nets.swin_transformer",)):
    """also route the decoders' plain `("instance", {"affine": True})` norms
    (factories.py:221-224 -> nn.InstanceNorm{1,2,3}d) through the same kernels.  The fast classes derive from
    torch's, so `isinstance(m, nn.InstanceNorm3d)` and every state-dict key stay as they were.  The modules named in
    `functional_users` get their `F` rebound to a proxy whose `instance_norm` is the fast one (`proj_out`)."""
    global _installed_plain
    factories = importlib.import_module(factories_module)
    classes = (norms.FastInstanceNorm1d, norms.FastInstanceNorm2d, norms.FastInstanceNorm3d)
    factories.Norm.add_factory_callable("instance", lambda dim: classes[dim - 1])
    _installed_plain = factories
    for name in functional_users or ():
        try:
            mod = importlib.import_module(name)
        except Exception:  # noqa: BLE001 - a net that is not importable here simply is not patched
            continue
        real = getattr(mod, "F", None)
        if real is not None and not isinstance(real, _FunctionalProxy):
            mod.F = _FunctionalProxy(real)
            _patched_functional.append((mod, real))
    return classes


def convert_plain(model):
    """Re-class already-constructed plain nn.InstanceNorm{1,2,3}d modules (no running stats) to the fast
    ones in place: parameters, buffers and state-dict keys are untouched."""
    import torch.nn as nn

    table = {nn.InstanceNorm1d: norms.FastInstanceNorm1d, nn.InstanceNorm2d: norms.FastInstanceNorm2d,
             nn.InstanceNorm3d: norms.FastInstanceNorm3d}
    count = 0
    for m in model.modules():
        if type(m) in table and not m.track_running_stats:
            m.__class__ = table[type(m)]
            count += 1
    return count


def convert_module(model, classes=None):
    """Swap already-constructed reference `_ConditionalInstanceNorm` modules in `model` for fast ones
    sharing the same parameters (state-dict keys unchanged).  For checkpoints built before install()."""
    import torch.nn as nn

    if classes is None:
        classes = _installed[2] if _installed is not None else (
            norms.FastConditionalInstanceNorm1d, norms.FastConditionalInstanceNorm2d,
            norms.FastConditionalInstanceNorm3d)
    by_norm = {nn.InstanceNorm1d: classes[0], nn.InstanceNorm2d: classes[1], nn.InstanceNorm3d: classes[2]}
    for name, child in list(model.named_children()):
        inner = getattr(child, "norms", None)
        if isinstance(inner, nn.ModuleList) and len(inner) and type(inner[0]) in by_norm \
                and hasattr(child, "num_styles") and not isinstance(child, norms.FastForwardMixin):
            cls = by_norm[type(inner[0])]
            new = cls.__new__(cls)
            nn.Module.__init__(new)
            new.num_styles = child.num_styles
            new.norms = inner
            new.train(child.training)
            setattr(model, name, new)
        else:
            convert_module(child, classes)
    return model
