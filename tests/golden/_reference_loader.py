"""Import the UNMODIFIED reference (``/root/reference/MicFormer``) in the build container.

Only the golden-vector generator and the optional ``reference``-marked CPU tests use this; nothing that
runs on the GPU box may (``/root/reference`` does not exist there).  The reference needs exactly one
third-party symbol that is absent from this image: ``timm.models.layers.DropPath``
(MicFormer/models/MICFormer_self.py:5).  A stub with timm's semantics is registered in ``sys.modules``.
"""
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("MICFORMER_REFERENCE", "/root/reference/MicFormer")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "MICFormer_self.py"))


class _DropPath(nn.Module):
    def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
        super().__init__()
        self.drop_prob = drop_prob
        self.scale_by_keep = scale_by_keep

    def forward(self, x):
        if self.drop_prob == 0.0 or not self.training:
            return x
        keep = 1 - self.drop_prob
        mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
        if keep > 0.0 and self.scale_by_keep:
            mask.div_(keep)
        return x * mask

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


def load_reference():
    """Returns (MICFormer_self module, dice module) of the untouched reference."""
    if not reference_available():
        raise FileNotFoundError(REFERENCE_ROOT)
    if "timm" not in sys.modules:
        timm = types.ModuleType("timm"); models = types.ModuleType("timm.models")
        layers = types.ModuleType("timm.models.layers")
        layers.DropPath = _DropPath
        timm.models = models; models.layers = layers
        sys.modules.update({"timm": timm, "timm.models": models, "timm.models.layers": layers})
    # the reference packages are called `models` / `loss`; import them under private names so they
    # cannot shadow anything else called `models` on sys.path
    import importlib.util

    def _load_pkg(alias, sub):
        path = os.path.join(REFERENCE_ROOT, sub)
        spec = importlib.util.spec_from_file_location(alias, os.path.join(path, "__init__.py")
                                                      if os.path.exists(os.path.join(path, "__init__.py")) else None,
                                                      submodule_search_locations=[path])
        if spec is None or spec.loader is None:
            pkg = types.ModuleType(alias); pkg.__path__ = [path]
            sys.modules[alias] = pkg
            return pkg
        pkg = importlib.util.module_from_spec(spec)
        sys.modules[alias] = pkg
        spec.loader.exec_module(pkg)
        return pkg

    if "_micref_models" not in sys.modules:
        _load_pkg("_micref_models", "models")
    if "_micref_loss" not in sys.modules:
        pkg = types.ModuleType("_micref_loss"); pkg.__path__ = [os.path.join(REFERENCE_ROOT, "loss")]
        sys.modules["_micref_loss"] = pkg
    import importlib
    m = importlib.import_module("_micref_models.MICFormer_self")
    d = importlib.import_module("_micref_loss.dice")
    return m, d
