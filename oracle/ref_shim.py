"""Import the UNMODIFIED reference model on CPU (TEST INFRASTRUCTURE ONLY).

Only usable where /root/reference exists (the build container, not the GPU box).
Used by tests/golden/make_golden.py to generate fixtures and by
tests/test_oracle_vs_reference.py to pin the numpy oracle.  Recipe: SURVEY.md §8(c).

Shims (none of them touch the arithmetic of the head):
  1. ``transformers`` is imported before ``timm`` is stubbed (it probes timm.__spec__).
  2. ``timm`` is absent in this image -> an empty stub module satisfies
     loadImageModelClassifier.py:3.
  3. loadModels.loadModelImageEncoder is replaced by ``(nn.Identity(), F)`` so the
     "image" argument *is* the backbone feature tensor [B, F]; pretrained backbones
     cannot be downloaded here and are outside the hot path anyway.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

REF_ROOT = os.environ.get("FB200_REFERENCE_ROOT", "/root/reference")
REF_MODELS = os.path.join(REF_ROOT, "src/scripts/benchmark/models")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_MODELS, "multimodalIntraInterModal.py"))


_cache = {}


def load():
    """Returns (multimodalIntraInterModal module, loadImageModelClassifier module)."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError("reference sources not present at " + REF_MODELS)
    import torch.nn as nn  # noqa: F401
    import transformers  # noqa: F401  (must precede the timm stub)
    if "timm" not in sys.modules:
        stub = types.ModuleType("timm")
        stub.list_models = lambda pretrained=True: []
        sys.modules["timm"] = stub
    if REF_MODELS not in sys.path:
        sys.path.insert(0, REF_MODELS)
    import loadImageModelClassifier as lim
    import multimodalIntraInterModal as mim
    _cache["mods"] = (mim, lim)
    return _cache["mods"]


def build_reference_model(mechanism, F, C, V=85, T=512, D=512, H=8, n=2,
                          text_model="one-hot-encoder", device="cpu"):
    """Construct the reference MultimodalModel with an identity backbone of width F."""
    import torch.nn as nn
    mim, lim = load()
    lim.loadModels.loadModelImageEncoder = staticmethod(
        lambda cnn_model_name, common_dim, backbone_train_mode="frozen", device="cpu": (nn.Identity(), F))
    model = mim.MultimodalModel(
        num_classes=C, num_heads=H, device=device, cnn_model_name="identity",
        text_model_name=text_model, common_dim=D, text_encoder_dim_output=T,
        vocab_size=V if V is not None else 91, attention_mecanism=mechanism, n=n)
    return model


def reference_forward(model, img_feat, text_in):
    """model(image, meta) for one-hot; for other text encoders continue from txt_feat
    with the reference's own sub-modules (the published forward is unreachable there:
    multimodalIntraInterModal.py:180-183 passes HF kwargs to TabTransformer - SURVEY §8c)."""
    if model.text_model_name == "one-hot-encoder":
        return model(img_feat, text_in)
    import torch.nn as nn

    class _Feed(nn.Module):
        def forward(self, x):
            return x

    # Re-enter the reference forward from line 185 by presenting txt_feat as the output
    # of a trivial "one-hot" text_fc: same downstream code, no arithmetic added.
    saved = (model.text_model_name, model.text_fc)
    model.text_model_name, model.text_fc = "one-hot-encoder", _Feed()
    try:
        return model(img_feat, text_in)
    finally:
        model.text_model_name, model.text_fc = saved


@contextlib.contextmanager
def injected_dropout(mask_queue):
    """Make every active nn.Dropout consume the next keep-mask from ``mask_queue``
    (a list of torch tensors, call order) instead of drawing from Philox."""
    import torch.nn as nn
    orig = nn.Dropout.forward
    q = list(mask_queue)

    def fwd(self, x):
        if not self.training or self.p == 0.0:
            return x
        m = q.pop(0).to(x.dtype).reshape(x.shape)
        return x * m / (1.0 - self.p)

    nn.Dropout.forward = fwd
    try:
        yield q
    finally:
        nn.Dropout.forward = orig
