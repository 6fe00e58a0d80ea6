"""Image / text encoder factory - STOCK PyTorch, outside the optimised path.

Mirrors the interface of the reference's ``loadModels`` (loadImageModelClassifier.py:41-203):
``loadModelImageEncoder(name, common_dim, backbone_train_mode) -> (module, cnn_dim_output)``.
Pretrained weights are required, like the reference (``pretrained=True`` everywhere).  Where they
cannot be loaded (no network in the build / bench containers) the architecture is built with random
init ONLY on explicit opt-in - ``FB200_ALLOW_RANDOM_BACKBONE=1`` or the ``random:`` name prefix, e.g.
``"random:resnet-50"`` - and a warning is emitted; otherwise the loading error is re-raised.
Extra names for head-only work: ``"identity:<F>"`` (the "image" is already the [B,F] feature).
"""
from __future__ import annotations

import os
import warnings

import torch.nn as nn

_TORCHVISION = {  # name -> (factory attr, attribute replaced by Identity, feature width)
    "resnet-18": ("resnet18", "fc", 512),
    "resnet-50": ("resnet50", "fc", 2048),
    "densenet169": ("densenet169", "classifier", 1664),
    "mobilenet-v2": ("mobilenet_v2", "classifier", 1280),
    "efficientnet-b0": ("efficientnet_b0", "classifier", 1280),
    "efficientnet-b7": ("efficientnet_b7", "classifier", 2560),
}


def set_backbone_train_mode(model, mode="frozen_weights", last_n_layers=1):
    """Same three modes (and the same ValueError) as loadImageModelClassifier.py:15-35."""
    params = list(model.parameters())
    for p in params:
        p.requires_grad = False
    if mode == "frozen_weights":
        return
    if mode == "unfrozen_weights":
        for p in params:
            p.requires_grad = True
    elif mode == "last_layer_unfrozen_weights":
        for p in params[-2 * last_n_layers:]:
            p.requires_grad = True
    else:
        raise ValueError(f"Invalid backbone_train_mode: {mode}")


def _random_init_allowed(explicit):
    return explicit or os.environ.get("FB200_ALLOW_RANDOM_BACKBONE", "0") == "1"


def _pretrained_or_optin(load_pretrained, load_random, what, explicit_random):
    """Pretrained weights, or - only on opt-in - a randomly initialised copy of the architecture (with a warning).
    A frozen random backbone trains the head on noise, so a silent fallback is never right."""
    if explicit_random:
        warnings.warn(f"{what}: random-init backbone requested explicitly (no pretrained weights)")
        return load_random()
    try:
        return load_pretrained()
    except Exception as exc:
        if not _random_init_allowed(False):
            raise RuntimeError(f"{what}: pretrained weights could not be loaded ({exc!r}); set FB200_ALLOW_RANDOM_BACKBONE=1 "
                               f"or use the 'random:' name prefix to build the architecture with random init") from exc
        warnings.warn(f"{what}: pretrained weights unavailable ({exc!r}); using RANDOM init (FB200_ALLOW_RANDOM_BACKBONE=1)")
        return load_random()


def _tv_model(factory, explicit_random=False):
    from torchvision import models
    fn = getattr(models, factory)
    return _pretrained_or_optin(lambda: fn(weights="DEFAULT"), lambda: fn(weights=None), f"torchvision.{factory}", explicit_random)


class loadModels:
    @staticmethod
    def loadModelImageEncoder(cnn_model_name, common_dim, backbone_train_mode="frozen", device="cpu"):
        if cnn_model_name.startswith("identity:"):
            return nn.Identity(), int(cnn_model_name.split(":", 1)[1])
        explicit_random = cnn_model_name.startswith("random:")
        if explicit_random:
            cnn_model_name = cnn_model_name.split(":", 1)[1]
        if cnn_model_name in _TORCHVISION:
            factory, head, width = _TORCHVISION[cnn_model_name]
            model = _tv_model(factory, explicit_random)
            setattr(model, head, nn.Identity())
            if cnn_model_name == "densenet169" and backbone_train_mode == "partial":
                for p in model.parameters():
                    p.requires_grad = False
                for p in model.features.denseblock4.parameters():
                    p.requires_grad = True
            else:
                set_backbone_train_mode(model, backbone_train_mode, last_n_layers=1)
            return model, width
        if cnn_model_name == "vgg16":
            model = _tv_model("vgg16", explicit_random)
            model.classifier = nn.Sequential(*list(model.classifier.children())[:-1])
            set_backbone_train_mode(model, backbone_train_mode, last_n_layers=1)
            return model, 4096
        try:
            import timm
        except ImportError:
            timm = None
        if timm is not None and cnn_model_name in timm.list_models():
            model = _pretrained_or_optin(lambda: timm.create_model(cnn_model_name, pretrained=True),
                                         lambda: timm.create_model(cnn_model_name, pretrained=False), f"timm.{cnn_model_name}", explicit_random)
            if hasattr(model, "reset_classifier"):
                model.reset_classifier(0)
            set_backbone_train_mode(model, backbone_train_mode)
            return model, int(model.num_features)
        raise ValueError(f"Backbone '{cnn_model_name}' não implementado.")

    @staticmethod
    def loadTextModelEncoder(text_model_encoder, train_mode="frozen_weights"):
        if text_model_encoder == "tab-transformer":
            return TabTransformer([10] * 82, num_continuous=4, output_dim=85), 85, 85
        if text_model_encoder in ("bert-base-uncased", "gpt2"):
            from transformers import AutoModel
            model = AutoModel.from_pretrained(text_model_encoder)
            for p in model.parameters():
                p.requires_grad = train_mode == "unfrozen_weights"
            return model, model.config.hidden_size, model.config.hidden_size
        raise ValueError(f"Text encoder '{text_model_encoder}' não suportado.")


from .tab_transformer import TabTransformer  # noqa: E402,F401  (fused sm_100a encoder behind the reference constructor; was stock PyTorch in round 1)
