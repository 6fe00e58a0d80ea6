"""fusion_b200 - B200-native (sm_100a) fusion head behind the reference's MultimodalModel API.

    import sys; sys.path.insert(0, ".../multimodal-model-skin-lesion-classifier_b200")
    from fusion_b200 import MultimodalModel, FusedCrossEntropyLoss

or, keeping the reference's own import line (train_pad_20.py:6):

    from models import multimodalIntraInterModal      # fusion_b200/compat adds this package
"""
from . import _lib, dp
from ._lib import Fb200Error
from .head import (FusedCrossEntropyLoss, FusedFocalLoss, FusedHeadFunction, FusedSoftTargetCrossEntropy, cross_entropy,
                   make_desc, softmax_argmax)
from .attention import FusedMHAFunction, MultiheadAttention
from .model import GraphedTrainStep, MultimodalModel
from .optim import FusedAdam
from .metadata import MetadataEncoder
from .tab_transformer import FusedLinearFunction, FusedTabEncoderFunction, TabTransformer

__all__ = ["TabTransformer", "FusedTabEncoderFunction", "FusedLinearFunction", "MetadataEncoder", "MultimodalModel", "MultiheadAttention", "FusedMHAFunction", "GraphedTrainStep", "FusedAdam", "FusedCrossEntropyLoss", "FusedFocalLoss", "FusedSoftTargetCrossEntropy", "softmax_argmax", "FusedHeadFunction", "cross_entropy", "make_desc", "Fb200Error", "_lib"]
