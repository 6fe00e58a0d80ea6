"""Same module name as the reference's backbone factory (loadImageModelClassifier.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion_b200.backbones import loadModels, set_backbone_train_mode  # noqa: E402,F401
