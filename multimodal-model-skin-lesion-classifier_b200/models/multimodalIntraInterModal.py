"""Same module name as the reference file so that the training scripts' import line
(train_pad_20.py:6, train_isic_2019.py:7, train_isic_2020.py:6) needs no edit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fusion_b200.model import MultimodalModel  # noqa: E402,F401
