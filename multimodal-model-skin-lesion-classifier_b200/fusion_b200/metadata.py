"""Metadata one-hot + StandardScaler with the dense vector built on the device.

Mirrors ``SkinLesionDataset.one_hot_encoding`` (models/skinLesionDatasets.py:133-176, and the ISIC-2019 / ISIC-2020
variants): categorical columns -> ``OneHotEncoder(handle_unknown='ignore')`` groups in column order, numerical columns
(NaN -> -1) -> ``StandardScaler``, ``np.hstack((categorical, numerical))``.  Fitting (unique strings, mean / std) and the
string -> code lookup are host work by nature; the [B, V] fp32 tensor the head consumes is produced by
``fb200_metadata_encode`` directly in HBM, so a batch travels as B x n_cat int32 codes + B x n_num values instead of a
B x V dense fp32 matrix (PAD-UFES-20: 22 columns instead of 85 floats), and there is no CPU path for it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


class MetadataEncoder:
    """``fit`` once per dataset, ``codes`` per batch on the host, ``transform`` on the device.

    ``categories`` follows scikit-learn: per column, the sorted unique values of the fitted data (as str);
    ``mean`` / ``scale`` are StandardScaler's mean_ / scale_ (population std, zero-variance columns scale by 1)."""

    def __init__(self, categories=None, mean=None, scale=None):
        self.categories = None if categories is None else [list(map(str, c)) for c in categories]
        self.mean = None if mean is None else np.asarray(mean, np.float64)
        self.scale = None if scale is None else np.asarray(scale, np.float64)
        self._dev = {}

    # -- fitting (host) ---------------------------------------------------------------
    def fit(self, categorical, numerical):
        """categorical: [N, n_cat] array-like of values (converted with str(), like .astype(str) at :144);
        numerical: [N, n_num] array-like of floats (NaN -> -1 like :148-152)."""
        cat = np.asarray(categorical, dtype=object)
        cat = np.array([[str(v) for v in row] for row in cat], dtype=object).reshape(len(cat), -1)
        self.categories = [sorted(set(cat[:, j].tolist())) for j in range(cat.shape[1])]
        num = self._clean_numeric(numerical)
        self.mean = num.mean(axis=0) if num.shape[1] else np.zeros(0)
        var = num.var(axis=0) if num.shape[1] else np.zeros(0)
        scale = np.sqrt(var)
        scale[scale == 0.0] = 1.0                      # sklearn.preprocessing._data._handle_zeros_in_scale
        self.scale = scale
        self._dev = {}
        return self

    @classmethod
    def from_sklearn(cls, ohe, scaler):
        """Adopt fitted scikit-learn objects (the pickles the reference stores under data/preprocess_data/)."""
        return cls([list(map(str, c)) for c in ohe.categories_], scaler.mean_, scaler.scale_)

    @staticmethod
    def _clean_numeric(numerical):
        num = np.asarray(numerical, dtype=np.float64)
        num = num.reshape(len(num), -1)
        return np.where(np.isnan(num), -1.0, num)

    @property
    def width(self):
        return sum(len(c) for c in self.categories) + len(self.mean)

    # -- per batch (host): strings -> codes --------------------------------------------
    def codes(self, categorical):
        cat = np.asarray(categorical, dtype=object)
        cat = cat.reshape(len(cat), -1)
        out = np.full(cat.shape, -1, np.int32)
        for j, cats in enumerate(self.categories):
            lut = {v: i for i, v in enumerate(cats)}
            out[:, j] = [lut.get(str(v), -1) for v in cat[:, j]]
        return out

    # -- per batch (device) --------------------------------------------------------------
    def _tables(self, device):
        t = self._dev.get(device)
        if t is None:
            sizes = [len(c) for c in self.categories]
            base = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int32) if sizes else np.zeros(0, np.int32)
            col_of = np.repeat(np.arange(len(sizes), dtype=np.int32), sizes)
            t = dict(col_of=torch.from_numpy(col_of).to(device), base=torch.from_numpy(base).to(device),
                     mean=torch.from_numpy(self.mean).to(device), scale=torch.from_numpy(self.scale).to(device),
                     cat_total=int(sum(sizes)))
            self._dev[device] = t
        return t

    def transform(self, codes, numerical, device="cuda"):
        """codes: int32 [B, n_cat] (``self.codes(...)``; host or device), numerical: [B, n_num] float64 raw values.
        Returns the [B, V] fp32 CUDA tensor ``model(image, metadata)`` takes."""
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.Fb200Error(-2, "MetadataEncoder.transform builds the tensor on the GPU (fusion_b200 has no CPU path)")
        if device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        t = self._tables(device)
        codes = torch.as_tensor(codes, dtype=torch.int32).to(device).contiguous()
        num = torch.as_tensor(self._clean_numeric(numerical) if not isinstance(numerical, torch.Tensor) else numerical, dtype=torch.float64)
        num = torch.where(torch.isnan(num), torch.full_like(num, -1.0), num).to(device).contiguous()
        B = codes.shape[0] if codes.numel() or not num.numel() else num.shape[0]
        n_cat, n_num = len(self.categories), len(self.mean)
        out = torch.empty(B, t["cat_total"] + n_num, dtype=torch.float32, device=device)
        p = lambda x: C.c_void_p(x.data_ptr()) if x.numel() else C.c_void_p(0)
        with torch.cuda.device(device):
            _lib.check(_lib.lib().fb200_metadata_encode(p(codes), p(t["col_of"]), p(t["base"]), p(num), p(t["mean"]), p(t["scale"]),
                                                        B, n_cat, t["cat_total"], n_num, C.c_void_p(out.data_ptr()),
                                                        C.c_void_p(torch.cuda.current_stream().cuda_stream)), "fb200_metadata_encode")
        return out
