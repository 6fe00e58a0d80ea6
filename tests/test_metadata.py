"""Metadata one-hot + StandardScaler: host-side fitting / code lookup against scikit-learn's own
OneHotEncoder + StandardScaler (the objects the reference uses, models/skinLesionDatasets.py:133-176) - CPU;
the dense [B, V] tensor built by the CUDA kernel against the golden vectors, bit-exact - GPU."""
import os

import numpy as np
import pytest

from fusion_b200.metadata import MetadataEncoder
from tests.golden import cases as C

G = np.load(os.path.join(C.GOLDEN_DIR, "metadata_pad20.npz"), allow_pickle=False)


def _categories():
    out, k = [], 0
    for n in G["n_categories"]:
        out.append([str(v) for v in G["categories_flat"][k:k + n]]); k += n
    return out


def _dense_from_codes(enc, codes, num):
    """numpy statement of the kernel, for the CPU-side checks of `codes` only."""
    sizes = [len(c) for c in enc.categories]
    out = np.zeros((len(codes), sum(sizes) + len(enc.mean)), np.float32)
    base = 0
    for j, n in enumerate(sizes):
        ok = codes[:, j] >= 0
        out[np.nonzero(ok)[0], base + codes[ok, j]] = 1.0
        base += n
    clean = np.where(np.isnan(num), -1.0, num)
    out[:, base:] = ((clean - enc.mean) / enc.scale).astype(np.float32)
    return out


def test_fit_reproduces_sklearn_categories_and_scaler():
    enc = MetadataEncoder().fit(G["fit_cat"], G["fit_num"])
    assert enc.categories == _categories()
    np.testing.assert_allclose(enc.mean, G["mean"], rtol=1e-13)
    np.testing.assert_allclose(enc.scale, G["scale"], rtol=1e-13)
    assert enc.width == G["dense_test"].shape[1] == 90


def test_codes_and_unknown_categories_match_reference_dense_vectors():
    enc = MetadataEncoder(_categories(), G["mean"], G["scale"])
    codes = enc.codes(G["test_cat"])
    assert codes.dtype == np.int32 and (codes < 0).any()                      # NEVER_SEEN / ZZZ -> -1 -> all-zero group
    assert np.array_equal(_dense_from_codes(enc, codes, G["test_num"]), G["dense_test"])
    assert np.array_equal(_dense_from_codes(enc, enc.codes(G["fit_cat"][:64]), G["fit_num"][:64]), G["dense_fit_head"])


def test_from_sklearn_objects():
    from sklearn.preprocessing import OneHotEncoder, StandardScaler
    ohe = OneHotEncoder(sparse_output=False, handle_unknown="ignore").fit(G["fit_cat"].astype(object))
    sc = StandardScaler().fit(np.where(np.isnan(G["fit_num"]), -1.0, G["fit_num"]))
    enc = MetadataEncoder.from_sklearn(ohe, sc)
    assert enc.categories == _categories()


def test_no_cpu_path():
    import fusion_b200 as fb
    enc = MetadataEncoder(_categories(), G["mean"], G["scale"])
    with pytest.raises(fb.Fb200Error):
        enc.transform(enc.codes(G["test_cat"]), G["test_num"], device="cpu")


@pytest.mark.gpu
def test_device_encoding_is_bit_identical_to_reference_pipeline():
    import torch
    enc = MetadataEncoder(_categories(), G["mean"], G["scale"])
    out = enc.transform(enc.codes(G["test_cat"]), G["test_num"])
    assert out.is_cuda and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), G["dense_test"])
    out = enc.transform(enc.codes(G["fit_cat"][:64]), G["fit_num"][:64])
    assert np.array_equal(out.cpu().numpy(), G["dense_fit_head"])
    # ragged / edge batches
    one = enc.transform(enc.codes(G["test_cat"][:1]), G["test_num"][:1])
    assert np.array_equal(one.cpu().numpy(), G["dense_test"][:1])
    big = enc.transform(np.tile(enc.codes(G["test_cat"]), (50, 1)), np.tile(G["test_num"], (50, 1)))
    assert np.array_equal(big.cpu().numpy(), np.tile(G["dense_test"], (50, 1)))


@pytest.mark.gpu
def test_encoded_metadata_feeds_the_head():
    import torch
    import fusion_b200 as fb
    enc = MetadataEncoder(_categories(), G["mean"], G["scale"])
    meta = enc.transform(enc.codes(G["test_cat"][:32]), G["test_num"][:32])
    model = fb.MultimodalModel(6, 8, "cuda", "identity:512", "one-hot-encoder", vocab_size=enc.width, attention_mecanism="weighted").cuda().eval()
    logits = model(torch.randn(32, 512, device="cuda"), meta)
    assert logits.shape == (32, 6) and torch.isfinite(logits).all()
