"""CPU: the drop-in MultimodalModel mirrors the reference's constructor, parameter names,
shapes and error behaviour, and fails loudly (no fallback) when asked to compute on the CPU."""
import inspect

import pytest
import torch

import fusion_b200 as fb
from oracle import head_oracle as ho
from oracle import ref_shim


def _mk(mech="crossattention", **kw):
    args = dict(num_classes=6, num_heads=8, device="cpu", cnn_model_name="identity:2048", text_model_name="one-hot-encoder",
                vocab_size=85, attention_mecanism=mech)
    args.update(kw)
    return fb.MultimodalModel(**args)


def test_constructor_signature_is_the_references():
    sig = inspect.signature(fb.MultimodalModel.__init__)
    names = list(sig.parameters)[1:13]
    assert names == ["num_classes", "num_heads", "device", "cnn_model_name", "text_model_name", "batch_size", "common_dim",
                     "text_encoder_dim_output", "vocab_size", "unfreeze_weights", "attention_mecanism", "n"]
    p = sig.parameters
    assert (p["batch_size"].default, p["common_dim"].default, p["text_encoder_dim_output"].default, p["vocab_size"].default,
            p["unfreeze_weights"].default, p["attention_mecanism"].default, p["n"].default) == (32, 512, 512, 91, "frozen_weights", "concatenation", 2)
    # positional use as in train_isic_2020.py:268
    fb.MultimodalModel(2, 8, "cpu", "identity:768", "one-hot-encoder", vocab_size=11)


def test_state_dict_keys_and_shapes():
    m = _mk()
    cfg = ho.HeadConfig("crossattention", F=2048, C=6, V=85)
    sd = m.state_dict()
    assert list(sd.keys()) == list(cfg.param_shapes().keys())
    for k, shp in cfg.param_shapes().items():
        assert tuple(sd[k].shape) == tuple(shp)
    assert len(sd) == 78


@pytest.mark.skipif(not ref_shim.available(), reason="reference sources not present")
@pytest.mark.parametrize("mech", ["crossattention", "no-metadata", ho.RG_ATT + "+metablock"])
def test_same_init_and_checkpoint_compat_as_reference(mech):
    """Same module creation order => identical weights under the same seed; strict load works both ways."""
    torch.manual_seed(7)
    ref = ref_shim.build_reference_model(mech, 2048, 6, V=85)
    torch.manual_seed(7)
    mine = _mk(mech)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k
    mine.load_state_dict(a, strict=True)
    ref.load_state_dict(b, strict=True)
    assert [n for n, _ in ref.named_modules()] == [n for n, _ in mine.named_modules()]


def test_unknown_mechanism_raises_reference_error():
    m = _mk("metablock-se")
    with pytest.raises(ValueError, match="Attention mechanism 'metablock-se' not implemented."):
        m(torch.zeros(2, 2048), torch.zeros(2, 85))


def test_no_cpu_fallback():
    m = _mk()
    with pytest.raises(fb.Fb200Error):
        m(torch.zeros(2, 2048), torch.zeros(2, 85))
    with pytest.raises(RuntimeError):          # Fb200Error is a RuntimeError: the loops' try/except keeps working
        fb.cross_entropy(torch.zeros(2, 6), torch.zeros(2, dtype=torch.long))


def test_tab_transformer_text_mode():
    m = fb.MultimodalModel(2, 8, "cpu", "identity:768", "tab-transformer", attention_mecanism="gfcam")
    assert m.text_fc is None and m.text_projector.weight.shape == (512, 85)
    # the encoder is the fused TabTransformer (reference constructor and state_dict); like the head it has no CPU path
    assert isinstance(m.text_encoder, fb.TabTransformer) and len(m.text_encoder.state_dict()) == 112
    x_cat = torch.randint(0, 10, (3, 82)); x_num = torch.randn(3, 4)
    with pytest.raises(fb.Fb200Error):
        m.text_encoder(x_cat, x_num)


def test_backbone_modes():
    from fusion_b200.backbones import loadModels
    # no network here: the pretrained weights the reference asks for cannot be fetched, and a silent random-init
    # fallback is refused - random init is an explicit opt-in ("random:" prefix or FB200_ALLOW_RANDOM_BACKBONE=1)
    import os
    if os.environ.get("FB200_ALLOW_RANDOM_BACKBONE", "0") != "1":
        with pytest.raises(RuntimeError, match="pretrained weights could not be loaded"):
            loadModels.loadModelImageEncoder("resnet-18", 512, "frozen_weights")
    with pytest.warns(UserWarning, match="random-init backbone"):
        enc, width = loadModels.loadModelImageEncoder("random:resnet-18", 512, "frozen_weights")
    assert width == 512 and not any(p.requires_grad for p in enc.parameters())
    with pytest.raises(ValueError), pytest.warns(UserWarning):
        loadModels.loadModelImageEncoder("random:resnet-18", 512, "false")        # conf/.env.test:8 pitfall, same error as the reference
    with pytest.raises(ValueError):
        loadModels.loadModelImageEncoder("not-a-backbone", 512, "frozen_weights")
