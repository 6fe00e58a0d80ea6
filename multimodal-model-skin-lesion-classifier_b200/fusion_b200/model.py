"""Drop-in ``MultimodalModel`` whose fusion head runs in libfb200 (sm_100a CUDA).

Same constructor, parameter names / shapes / initialisation, fusion strings and
``forward(image, text_metadata) -> logits`` as the reference class
(src/scripts/benchmark/models/multimodalIntraInterModal.py:13-416), so the
train_pad_20 / train_isic_2019 / train_isic_2020 loops, ``state_dict`` checkpoints
(api.py:117 loads strictly) and ``.image_encoder`` users keep working.  The torch.nn
sub-modules created here are PARAMETER CONTAINERS ONLY - their ``forward`` is never
called; ``forward`` hands raw device pointers of their weights to the C ABI.
"""
from __future__ import annotations

import ctypes as C
import os

import torch
import torch.nn as nn

from . import _lib
from .backbones import loadModels
from .head import FusedHeadFunction, ParamTable, _check_input, _mask_table, _ptr, _stream, denom_arg, make_desc, param_table

RG_ATT = "att-intramodal+residual+cross-attention-metadados"


def _mlp_container(first_in, dim, num_classes, p):
    # index layout 0 Linear, 1 LayerNorm, 2 ReLU, 3 Dropout, 4 Linear, 5 LayerNorm, 6 ReLU, 7 Dropout, 8 Linear
    widths = [(first_in, dim), (dim, dim // 2)]
    mods = []
    for i, o in widths:
        mods += [nn.Linear(i, o), nn.LayerNorm(o), nn.ReLU(), nn.Dropout(p)]
    mods.append(nn.Linear(dim // 2, num_classes))
    return nn.Sequential(*mods)


class _GatedResidualParams(nn.Module):
    """Parameters of GatedAlteredResidualBlock (gatedResidualBlock.py:4-10), same names."""

    def __init__(self, dim, dropout=0.1):
        super().__init__()
        self.norm = nn.LayerNorm(dim)
        self.attn = nn.MultiheadAttention(embed_dim=dim, num_heads=8, batch_first=False)
        self.dropout = nn.Dropout(dropout)
        self.gate_linear = nn.Linear(dim, dim)


class _MetaBlockParams(nn.Module):
    """Parameters of MetaBlock(V_dim, U_dim) (metablock.py:9-20), same names."""

    def __init__(self, v_dim, u_dim):
        super().__init__()
        self.fb = nn.Sequential(nn.Linear(u_dim, v_dim), nn.LayerNorm(v_dim))
        self.gb = nn.Sequential(nn.Linear(u_dim, v_dim), nn.LayerNorm(v_dim))


class MultimodalModel(nn.Module):
    def __init__(self, num_classes, num_heads, device, cnn_model_name, text_model_name, batch_size=32,
                 common_dim=512, text_encoder_dim_output=512, vocab_size=91, unfreeze_weights="frozen_weights",
                 attention_mecanism="concatenation", n=2, compute_dtype="fp32", engine_flags=0, backend="b200"):
        super().__init__()
        if backend not in ("b200", "reference"):
            raise ValueError(f"backend must be 'b200' or 'reference', got {backend!r}")
        self.backend = backend               # "reference": stock torch.nn composition for double-backward callers (reference_backend.py)
        self.device = device
        self.common_dim, self.num_heads, self.n = common_dim, num_heads, n
        self.attention_mecanism = attention_mecanism
        self.vocab_size, self.num_classes = vocab_size, num_classes
        self.cnn_model_name, self.text_model_name = cnn_model_name, text_model_name
        self.unfreeze_weights = unfreeze_weights
        self.text_encoder_dim_output = text_encoder_dim_output
        self.compute_dtype, self.engine_flags = compute_dtype, engine_flags
        D = common_dim

        # creation order == reference order, so torch.manual_seed(k) gives identical initial weights
        self.image_encoder, self.cnn_dim_output = loadModels.loadModelImageEncoder(
            cnn_model_name=cnn_model_name, common_dim=D, backbone_train_mode=unfreeze_weights)
        self.image_projector = nn.Linear(self.cnn_dim_output, D)
        if text_model_name == "one-hot-encoder":
            self.text_fc = nn.Sequential(nn.Linear(vocab_size, 256), nn.ReLU(), nn.Linear(256, 512), nn.ReLU(),
                                         nn.Linear(512, self.text_encoder_dim_output))
            self.text_encoder = None
        else:
            self.text_encoder, self.text_encoder_dim_output, _ = loadModels.loadTextModelEncoder(
                text_model_encoder=text_model_name, train_mode=unfreeze_weights)
            self.text_fc = None
        self.text_projector = nn.Linear(self.text_encoder_dim_output, D)
        for name in ("image_self_attention", "text_self_attention", "image_cross_attention", "text_cross_attention"):
            setattr(self, name, nn.MultiheadAttention(embed_dim=D, num_heads=num_heads, batch_first=False))
        self.img_gate = nn.Linear(D, D)
        self.txt_gate = nn.Linear(D, D)
        mb_common = attention_mecanism == RG_ATT + "+metablock"
        self.meta_block = _MetaBlockParams(
            v_dim=D if mb_common else self.cnn_dim_output,
            u_dim=D if attention_mecanism in (RG_ATT + "+metablock", "metablock-se") else self.text_encoder_dim_output)
        self.image_residual = _GatedResidualParams(D)
        self.text_residual = _GatedResidualParams(D)
        self.fc_fusion = _mlp_container(D * (1 if attention_mecanism == "no-metadata" else n), D, num_classes, 0.5)
        self.fc_visual_only = nn.Linear(self.cnn_dim_output, num_classes)
        self.fc_fusion_proj_feat2output = nn.Linear(D, num_classes)
        self.fc_mlp_module_after_metablock_fusion_module = _mlp_container(self.cnn_dim_output, D, num_classes, 0.3)

        self._slot_names = _lib.param_names()
        self._step = 0                       # Philox offset: advances once per training forward
        self._injected_masks = None          # tests: {site: uint8 keep-mask}
        self._rng_state = None               # device {seed, offset} for the fused / graph-captured train step

    # ------------------------------------------------------------------ plumbing
    def _cfg(self, train):
        return dict(mechanism=self.attention_mecanism, F=self.cnn_dim_output,
                    V=self.vocab_size if self.text_model_name == "one-hot-encoder" else 0,
                    T=self.text_encoder_dim_output, D=self.common_dim, H=self.num_heads, C=self.num_classes, n=self.n,
                    text_mode=0 if self.text_model_name == "one-hot-encoder" else 1,
                    dtype=self.compute_dtype, train=train, flags=self.engine_flags)

    def _params_in_slot_order(self):
        """The 78 head parameters in slot order (None for absent ones).  Cached: walking named_parameters() (backbone included)
        on every pass costs more host time than the step kernel's launch; `_apply` (.to / .cuda / .float) keeps nn.Parameter
        objects and replaces their .data, load_state_dict copies in place, so object identity is a valid cache key - a
        parameter object that was REPLACED shows up as a different id and rebuilds the list."""
        cache = self.__dict__.get("_slot_cache")
        if cache is not None and all(p is q for p, q in zip(cache[1], (self._parameters_of(m).get(n) for m, n in cache[0]))):
            return cache[2]
        named = dict(self.named_parameters())
        owners = []
        for k in self._slot_names:
            mod_path, _, leaf = k.rpartition(".")
            try:
                owners.append((self.get_submodule(mod_path) if mod_path else self, leaf))
            except AttributeError:
                owners.append((None, leaf))
        plist = [named.get(k) for k in self._slot_names]
        self.__dict__["_slot_cache"] = (owners, list(plist), plist)
        return plist

    @staticmethod
    def _parameters_of(module):
        return module._parameters if module is not None else {}

    def _live_slots(self, desc):
        key = _lib.desc_key(desc)[:10]
        hit = self.__dict__.setdefault("_live_cache", {}).get(key)
        if hit is None:
            _, offs = _lib.grad_layout(desc)
            hit = self.__dict__["_live_cache"][key] = tuple(sorted(offs))
        return hit

    def inject_dropout_masks(self, masks):
        """Parity tests: use explicit {0,1} keep-masks instead of the in-kernel Philox stream."""
        self._injected_masks = masks

    def _encode(self, image, text_metadata):
        image = image.to(self.device)
        img_feat = self.image_encoder(image)
        if img_feat.dim() == 4:
            img_feat = img_feat.mean(dim=(-2, -1))
        if self.text_model_name == "one-hot-encoder":
            text_in = text_metadata.to(self.device)
        elif isinstance(text_metadata, torch.Tensor):
            text_in = text_metadata.to(self.device)                      # pre-computed encoder features [B,T]
        elif isinstance(text_metadata, (tuple, list)):
            text_in = self.text_encoder(*[t.to(self.device) for t in text_metadata])   # TabTransformer(x_cat, x_num)
        else:                                                            # HF tokenizer dict (reference :180-183)
            ids = text_metadata["input_ids"].squeeze(1).to(self.device)
            att = text_metadata["attention_mask"].squeeze(1).to(self.device)
            text_in = self.text_encoder(input_ids=ids, attention_mask=att).last_hidden_state[:, 0, :]
        if self.backend == "reference":
            return img_feat, text_in                                     # stock torch ops: any floating dtype
        return img_feat.float(), text_in.float()

    # ------------------------------------------------------------------ reference API
    def forward(self, image, text_metadata):
        if _lib.mechanism_id(self.attention_mecanism) < 0:
            raise ValueError(f"Attention mechanism '{self.attention_mecanism}' not implemented.")
        img_feat, text_in = self._encode(image, text_metadata)
        if self.backend == "reference":
            from .reference_backend import reference_forward
            return reference_forward(self, img_feat, text_in)
        train = bool(self.training)
        if train:
            self._step += 1
        cfg = self._cfg(train)
        params = self._params_in_slot_order()
        desc = make_desc(cfg["mechanism"], img_feat.shape[0], cfg["F"], cfg["V"], cfg["T"], cfg["D"], cfg["H"], cfg["C"], cfg["n"],
                         cfg["text_mode"], cfg["dtype"], train, cfg["flags"])
        slots = tuple(s for s in self._live_slots(desc) if params[s] is not None)
        return FusedHeadFunction.apply(img_feat, text_in, cfg, self._injected_masks,
                                       torch.initial_seed() & 0xFFFFFFFFFFFFFFFF, self._step, slots,
                                       [None if p is None else p.detach() for p in params], *[params[s] for s in slots])

    # ------------------------------------------------------------------ fused train step (opt-in)
    def forward_loss(self, image, text_metadata, label, class_weights=None, denom=None, zero_grad=True, mid_event=None, flat_out=None):
        """forward + weighted CE + backward of the head in ONE library call.

        Writes ``.grad`` of every live head parameter (views of one flat buffer, summed into
        existing grads unless ``zero_grad``), back-propagates into the backbone when it has
        trainable parameters, and returns (loss, logits) as device tensors without syncing.
        ``denom``: device scalar with the global sum of class weights (data parallel runs).
        ``flat_out``: caller-owned fp32 tensor of at least ``_lib.grad_layout(desc)[0]`` elements that receives the flat
        gradient buffer (data parallel: a symmetric-memory buffer, dp.SymmetricGradBucket, so that the all-reduce kernel
        works on it in place).
        ``mid_event``: a recorded-once ``torch.cuda.Event``; the library records it as soon as every gradient below
        ``_lib.dp_bucket_split(desc)`` in ``flat_grad`` is final (fb200_head_train_step_dp) - dp.BucketedAllReduce
        all-reduces that bucket while the remaining weight gradients are still being computed."""
        L = _lib.lib()
        img_feat, text_in = self._encode(image, text_metadata)
        train = bool(self.training)
        if train:
            self._step += 1
        need_dimg, need_dtxt = bool(img_feat.requires_grad), bool(text_in.requires_grad)
        cfg = self._cfg(train)
        flags = cfg["flags"] | (_lib.FLAG_NEED_DIMG if need_dimg else 0) | (_lib.FLAG_NEED_DTEXT if need_dtxt else 0)
        B = img_feat.shape[0]
        desc = make_desc(cfg["mechanism"], B, cfg["F"], cfg["V"], cfg["T"], cfg["D"], cfg["H"], cfg["C"], cfg["n"],
                         cfg["text_mode"], cfg["dtype"], train, flags)
        x = _check_input(img_feat.detach(), "img_feat", cfg["F"])
        t = _check_input(text_in.detach(), "text_metadata", cfg["T"] if cfg["text_mode"] else cfg["V"])
        params = self._params_in_slot_order()
        table = param_table([p.detach() if p is not None else None for p in params])
        dev = x.device
        ws = torch.empty(_lib.workspace_bytes(desc), dtype=torch.uint8, device=dev)
        total, offs = _lib.grad_layout(desc)
        if flat_out is not None:
            if flat_out.dtype != torch.float32 or flat_out.device != dev or flat_out.numel() < total or not flat_out.is_contiguous():
                raise ValueError(f"flat_out must be a contiguous fp32 tensor on {dev} with at least {total} elements")
            flat = flat_out
        else:
            flat = torch.empty(max(total, 1), dtype=torch.float32, device=dev)
        logits = torch.empty(B, cfg["C"], dtype=torch.float32, device=dev)
        loss_out = torch.empty(3, dtype=torch.float32, device=dev)
        d_img = torch.empty_like(x) if need_dimg else None
        d_txt = torch.empty_like(t) if need_dtxt else None
        y = label.to(device=dev, dtype=torch.int64).contiguous()
        w = None if class_weights is None else class_weights.to(device=dev, dtype=torch.float32).contiguous()
        denom = denom_arg(denom, dev)
        marr, _keep = _mask_table(self._injected_masks)
        if self._rng_state is None or self._rng_state.device != dev:
            # Philox key lives on the device so that a captured CUDA graph draws new masks on every replay
            self._rng_state = torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, 1], dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            ev = C.c_void_p(mid_event.cuda_event) if mid_event is not None else C.c_void_p(0)
            _lib.check(L.fb200_head_train_step_dp(C.byref(desc), table.arr, _ptr(x), _ptr(t), _ptr(y), _ptr(w), _ptr(denom), marr,
                                                  0, 0, _ptr(self._rng_state),
                                                  _ptr(logits), _ptr(loss_out), _ptr(flat), _ptr(d_img), _ptr(d_txt), _ptr(ws), _stream(), ev),
                       "fb200_head_train_step")
            if train:
                _lib.check(L.fb200_rng_advance(_ptr(self._rng_state), 1, _stream()), "fb200_rng_advance")
        for s, p in enumerate(params):
            if p is None or s not in offs or not p.requires_grad:
                continue
            g = flat[offs[s]: offs[s] + p.numel()].view(p.shape)
            if p.grad is None or zero_grad:
                p.grad = g
            else:
                p.grad = p.grad + g
        self.flat_grad = flat                # one contiguous bucket for the DP all-reduce
        self.last_desc = desc
        if need_dimg:
            img_feat.backward(d_img)
        if need_dtxt:
            text_in.backward(d_txt)
        return loss_out[0], logits


class GraphedTrainStep:
    """One CUDA graph = one whole train step of the head (weight prep, forward, weighted CE,
    backward, Philox advance) on fixed device buffers: ~70 kernel launches become one
    cudaGraphLaunch, which is what makes small batches (B = 32: microseconds of GPU work)
    anything but launch-bound.  Gradients land in ``model.flat_grad`` / the parameters' ``.grad``
    views, the loss in ``self.loss`` (device scalar), logits in ``self.logits``.

        step = GraphedTrainStep(model, x, meta, y, class_weights)   # x, meta, y: static CUDA tensors
        x.copy_(next_x); ...; step.run(); optimizer.step()
    """

    def __init__(self, model, image, text_metadata, label, class_weights=None, denom=None, warmup=2, mid_event=None, flat_out=None,
                 after_step=None):
        """``after_step``: optional callable run inside the capture right after the step (data parallel: the gradient
        all-reduce, so that one graph launch = step + collective)."""
        if any(p.requires_grad for p in model.image_encoder.parameters()):
            raise ValueError("GraphedTrainStep captures the head only: use a frozen backbone (or feed features)")
        self.model, self.args = model, (image, text_metadata, label, class_weights, denom, True, mid_event, flat_out)
        self.mid_event = mid_event
        # Engine for the captured step.  Eager calls of small fp32 batches (<= 32 rows) run the persistent step kernel: ONE launch
        # per pass is what a host-bound loop wants (B = 32, eager forward_loss: 0.29 ms against 0.44 ms with per-op launches).
        # Inside a graph launches are free, and the per-op tcgen05 kernels with cluster split-K are faster (0.178 against 0.194 ms,
        # cfg3a 0.162 against 0.184): capture those unless the caller chose an engine.
        chosen = _lib.FLAG_FORCE_SIMT | _lib.FLAG_FORCE_TC | _lib.FLAG_NO_MEGA | _lib.FLAG_FORCE_MEGA
        saved_flags = model.engine_flags
        if model.compute_dtype == "fp32" and not (saved_flags & chosen) and os.environ.get("FB200_GRAPH_ENGINE", "tc") == "tc":
            model.engine_flags = saved_flags | _lib.FLAG_FORCE_TC
        try:
            side = torch.cuda.Stream(device=image.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    model.forward_loss(*self.args)
                    if after_step is not None:
                        after_step(model.flat_grad)
            torch.cuda.current_stream().wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss, self.logits = model.forward_loss(*self.args)
                if after_step is not None:
                    after_step(model.flat_grad)
        finally:
            model.engine_flags = saved_flags
        self.flat_grad = model.flat_grad
        self.desc = model.last_desc

    def run(self):
        self.graph.replay()
        return self.loss
