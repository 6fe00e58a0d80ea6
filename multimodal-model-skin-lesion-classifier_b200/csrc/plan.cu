// plan.cu - lowers a fb200_desc to the op program (host only).
#include "plan.h"
#include <cstring>
#include <cstdlib>
#include <algorithm>

namespace fb200 {

const char* const kSlotNames[NUM_SLOTS] = {
  "image_projector.weight", "image_projector.bias",
  "text_fc.0.weight", "text_fc.0.bias", "text_fc.2.weight", "text_fc.2.bias", "text_fc.4.weight", "text_fc.4.bias",
  "text_projector.weight", "text_projector.bias",
  "image_self_attention.in_proj_weight", "image_self_attention.in_proj_bias", "image_self_attention.out_proj.weight", "image_self_attention.out_proj.bias",
  "text_self_attention.in_proj_weight", "text_self_attention.in_proj_bias", "text_self_attention.out_proj.weight", "text_self_attention.out_proj.bias",
  "image_cross_attention.in_proj_weight", "image_cross_attention.in_proj_bias", "image_cross_attention.out_proj.weight", "image_cross_attention.out_proj.bias",
  "text_cross_attention.in_proj_weight", "text_cross_attention.in_proj_bias", "text_cross_attention.out_proj.weight", "text_cross_attention.out_proj.bias",
  "img_gate.weight", "img_gate.bias", "txt_gate.weight", "txt_gate.bias",
  "meta_block.fb.0.weight", "meta_block.fb.0.bias", "meta_block.fb.1.weight", "meta_block.fb.1.bias",
  "meta_block.gb.0.weight", "meta_block.gb.0.bias", "meta_block.gb.1.weight", "meta_block.gb.1.bias",
  "image_residual.norm.weight", "image_residual.norm.bias", "image_residual.attn.in_proj_weight", "image_residual.attn.in_proj_bias",
  "image_residual.attn.out_proj.weight", "image_residual.attn.out_proj.bias", "image_residual.gate_linear.weight", "image_residual.gate_linear.bias",
  "text_residual.norm.weight", "text_residual.norm.bias", "text_residual.attn.in_proj_weight", "text_residual.attn.in_proj_bias",
  "text_residual.attn.out_proj.weight", "text_residual.attn.out_proj.bias", "text_residual.gate_linear.weight", "text_residual.gate_linear.bias",
  "fc_fusion.0.weight", "fc_fusion.0.bias", "fc_fusion.1.weight", "fc_fusion.1.bias", "fc_fusion.4.weight", "fc_fusion.4.bias",
  "fc_fusion.5.weight", "fc_fusion.5.bias", "fc_fusion.8.weight", "fc_fusion.8.bias",
  "fc_visual_only.weight", "fc_visual_only.bias",
  "fc_fusion_proj_feat2output.weight", "fc_fusion_proj_feat2output.bias",
  "fc_mlp_module_after_metablock_fusion_module.0.weight", "fc_mlp_module_after_metablock_fusion_module.0.bias",
  "fc_mlp_module_after_metablock_fusion_module.1.weight", "fc_mlp_module_after_metablock_fusion_module.1.bias",
  "fc_mlp_module_after_metablock_fusion_module.4.weight", "fc_mlp_module_after_metablock_fusion_module.4.bias",
  "fc_mlp_module_after_metablock_fusion_module.5.weight", "fc_mlp_module_after_metablock_fusion_module.5.bias",
  "fc_mlp_module_after_metablock_fusion_module.8.weight", "fc_mlp_module_after_metablock_fusion_module.8.bias",
};

static Shape lin(int64_t o, int64_t i, bool weight) { return weight ? Shape{o, i, true} : Shape{o, 0, true}; }

Shape slot_shape(const fb200_desc& d, int s) {
  const int64_t D = d.D, F = d.F, C = d.C, T = d.T, V = d.V;
  // meta_block(V_dim, U_dim): multimodalIntraInterModal.py:112-115
  const bool mb_common = d.mechanism == FB200_RGATT_FULL_METABLOCK;
  const int64_t mbV = mb_common ? D : F, mbU = mb_common ? D : T;
  const int64_t fuse_in = (d.mechanism == FB200_NO_METADATA ? 1 : d.n) * D;    // :124-126
  auto mha = [&](int k) -> Shape {    // in_w, in_b, out_w, out_b
    switch (k) { case 0: return {3 * D, D, true}; case 1: return {3 * D, 0, true}; case 2: return {D, D, true}; default: return {D, 0, true}; }
  };
  auto mlp = [&](int k, int64_t first_in) -> Shape {
    switch (k) {
      case 0: return {D, first_in, true}; case 1: return {D, 0, true};
      case 2: case 3: return {D, 0, true};
      case 4: return {D / 2, D, true}; case 5: return {D / 2, 0, true};
      case 6: case 7: return {D / 2, 0, true};
      case 8: return {C, D / 2, true}; default: return {C, 0, true};
    }
  };
  if (s < 0 || s >= NUM_SLOTS) return {0, 0, false};
  if (s <= S_IMGPROJ_B) return lin(D, F, s == S_IMGPROJ_W);
  if (s <= S_TFC4_B) {
    if (d.text_mode != 0) return {0, 0, false};
    if (s <= S_TFC0_B) return lin(256, V, s == S_TFC0_W);
    if (s <= S_TFC2_B) return lin(512, 256, s == S_TFC2_W);
    return lin(T, 512, s == S_TFC4_W);
  }
  if (s <= S_TXTPROJ_B) return lin(D, T, s == S_TXTPROJ_W);
  if (s < S_IMGGATE_W) return mha((s - S_ISA) % 4);
  if (s <= S_TXTGATE_B) return lin(D, D, (s - S_IMGGATE_W) % 2 == 0);
  if (s < S_IRES) {
    int k = (s - S_MB_FB_W) % 4;
    if (k == 0) return {mbV, mbU, true};
    return {mbV, 0, true};
  }
  if (s < S_FUSION) {
    int k = (s - S_IRES) % 8;
    if (k < 2) return {D, 0, true};
    if (k < 6) return mha(k - 2);
    return lin(D, D, k == 6);
  }
  if (s < S_VISONLY_W) return mlp(s - S_FUSION, fuse_in);
  if (s <= S_VISONLY_B) return lin(C, F, s == S_VISONLY_W);
  if (s <= S_F2O_B) return lin(C, D, s == S_F2O_W);
  return mlp(s - S_MBMLP, F);
}

namespace {

struct Builder {
  Plan& p;
  const fb200_desc& d;
  explicit Builder(Plan& pl) : p(pl), d(pl.d) {}

  int new_act(int cols, int ext = 0) {
    Act a; a.cols = cols; a.ext = ext;
    p.acts.push_back(a);
    return (int)p.acts.size() - 1;
  }
  View whole(int buf) { View v; v.buf = buf; v.col0 = 0; v.cols = p.acts[buf].cols; return v; }
  View half(int buf, int which, int cols) { View v; v.buf = buf; v.col0 = which * cols; v.cols = cols; return v; }
  void touch(int slot) { p.live[slot] = true; }

  // y = x W^T + b (optionally ReLU).  `dst` lets the producer write straight into a
  // concatenation half (torch.cat at :212-228 never materialises separately).
  int cast_of[3] = {-1, -1, -1};       // operand-format copies of the external inputs (built once)
  View linear(View x, int w_slot, int out_cols, bool relu = false, const View* dst = nullptr, int w_row0 = 0) {
    Op o; o.kind = OP_LINEAR; o.w_slot = w_slot; o.b_slot = w_slot + 1; o.w_row0 = w_row0; o.relu = relu ? 1 : 0;
    o.dx_view = x;
    // tcgen05 takes 16-byte global strides: both widths multiples of 8 (metadata widths 85/13/11 and the
    // class count stay on the FFMA kernel)
    o.engine = (p.use_tc && x.cols % 8 == 0 && out_cols % 8 == 0 && x.cols >= 32 && out_cols >= 32) ? 1 : 0;
    const int ext = p.acts[x.buf].ext;
    if (o.engine == 1 && ext != 0 && p.fmt == FMT_BF16) {   // bf16 operands: convert the fp32 input once (fp32-strict reads it as is)
      if (cast_of[ext] < 0) {
        cast_of[ext] = new_act(p.acts[x.buf].cols);
        Op c; c.kind = OP_CAST; c.in0 = whole(x.buf); c.out = whole(cast_of[ext]);
        p.ops.push_back(c);
      }
      View xc = x; xc.buf = cast_of[ext];
      x = xc;
    }
    o.in0 = x;
    if (o.engine == 1 && p.fmt == FMT_BF16) {       // bf16 copy of W, refreshed at the top of every forward
      int found = -1;
      for (size_t i = 0; i < p.wprep.size(); ++i) if (p.wprep[i].slot == w_slot && p.wprep[i].row0 == w_row0) found = (int)i;
      if (found < 0) { p.wprep.push_back(WPrep{w_slot, w_row0, out_cols, x.cols, 0}); found = (int)p.wprep.size() - 1; }
      o.wprep = found;
    }
    o.out = dst ? *dst : whole(new_act(out_cols));
    if (relu) p.acts[o.out.buf].relu_out = true;
    touch(w_slot); touch(w_slot + 1);
    p.ops.push_back(o);
    return o.out;
  }
  // nn.MultiheadAttention at S_q = S_kv = 1: out_proj(v_proj(kv)); base = first of its 4 slots
  View attn(int base, View kv, const View* dst = nullptr) {
    View a = linear(kv, base, d.D, false, nullptr, 2 * d.D);
    return linear(a, base + 2, d.D, false, dst);
  }
  View gate(View x, View z, const View* dst) {
    Op o; o.kind = OP_GATE; o.in0 = x; o.in1 = z; o.out = dst ? *dst : whole(new_act(x.cols));
    p.ops.push_back(o);
    return o.out;
  }
  // GatedAlteredResidualBlock.forward(q, k, v) with k = v (gatedResidualBlock.py:12-17); base = S_IRES / S_TRES
  View residual(int base, View q, View kv, int site, const View* dst = nullptr) {
    View a = attn(base + 2, kv);
    View z = linear(q, base + 6, d.D);
    Op o; o.kind = OP_GRB; o.in0 = q; o.in1 = a; o.in2 = z; o.out = dst ? *dst : whole(new_act(d.D));
    o.ln_w[0] = base; o.ln_b[0] = base + 1; o.site = site; o.p = 0.1f;
    touch(base); touch(base + 1);
    p.drop_p[site] = 0.1f; p.drop_cols[site] = d.D;
    p.ops.push_back(o);
    return o.out;
  }
  View metablock(View v, View u) {
    View f = linear(u, S_MB_FB_W, v.cols);
    View g = linear(u, S_MB_GB_W, v.cols);
    Op o; o.kind = OP_META; o.in0 = v; o.in1 = f; o.in2 = g; o.out = whole(new_act(v.cols));
    o.ln_w[0] = S_MB_FB_LNW; o.ln_b[0] = S_MB_FB_LNB; o.ln_w[1] = S_MB_GB_LNW; o.ln_b[1] = S_MB_GB_LNB;
    for (int s = S_MB_FB_LNW; s <= S_MB_FB_LNB; ++s) touch(s);
    for (int s = S_MB_GB_LNW; s <= S_MB_GB_LNB; ++s) touch(s);
    p.ops.push_back(o);
    return o.out;
  }
  View lnrd(View x, int lnw, int site, float pdrop) {
    Op o; o.kind = OP_LNRD; o.in0 = x; o.out = whole(new_act(x.cols));
    o.ln_w[0] = lnw; o.ln_b[0] = lnw + 1; o.site = site; o.p = pdrop;
    touch(lnw); touch(lnw + 1);
    p.drop_p[site] = pdrop; p.drop_cols[site] = x.cols;
    p.ops.push_back(o);
    return o.out;
  }
  // fc_mlp_module / fc_mlp_module_after_metablock (:134-160); base = S_FUSION / S_MBMLP
  View mlp(int base, View x, float pdrop, const View& logits) {
    View h = linear(x, base + 0, d.D);
    h = lnrd(h, base + 2, FB200_DROP_FC1, pdrop);
    h = linear(h, base + 4, d.D / 2);
    h = lnrd(h, base + 6, FB200_DROP_FC2, pdrop);
    return linear(h, base + 8, d.C, false, &logits);
  }
};

}  // namespace

int build_plan(const fb200_desc& d, Plan& p) {
  p = Plan();
  p.d = d;
  if (d.B < 1 || d.F < 1 || d.C < 1 || d.D < 8 || d.T < 1 || d.H < 1) { p.error = "non-positive dimension"; return FB200_EBADARG; }
  if (d.mechanism < 0 || d.mechanism >= FB200_NUM_MECHANISMS) { p.error = "unknown mechanism"; return FB200_EBADARG; }
  if (d.text_mode == 0 && d.V < 1) { p.error = "V must be >= 1 for one-hot metadata"; return FB200_EBADARG; }
  if (d.D % d.H != 0) { p.error = "embed_dim must be divisible by num_heads"; return FB200_EBADARG; }
  if (d.D % 8 != 0) { p.error = "common_dim must be a multiple of 8 (GatedAlteredResidualBlock uses 8 heads)"; return FB200_EUNSUPPORTED; }
  if (d.F % 4 != 0 || d.F > 4096 || d.D > 4096) { p.error = "F must be a multiple of 4 and F, D <= 4096"; return FB200_EUNSUPPORTED; }
  if (d.dtype != FB200_F32 && d.dtype != FB200_BF16) { p.error = "dtype"; return FB200_EBADARG; }
  if (d.mechanism != FB200_NO_METADATA && d.mechanism != FB200_NO_METADATA_WITHOUT_MLP && d.mechanism != FB200_METABLOCK &&
      d.mechanism != FB200_RGATT2FUSEFEATURES && d.mechanism != FB200_RGATT_FULL_RGATT2FUSE && d.mechanism != FB200_RGATT_FULL_METABLOCK && d.n != 2) {
    p.error = "fusion strings that concatenate two modalities need n = 2"; return FB200_EBADARG;
  }
  // engine policy: bf16 always rides the tensor cores; fp32 switches to the 3xTF32 tensor path above 32 rows (measured
  // with the r01 kernels: 0.365 vs 0.371 ms per step at B=32, 0.374 vs 0.402 at 64, 0.391 vs 0.494 at 128; up to 32 rows
  // the exact FFMA kernel with split-K fix-up costs the same and is bit-faithful fp32)
  // fp32 batches up to 64 rows (the reference's own BATCH_SIZE=32, conf/.env.test:2): one persistent cooperative kernel
  // walks the whole op program (mega.cuh) - exact FFMA arithmetic, one launch per pass.  FB200_MEGA=0 / FB200_FLAG_NO_MEGA
  // keep the per-op kernels (A/B measurements, and the FFMA / tcgen05 engines stay covered by the tests through FORCE_*).
  static const bool mega_env_off = [] { const char* e = getenv("FB200_MEGA"); return e && e[0] == '0'; }();
  // r02d: with cluster split-K the tcgen05 path takes 33 .. 64 rows in 0.21 ms against 0.25 - 0.26 ms for the step kernel (which
  // walks two 32-row groups there), so the step kernel keeps batches up to 32 rows; FB200_FLAG_FORCE_MEGA restores its full range.
  static const int mega_rows_env = [] { const char* e = getenv("FB200_MEGA_ROWS"); return e ? atoi(e) : 32; }();   // A/B: largest batch of the step kernel
  static const int tc_min_env = [] { const char* e = getenv("FB200_TC_MIN"); return e ? atoi(e) : 32; }();         // A/B: tcgen05 above this many rows
  const int mega_rows = (d.flags & FB200_FLAG_FORCE_MEGA) ? 64 : mega_rows_env;
  p.use_mega = d.dtype == FB200_F32 && d.B <= mega_rows && !(d.flags & (FB200_FLAG_FORCE_SIMT | FB200_FLAG_FORCE_TC | FB200_FLAG_NO_MEGA)) && !mega_env_off;
  p.use_tc = !p.use_mega && !(d.flags & FB200_FLAG_FORCE_SIMT) && ((d.flags & FB200_FLAG_FORCE_TC) || d.dtype == FB200_BF16 || d.B > tc_min_env);
  p.fmt = d.dtype == FB200_BF16 ? FMT_BF16 : FMT_F32;    // fp32-strict keeps everything fp32 in memory (hi/lo split happens in smem)

  Builder b(p);
  const int D = d.D;
  const int X = b.new_act(d.F, 1);
  const int TIN = b.new_act(d.text_mode == 0 ? d.V : d.T, 2);
  const int LOG = b.new_act(d.C, 3);
  const View x = b.whole(X), tin = b.whole(TIN), logits = b.whole(LOG);
  p.logits = logits;
  const int m = d.mechanism;

  // lazily built common prefix (:172-197); dead branches of the reference forward are skipped:
  // their parameters keep .grad = None there, and they do not influence the logits.
  View p_img, p_txt, txt_feat, img_att, txt_att;
  auto need_pimg = [&](const View* dst = nullptr) { if (p_img.buf < 0) p_img = b.linear(x, S_IMGPROJ_W, D, false, dst); return p_img; };
  auto need_txtfeat = [&]() {
    if (txt_feat.buf < 0) {
      if (d.text_mode == 0) {
        View h = b.linear(tin, S_TFC0_W, 256, true);
        h = b.linear(h, S_TFC2_W, 512, true);
        txt_feat = b.linear(h, S_TFC4_W, d.T);
      } else txt_feat = tin;
    }
    return txt_feat;
  };
  auto need_ptxt = [&](const View* dst = nullptr) { if (p_txt.buf < 0) p_txt = b.linear(need_txtfeat(), S_TXTPROJ_W, D, false, dst); return p_txt; };
  auto need_iatt = [&](const View* dst = nullptr) { if (img_att.buf < 0) img_att = b.attn(S_ISA, need_pimg(), dst); return img_att; };
  auto need_tatt = [&](const View* dst = nullptr) { if (txt_att.buf < 0) txt_att = b.attn(S_TSA, need_ptxt(), dst); return txt_att; };

  auto cat_buf = [&]() { return b.new_act(2 * D); };

  switch (m) {
    case FB200_NO_METADATA:
      b.mlp(S_FUSION, need_pimg(), 0.5f, logits);
      break;
    case FB200_NO_METADATA_WITHOUT_MLP:
      b.linear(x, S_VISONLY_W, d.C, false, &logits);
      break;
    case FB200_CONCATENATION: {
      int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
      need_pimg(&l); need_ptxt(&r);
      b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
    } break;
    case FB200_ATT_INTRAMODAL: {
      int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
      need_iatt(&l); need_tatt(&r);
      b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
    } break;
    case FB200_CROSSATTENTION: case FB200_GFCAM: case FB200_CROSS_WEIGHTS_AFTER_CROSSATT: {
      int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
      View ia = need_iatt(), ta = need_tatt();
      const bool plain = (m == FB200_CROSSATTENTION);
      View ic = b.attn(S_ICA, ta, plain ? &l : nullptr);
      View tc = b.attn(S_TCA, ia, plain ? &r : nullptr);
      if (!plain) {
        View zi = b.linear(ic, S_IMGGATE_W, D), zt = b.linear(tc, S_TXTGATE_W, D);
        if (m == FB200_GFCAM) { b.gate(ic, zi, &l); b.gate(tc, zt, &r); }
        else                  { b.gate(ic, zt, &l); b.gate(tc, zi, &r); }     // :231-235 swaps the gates
      }
      b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
    } break;
    case FB200_WEIGHTED: {
      int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
      View pi = need_pimg(), pt = need_ptxt();
      View zi = b.linear(pi, S_IMGGATE_W, D), zt = b.linear(pt, S_TXTGATE_W, D);
      b.gate(pi, zi, &l); b.gate(pt, zt, &r);
      b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
    } break;
    case FB200_METABLOCK: {
      View y = b.metablock(x, need_txtfeat());
      b.mlp(S_MBMLP, y, 0.3f, logits);
    } break;
    case FB200_RGATT2FUSEFEATURES: {
      View r = b.residual(S_IRES, need_ptxt(), need_pimg(), FB200_DROP_IMG_RES);
      b.linear(r, S_F2O_W, d.C, false, &logits);
    } break;
    case FB200_RG_ATT: case FB200_ATT_INTRAMODAL_RESIDUAL: {
      int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
      View pi = need_pimg(), pt = need_ptxt();
      View kvi = (m == FB200_RG_ATT) ? pt : need_iatt();
      View kvt = (m == FB200_RG_ATT) ? pi : need_tatt();
      b.residual(S_IRES, pi, kvi, FB200_DROP_IMG_RES, &l);
      b.residual(S_TRES, pt, kvt, FB200_DROP_TXT_RES, &r);
      b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
    } break;
    case FB200_CROSS_ATTENTION_ONLY: {
      int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
      View pi = need_pimg(), pt = need_ptxt();
      b.attn(S_ICA, pt, &l); b.attn(S_TCA, pi, &r);
      b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
    } break;
    case FB200_RESIDUAL_CROSSATT: case FB200_RGATT_FULL: case FB200_RGATT_FULL_RGATT2FUSE:
    case FB200_RGATT_FULL_METABLOCK: case FB200_RGATT_FULL_INTRAMODAL_RES: {
      View pi = need_pimg(), pt = need_ptxt();
      View kvi = (m == FB200_RESIDUAL_CROSSATT) ? pi : need_iatt();
      View kvt = (m == FB200_RESIDUAL_CROSSATT) ? pt : need_tatt();
      View ir = b.residual(S_IRES, pi, kvi, FB200_DROP_IMG_RES);
      View tr = b.residual(S_TRES, pt, kvt, FB200_DROP_TXT_RES);
      if (m == FB200_RESIDUAL_CROSSATT || m == FB200_RGATT_FULL) {
        int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
        b.attn(S_ICA, tr, &l); b.attn(S_TCA, ir, &r);
        b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
      } else {
        View ic = b.attn(S_ICA, tr), tc = b.attn(S_TCA, ir);
        if (m == FB200_RGATT_FULL_RGATT2FUSE) {
          View r2 = b.residual(S_IRES, tc, ic, FB200_DROP_IMG_RES2);
          b.linear(r2, S_F2O_W, d.C, false, &logits);
        } else if (m == FB200_RGATT_FULL_METABLOCK) {
          View y = b.metablock(ic, tc);
          b.linear(y, S_F2O_W, d.C, false, &logits);
        } else {
          int cb = cat_buf(); View l = b.half(cb, 0, D), r = b.half(cb, 1, D);
          View ia2 = b.attn(S_ISA, ic), ta2 = b.attn(S_TSA, tc);
          b.residual(S_IRES, ic, ia2, FB200_DROP_IMG_RES2, &l);
          b.residual(S_TRES, tc, ta2, FB200_DROP_TXT_RES2, &r);
          b.mlp(S_FUSION, b.whole(cb), 0.5f, logits);
        }
      }
    } break;
    default:
      p.error = "mechanism not implemented"; return FB200_EUNSUPPORTED;
  }

  // ---- lanes: the image chain and the metadata chain of a fusion string are independent until the concatenation
  //      (or the first op that reads both).  Ops whose inputs derive from the metadata input alone get lane 1; the
  //      executor launches them on a side stream: at large batches the two chains of 128-CTA GEMMs fill each other's
  //      idle SMs and launch gaps, at small batches (launch-latency-bound kernels of a few CTAs) they simply run side by side.
  {
    std::vector<int> color(p.acts.size(), 0);          // bit 0: image input, bit 1: metadata input
    color[X] = 1; color[TIN] = 2;
    int n1 = 0;
    for (auto& o : p.ops) {
      int c = 0;
      for (const View* v : {&o.in0, &o.in1, &o.in2}) if (v->buf >= 0) c |= color[v->buf];
      o.lane = (c == 2) ? 1 : 0;
      n1 += o.lane;
      color[o.out.buf] |= c;
    }
    static const bool env_off = [] { const char* e = getenv("FB200_LANES"); return e && e[0] == '0'; }();   // A/B measurements
    p.two_lanes = n1 > 0 && n1 < (int)p.ops.size() && !(d.flags & FB200_FLAG_ONE_STREAM) && !env_off && !p.use_mega;
    if (!p.two_lanes) for (auto& o : p.ops) o.lane = 0;
  }

  // ---- validate slots exist, lay out the flat gradient buffer (slot order)
  int64_t off = 0;
  for (int s = 0; s < NUM_SLOTS; ++s) {
    p.goff[s] = -1;
    if (!p.live[s]) continue;
    Shape sh = slot_shape(d, s);
    if (!sh.present) { p.error = std::string("parameter absent in this configuration: ") + kSlotNames[s]; return FB200_EBADARG; }
    p.goff[s] = off;
    int64_t n = sh.rows * (sh.cols ? sh.cols : 1);
    off += (n + 3) & ~int64_t(3);                      // keep every gradient 16-byte aligned
  }
  p.grad_elems = off;

  // ---- data-parallel bucket split (fb200_head_train_step_dp): the weight gradients of the tcgen05 Linears are all produced by
  //      the grouped launch that ends the backward pass.  Split them by offset into two halves of about equal tile count:
  //      the caller all-reduces everything below dp_split while the second half is still being computed.
  {
    struct W { int64_t off; int tiles; };
    std::vector<W> ws;
    bool twice = false;
    for (auto& o : p.ops) if (o.kind == OP_LINEAR && o.engine == 1) {
      const int64_t off = p.goff[o.w_slot] + (int64_t)o.w_row0 * o.in0.cols;
      for (auto& w : ws) if (w.off == off) twice = true;
      ws.push_back(W{off, ((o.out.cols + 127) / 128) * ((o.in0.cols + 127) / 128)});
    }
    p.dp_split = 0;
    if (!twice && ws.size() >= 2) {
      std::sort(ws.begin(), ws.end(), [](const W& a, const W& b) { return a.off < b.off; });
      int total = 0; for (auto& w : ws) total += w.tiles;
      int acc = 0;
      for (size_t i = 0; i + 1 < ws.size(); ++i) {
        acc += ws[i].tiles;
        if (2 * acc >= total) { p.dp_split = ws[i + 1].off; break; }
      }
    }
  }

  // ---- per-buffer storage format: only GEMM operands pay for the operand format (bf16 / tf32 pair);
  //      everything the row kernels exchange among themselves stays fp32
  for (auto& o : p.ops) if (o.kind == OP_LINEAR) {
    if (!p.acts[o.in0.buf].ext) p.acts[o.in0.buf].vfmt = p.fmt;
    if (!p.acts[o.out.buf].ext) p.acts[o.out.buf].gfmt = p.fmt;
  }
  // ---- workspace layout: values, then gradients, then row statistics
  auto align = [](size_t v) { return (v + 255) & ~size_t(255); };
  auto fbytes = [](int fmt) -> size_t { return fmt == FMT_BF16 ? 2 : (fmt == FMT_PAIR ? 8 : 4); };
  size_t cur = 0;
  for (auto& a : p.acts) {
    if (a.ext) continue;
    if (a.cols % 4 != 0) { p.error = "internal activation width must be a multiple of 4"; return FB200_EUNSUPPORTED; }
    a.off = cur; cur = align(cur + (size_t)d.B * a.cols * fbytes(a.vfmt));
  }
  for (auto& a : p.acts) {
    if (a.ext) continue;
    a.goff = cur; cur = align(cur + (size_t)d.B * a.cols * fbytes(a.gfmt));
  }
  for (auto& w : p.wprep) { w.off = cur; cur = align(cur + (size_t)w.rows * w.cols * fbytes(p.fmt)); }
  for (auto& o : p.ops) {
    if (o.kind == OP_LNRD || o.kind == OP_GRB) { o.stats_off = cur; cur = align(cur + (size_t)d.B * 2 * sizeof(float)); }
    if (o.kind == OP_META) { o.stats_off = cur; cur = align(cur + (size_t)d.B * 4 * sizeof(float)); }
  }
  // small batches on the FFMA path: split-K fix-up scratch (partial tiles) + per-tile arrival counters
  if (!p.use_tc || d.B <= 128) { p.splitk_off = cur; p.splitk_bytes = (size_t)16 << 20; cur = align(cur + p.splitk_bytes); p.counters_off = cur; cur = align(cur + 4096 * sizeof(unsigned)); }
  p.mega_bar_off = cur; cur = align(cur + 256);      // step-kernel barrier / fused-tail scalar scratch
  // tail: dlogits of the fused train step (exec.cu addresses it from the end)
  cur = align(cur + (size_t)d.B * d.C * sizeof(float));
  p.ws_bytes = cur + 256;

  // ---- algorithmic work: FLOPs of the live GEMMs, minimal HBM bytes
  const bool need_dimg = d.flags & FB200_FLAG_NEED_DIMG, need_dtxt = d.flags & FB200_FLAG_NEED_DTEXT;
  double macs_fwd = 0, macs_dx = 0; int64_t plive = 0;
  bool counted[NUM_SLOTS] = {};
  for (auto& o : p.ops) {
    if (o.kind == OP_LINEAR) {
      double mk = (double)o.in0.cols * o.out.cols;
      macs_fwd += mk;
      const int ext = p.acts[o.dx_view.buf].ext;
      if (ext == 0 || (ext == 1 && need_dimg) || (ext == 2 && need_dtxt)) macs_dx += mk;
      if (!counted[o.w_slot]) { counted[o.w_slot] = true; plive += (int64_t)o.in0.cols * o.out.cols + o.out.cols; }   // only the V third of in_proj is live
    } else if (o.kind != OP_CAST) {
      for (int k = 0; k < 2; ++k) if (o.ln_w[k] >= 0 && !counted[o.ln_w[k]]) { counted[o.ln_w[k]] = true; plive += 2 * (int64_t)o.out.cols; }
    }
  }
  p.live_params = plive;
  p.flops = 2.0 * d.B * (2.0 * macs_fwd + macs_dx);
  const double tin_cols = p.acts[TIN].cols;
  p.bytes = 3.0 * plive * 4 + (double)d.B * (d.F + tin_cols) * 4 + (need_dimg ? (double)d.B * d.F * 4 : 0) + 2.0 * d.B * d.C * 4;
  return FB200_OK;
}

}  // namespace fb200
