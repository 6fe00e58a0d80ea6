// plan.h - the fusion head as a small op program.
//
// A fb200_desc is lowered once per call (microseconds, host only) into a list of ops over
// numbered activation buffers; the forward executor walks the list, the backward executor
// walks it in reverse.  Each fusion string of MultimodalModel.forward
// (multimodalIntraInterModal.py:205-416) is ~5 lines of builder calls, and S = 1 attention
// is lowered to its two live GEMMs (V-projection, output projection): softmax over one key
// is exactly 1, W_q / W_k only ever receive zero gradients (SURVEY.md "Facts").
#pragma once
#include <vector>
#include <string>
#include "common.cuh"

namespace fb200 {

constexpr int NUM_SLOTS = 78;

// parameter slots, reference state_dict order (multimodalIntraInterModal.py:55-160)
enum Slot : int {
  S_IMGPROJ_W = 0, S_IMGPROJ_B,
  S_TFC0_W, S_TFC0_B, S_TFC2_W, S_TFC2_B, S_TFC4_W, S_TFC4_B,
  S_TXTPROJ_W, S_TXTPROJ_B,
  S_ISA = 10,   // image_self_attention : in_proj_weight, in_proj_bias, out_proj.weight, out_proj.bias
  S_TSA = 14,   // text_self_attention
  S_ICA = 18,   // image_cross_attention
  S_TCA = 22,   // text_cross_attention
  S_IMGGATE_W = 26, S_IMGGATE_B, S_TXTGATE_W, S_TXTGATE_B,
  S_MB_FB_W = 30, S_MB_FB_B, S_MB_FB_LNW, S_MB_FB_LNB,
  S_MB_GB_W = 34, S_MB_GB_B, S_MB_GB_LNW, S_MB_GB_LNB,
  S_IRES = 38,  // image_residual : norm.weight, norm.bias, attn.{in_w,in_b,out_w,out_b}, gate_linear.{weight,bias}
  S_TRES = 46,  // text_residual
  S_FUSION = 54,   // fc_fusion : 0.w 0.b 1.w 1.b 4.w 4.b 5.w 5.b 8.w 8.b
  S_VISONLY_W = 64, S_VISONLY_B,
  S_F2O_W = 66, S_F2O_B,
  S_MBMLP = 68,    // fc_mlp_module_after_metablock_fusion_module (same 10-slot layout as fc_fusion)
};

extern const char* const kSlotNames[NUM_SLOTS];

struct Shape { int64_t rows, cols; bool present; };   // cols = 0 -> 1-D
Shape slot_shape(const fb200_desc& d, int slot);

enum OpKind : int { OP_LINEAR = 0, OP_LNRD, OP_GATE, OP_GRB, OP_META, OP_CAST };

struct View { int buf = -1; int col0 = 0; int cols = 0; };

struct Act {             // one activation buffer (and its gradient twin)
  int cols = 0;
  int ext = 0;           // 0: workspace, 1: img_feat, 2: text_in, 3: logits (gradient = dlogits)
  bool relu_out = false; // produced by Linear+ReLU: the gradient written into it must be masked by [value > 0]
  int vfmt = FMT_F32;    // storage format of the value: the GEMM operand format iff some Linear reads it, else fp32
  int gfmt = FMT_F32;    // storage format of the gradient: the operand format iff some Linear wrote the value (dY operand)
  size_t off = 0;        // byte offset of the value in the workspace
  size_t goff = 0;       // byte offset of the gradient
};

struct Op {
  int kind;
  View in0, in1, in2, out;     // LINEAR: in0=x ; GATE: in0=x in1=z ; GRB: in0=q in1=a in2=z ; META: in0=v in1=f in2=g
  int w_slot = -1, b_slot = -1; int w_row0 = 0;     // LINEAR: W = slot rows [w_row0, w_row0+out.cols), K = in0.cols
  int relu = 0;
  int ln_w[2] = {-1, -1}, ln_b[2] = {-1, -1};       // LayerNorm affine slots (META uses both pairs)
  int site = -1; float p = 0.f;                     // dropout site
  size_t stats_off = 0;                             // fp32 row statistics in the workspace
  int engine = 0;                                   // LINEAR: 0 SIMT, 1 tcgen05
  View dx_view;                                     // LINEAR: where dX goes (the fp32 external input when in0 is its operand-format copy)
  int wprep = -1;                                   // LINEAR on tcgen05: index into Plan::wprep (operand-format copy of W)
  int lane = 0;                                     // 1: depends on the metadata input only -> may run on the side stream
};

struct WPrep { int slot, row0, rows, cols; size_t off; };   // operand-format weight copy living in the workspace

struct Plan {
  fb200_desc d;
  bool use_tc = false;               // tcgen05 GEMMs enabled for this call
  bool two_lanes = false;            // image chain and metadata chain are launched on two streams (exec.cu)
  bool use_mega = false;             // small fp32 batches: the whole pass runs as ONE persistent cooperative kernel (mega.cuh)
  size_t mega_bar_off = 0;           // grid-barrier counter of that kernel in the workspace
  std::vector<WPrep> wprep;
  int fmt = FMT_F32;                 // GEMM operand format of workspace activations (F32 / PAIR / BF16)
  std::vector<Act> acts;
  std::vector<Op> ops;
  View logits;
  bool live[NUM_SLOTS] = {};
  int64_t goff[NUM_SLOTS];           // element offset in the flat gradient buffer, -1 if not live
  int64_t grad_elems = 0;
  int64_t dp_split = 0;              // element offset in the flat gradient buffer: tcgen05 weight gradients below it are launched
                                     // first (then fb200_head_train_step_dp records its event), the others after; 0 = no split
  size_t ws_bytes = 0;
  size_t splitk_off = 0, splitk_bytes = 0, counters_off = 0;   // FFMA split-K fix-up scratch (small batches)
  float drop_p[FB200_NUM_DROPOUT_SITES] = {};
  int drop_cols[FB200_NUM_DROPOUT_SITES] = {};
  double flops = 0, bytes = 0; int64_t live_params = 0;
  int fwd_launches = 0, bwd_launches = 0;
  std::string error;
};

// Returns FB200_OK or an error status; never throws.
int build_plan(const fb200_desc& d, Plan& p);

}  // namespace fb200
