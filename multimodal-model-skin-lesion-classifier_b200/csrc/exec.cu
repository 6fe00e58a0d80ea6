// exec.cu - forward / backward executors over the op program, and the C ABI.
#include "plan.h"
#include "simt_gemm.cuh"
#include "rowwise.cuh"
#include "tc_gemm.cuh"
#include "attention.cuh"
#include "attention_tc.cuh"
#include "mega.cuh"
#include "dp_comm.cuh"
#include <cstdio>
#include <cstring>
#include <mutex>
#include <unordered_set>
#include <unordered_map>
#include <algorithm>
#include <cmath>

namespace fb200 {

struct DeviceInfo { int num_sms = 0; int cc_major = 0; };
static int get_device_info(DeviceInfo& out) {
  static std::mutex mu;
  static DeviceInfo cache[64];
  static bool have[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return FB200_ECUDA;
  std::lock_guard<std::mutex> g(mu);
  if (!have[dev]) {
    int sms = 0, maj = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return FB200_ECUDA;
    if (cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return FB200_ECUDA;
    cache[dev].num_sms = sms; cache[dev].cc_major = maj; have[dev] = true;
  }
  out = cache[dev];
  return FB200_OK;
}

static bool is_device_ptr(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// cudaPointerGetAttributes costs about a microsecond; parameters, class weights and masks are the same allocations call after
// call, so pointers that passed once are remembered (a freed-and-reused address is still a device address or the
// allocator would not have handed it out to a CUDA tensor - the check guards against HOST pointers, not lifetimes).
static bool is_device_ptr_cached(const void* p) {
  static std::mutex mu;
  static std::unordered_set<const void*> seen;
  {
    std::lock_guard<std::mutex> g(mu);
    if (seen.count(p)) return true;
  }
  if (!is_device_ptr(p)) return false;
  std::lock_guard<std::mutex> g(mu);
  if (seen.size() < (size_t)1 << 16) seen.insert(p);
  return true;
}

#define CUDA_OK(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) { cudaGetLastError(); return FB200_ECUDA; } } while (0)

struct Ctx {
  const Plan& p;
  const void* const* params;
  const void* img; const void* txt;
  void* logits; const void* dlogits;
  void* d_img; void* d_txt;
  float* grads;
  const uint8_t* const* masks; uint64_t seed, offset;
  const uint64_t* rng_state;
  char* ws;
  cudaStream_t st;
  DeviceInfo dev;
  bool gemm_only = false;     // debug replay: launch only the GEMM kernels of the step (bench.py times them with CUDA events)
  bool dry_run = false;       // host only: build the step-kernel program without touching CUDA (fb200_mega_program_info)
  // two-lane execution (Plan::two_lanes): lane 0 = the caller's stream, lane 1 = a side stream for the metadata chain
  cudaStream_t lane_st[2] = {nullptr, nullptr};
  void* mid_event = nullptr;  // fb200_head_train_step_dp: recorded once every gradient below Plan::dp_split is final
  // persistent step kernel (Plan::use_mega): the executors emit ops into `mega` instead of launching kernels.  Stage in which
  // a buffer is complete: vready per activation value, gready per activation gradient, pready per parameter gradient.
  // large batches, fused train step: the classifier tail (last LayerNorm -> head -> cross entropy -> their backward) is ONE launch
  bool tail_fuse = false; int tail_ln_op = -1, tail_head_op = -1; bool tail_have_fwd = false;
  TailArgs tail{};
  std::vector<GemmArgs> deferred_simt_dw;      // FFMA weight gradients of a tcgen05 step: launched beside the grouped dW, not in the dX chain
  MegaBuilder* mega = nullptr;
  std::vector<int> vready, gready; int pready[NUM_SLOTS] = {};
  int vr(int b) const { return b >= 0 ? vready[b] : 0; }
  int gr(int b) const { return b >= 0 ? gready[b] : 0; }
  // Row ops that only depend on the SAME rows of the row op emitted right before them (LayerNorm -> classifier head -> cross
  // entropy -> their backward) run chained inside one task of the step kernel instead of one grid-wide stage each.
  struct LastRow { size_t nrow = 0, ngemm = 0; int stage = -1, tiles = 0; int vbuf = -1; int gbuf[3] = {-1, -1, -1}; } lastrow;
  struct Dep { int ready; int buf; bool is_grad; };
  // stage of a warp-per-row op with the given dependencies; `chain` = it can run in the previous row op's task
  int row_stage(std::initializer_list<Dep> deps, int tiles, bool mappable, bool& chain) const {
    int all = 0, other = 0; bool hit = false;
    for (const Dep& d : deps) {
      if (d.buf < 0) continue;
      all = std::max(all, d.ready);
      const bool from_last = d.is_grad ? (d.buf == lastrow.gbuf[0] || d.buf == lastrow.gbuf[1] || d.buf == lastrow.gbuf[2]) : (d.buf == lastrow.vbuf);
      if (from_last && d.ready == lastrow.stage) hit = true; else other = std::max(other, d.ready);
    }
    chain = mega && mappable && hit && lastrow.stage >= 1 && lastrow.nrow == mega->r.size() && lastrow.ngemm == mega->g.size() &&
            tiles == lastrow.tiles && other <= lastrow.stage - 1;
    return chain ? lastrow.stage : 1 + all;
  }
  void row_emitted(int stage, int tiles, bool mappable, int vbuf, int g0 = -1, int g1 = -1, int g2 = -1) {
    lastrow.nrow = mega->r.size(); lastrow.ngemm = mega->g.size(); lastrow.stage = mappable ? stage : -1; lastrow.tiles = tiles;
    lastrow.vbuf = vbuf; lastrow.gbuf[0] = g0; lastrow.gbuf[1] = g1; lastrow.gbuf[2] = g2;
  }

  TRef value(const View& v) const {
    const Act& a = p.acts[v.buf];
    if (a.ext == 1) return make_ref((float*)img + v.col0, a.cols, FMT_F32);
    if (a.ext == 2) return make_ref((float*)txt + v.col0, a.cols, FMT_F32);
    if (a.ext == 3) return make_ref((float*)logits + v.col0, a.cols, FMT_F32);
    return make_ref(ws + a.off + (size_t)v.col0 * fmt_bytes(a.vfmt), a.cols, a.vfmt, (int64_t)p.d.B * a.cols);
  }
  TRef grad(const View& v) const {
    const Act& a = p.acts[v.buf];
    if (a.ext == 1) return make_ref(d_img ? (float*)d_img + v.col0 : nullptr, a.cols, FMT_F32);
    if (a.ext == 2) return make_ref(d_txt ? (float*)d_txt + v.col0 : nullptr, a.cols, FMT_F32);
    if (a.ext == 3) return make_ref((float*)dlogits + v.col0, a.cols, FMT_F32);
    return make_ref(ws + a.goff + (size_t)v.col0 * fmt_bytes(a.gfmt), a.cols, a.gfmt, (int64_t)p.d.B * a.cols);
  }
  const float* param(int slot, int64_t elem_off = 0) const { return (const float*)params[slot] + elem_off; }
  float* pgrad(int slot, int64_t elem_off = 0) const { return grads + p.goff[slot] + elem_off; }
  void enable_fixup(GemmArgs& g, int lane = 0) const {       // small-batch FFMA GEMMs: split K across CTAs, last slice fixes up
    if (p.splitk_bytes == 0 || g.M > 128) return;
    // each lane owns half of the scratch and of the arrival counters: GEMMs of the two lanes run concurrently
    const size_t half = p.splitk_bytes / 2;
    g.partial = (float*)(ws + p.splitk_off + (size_t)lane * half); g.partial_bytes = half;
    g.counters = (unsigned*)(ws + p.counters_off) + lane * 2048;
  }
  DropSpec drop(const Op& o) const {
    DropSpec s; s.mask = nullptr; s.seed = seed; s.offset = offset; s.state = rng_state; s.p = o.p; s.site = o.site; s.active = 0;
    if (p.d.train && o.site >= 0 && o.p > 0.f) { s.active = 1; if (masks && masks[o.site]) s.mask = masks[o.site]; }
    return s;
  }
};


// One launch converts every weight the tcgen05 GEMMs read into the operand format (bf16, or
// tf32 hi/lo planes): weights change every optimizer step, so this runs at the top of forward.
struct PrepSeg { const float* src; void* dst; int64_t n4; int64_t plane; };
struct PrepArgs { PrepSeg seg[32]; int nseg; int fmt; };
__global__ void __launch_bounds__(256) wprep_kernel(const PrepArgs a) { pdl_sync();
  const PrepSeg sg = a.seg[blockIdx.y];
  const TRef out = make_ref(sg.dst, 0, a.fmt, sg.plane);          // flat view: row 0, column = element index
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < sg.n4; i += (int64_t)gridDim.x * blockDim.x)
    st4(out, 0, (int)(i * 4), __ldg((const float4*)sg.src + i));
}

static TcOperand tc_operand(const TRef& r, int inner, int outer) {
  TcOperand o; o.base = r.p; o.plane_elems = r.plane; o.ld = r.ld; o.inner = inner; o.outer = outer; return o;
}

// ------------------------------------------------------------------------------- lanes
// Side stream and event pool of the calling host thread (forward runs on the main thread, backward on autograd's
// worker).  Events only order work between the two lanes INSIDE one call: every call forks the side stream from the
// caller's stream and joins it back before returning, so callers (and CUDA-graph capture) see one stream.
struct LaneRes {
  cudaStream_t side[64] = {};
  std::vector<cudaEvent_t> pool[64];
};
static thread_local LaneRes t_lanes;
struct LaneSync {
  bool on = false; int dev = 0; size_t next = 0;
  cudaStream_t st[2] = {nullptr, nullptr};
  std::vector<cudaEvent_t> ev[2];        // per buffer: last event recorded on each lane (nullptr: none)
  int begin(Ctx& c, size_t nbuf) {
    on = false;
    c.lane_st[0] = c.lane_st[1] = c.st;
    if (!c.p.two_lanes) return FB200_OK;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return FB200_ECUDA;
    if (!t_lanes.side[dev]) CUDA_OK(cudaStreamCreateWithFlags(&t_lanes.side[dev], cudaStreamNonBlocking));
    st[0] = c.st; st[1] = t_lanes.side[dev];
    ev[0].assign(nbuf, nullptr); ev[1].assign(nbuf, nullptr);
    next = 0; on = true;
    cudaEvent_t e; int rc = record(0, e); if (rc != FB200_OK) return rc;      // fork: the side stream starts after everything queued so far
    CUDA_OK(cudaStreamWaitEvent(st[1], e, 0));
    c.lane_st[1] = st[1];
    return FB200_OK;
  }
  int record(int lane, cudaEvent_t& e) {
    auto& pool = t_lanes.pool[dev];
    if (next == pool.size()) { cudaEvent_t n; CUDA_OK(cudaEventCreateWithFlags(&n, cudaEventDisableTiming)); pool.push_back(n); }
    e = pool[next++];
    CUDA_OK(cudaEventRecord(e, st[lane]));
    return FB200_OK;
  }
  // before an op on `lane` touches buffer b: wait for the other lane's last op on it
  int wait_for(int lane, int b) {
    if (!on || b < 0) return FB200_OK;
    cudaEvent_t e = ev[1 - lane][b];
    if (e) CUDA_OK(cudaStreamWaitEvent(st[lane], e, 0));
    return FB200_OK;
  }
  // after an op on `lane` touched the buffers bs
  int touched(int lane, std::initializer_list<int> bs) {
    if (!on) return FB200_OK;
    cudaEvent_t e; int rc = record(lane, e); if (rc != FB200_OK) return rc;
    for (int b : bs) if (b >= 0) ev[lane][b] = e;
    return FB200_OK;
  }
  int join(Ctx& c) {
    if (!on) return FB200_OK;
    cudaEvent_t e; int rc = record(1, e); if (rc != FB200_OK) return rc;
    CUDA_OK(cudaStreamWaitEvent(st[0], e, 0));
    on = false; c.st = st[0];
    return FB200_OK;
  }
};

// ------------------------------------------------------------------------------- forward
// The [N,K] weight operand of a tcgen05 Linear: the bf16 copy in the workspace, or the fp32 master itself.
static TRef weight_operand(const Ctx& c, const Op& o) {
  if (o.wprep >= 0) { const WPrep& w = c.p.wprep[o.wprep]; return make_ref(c.ws + w.off, w.cols, c.p.fmt, (int64_t)w.rows * w.cols); }
  return make_ref((void*)c.param(o.w_slot, (int64_t)o.w_row0 * o.in0.cols), o.in0.cols, FMT_F32);
}

static int run_forward(Ctx& c) {
  const Plan& p = c.p; const int B = p.d.B;
  if (c.mega) { c.vready.assign(p.acts.size(), 0); }
  if (p.splitk_bytes && !c.gemm_only && !c.mega) CUDA_OK(cudaMemsetAsync(c.ws + p.counters_off, 0, 4096 * sizeof(unsigned), c.st));
  // bf16: operand-format copies of the weights (wprep_kernel), on the caller's stream before the lanes fork.  FB200_WPREP_SIDE=1
  // makes them on the SIDE lane instead (the main lane starts with the format copy of the image features and only its first
  // tcgen05 GEMM needs the copies; the side lane starts with the FFMA GEMM of the raw metadata).  Measured mixed on the same box
  // (r02e, profiles/r02e_wprep.txt: cfg5 B = 4096 0.516 -> 0.511 ms, cfg4b 0.448 -> 0.452, B = 256 +0.3-1.4 %), so it stays off.
  auto launch_wprep = [&](cudaStream_t st) -> int {
    for (size_t base = 0; base < p.wprep.size(); base += 32) {
      PrepArgs a{}; a.fmt = p.fmt; a.nseg = 0;
      for (size_t i = base; i < p.wprep.size() && a.nseg < 32; ++i) {
        const WPrep& w = p.wprep[i];
        PrepSeg& sg = a.seg[a.nseg++];
        sg.src = c.param(w.slot, (int64_t)w.row0 * w.cols); sg.dst = c.ws + w.off; sg.n4 = (int64_t)w.rows * w.cols / 4; sg.plane = (int64_t)w.rows * w.cols;
      }
      if (!c.gemm_only) pdl_launch(wprep_kernel, dim3(64, a.nseg), 256, 0, st, a);
      CUDA_OK(cudaGetLastError());
    }
    return FB200_OK;
  };
  static const bool wprep_side_env = [] { const char* e = getenv("FB200_WPREP_SIDE"); return e && atoi(e) != 0; }();
  const bool wprep_side = wprep_side_env && p.two_lanes && !p.wprep.empty() && !c.mega;
  if (!wprep_side) { int rc = launch_wprep(c.st); if (rc != FB200_OK) return rc; }
  LaneSync ls; const cudaStream_t main_st = c.st;
  { int rc = ls.begin(c, p.acts.size()); if (rc != FB200_OK) return rc; }
  cudaEvent_t wprep_done = nullptr;           // recorded on the side lane; the main lane waits for it before its first tcgen05 GEMM
  if (wprep_side) {
    if (!ls.on) return FB200_ECUDA;
    int rc = launch_wprep(c.lane_st[1]); if (rc != FB200_OK) return rc;
    rc = ls.record(1, wprep_done); if (rc != FB200_OK) return rc;
  }
  // ops up to the last one of the side lane may have a GEMM of the other lane running beside them (cluster split-K hint)
  int last_side = -1;
  if (p.two_lanes) for (int i = 0; i < (int)p.ops.size(); ++i) if (p.ops[i].lane == 1) last_side = i;
  int op_index = -1;
  for (const Op& o : p.ops) {
    ++op_index;
    c.st = c.lane_st[o.lane];
    for (int b : {o.in0.buf, o.in1.buf, o.in2.buf}) { int rc = ls.wait_for(o.lane, b); if (rc != FB200_OK) return rc; }
    if (wprep_done && o.lane == 0 && o.kind == OP_LINEAR && o.engine == 1) { CUDA_OK(cudaStreamWaitEvent(ls.st[0], wprep_done, 0)); wprep_done = nullptr; }
    const int mst = c.mega ? 1 + std::max(c.vr(o.in0.buf), std::max(c.vr(o.in1.buf), c.vr(o.in2.buf))) : 0;   // stage of this op in the step kernel
    int mout = mst;                                                                                            // stage after which its output is complete
    switch (o.kind) {
      case OP_CAST: {
        const TRef src = c.value(o.in0);
        if (!c.gemm_only) pdl_launch(convert_kernel, c.dev.num_sms * 4, 256, 0, c.st, (const float*)src.p, src.ld, c.value(o.out), B, o.out.cols);
      } break;
      case OP_LINEAR: {
        if (o.engine == 1) {
          TcGemmArgs t{};
          t.kind = p.fmt == FMT_BF16 ? 0 : 1; t.a_mn = 0; t.b_mn = 0;
          t.A = tc_operand(c.value(o.in0), o.in0.cols, B);
          t.B = tc_operand(weight_operand(c, o), o.in0.cols, o.out.cols);
          t.M = B; t.N = o.out.cols; t.K = o.in0.cols;
          t.ep.C = c.value(o.out); t.ep.bias = c.param(o.b_slot, o.w_row0); t.ep.relu = o.relu; t.ep.mask_src.p = nullptr;
          t.ep.accumulate = 0; t.ep.atomic = 0; t.ep.colsum = nullptr; t.allow_split = 0; t.alone = op_index > last_side;
          int rc = launch_tc_gemm(t, c.dev.num_sms, c.st);
          if (rc != FB200_OK) return rc;
          break;
        }
        if (smalln_ok(o.out.cols, o.in0.cols) && c.value(o.out).fmt == FMT_F32 && !o.relu && o.in0.cols % 4 == 0 && c.value(o.in0).ld % 4 == 0) {
          // classifier head: warp-per-row kernel instead of a >90% padded GEMM tile
          SmallNArgs a{}; a.x = c.value(o.in0); a.y = c.value(o.out); a.W = c.param(o.w_slot, (int64_t)o.w_row0 * o.in0.cols);
          a.bias = c.param(o.b_slot, o.w_row0); a.B = B; a.K = o.in0.cols; a.C = o.out.cols;
          if (c.mega) {
            const int tiles = (B + ROW_WARPS - 1) / ROW_WARPS; bool chain;
            mout = c.row_stage({{c.vr(o.in0.buf), o.in0.buf, false}}, tiles, true, chain);
            c.mega->add_row(MR_SMALLN_FWD, mout, tiles, chain).u.smalln = a;
            c.row_emitted(mout, tiles, true, o.out.buf);
            break;
          }
          if (c.tail_fuse && (int)(&o - p.ops.data()) == c.tail_head_op) { c.tail.head_fwd = a; break; }      // runs inside the tail kernel
          CUDA_OK(launch_smalln_fwd(a, c.dev.num_sms, c.st));
          break;
        }
        GemmArgs g{};
        g.A = c.value(o.in0); g.B = make_ref((void*)c.param(o.w_slot, (int64_t)o.w_row0 * o.in0.cols), o.in0.cols, FMT_F32);
        g.C = c.value(o.out); g.M = B; g.N = o.out.cols; g.K = o.in0.cols; g.a_kc = 1; g.b_kc = 1;
        g.bias = c.param(o.b_slot, o.w_row0); g.relu = o.relu; g.mask_src.p = nullptr; g.accumulate = 0; g.split_k = 1; g.colsum_a = nullptr;
        if (c.mega) { mout = c.mega->add_gemm(g, mst); break; }
        c.enable_fixup(g, o.lane);
        CUDA_OK(launch_simt_gemm(g, c.dev.num_sms, c.st));
      } break;
      case OP_LNRD: {
        LnrdArgs a{}; a.x = c.value(o.in0); a.y = c.value(o.out); a.gamma = c.param(o.ln_w[0]); a.beta = c.param(o.ln_b[0]);
        a.stats = (float*)(c.ws + o.stats_off); a.drop = c.drop(o); a.B = B; a.N = o.out.cols;
        if (c.mega) {
          const int tiles = MegaBuilder::row_tiles(B, a.N); bool chain;
          mout = c.row_stage({{c.vr(o.in0.buf), o.in0.buf, false}}, tiles, a.N <= 512, chain);
          c.mega->add_row(MR_LNRD_FWD, mout, tiles, chain).u.lnrd = a;
          c.row_emitted(mout, tiles, a.N <= 512, o.out.buf);
          break;
        }
        if (c.tail_fuse && (int)(&o - p.ops.data()) == c.tail_ln_op) { c.tail.ln_fwd = a; c.tail_have_fwd = true; break; }   // runs inside the tail kernel
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(lnrd_fwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
      } break;
      case OP_GATE: {
        GateArgs a{}; a.x = c.value(o.in0); a.z = c.value(o.in1); a.y = c.value(o.out); a.B = B; a.N = o.out.cols;
        if (c.mega) { c.mega->add_row(MR_GATE_FWD, mst, MegaBuilder::row_tiles(B, a.N)).u.gate = a; break; }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(gate_fwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
      } break;
      case OP_GRB: {
        GrbArgs a{}; a.q = c.value(o.in0); a.a = c.value(o.in1); a.z = c.value(o.in2); a.y = c.value(o.out);
        a.gamma = c.param(o.ln_w[0]); a.beta = c.param(o.ln_b[0]); a.stats = (float*)(c.ws + o.stats_off);
        a.drop = c.drop(o); a.B = B; a.N = o.out.cols;
        if (c.mega) { c.mega->add_row(MR_GRB_FWD, mst, MegaBuilder::row_tiles(B, a.N)).u.grb = a; break; }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(grb_fwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
      } break;
      case OP_META: {
        MetaArgs a{}; a.v = c.value(o.in0); a.f = c.value(o.in1); a.g = c.value(o.in2); a.y = c.value(o.out);
        a.gamma_f = c.param(o.ln_w[0]); a.beta_f = c.param(o.ln_b[0]); a.gamma_g = c.param(o.ln_w[1]); a.beta_g = c.param(o.ln_b[1]);
        a.stats = (float*)(c.ws + o.stats_off); a.B = B; a.N = o.out.cols;
        if (c.mega) { c.mega->add_row(MR_META_FWD, mst, MegaBuilder::row_tiles(B, a.N)).u.meta = a; break; }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(meta_fwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
      } break;
      default: return FB200_EBADARG;
    }
    if (!c.dry_run) CUDA_OK(cudaGetLastError());
    if (c.mega) c.vready[o.out.buf] = std::max(c.vready[o.out.buf], mout);
    { int rc = ls.touched(o.lane, {o.out.buf}); if (rc != FB200_OK) return rc; }
  }
  c.st = main_st;
  return ls.join(c);
}

// ------------------------------------------------------------------------------- backward
static int run_backward(Ctx& c) {
  const Plan& p = c.p; const int B = p.d.B;
  if (c.mega) {                     // a stand-alone backward pass starts with every value and dlogits complete (stage 0)
    if (c.vready.empty()) c.vready.assign(p.acts.size(), 0);
    if (c.gready.empty()) c.gready.assign(p.acts.size(), 0);
  }
  if (!c.gemm_only && !c.dry_run) CUDA_OK(cudaMemsetAsync(c.grads, 0, (size_t)p.grad_elems * sizeof(float), c.st));
  std::vector<char> gwritten(p.acts.size(), 0);       // has the gradient buffer been written yet?
  std::vector<char> pwritten(NUM_SLOTS, 0);           // has this weight gradient been written yet?
  gwritten[p.logits.buf] = 1;
  // Views that alias one buffer (concat halves) are tracked per (buffer, col0) through a small map
  std::vector<std::pair<long long, char>> vw;
  auto vkey = [](const View& v) { return ((long long)v.buf << 32) | (unsigned)v.col0; };
  auto is_written = [&](const View& v) { if (gwritten[v.buf]) return true; for (auto& e : vw) if (e.first == vkey(v)) return e.second != 0; return false; };
  auto set_written = [&](const View& v) {
    if (v.col0 == 0 && v.cols == p.acts[v.buf].cols) { gwritten[v.buf] = 1; return; }
    for (auto& e : vw) if (e.first == vkey(v)) { e.second = 1; return; }
    vw.push_back({vkey(v), 1});
  };
  auto grad_wanted = [&](const View& v) {
    const int ext = p.acts[v.buf].ext;
    if (ext == 1) return c.d_img != nullptr;
    if (ext == 2) return c.d_txt != nullptr;
    return true;
  };

  struct MegaDw { GemmArgs g; int slot; int dy_ready; };
  std::vector<MegaDw> mega_dw;
  std::vector<ColsumSeg> colsums;
  std::vector<std::vector<TcGroupProblem>> dw_round;   // [k]: k-th application of a weight (k > 0 accumulates, in launch order)
  std::vector<int> tc_uses(NUM_SLOTS, 0);
  LaneSync ls; const cudaStream_t main_st = c.st;
  // event slots: one per activation buffer, then one per parameter slot (FFMA weight gradients are non-atomic
  // read-modify-writes of the parameter's gradient slice: a weight applied on both lanes must be ordered too)
  const int nbuf = (int)p.acts.size();
  { int rc = ls.begin(c, p.acts.size() + NUM_SLOTS); if (rc != FB200_OK) return rc; }      // after the zero fill of the gradient buffer
  int last_side_bwd = -1;            // (same hint as in the forward pass: the backward of these ops shares the chip with the side lane)
  if (p.two_lanes) for (int i = 0; i < (int)p.ops.size(); ++i) if (p.ops[i].lane == 1) last_side_bwd = i;
  for (int oi = (int)p.ops.size() - 1; oi >= 0; --oi) {
    const Op& o = p.ops[oi];
    // the gradient of a view that lives inside a fully written buffer counts as written
    if (o.kind == OP_CAST) continue;                  // format copy of an input: its gradient goes straight to the input (dx_view)
    if (!is_written(o.out)) return FB200_EBADARG;
    c.st = c.lane_st[o.lane];
    // gradient buffers this op reads (out) or writes / accumulates into (its inputs): order after the other lane's last touch
    const int wslot_ev = (o.kind == OP_LINEAR && o.engine == 0 && o.w_slot >= 0) ? nbuf + o.w_slot : -1;
    for (int b : {o.out.buf, o.in0.buf, o.in1.buf, o.in2.buf, o.dx_view.buf, wslot_ev}) { int rc = ls.wait_for(o.lane, b); if (rc != FB200_OK) return rc; }
    switch (o.kind) {
      case OP_CAST: break;
      case OP_LINEAR: {
        const int N = o.out.cols, K = o.in0.cols;
        if (o.engine == 1) {
          const int kind = p.fmt == FMT_BF16 ? 0 : 1;
          // dW[N,K] (+)= dY^T X : both operands MN-major views of the row-major activations.  Deferred: every
          // weight gradient of the step goes into one grouped launch after the dX chain (dY / X stay in the workspace).
          {
            TcGroupProblem gp{};
            gp.A = tc_operand(c.grad(o.out), N, B); gp.B = tc_operand(c.value(o.in0), K, B);
            gp.M = N; gp.N = K; gp.C = c.pgrad(o.w_slot, (int64_t)o.w_row0 * K); gp.ldc = K;
            const int use = tc_uses[o.w_slot]++;             // the same weight applied again: later rounds accumulate
            gp.accumulate = (use > 0 || pwritten[o.w_slot]) ? 1 : 0;
            if ((int)dw_round.size() <= use) dw_round.resize(use + 1);
            dw_round[use].push_back(gp);
            pwritten[o.w_slot] = 1;
          }
          int rc = FB200_OK;
          colsums.push_back(ColsumSeg{c.grad(o.out), N, c.pgrad(o.b_slot, o.w_row0)});   // db: batched after the loop
          if (grad_wanted(o.dx_view)) {
            // dX[B,K] (+)= dY W : W read MN-major from the same operand-format copy the forward used
            TcGemmArgs h{};
            h.kind = kind; h.a_mn = 0; h.b_mn = 1;
            h.A = tc_operand(c.grad(o.out), N, B); h.B = tc_operand(weight_operand(c, o), K, N);
            h.M = B; h.N = K; h.K = N;
            h.ep.C = c.grad(o.dx_view); h.ep.bias = nullptr; h.ep.relu = 0; h.ep.mask_src.p = nullptr;
            if (p.acts[o.in0.buf].relu_out) h.ep.mask_src = c.value(o.in0);
            h.ep.accumulate = is_written(o.dx_view); h.ep.atomic = 0; h.ep.colsum = nullptr; h.allow_split = 0; h.alone = oi > last_side_bwd;
            rc = launch_tc_gemm(h, c.dev.num_sms, c.st);
            if (rc != FB200_OK) return rc;
            set_written(o.dx_view);
          }
          break;
        }
        if (smalln_ok(N, K) && c.grad(o.out).fmt == FMT_F32 && !o.relu && c.value(o.in0).ld % 4 == 0) {
          SmallNArgs a{}; a.x = c.value(o.in0); a.dy = c.grad(o.out); a.W = c.param(o.w_slot, (int64_t)o.w_row0 * K);
          a.dW = c.pgrad(o.w_slot, (int64_t)o.w_row0 * K); a.db = c.pgrad(o.b_slot, o.w_row0); a.B = B; a.K = K; a.C = N;
          a.dx.p = nullptr; a.mask_src.p = nullptr;
          if (grad_wanted(o.dx_view)) {
            a.dx = c.grad(o.dx_view); a.dx_accumulate = is_written(o.dx_view);
            if (p.acts[o.in0.buf].relu_out) a.mask_src = c.value(o.in0);
          }
          if (c.tail_fuse && oi == c.tail_head_op) {                // ran inside the tail kernel
            c.tail.head_bwd = a; pwritten[o.w_slot] = 1;
            if (a.dx.p) set_written(o.dx_view);
            break;
          }
          if (c.mega) {
            const int tiles = (B + ROW_WARPS - 1) / ROW_WARPS; bool chain;
            const int st = c.row_stage({{c.gr(o.out.buf), o.out.buf, true}, {a.dx.p ? c.gr(o.dx_view.buf) : 0, a.dx.p ? o.dx_view.buf : -1, true}}, tiles, true, chain);
            c.mega->add_row(MR_SMALLN_BWD, st, tiles, chain).u.smalln = a;
            if (a.dx.p) { c.gready[o.dx_view.buf] = st; set_written(o.dx_view); }
            c.row_emitted(st, tiles, true, -1, a.dx.p ? o.dx_view.buf : -1);
            pwritten[o.w_slot] = 1;
            break;
          }
          // (r02d, measured and removed: running only dX here and leaving dW / db to a deferred FFMA GEMM beside the grouped weight
          // gradients made the step SLOWER - B = 256: 0.267 vs 0.254 ms, B = 1024: 0.403 vs 0.389, B = 4096: no change - the
          // 6-row TN GEMM is a worse reduction kernel than this one and lengthens the side lane)
          CUDA_OK(launch_smalln_bwd(a, c.dev.num_sms, c.st));
          pwritten[o.w_slot] = 1;
          if (a.dx.p) set_written(o.dx_view);
          break;
        }
        // dW[N,K] (+)= dY^T X ; db[N] += colsum(dY)
        GemmArgs g{};
        g.A = c.grad(o.out); g.a_kc = 0; g.B = c.value(o.in0); g.b_kc = 0;
        g.C = make_ref(c.pgrad(o.w_slot, (int64_t)o.w_row0 * K), K, FMT_F32);
        g.M = N; g.N = K; g.K = B; g.bias = nullptr; g.relu = 0; g.mask_src.p = nullptr;
        g.accumulate = pwritten[o.w_slot]; g.split_k = 0; g.colsum_a = c.pgrad(o.b_slot, o.w_row0);
        if (c.mega) {               // the weight gradient has no consumer: all of them run after the dX chain, as the last stage(s)
          mega_dw.push_back({g, o.w_slot, c.gr(o.out.buf)});
        } else if (p.use_tc && !c.gemm_only) {
          c.deferred_simt_dw.push_back(g);   // e.g. text_fc.0 (K = 85 is not TMA-legal): beside the grouped tcgen05 weight gradients, on the side lane
        } else CUDA_OK(launch_simt_gemm(g, c.dev.num_sms, c.st));
        pwritten[o.w_slot] = 1;
        // dX[B,K] (+)= dY W, masked by the producer's ReLU when the input came out of Linear+ReLU
        if (grad_wanted(o.dx_view)) {
          GemmArgs h{};
          h.A = c.grad(o.out); h.a_kc = 1; h.B = make_ref((void*)c.param(o.w_slot, (int64_t)o.w_row0 * K), K, FMT_F32); h.b_kc = 0;
          h.C = c.grad(o.dx_view); h.M = B; h.N = K; h.K = N; h.bias = nullptr; h.relu = 0;
          h.mask_src.p = nullptr;
          if (p.acts[o.in0.buf].relu_out) h.mask_src = c.value(o.in0);
          h.accumulate = is_written(o.dx_view); h.split_k = 1; h.colsum_a = nullptr;
          if (c.mega) {
            const int st = 1 + std::max(c.gr(o.out.buf), c.gr(o.dx_view.buf));
            c.mega->add_gemm(h, st); c.gready[o.dx_view.buf] = st;
          } else {
            c.enable_fixup(h, o.lane);
            CUDA_OK(launch_simt_gemm(h, c.dev.num_sms, c.st));
          }
          set_written(o.dx_view);
        }
      } break;
      case OP_LNRD: {
        LnrdArgs a{}; a.x = c.value(o.in0); a.y = c.value(o.out); a.dy = c.grad(o.out); a.dx = c.grad(o.in0);
        a.gamma = c.param(o.ln_w[0]); a.stats = (float*)(c.ws + o.stats_off);
        a.dgamma = c.pgrad(o.ln_w[0]); a.dbeta = c.pgrad(o.ln_b[0]); a.drop = c.drop(o); a.B = B; a.N = o.out.cols;
        if (is_written(o.in0)) return FB200_EUNSUPPORTED;     // LN input has exactly one consumer in every program
        if (c.tail_fuse && oi == c.tail_ln_op) {                // the tail kernel: launched HERE, at the position of its last member
          c.tail.ln_bwd = a;
          CUDA_OK(launch_tail_chain(c.tail, c.dev.num_sms, c.st));
          set_written(o.in0); break;
        }
        if (c.mega) {
          const int tiles = MegaBuilder::row_tiles(B, a.N); bool chain;
          const int st = c.row_stage({{c.gr(o.out.buf), o.out.buf, true}, {c.gr(o.in0.buf), o.in0.buf, true}}, tiles, a.N <= 512, chain);
          c.mega->add_row(MR_LNRD_BWD, st, tiles, chain).u.lnrd = a; c.gready[o.in0.buf] = st;
          c.row_emitted(st, tiles, a.N <= 512, -1, o.in0.buf);
          set_written(o.in0); break;
        }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(lnrd_bwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
        set_written(o.in0);
      } break;
      case OP_GATE: {
        GateArgs a{}; a.x = c.value(o.in0); a.z = c.value(o.in1); a.dy = c.grad(o.out); a.dz = c.grad(o.in1); a.dx = c.grad(o.in0);
        a.dx_accumulate = is_written(o.in0); a.B = B; a.N = o.out.cols;
        if (is_written(o.in1)) return FB200_EUNSUPPORTED;
        if (c.mega) {
          const int st = 1 + std::max(c.gr(o.out.buf), std::max(c.gr(o.in0.buf), c.gr(o.in1.buf)));
          c.mega->add_row(MR_GATE_BWD, st, MegaBuilder::row_tiles(B, a.N)).u.gate = a; c.gready[o.in0.buf] = c.gready[o.in1.buf] = st;
          set_written(o.in0); set_written(o.in1); break;
        }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(gate_bwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
        set_written(o.in0); set_written(o.in1);
      } break;
      case OP_GRB: {
        GrbArgs a{}; a.q = c.value(o.in0); a.a = c.value(o.in1); a.z = c.value(o.in2); a.dy = c.grad(o.out);
        a.da = c.grad(o.in1); a.dz = c.grad(o.in2); a.dq = c.grad(o.in0); a.dq_accumulate = is_written(o.in0);
        a.gamma = c.param(o.ln_w[0]); a.stats = (float*)(c.ws + o.stats_off);
        a.dgamma = c.pgrad(o.ln_w[0]); a.dbeta = c.pgrad(o.ln_b[0]); a.drop = c.drop(o); a.B = B; a.N = o.out.cols;
        if (is_written(o.in1) || is_written(o.in2)) return FB200_EUNSUPPORTED;
        if (c.mega) {
          const int st = 1 + std::max(std::max(c.gr(o.out.buf), c.gr(o.in0.buf)), std::max(c.gr(o.in1.buf), c.gr(o.in2.buf)));
          c.mega->add_row(MR_GRB_BWD, st, MegaBuilder::row_tiles(B, a.N)).u.grb = a;
          c.gready[o.in0.buf] = c.gready[o.in1.buf] = c.gready[o.in2.buf] = st;
          set_written(o.in0); set_written(o.in1); set_written(o.in2); break;
        }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(grb_bwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
        set_written(o.in0); set_written(o.in1); set_written(o.in2);
      } break;
      case OP_META: {
        MetaArgs a{}; a.v = c.value(o.in0); a.f = c.value(o.in1); a.g = c.value(o.in2); a.y = c.value(o.out); a.dy = c.grad(o.out);
        a.df = c.grad(o.in1); a.dg = c.grad(o.in2);
        a.dv.p = nullptr;
        if (grad_wanted(o.in0)) { a.dv = c.grad(o.in0); a.dv_accumulate = is_written(o.in0); }
        a.gamma_f = c.param(o.ln_w[0]); a.beta_f = c.param(o.ln_b[0]); a.gamma_g = c.param(o.ln_w[1]); a.beta_g = c.param(o.ln_b[1]);
        a.stats = (float*)(c.ws + o.stats_off);
        a.dgamma_f = c.pgrad(o.ln_w[0]); a.dbeta_f = c.pgrad(o.ln_b[0]); a.dgamma_g = c.pgrad(o.ln_w[1]); a.dbeta_g = c.pgrad(o.ln_b[1]);
        a.B = B; a.N = o.out.cols;
        if (is_written(o.in1) || is_written(o.in2)) return FB200_EUNSUPPORTED;
        if (c.mega) {
          const int st = 1 + std::max(std::max(c.gr(o.out.buf), a.dv.p ? c.gr(o.in0.buf) : 0), std::max(c.gr(o.in1.buf), c.gr(o.in2.buf)));
          c.mega->add_row(MR_META_BWD, st, MegaBuilder::row_tiles(B, a.N)).u.meta = a;
          if (a.dv.p) { c.gready[o.in0.buf] = st; set_written(o.in0); }
          c.gready[o.in1.buf] = c.gready[o.in2.buf] = st;
          set_written(o.in1); set_written(o.in2); break;
        }
        const int grid = row_grid_for(B, a.N, c.dev.num_sms);
#define CALL(NV, TPR) pdl_launch(meta_bwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, c.st, a)
        if (!c.gemm_only) FB200_ROW_DISPATCH(a.N, CALL);
#undef CALL
        if (a.dv.p) set_written(o.in0);
        set_written(o.in1); set_written(o.in2);
      } break;
      default: return FB200_EBADARG;
    }
    if (!c.dry_run) CUDA_OK(cudaGetLastError());
    { int rc = ls.touched(o.lane, {o.out.buf, o.in0.buf, o.in1.buf, o.in2.buf, o.dx_view.buf, wslot_ev}); if (rc != FB200_OK) return rc; }
  }
  c.st = main_st;
  { int rc = ls.join(c); if (rc != FB200_OK) return rc; }
  if (c.mega) {
    // Weight gradients interleaved with the dX chain doubled the tasks of every backward stage (two rounds per CTA, each stage
    // as long as its slowest CTA); as ONE final stage they are ~7 back-to-back tasks per CTA of the same hot loop.
    int last = 0;
    for (int b = 0; b < (int)c.gready.size(); ++b) last = std::max(last, c.gready[b]);
    for (auto& m : mega_dw) last = std::max(last, m.dy_ready);
    for (auto& m : mega_dw) {
      const int st = std::max(last + 1, c.pready[m.slot] + 1);       // a weight applied twice accumulates: its second gradient runs a stage later
      c.mega->add_gemm(m.g, st); c.pready[m.slot] = st;
    }
  }
  // Every dY is written now.  The bias gradients (HBM-bound column sums over all dY buffers) go to the side stream and run
  // next to the grouped weight-gradient launch (tensor-bound) instead of after it.
  { int rc = ls.begin(c, 0); if (rc != FB200_OK) return rc; }
  c.st = c.lane_st[1];
  for (auto& g : c.deferred_simt_dw) CUDA_OK(launch_simt_gemm(g, c.dev.num_sms, c.st));
  c.deferred_simt_dw.clear();
  for (size_t base = 0; base < colsums.size(); base += 24) {
    ColsumBatch cb{}; cb.B = B; cb.nseg = 0;
    for (size_t i = base; i < colsums.size() && cb.nseg < 24; ++i) cb.seg[cb.nseg++] = colsums[i];
    int gx = (B + 63) / 64; if (gx > 8 * c.dev.num_sms / cb.nseg + 1) gx = 8 * c.dev.num_sms / cb.nseg + 1; if (gx < 1) gx = 1;   // ~64 rows per CTA
    if (!c.gemm_only) pdl_launch(colsum_batch_kernel, dim3(gx, cb.nseg), 256, 0, c.st, cb);
    CUDA_OK(cudaGetLastError());
  }
  c.st = main_st;
  // weight gradients of all tcgen05 Linears, grouped (round 1 only exists for weights applied twice).  With a data-parallel
  // caller's event: first the problems below Plan::dp_split, then the event (every gradient below the split is final: the
  // row kernels and the column sums are done), then the rest - the caller all-reduces the first bucket under the second half.
  auto launch_dw = [&](const std::vector<TcGroupProblem>& v) -> int {
    for (size_t base = 0; base < v.size(); base += TC_MAX_GROUP) {
      const int n = (int)std::min<size_t>(TC_MAX_GROUP, v.size() - base);
      int rc = launch_tc_grouped_tn(p.fmt == FMT_BF16 ? 0 : 1, v.data() + base, n, B, c.st);
      if (rc != FB200_OK) return rc;
    }
    return FB200_OK;
  };
  auto record_mid = [&]() -> int {
    if (!c.mid_event) return FB200_OK;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    CUDA_OK(cudaStreamIsCapturing(c.st, &cap));
    CUDA_OK(cudaEventRecordWithFlags((cudaEvent_t)c.mid_event, c.st, cap == cudaStreamCaptureStatusActive ? cudaEventRecordExternal : cudaEventRecordDefault));
    return FB200_OK;
  };
  const bool split = c.mid_event && p.dp_split > 0 && dw_round.size() == 1;
  if (split) {
    std::vector<TcGroupProblem> first, second;
    for (auto& gp : dw_round[0]) ((gp.C - c.grads) < p.dp_split ? first : second).push_back(gp);
    { int rc = launch_dw(first); if (rc != FB200_OK) return rc; }
    { int rc = ls.join(c); if (rc != FB200_OK) return rc; }
    { int rc = record_mid(); if (rc != FB200_OK) return rc; }
    { int rc = launch_dw(second); if (rc != FB200_OK) return rc; }
  } else {
    for (size_t round = 0; round < dw_round.size(); ++round) { int rc = launch_dw(dw_round[round]); if (rc != FB200_OK) return rc; }
    { int rc = ls.join(c); if (rc != FB200_OK) return rc; }
    { int rc = record_mid(); if (rc != FB200_OK) return rc; }
  }
  if (c.dry_run) return FB200_OK;
  // inputs nobody differentiated through still owe the caller a defined gradient
  if (c.d_img && !gwritten[0]) CUDA_OK(cudaMemsetAsync(c.d_img, 0, (size_t)B * p.d.F * sizeof(float), c.st));
  if (c.d_txt && !gwritten[1]) CUDA_OK(cudaMemsetAsync(c.d_txt, 0, (size_t)B * p.acts[1].cols * sizeof(float), c.st));
  return FB200_OK;
}

static int check_common(const fb200_desc* d, const void* const* params, const void* img, const void* txt, const void* ws, Plan& plan, DeviceInfo& dev) {
  if (!d || !params || !img || !txt || !ws) return FB200_EBADARG;
  int rc = build_plan(*d, plan);
  if (rc != FB200_OK) return rc;
  rc = get_device_info(dev);
  if (rc != FB200_OK) return rc;
  if (dev.cc_major < 10) return FB200_EUNSUPPORTED;          // sm_100a only: no other architecture is built
  if (!is_device_ptr(img) || !is_device_ptr(txt) || !is_device_ptr(ws)) return FB200_EUNSUPPORTED;   // no CPU path
  for (int s = 0; s < NUM_SLOTS; ++s) {
    if (!plan.live[s]) continue;
    if (!params[s]) return FB200_EBADARG;
    if (((uintptr_t)params[s]) & 15) return FB200_EALIGN;
    if (!is_device_ptr_cached(params[s])) return FB200_EUNSUPPORTED;
  }
  if ((((uintptr_t)img) & 15) || (((uintptr_t)ws) & 255)) return FB200_EALIGN;
  // text_in rows are read with 128-bit loads whenever its width allows it: the base must then be 16-byte aligned
  // (widths 85 / 13 / 11 take the scalar loaders of the FFMA kernel, any 4-byte alignment)
  if (((uintptr_t)txt) & 3) return FB200_EALIGN;
  if (plan.acts[1].cols % 4 == 0 && (((uintptr_t)txt) & 15)) return FB200_EALIGN;
  return FB200_OK;
}

// optional device pointers of a call (labels, class weights, denominator, dropout masks, rng state): NULL is fine, a host
// pointer is refused with a status instead of an illegal-address fault that would kill the context
static int check_optional_device(std::initializer_list<const void*> ptrs, const uint8_t* const* masks) {
  for (const void* q : ptrs) if (q && !is_device_ptr_cached(q)) return FB200_EUNSUPPORTED;
  if (masks) for (int i = 0; i < FB200_NUM_DROPOUT_SITES; ++i) if (masks[i] && !is_device_ptr_cached(masks[i])) return FB200_EUNSUPPORTED;
  return FB200_OK;
}

// Persistent step kernel: emit the weighted cross entropy as a row task (train step) and launch the finished program.
static void mega_emit_ce(Ctx& c, const void* logits, const int64_t* labels, const float* class_w, const float* denom, float* loss_out, void* dlogits) {
  const Plan& p = c.p;
  const int tiles = (p.d.B + ROW_WARPS - 1) / ROW_WARPS; bool chain;
  const int st = c.row_stage({{c.vr(p.logits.buf), p.logits.buf, false}}, tiles, true, chain);
  c.mega->add_row(MR_CE, st, tiles, chain).u.ce = CeArgs{(const float*)logits, labels, class_w, denom, loss_out, (float*)dlogits, p.d.B, p.d.C, nullptr};
  c.gready.assign(p.acts.size(), 0);
  c.gready[p.logits.buf] = st;
  c.row_emitted(st, tiles, true, -1, p.logits.buf);
}
static int mega_finish(Ctx& c) {
  static thread_local MegaProg prog;             // 25 KB: not on the stack of autograd's worker thread
  int rc = c.mega->finalize(prog, (unsigned*)(c.ws + c.p.mega_bar_off));
  if (rc != FB200_OK) return rc;
  return mega_launch(prog, c.dev.num_sms, c.st);
}
static void mega_begin(Ctx& c, MegaBuilder& mb) {
  if (!c.p.use_mega || c.gemm_only) return;
  mb.scratch = c.ws + c.p.splitk_off; mb.scratch_bytes = c.p.splitk_bytes;
  c.mega = &mb;
}

static int ce_launch(const void* logits, const int64_t* labels, const float* class_w, const float* denom, int B, int C,
                     float* loss_out, void* dlogits, cudaStream_t st) {
  if (!logits || !labels || !loss_out || B < 1 || C < 1) return FB200_EBADARG;
  CUDA_OK(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), st));
  int rows_per_cta = 256 / 8;
  int grid = (B + rows_per_cta - 1) / rows_per_cta; if (grid > 1184) grid = 1184;
  pdl_launch(ce_pass1_kernel, grid, 256, 0, st, (const float*)logits, labels, class_w, B, C, loss_out, (float*)dlogits);
  CUDA_OK(cudaGetLastError());
  int n = B * C; int g2 = (n + 255) / 256; if (g2 > 1184) g2 = 1184;
  pdl_launch(ce_pass2_kernel, g2, 256, 0, st, denom, n, loss_out, (float*)dlogits);
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

}  // namespace fb200

using namespace fb200;

// =================================================================================== C ABI
extern "C" {

int fb200_version(void) { return FB200_VERSION; }

const char* fb200_strerror(int s) {
  switch (s) {
    case FB200_OK: return "ok";
    case FB200_EBADARG: return "fb200: bad argument (null pointer or inconsistent descriptor)";
    case FB200_EUNSUPPORTED: return "fb200: unsupported shape, dtype or device (this library is CUDA sm_100a only; there is no CPU path)";
    case FB200_EALIGN: return "fb200: pointer or leading dimension violates the 16-byte alignment contract";
    case FB200_ECUDA: return "fb200: CUDA runtime error";
    case FB200_EABSENT: return "fb200: parameter slot absent in this configuration";
    default: return "fb200: unknown status";
  }
}

static const char* const kMechNames[FB200_NUM_MECHANISMS] = {
  "no-metadata", "no-metadata-without-mlp", "concatenation", "crossattention", "weighted", "gfcam",
  "cross-weights-after-crossattention", "metablock", "rg-att2fusefeatures", "rg-att", "att-intramodal",
  "att-intramodal+residual", "cross-attention-only", "residual+cross-attention-metadados",
  "att-intramodal+residual+cross-attention-metadados",
  "att-intramodal+residual+cross-attention-metadados+rg-att2fusefeatures",
  "att-intramodal+residual+cross-attention-metadados+metablock",
  "att-intramodal+residual+cross-attention-metadados+att-intramodal+residual",
};
int fb200_mechanism_from_string(const char* s) {
  if (!s) return -1;
  for (int i = 0; i < FB200_NUM_MECHANISMS; ++i) if (std::strcmp(s, kMechNames[i]) == 0) return i;
  return -1;
}
const char* fb200_mechanism_string(int m) { return (m >= 0 && m < FB200_NUM_MECHANISMS) ? kMechNames[m] : nullptr; }
int fb200_num_params(void) { return NUM_SLOTS; }
const char* fb200_param_name(int slot) { return (slot >= 0 && slot < NUM_SLOTS) ? kSlotNames[slot] : nullptr; }

int fb200_param_shape(const fb200_desc* d, int slot, int64_t* rows, int64_t* cols) {
  if (!d || !rows || !cols || slot < 0 || slot >= NUM_SLOTS) return FB200_EBADARG;
  Shape s = slot_shape(*d, slot);
  if (!s.present) return FB200_EABSENT;
  *rows = s.rows; *cols = s.cols;
  return FB200_OK;
}
int64_t fb200_grad_offset(const fb200_desc* d, int slot) {
  if (!d || slot < 0 || slot >= NUM_SLOTS) return -1;
  Plan p; if (build_plan(*d, p) != FB200_OK) return -1;
  return p.goff[slot];
}
int64_t fb200_grad_elems(const fb200_desc* d) {
  if (!d) return -1;
  Plan p; if (build_plan(*d, p) != FB200_OK) return -1;
  return p.grad_elems;
}
int fb200_workspace_bytes(const fb200_desc* d, size_t* bytes) {
  if (!d || !bytes) return FB200_EBADARG;
  Plan p; int rc = build_plan(*d, p); if (rc != FB200_OK) return rc;
  *bytes = p.ws_bytes;
  return FB200_OK;
}
float fb200_dropout_p(const fb200_desc* d, int site) {
  if (!d || site < 0 || site >= FB200_NUM_DROPOUT_SITES) return 0.f;
  Plan p; if (build_plan(*d, p) != FB200_OK) return 0.f;
  return p.drop_p[site];
}
int fb200_dropout_shape(const fb200_desc* d, int site, int64_t* rows, int64_t* cols) {
  if (!d || !rows || !cols || site < 0 || site >= FB200_NUM_DROPOUT_SITES) return FB200_EBADARG;
  Plan p; int rc = build_plan(*d, p); if (rc != FB200_OK) return rc;
  if (p.drop_p[site] == 0.f) return FB200_EABSENT;
  *rows = d->B; *cols = p.drop_cols[site];
  return FB200_OK;
}
int fb200_algorithmic_work(const fb200_desc* d, double* flops, double* bytes, int64_t* live_params) {
  if (!d) return FB200_EBADARG;
  Plan p; int rc = build_plan(*d, p); if (rc != FB200_OK) return rc;
  if (flops) *flops = p.flops;
  if (bytes) *bytes = p.bytes;
  if (live_params) *live_params = p.live_params;
  return FB200_OK;
}
int fb200_launch_count(const fb200_desc* d, int* forward, int* backward) {
  if (!d) return FB200_EBADARG;
  Plan p; int rc = build_plan(*d, p); if (rc != FB200_OK) return rc;
  int f = 0, b = 0;
  if (p.use_mega) {                  // one persistent kernel per pass (a fused train step is ONE launch: the caller counts forward only)
    if (forward) *forward = 1;
    if (backward) *backward = 1;
    return FB200_OK;
  }
  const bool need_dimg = d->flags & FB200_FLAG_NEED_DIMG, need_dtxt = d->flags & FB200_FLAG_NEED_DTEXT;
  f += (int)((p.wprep.size() + 31) / 32);
  int ntc = 0, rounds = 0; int uses[NUM_SLOTS] = {};
  for (auto& o : p.ops) {
    f += 1;
    if (o.kind == OP_LINEAR) {
      const int ext = p.acts[o.dx_view.buf].ext;
      if (o.engine == 1) { ++ntc; int u = ++uses[o.w_slot]; if (u > rounds) rounds = u; }
      if (o.engine == 0 && smalln_ok(o.out.cols, o.in0.cols) && !o.relu && p.acts[o.out.buf].ext == 3) { b += 1; continue; }
      b += (o.engine == 1 ? 0 : 1) + ((ext == 0 || (ext == 1 && need_dimg) || (ext == 2 && need_dtxt)) ? 1 : 0);
    } else if (o.kind != OP_CAST) b += 1;
  }
  b += (ntc + 23) / 24 + rounds;      // batched bias-gradient kernel(s) + grouped weight-gradient launch(es)
  if (forward) *forward = f;
  if (backward) *backward = b;
  return FB200_OK;
}

int fb200_list_gemms(const fb200_desc* d, int32_t* out, int cap) {
  if (!d || !out || cap < 1) return FB200_EBADARG;
  Plan p; int rc = build_plan(*d, p); if (rc != FB200_OK) return rc;
  const bool need_dimg = d->flags & FB200_FLAG_NEED_DIMG, need_dtxt = d->flags & FB200_FLAG_NEED_DTEXT;
  int n = 0;
  auto put = [&](int layout, int engine, int M, int N, int K) {
    if (n < cap) { int32_t* e = out + 5 * n; e[0] = layout; e[1] = engine; e[2] = M; e[3] = N; e[4] = K; }
    ++n;
  };
  for (auto& o : p.ops) {
    if (o.kind != OP_LINEAR) continue;
    if (o.engine == 0 && smalln_ok(o.out.cols, o.in0.cols) && !o.relu && p.acts[o.out.buf].ext == 3) continue;   // classifier head: warp-per-row kernels, not a GEMM
    const int eng = o.engine == 0 ? 0 : (d->dtype == FB200_BF16 ? 2 : 1);
    put(0, eng, d->B, o.out.cols, o.in0.cols);
    put(2, eng, o.out.cols, o.in0.cols, d->B);
    const int ext = p.acts[o.dx_view.buf].ext;
    if (ext == 0 || (ext == 1 && need_dimg) || (ext == 2 && need_dtxt)) put(1, eng, d->B, o.in0.cols, o.out.cols);
  }
  return n < cap ? n : cap;
}

__global__ void rng_advance_kernel(uint64_t* state, uint64_t inc) { pdl_sync(); state[1] += inc; }

/* debug: device buffer of >= 8*k_blocks int64 receiving pipeline time stamps of CTA (0,0,0) of every tcgen05 GEMM launched afterwards; NULL disables */
int fb200_debug_set_pdl(int on) { const int prev = pdl_enabled() ? 1 : 0; pdl_flag() = on ? 1 : 0; return prev; }
int fb200_debug_tc_trace(void* device_buf) { tc_trace_buffer() = (long long*)device_buf; return FB200_OK; }
int fb200_debug_tc_timeline(void* device_buf, int max_launches) {
  TcTimeline& t = tc_timeline();
  const int recorded = t.next;
  t.buf = (long long*)device_buf; t.max_launches = device_buf ? max_launches : 0; t.next = 0;
  return recorded;
}
/* debug: device buffer of >= 2 + 2 * stages int64 receiving clock64 stamps of CTA 0 of the persistent step kernel
 * ([0] entry, [1 + 2s] own tasks of stage s done, [2 + 2s] barrier after stage s passed); NULL disables */
/* host only (no CUDA call): the program the persistent step kernel would run for `d` - pass 0 forward, 1 backward, 2 fused train
 * step.  out[0] stages, out[1] GEMM ops, out[2] row ops, out[3] tile tasks; FB200_EUNSUPPORTED when `d` does not take that path. */
int fb200_mega_program_info(const fb200_desc* d, int pass, int* out) {
  if (!d || !out || pass < 0 || pass > 2) return FB200_EBADARG;
  Plan plan; int rc = build_plan(*d, plan); if (rc != FB200_OK) return rc;
  if (!plan.use_mega) return FB200_EUNSUPPORTED;
  // fake, never dereferenced on the host: distinct 256-byte aligned addresses
  std::vector<const void*> params(NUM_SLOTS, nullptr);
  for (int s = 0; s < NUM_SLOTS; ++s) if (plan.live[s]) params[s] = (const void*)(uintptr_t)(0x100000000ull + (uint64_t)s * 0x4000000ull);
  char* base = (char*)(uintptr_t)0x4000000000ull;
  DeviceInfo dev; dev.num_sms = 148; dev.cc_major = 10;
  const bool nd = d->flags & FB200_FLAG_NEED_DIMG, nt = d->flags & FB200_FLAG_NEED_DTEXT;
  Ctx c{plan, params.data(), base + 0x1000000000ull, base + 0x1100000000ull, base + 0x1200000000ull, base + 0x1300000000ull,
        nd ? base + 0x1400000000ull : nullptr, nt ? base + 0x1500000000ull : nullptr, (float*)(base + 0x1600000000ull), nullptr, 0, 0, nullptr,
        base, nullptr, dev};
  c.dry_run = true;
  MegaBuilder mb; mb.scratch = c.ws + plan.splitk_off; mb.scratch_bytes = plan.splitk_bytes; c.mega = &mb;
  if (pass != 1) { rc = run_forward(c); if (rc != FB200_OK) return rc; }
  if (pass == 2) mega_emit_ce(c, c.logits, (const int64_t*)(base + 0x1700000000ull), nullptr, nullptr, (float*)(base + 0x1800000000ull), (void*)c.dlogits);
  if (pass != 0) { rc = run_backward(c); if (rc != FB200_OK) return rc; }
  static thread_local MegaProg prog;
  rc = mb.finalize(prog, (unsigned*)(c.ws + plan.mega_bar_off)); if (rc != FB200_OK) return rc;
  int tasks = 0;
  for (int i = 0; i < prog.ngemm; ++i) tasks += prog.g[i].tiles;
  for (int i = 0; i < prog.nrow; ++i) tasks += prog.r[i].chain ? 0 : prog.r[i].tiles;
  out[0] = prog.nstages; out[1] = prog.ngemm; out[2] = prog.nrow; out[3] = tasks;
  return FB200_OK;
}
int fb200_debug_mega_trace(void* device_buf) { mega_trace_buffer() = (long long*)device_buf; return FB200_OK; }
/* debug: the step kernel with `nstages` EMPTY stages - the cost of the launch and of the grid barriers alone */
int fb200_debug_mega_barriers(int nstages, void* ws256, void* stream) {
  if (nstages < 1 || nstages > MEGA_MAX_STAGES || !ws256) return FB200_EBADARG;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  static thread_local MegaProg prog;
  for (int s = 0; s < nstages; ++s) prog.st[s] = MStage{0, 0, 0, 0};
  prog.nstages = nstages; prog.ngemm = 0; prog.nrow = 0; prog.barrier = (unsigned*)ws256; prog.trace = mega_trace_buffer(); prog.pad_ = 0;
  return mega_launch(prog, dev.num_sms, (cudaStream_t)stream);
}



int fb200_aux_loss(int kind, const void* logits, const void* targets, const float* weight, float gamma, int B, int C,
                   float* loss_out, void* dlogits, void* stream) {
  if (!logits || !targets || !loss_out || B < 1 || C < 1 || (kind != 1 && kind != 2)) return FB200_EBADARG;
  if (!is_device_ptr(logits)) return FB200_EUNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
  int grid = (B + 31) / 32; if (grid > 1184) grid = 1184;
  pdl_launch(aux_loss_kernel, grid, 256, 0, st, kind, (const float*)logits, kind == 1 ? (const int64_t*)targets : nullptr,
                                        kind == 2 ? (const float*)targets : nullptr, weight, gamma, B, C, loss_out, (float*)dlogits);
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

int fb200_metadata_encode(const int32_t* codes, const int32_t* col_of, const int32_t* col_base, const double* numeric,
                          const double* mean, const double* scale, int B, int n_cat, int cat_total, int n_num, float* out, void* stream) {
  if (!out || B < 1 || n_cat < 0 || cat_total < 0 || n_num < 0 || cat_total + n_num < 1) return FB200_EBADARG;
  if ((cat_total > 0 && (!codes || !col_of || !col_base || n_cat < 1)) || (n_num > 0 && (!numeric || !mean || !scale))) return FB200_EBADARG;
  if (!is_device_ptr(out)) return FB200_EUNSUPPORTED;
  { int rc = check_optional_device({codes, col_of, col_base, numeric, mean, scale}, nullptr); if (rc != FB200_OK) return rc; }
  const int64_t total = (int64_t)B * (cat_total + n_num);
  int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
  pdl_launch(metadata_encode_kernel, grid, 256, 0, (cudaStream_t)stream, codes, col_of, col_base, numeric, mean, scale, B, n_cat, cat_total, n_num, out);
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

int fb200_softmax_argmax(const void* logits, int B, int C, void* probs, int64_t* pred, void* stream) {
  if (!logits || B < 1 || C < 1 || (!probs && !pred)) return FB200_EBADARG;
  if (!is_device_ptr(logits)) return FB200_EUNSUPPORTED;
  int grid = (B + 31) / 32; if (grid > 1184) grid = 1184;
  pdl_launch(softmax_argmax_kernel, grid, 256, 0, (cudaStream_t)stream, (const float*)logits, B, C, (float*)probs, pred);
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

int fb200_adam_step(int ntensors, void* const* params, const void* const* grads, void* const* exp_avg, void* const* exp_avg_sq,
                    const int64_t* numel, float lr, float beta1, float beta2, float eps, float weight_decay, int64_t step,
                    float grad_scale, void* stream) {
  if (ntensors < 0 || (ntensors && (!params || !grads || !exp_avg || !exp_avg_sq || !numel)) || step < 1) return FB200_EBADARG;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  const float bc1 = 1.f - (float)std::pow((double)beta1, (double)step), bc2 = 1.f - (float)std::pow((double)beta2, (double)step);
  for (int base = 0; base < ntensors; base += 48) {
    AdamBatch a{}; a.nseg = 0; a.lr = lr; a.b1 = beta1; a.b2 = beta2; a.eps = eps; a.wd = weight_decay; a.bc1 = bc1; a.bc2 = bc2; a.grad_scale = grad_scale;
    int64_t maxn = 0;
    for (int i = base; i < ntensors && a.nseg < 48; ++i) {
      if (!params[i] || !grads[i] || !exp_avg[i] || !exp_avg_sq[i] || numel[i] < 0) return FB200_EBADARG;
      if (numel[i] == 0) continue;
      if (a.nseg == 0 && !is_device_ptr(params[i])) return FB200_EUNSUPPORTED;
      a.seg[a.nseg++] = AdamSeg{(float*)params[i], (const float*)grads[i], (float*)exp_avg[i], (float*)exp_avg_sq[i], numel[i]};
      if (numel[i] > maxn) maxn = numel[i];
    }
    if (a.nseg == 0) continue;
    int gx = (int)((maxn / 4 + 255) / 256); if (gx > dev.num_sms * 2) gx = dev.num_sms * 2; if (gx < 1) gx = 1;
    pdl_launch(adam_kernel, dim3(gx, a.nseg), 256, 0, (cudaStream_t)stream, a);
    CUDA_OK(cudaGetLastError());
  }
  return FB200_OK;
}

int fb200_rng_advance(void* rng_state, uint64_t increment, void* stream) {
  if (!rng_state) return FB200_EBADARG;
  if (!is_device_ptr(rng_state)) return FB200_EUNSUPPORTED;
  pdl_launch(rng_advance_kernel, 1, 1, 0, (cudaStream_t)stream, (uint64_t*)rng_state, increment);
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}


/* debug / measurement aid: launch ONLY the GEMM kernels of one train step under desc (forward, dX chain, grouped
 * weight gradients) on the given buffers - same kernels, grids and operands as fb200_head_train_step issues. */
int fb200_debug_gemm_replay(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                            void* logits, void* grads, void* ws, void* stream) {
  Plan plan; DeviceInfo dev;
  if (!d) return FB200_EBADARG;
  fb200_desc dd = *d; dd.flags |= FB200_FLAG_NO_MEGA;          // the replay times the per-op GEMM kernels
  int rc = check_common(&dd, params, img_feat, text_in, ws, plan, dev);
  if (rc != FB200_OK) return rc;
  if (!logits || !grads) return FB200_EBADARG;
  char* w = (char*)ws;
  float* dlog = (float*)(w + plan.ws_bytes - (((size_t)d->B * d->C * sizeof(float) + 255) & ~size_t(255)) - 256);
  Ctx c{plan, params, img_feat, text_in, logits, dlog, nullptr, nullptr, (float*)grads, nullptr, 0, 0, nullptr, w, (cudaStream_t)stream, dev};
  c.gemm_only = true;
  rc = run_forward(c);
  if (rc != FB200_OK) return rc;
  return run_backward(c);
}


int fb200_grad_live_ranges(const fb200_desc* d, int64_t* out, int cap) {
  if (!d || !out || cap < 1) return FB200_EBADARG;
  Plan p; int rc = build_plan(*d, p); if (rc != FB200_OK) return rc;
  // per slot: rows that can be non-zero.  in_proj_weight / in_proj_bias of an S=1 attention only ever get their V third.
  std::vector<std::pair<int64_t, int64_t>> r;
  bool is_inproj[NUM_SLOTS] = {};
  for (int base : {(int)S_ISA, (int)S_TSA, (int)S_ICA, (int)S_TCA, (int)S_IRES + 2, (int)S_TRES + 2}) { is_inproj[base] = true; is_inproj[base + 1] = true; }
  for (int s = 0; s < NUM_SLOTS; ++s) {
    if (p.goff[s] < 0) continue;
    Shape sh = slot_shape(*d, s);
    const int64_t n = sh.rows * (sh.cols ? sh.cols : 1);
    int64_t b = p.goff[s], e = p.goff[s] + n;
    if (is_inproj[s]) b += 2 * (n / 3);
    if (!r.empty() && r.back().second >= b - 4) r.back().second = e;      // merge neighbours (alignment padding included)
    else r.push_back({b, e});
  }
  int n = 0;
  for (auto& x : r) { if (n < cap) { out[2 * n] = x.first; out[2 * n + 1] = x.second; } ++n; }
  return n;
}

int fb200_head_forward(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                       const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state, void* logits, void* ws, void* stream) {
  Plan plan; DeviceInfo dev;
  int rc = check_common(d, params, img_feat, text_in, ws, plan, dev);
  if (rc != FB200_OK) return rc;
  if (!logits) return FB200_EBADARG;
  rc = check_optional_device({rng_state, logits}, masks); if (rc != FB200_OK) return rc;
  Ctx c{plan, params, img_feat, text_in, logits, nullptr, nullptr, nullptr, nullptr, masks, seed, offset, (const uint64_t*)rng_state, (char*)ws, (cudaStream_t)stream, dev};
  MegaBuilder mb; mega_begin(c, mb);
  rc = run_forward(c);
  if (rc != FB200_OK || !c.mega) return rc;
  return mega_finish(c);
}

int fb200_head_backward(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                        const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state, const void* dlogits, void* grads,
                        void* d_img_feat, void* d_text_in, void* ws, void* stream) {
  Plan plan; DeviceInfo dev;
  int rc = check_common(d, params, img_feat, text_in, ws, plan, dev);
  if (rc != FB200_OK) return rc;
  if (!dlogits || !grads) return FB200_EBADARG;
  if ((d->flags & FB200_FLAG_NEED_DIMG) && !d_img_feat) return FB200_EBADARG;
  if ((d->flags & FB200_FLAG_NEED_DTEXT) && !d_text_in) return FB200_EBADARG;
  rc = check_optional_device({rng_state, dlogits, grads, d_img_feat, d_text_in}, masks); if (rc != FB200_OK) return rc;
  Ctx c{plan, params, img_feat, text_in, nullptr, dlogits,
        (d->flags & FB200_FLAG_NEED_DIMG) ? d_img_feat : nullptr, (d->flags & FB200_FLAG_NEED_DTEXT) ? d_text_in : nullptr,
        (float*)grads, masks, seed, offset, (const uint64_t*)rng_state, (char*)ws, (cudaStream_t)stream, dev};
  MegaBuilder mb; mega_begin(c, mb);
  rc = run_backward(c);
  if (rc != FB200_OK || !c.mega) return rc;
  return mega_finish(c);
}

int fb200_cross_entropy(const void* logits, const int64_t* labels, const float* class_w, const float* denom, int B, int C,
                        float* loss_out, void* dlogits, void* stream) {
  if (!logits || !is_device_ptr(logits)) return logits ? FB200_EUNSUPPORTED : FB200_EBADARG;
  { int rc = check_optional_device({labels, class_w, denom, loss_out, dlogits}, nullptr); if (rc != FB200_OK) return rc; }
  return ce_launch(logits, labels, class_w, denom, B, C, loss_out, dlogits, (cudaStream_t)stream);
}

int fb200_head_train_step(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                          const int64_t* labels, const float* class_w, const float* denom,
                          const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                          void* logits, float* loss_out, void* grads, void* d_img_feat, void* d_text_in, void* ws, void* stream) {
  return fb200_head_train_step_dp(d, params, img_feat, text_in, labels, class_w, denom, masks, seed, offset, rng_state,
                                  logits, loss_out, grads, d_img_feat, d_text_in, ws, stream, nullptr);
}

int fb200_dp_allreduce(void* multicast_ptr, void* const* peer_ptrs, const int64_t* ranges, int nranges, int rank, int world, int max_ctas,
                       void* const* signal_pads, void* state, int slot, void* stream) {
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  return dp_allreduce_launch((float*)multicast_ptr, (float* const*)peer_ptrs, ranges, nranges, rank, world, dev.num_sms, max_ctas,
                             (uint32_t* const*)signal_pads, (uint32_t*)state, slot, (cudaStream_t)stream);
}

int64_t fb200_dp_bucket_split(const fb200_desc* d) {
  if (!d) return -1;
  Plan p; if (build_plan(*d, p) != FB200_OK) return -1;
  return p.dp_split;
}

int fb200_head_train_step_dp(const fb200_desc* d, const void* const* params, const void* img_feat, const void* text_in,
                             const int64_t* labels, const float* class_w, const float* denom,
                             const uint8_t* const* masks, uint64_t seed, uint64_t offset, const void* rng_state,
                             void* logits, float* loss_out, void* grads, void* d_img_feat, void* d_text_in, void* ws, void* stream,
                             void* mid_event) {
  Plan plan; DeviceInfo dev;
  int rc = check_common(d, params, img_feat, text_in, ws, plan, dev);
  if (rc != FB200_OK) return rc;
  if (!logits || !labels || !loss_out || !grads) return FB200_EBADARG;
  if ((d->flags & FB200_FLAG_NEED_DIMG) && !d_img_feat) return FB200_EBADARG;
  if ((d->flags & FB200_FLAG_NEED_DTEXT) && !d_text_in) return FB200_EBADARG;
  rc = check_optional_device({labels, class_w, denom, rng_state, logits, loss_out, grads, d_img_feat, d_text_in}, masks); if (rc != FB200_OK) return rc;
  // dlogits live at the tail of the workspace (fb200_workspace_bytes reserves them)
  char* w = (char*)ws;
  float* dlog = (float*)(w + plan.ws_bytes - (((size_t)d->B * d->C * sizeof(float) + 255) & ~size_t(255)) - 256);
  Ctx c{plan, params, img_feat, text_in, logits, dlog,
        (d->flags & FB200_FLAG_NEED_DIMG) ? d_img_feat : nullptr, (d->flags & FB200_FLAG_NEED_DTEXT) ? d_text_in : nullptr,
        (float*)grads, masks, seed, offset, (const uint64_t*)rng_state, w, (cudaStream_t)stream, dev};
  c.mid_event = mid_event;
  MegaBuilder mb; mega_begin(c, mb);
  if (!c.mega && !c.gemm_only) {
    // the program ends with LayerNorm+ReLU+dropout (<= 512 wide) -> C-wide classifier head into the logits: fuse the tail
    // Measured on B200 (cfg2, B = 4096): 0.758 ms per step with the fused tail against 0.737 ms with its six kernels - one
    // wave of 512 CTAs walking five cold bodies in turn loses to six fully parallel launches overlapped by PDL.  Kept behind
    // FB200_TAIL_FUSE=1 (the persistent step kernel chains the same bodies for small batches, where launches dominate).
    static const bool on = [] { const char* e = getenv("FB200_TAIL_FUSE"); return e && e[0] == '1'; }();
    const int n = (int)plan.ops.size();
    if (on && n >= 2) {
      const Op& hd = plan.ops[n - 1]; const Op& ln = plan.ops[n - 2];
      if (hd.kind == OP_LINEAR && hd.engine == 0 && plan.acts[hd.out.buf].ext == 3 && smalln_ok(hd.out.cols, hd.in0.cols) && !hd.relu &&
          ln.kind == OP_LNRD && ln.out.buf == hd.in0.buf && ln.out.cols <= 512 && hd.in0.cols % 4 == 0) {
        c.tail_fuse = true; c.tail_ln_op = n - 2; c.tail_head_op = n - 1;
      }
    }
  }
  rc = run_forward(c);
  if (rc != FB200_OK) return rc;
  if (c.mega) {
    CUDA_OK(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), c.st));      // the row-parallel cross entropy accumulates the loss with atomics
    mega_emit_ce(c, logits, labels, class_w, denom, loss_out, dlog);
    rc = run_backward(c);
    if (rc != FB200_OK) return rc;
    return mega_finish(c);
  }
  if (c.tail_fuse && c.tail_have_fwd) {
    // cross entropy inside the tail kernel: loss_out accumulates by atomics, the batch's own weight sum comes from a 1-CTA kernel
    float* den_local = (float*)(w + plan.mega_bar_off);          // 256 spare bytes of the workspace (unused outside the step kernel)
    CUDA_OK(cudaMemsetAsync(loss_out, 0, 3 * sizeof(float), c.st));
    pdl_launch(ce_den_kernel, 1, 256, 0, c.st, labels, class_w, d->B, d->C, den_local);
    CUDA_OK(cudaGetLastError());
    c.tail.ce = CeArgs{(const float*)logits, labels, class_w, denom, loss_out, dlog, d->B, d->C, den_local};
    return run_backward(c);
  }
  rc = ce_launch(logits, labels, class_w, denom, d->B, d->C, loss_out, dlog, c.st);
  if (rc != FB200_OK) return rc;
  return run_backward(c);
}

// ---- primitives -----------------------------------------------------------------------
int fb200_gemm_workspace_bytes(int layout, int engine, int M, int N, int K, size_t* bytes) {
  if (!bytes || layout < 0 || layout > 2 || engine < 0 || engine > 2 || M < 1 || N < 1 || K < 1) return FB200_EBADARG;
  *bytes = 256;
  if (engine != 0) return tc_gemm_workspace_bytes(layout, engine, M, N, K, bytes);
  return FB200_OK;
}

int fb200_gemm(int layout, int engine, int M, int N, int K, const float* A, int lda, const float* B, int ldb,
               float* C, int ldc, const float* bias, int relu, int accumulate, void* ws, size_t ws_bytes, void* stream) {
  if (!A || !B || !C || layout < 0 || layout > 2 || M < 1 || N < 1 || K < 1) return FB200_EBADARG;
  if (!is_device_ptr(A) || !is_device_ptr(C)) return FB200_EUNSUPPORTED;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  if (dev.cc_major < 10) return FB200_EUNSUPPORTED;
  if (engine == 0) {
    GemmArgs g{};
    g.A = make_ref((void*)A, lda, FMT_F32); g.B = make_ref((void*)B, ldb, FMT_F32); g.C = make_ref(C, ldc, FMT_F32);
    g.M = M; g.N = N; g.K = K;
    g.a_kc = (layout == 2) ? 0 : 1;            // TN: A is [K,M]
    g.b_kc = (layout == 0) ? 1 : 0;            // NT: B is [N,K]
    g.bias = bias; g.relu = relu; g.mask_src.p = nullptr; g.accumulate = accumulate; g.split_k = 1; g.colsum_a = nullptr;
    if (layout == 2 && !bias && !relu && ldc == N) {     // weight-gradient shape: same auto split-K as the train step
      g.split_k = 0;
      if (!accumulate) CUDA_OK(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), (cudaStream_t)stream));
    }
    CUDA_OK(launch_simt_gemm(g, dev.num_sms, (cudaStream_t)stream));
    return FB200_OK;
  }
  return tc_gemm_f32(layout, engine, M, N, K, A, lda, B, ldb, C, ldc, bias, relu, accumulate, ws, ws_bytes, dev.num_sms, (cudaStream_t)stream);
}

// ---- multi-head attention on token sequences -------------------------------------------------------
namespace {
struct MhaLayout { size_t q, k, v, o, lse, delta, dO, dq, dk, dv, pool, dpool, total; };
MhaLayout mha_layout(const fb200_mha_desc& d) {
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  const size_t nq = (size_t)d.Sq * d.B * d.D * sizeof(float), nk = (size_t)d.Skv * d.B * d.D * sizeof(float);
  const size_t ns = (size_t)d.B * d.H * d.Sq * sizeof(float);
  MhaLayout L; size_t c = 0;
  L.q = c; c = al(c + nq); L.k = c; c = al(c + nk); L.v = c; c = al(c + nk); L.o = c; c = al(c + nq);
  L.lse = c; c = al(c + ns); L.delta = c; c = al(c + ns);
  L.dO = c; c = al(c + nq); L.dq = c; c = al(c + nq); L.dk = c; c = al(c + nk); L.dv = c; c = al(c + nk);
  const size_t np = (size_t)d.B * d.D * sizeof(float);
  L.pool = c; c = al(c + np); L.dpool = c; c = al(c + np);
  L.total = c + 256;
  return L;
}
int mha_check(const fb200_mha_desc* d) {
  if (!d || d->Sq < 1 || d->Skv < 1 || d->B < 1 || d->D < 4 || d->H < 1) return FB200_EBADARG;
  if (d->D % d->H != 0) return FB200_EBADARG;                  // nn.MultiheadAttention asserts embed_dim % num_heads == 0
  if (d->D % 4 != 0 || d->D / d->H > 256) return FB200_EUNSUPPORTED;
  return FB200_OK;
}
// fp32 GEMM on the engine that fits: tcgen05 3xTF32 once the row dimension exceeds 32 and the strides are TMA-legal, else FFMA
int gemm_auto(int layout, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
              const float* bias, int accumulate, void* stream, int relu = 0) {
  const int rows = layout == 2 ? K : M;
  const bool aligned = !((((uintptr_t)A) | ((uintptr_t)B) | ((uintptr_t)C)) & 15) && lda % 4 == 0 && ldb % 4 == 0 && ldc % 4 == 0;
  const int engine = (rows > 32 && aligned && tc_shape_ok(layout, M, N, K)) ? 1 : 0;
  return fb200_gemm(layout, engine, M, N, K, A, lda, B, ldb, C, ldc, bias, relu, accumulate, nullptr, 0, stream);
}
int colsum_rows(const float* const* xs, float* const* dsts, int n, int rows, int N, int num_sms, cudaStream_t st) {
  ColsumBatch cb{}; cb.B = rows; cb.nseg = n;
  for (int i = 0; i < n; ++i) cb.seg[i] = ColsumSeg{make_ref((void*)xs[i], N, FMT_F32), N, dsts[i]};
  int gx = (rows + 63) / 64; if (gx > 8 * num_sms / n + 1) gx = 8 * num_sms / n + 1; if (gx < 1) gx = 1;
  pdl_launch(colsum_batch_kernel, dim3(gx, n), 256, 0, st, cb);
  return cudaGetLastError() == cudaSuccess ? FB200_OK : FB200_ECUDA;
}
}  // namespace

}  // extern "C" paused: a device kernel
namespace fb200 {
// column sums of a row-major matrix of any width / stride (bias gradient of an output layer whose width is not a multiple of 4)
__global__ void __launch_bounds__(256) colsum_any_kernel(const float* __restrict__ y, int ld, int M, int N, float* __restrict__ out) { pdl_sync();
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, rg = threadIdx.x >> 5, c = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (c < N) for (int r = rg; r < M; r += 8) s += __ldg(y + (size_t)r * ld + c);
  red[rg][lane] = s;
  __syncthreads();
  if (rg == 0 && c < N) { for (int g = 1; g < 8; ++g) s += red[g][lane]; out[c] = s; }
}
}  // namespace fb200
extern "C" {
// ---- nn.Linear with row strides (the GEMMs around the TabTransformer encoder: tab_transformer.py:30, :33-38) -----------------
int fb200_linear_forward(int M, int N, int K, const float* x, int ldx, const float* W, const float* bias, int relu,
                         float* y, int ldy, void* stream) {
  if (!x || !W || !y || M < 1 || N < 1 || K < 1 || ldx < K || ldy < N) return FB200_EBADARG;
  if (!is_device_ptr(W) || (bias && !is_device_ptr(bias))) return FB200_EUNSUPPORTED;          // x and y are checked by fb200_gemm
  return gemm_auto(0, M, N, K, x, ldx, W, K, y, ldy, bias, 0, stream, relu ? 1 : 0);
}

int fb200_linear_backward(int M, int N, int K, const float* x, int ldx, const float* W, const float* dy, int lddy,
                          float* dx, int lddx, float* dW, float* db, void* stream) {
  if (!x || !W || !dy || !dW || M < 1 || N < 1 || K < 1 || ldx < K || lddy < N || (dx && lddx < K)) return FB200_EBADARG;
  if (!is_device_ptr(W) || !is_device_ptr(x) || (db && !is_device_ptr(db))) return FB200_EUNSUPPORTED;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  rc = gemm_auto(2, N, K, M, dy, lddy, x, ldx, dW, K, nullptr, 0, stream); if (rc != FB200_OK) return rc;      // dW = dy^T x
  if (dx) { rc = gemm_auto(1, M, K, N, dy, lddy, W, K, dx, lddx, nullptr, 0, stream); if (rc != FB200_OK) return rc; }   // dx = dy W
  if (db) {
    if (N % 4 == 0 && lddy % 4 == 0 && !(((uintptr_t)dy | (uintptr_t)db) & 15)) {
      CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * sizeof(float), (cudaStream_t)stream));
      const float* xs[1] = {dy}; float* ds[1] = {db};
      ColsumBatch cb{}; cb.B = M; cb.nseg = 1; cb.seg[0] = ColsumSeg{make_ref((void*)dy, lddy, FMT_F32), N, db};
      int gx = (M + 63) / 64; if (gx > 8 * dev.num_sms + 1) gx = 8 * dev.num_sms + 1;
      (void)xs; (void)ds;
      pdl_launch(colsum_batch_kernel, dim3(gx, 1), 256, 0, (cudaStream_t)stream, cb);
    } else {
      pdl_launch(colsum_any_kernel, (N + 31) / 32, 256, 0, (cudaStream_t)stream, dy, lddy, M, N, db);      // e.g. the 85-wide output layer
    }
    CUDA_OK(cudaGetLastError());
  }
  return FB200_OK;
}

int fb200_mha_workspace_bytes(const fb200_mha_desc* d, size_t* bytes) {
  int rc = mha_check(d); if (rc != FB200_OK) return rc;
  if (!bytes) return FB200_EBADARG;
  *bytes = mha_layout(*d).total;
  return FB200_OK;
}

int fb200_mha_forward(const fb200_mha_desc* d, const float* query, const float* key, const float* value,
                      const float* in_proj_weight, const float* in_proj_bias, const float* out_proj_weight, const float* out_proj_bias,
                      float* out, void* ws, void* stream) {
  int rc = mha_check(d); if (rc != FB200_OK) return rc;
  if (!query || !key || !value || !in_proj_weight || !in_proj_bias || !out_proj_weight || !out_proj_bias || !out || !ws) return FB200_EBADARG;
  if (!is_device_ptr(query) || !is_device_ptr(ws)) return FB200_EUNSUPPORTED;
  if ((((uintptr_t)ws) & 255)) return FB200_EALIGN;
  const MhaLayout L = mha_layout(*d);
  char* w = (char*)ws;
  const int D = d->D, Mq = d->Sq * d->B, Mk = d->Skv * d->B;
  float* Qp = (float*)(w + L.q); float* Kp = (float*)(w + L.k); float* Vp = (float*)(w + L.v); float* Oc = (float*)(w + L.o);
  // packed in_proj_weight = [W_q; W_k; W_v] (torch.nn.MultiheadAttention)
  rc = gemm_auto(0, Mq, D, D, query, D, in_proj_weight, D, Qp, D, in_proj_bias, 0, stream); if (rc != FB200_OK) return rc;
  rc = gemm_auto(0, Mk, D, D, key, D, in_proj_weight + (size_t)D * D, D, Kp, D, in_proj_bias + D, 0, stream); if (rc != FB200_OK) return rc;
  rc = gemm_auto(0, Mk, D, D, value, D, in_proj_weight + (size_t)2 * D * D, D, Vp, D, in_proj_bias + 2 * D, 0, stream); if (rc != FB200_OK) return rc;
  AttnArgs a{};
  a.Q = Qp; a.K = Kp; a.V = Vp; a.ldq = a.ldk = a.ldv = D; a.O = Oc; a.ldo = D; a.lse = (float*)(w + L.lse);
  a.Sq = d->Sq; a.Sk = d->Skv; a.B = d->B; a.H = d->H; a.hd = D / d->H; a.scale = 1.0f / sqrtf((float)a.hd);
  CUDA_OK(attn_tc_ok(a) ? launch_attn_tc_fwd(a, (cudaStream_t)stream) : launch_attn_fwd(a, (cudaStream_t)stream));   // tensor-core core for hd 16 / 32 / 64
  if (d->flags & FB200_MHA_POOL_MEAN) {          // out [B, D] = out_proj(mean over the query tokens): the pooling commutes with the projection
    float* pooled = (float*)(w + L.pool);
    int grid = (d->B * D / 4 + 255) / 256; if (grid > 1184) grid = 1184;
    pdl_launch(attn_mean_pool_kernel, grid, 256, 0, (cudaStream_t)stream, (const float*)Oc, d->Sq, d->B, D, pooled);
    CUDA_OK(cudaGetLastError());
    return gemm_auto(0, d->B, D, D, pooled, D, out_proj_weight, D, out, D, out_proj_bias, 0, stream);
  }
  return gemm_auto(0, Mq, D, D, Oc, D, out_proj_weight, D, out, D, out_proj_bias, 0, stream);
}

int fb200_mha_backward(const fb200_mha_desc* d, const float* query, const float* key, const float* value,
                       const float* in_proj_weight, const float* out_proj_weight, const float* dout,
                       float* dquery, float* dkey, float* dvalue, float* d_in_proj_weight, float* d_in_proj_bias,
                       float* d_out_proj_weight, float* d_out_proj_bias, void* ws, void* stream) {
  int rc = mha_check(d); if (rc != FB200_OK) return rc;
  if (!query || !key || !value || !in_proj_weight || !out_proj_weight || !dout || !d_in_proj_weight || !d_in_proj_bias ||
      !d_out_proj_weight || !d_out_proj_bias || !ws) return FB200_EBADARG;
  if (!is_device_ptr(dout) || !is_device_ptr(ws)) return FB200_EUNSUPPORTED;
  DeviceInfo dev; rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const MhaLayout L = mha_layout(*d);
  char* w = (char*)ws;
  const int D = d->D, Mq = d->Sq * d->B, Mk = d->Skv * d->B;
  float* Qp = (float*)(w + L.q); float* Kp = (float*)(w + L.k); float* Vp = (float*)(w + L.v); float* Oc = (float*)(w + L.o);
  float* dOc = (float*)(w + L.dO); float* dQp = (float*)(w + L.dq); float* dKp = (float*)(w + L.dk); float* dVp = (float*)(w + L.dv);
  const bool pool = d->flags & FB200_MHA_POOL_MEAN;
  if (pool) {
    // dout is [B, D]: dW_o = dout^T pooled, dPooled = dout W_o, then every token row of O receives dPooled / S_q
    float* pooled = (float*)(w + L.pool); float* dpooled = (float*)(w + L.dpool);
    rc = gemm_auto(2, D, D, d->B, dout, D, pooled, D, d_out_proj_weight, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
    rc = gemm_auto(1, d->B, D, D, dout, D, out_proj_weight, D, dpooled, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
    pdl_launch(attn_mean_pool_bwd_kernel, 1184, 256, 0, st, (const float*)dpooled, d->Sq, d->B, D, dOc);
    CUDA_OK(cudaGetLastError());
  } else {
  // out_proj: dW_o = dout^T O, db_o = colsum(dout), dO = dout W_o
  rc = gemm_auto(2, D, D, Mq, dout, D, Oc, D, d_out_proj_weight, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
  rc = gemm_auto(1, Mq, D, D, dout, D, out_proj_weight, D, dOc, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
  }
  AttnArgs a{};
  a.Q = Qp; a.K = Kp; a.V = Vp; a.ldq = a.ldk = a.ldv = D; a.O = Oc; a.ldo = D; a.lse = (float*)(w + L.lse); a.delta = (float*)(w + L.delta);
  a.dO = dOc; a.lddo = D; a.dQ = dQp; a.dK = dKp; a.dV = dVp; a.lddq = a.lddk = a.lddv = D;
  a.Sq = d->Sq; a.Sk = d->Skv; a.B = d->B; a.H = d->H; a.hd = D / d->H; a.scale = 1.0f / sqrtf((float)a.hd);
  CUDA_OK(attn_tc_ok(a) ? launch_attn_tc_bwd(a, st) : launch_attn_bwd(a, st));
  // in_proj: dW = [dQ^T query; dK^T key; dV^T value], db = column sums, and the input gradients
  rc = gemm_auto(2, D, D, Mq, dQp, D, query, D, d_in_proj_weight, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
  rc = gemm_auto(2, D, D, Mk, dKp, D, key, D, d_in_proj_weight + (size_t)D * D, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
  rc = gemm_auto(2, D, D, Mk, dVp, D, value, D, d_in_proj_weight + (size_t)2 * D * D, D, nullptr, 0, stream); if (rc != FB200_OK) return rc;
  CUDA_OK(cudaMemsetAsync(d_in_proj_bias, 0, (size_t)3 * D * sizeof(float), st));
  CUDA_OK(cudaMemsetAsync(d_out_proj_bias, 0, (size_t)D * sizeof(float), st));
  if (pool) {
    { const float* xs[1] = {dout}; float* ds[1] = {d_out_proj_bias}; rc = colsum_rows(xs, ds, 1, d->B, D, dev.num_sms, st); if (rc != FB200_OK) return rc; }
    { const float* xs[1] = {dQp}; float* ds[1] = {d_in_proj_bias}; rc = colsum_rows(xs, ds, 1, Mq, D, dev.num_sms, st); if (rc != FB200_OK) return rc; }
  } else {
    const float* xs[2] = {dout, dQp}; float* ds[2] = {d_out_proj_bias, d_in_proj_bias};
    rc = colsum_rows(xs, ds, 2, Mq, D, dev.num_sms, st); if (rc != FB200_OK) return rc;
  }
  { const float* xs[2] = {dKp, dVp}; float* ds[2] = {d_in_proj_bias + D, d_in_proj_bias + 2 * D};
    rc = colsum_rows(xs, ds, 2, Mk, D, dev.num_sms, st); if (rc != FB200_OK) return rc; }
  if (dquery) { rc = gemm_auto(1, Mq, D, D, dQp, D, in_proj_weight, D, dquery, D, nullptr, 0, stream); if (rc != FB200_OK) return rc; }
  if (dkey) { rc = gemm_auto(1, Mk, D, D, dKp, D, in_proj_weight + (size_t)D * D, D, dkey, D, nullptr, 0, stream); if (rc != FB200_OK) return rc; }
  if (dvalue) { rc = gemm_auto(1, Mk, D, D, dVp, D, in_proj_weight + (size_t)2 * D * D, D, dvalue, D, nullptr, 0, stream); if (rc != FB200_OK) return rc; }
  return FB200_OK;
}

int fb200_ln_relu_dropout_fwd(const float* x, const float* gamma, const float* beta, const uint8_t* mask, float p, int train,
                              uint64_t seed, uint64_t offset, int site, int B, int N, float* y, float* stats, void* stream) {
  if (!x || !gamma || !beta || !y || !stats || B < 1 || N < 4 || N % 4 || N > 4096) return FB200_EBADARG;
  if (!is_device_ptr(x)) return FB200_EUNSUPPORTED;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  LnrdArgs a{}; a.x = make_ref((void*)x, N); a.y = make_ref(y, N); a.gamma = gamma; a.beta = beta; a.stats = stats;
  a.drop.mask = mask; a.drop.seed = seed; a.drop.offset = offset; a.drop.state = nullptr; a.drop.p = p; a.drop.site = site; a.drop.active = (train && p > 0.f) ? 1 : 0;
  a.B = B; a.N = N;
  const int grid = row_grid_for(B, N, dev.num_sms);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(NV, TPR) pdl_launch(lnrd_fwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, st, a)
  FB200_ROW_DISPATCH(N, CALL);
#undef CALL
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

int fb200_ln_relu_dropout_bwd(const float* x, const float* y, const float* gamma, const float* stats, const float* dy, float p, int train,
                              int B, int N, float* dx, float* dgamma, float* dbeta, void* stream) {
  if (!x || !y || !gamma || !stats || !dy || !dx || !dgamma || !dbeta || B < 1 || N < 4 || N % 4 || N > 4096) return FB200_EBADARG;
  if (!is_device_ptr(x)) return FB200_EUNSUPPORTED;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_OK(cudaMemsetAsync(dgamma, 0, N * sizeof(float), st));
  CUDA_OK(cudaMemsetAsync(dbeta, 0, N * sizeof(float), st));
  LnrdArgs a{}; a.x = make_ref((void*)x, N); a.y = make_ref((void*)y, N); a.dy = make_ref((void*)dy, N); a.dx = make_ref(dx, N);
  a.gamma = gamma; a.stats = (float*)stats; a.dgamma = dgamma; a.dbeta = dbeta;
  a.drop.mask = nullptr; a.drop.state = nullptr; a.drop.p = p; a.drop.active = (train && p > 0.f) ? 1 : 0; a.B = B; a.N = N;
  const int grid = row_grid_for(B, N, dev.num_sms);
#define CALL(NV, TPR) pdl_launch(lnrd_bwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, st, a)
  FB200_ROW_DISPATCH(N, CALL);
#undef CALL
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

int fb200_metablock_fwd(const float* v, const float* f, const float* g, const float* gamma_f, const float* beta_f,
                        const float* gamma_g, const float* beta_g, int B, int N, float* y, float* stats, void* stream) {
  if (!v || !f || !g || !gamma_f || !beta_f || !gamma_g || !beta_g || !y || !stats || B < 1 || N < 4 || N % 4 || N > 4096) return FB200_EBADARG;
  if (!is_device_ptr(v)) return FB200_EUNSUPPORTED;
  DeviceInfo dev; int rc = get_device_info(dev); if (rc != FB200_OK) return rc;
  MetaArgs a{}; a.v = make_ref((void*)v, N); a.f = make_ref((void*)f, N); a.g = make_ref((void*)g, N); a.y = make_ref(y, N);
  a.gamma_f = gamma_f; a.beta_f = beta_f; a.gamma_g = gamma_g; a.beta_g = beta_g; a.stats = stats; a.B = B; a.N = N;
  const int grid = row_grid_for(B, N, dev.num_sms);
  cudaStream_t st = (cudaStream_t)stream;
#define CALL(NV, TPR) pdl_launch(meta_fwd_kernel<NV, TPR>, grid, ROW_WARPS * 32, 0, st, a)
  FB200_ROW_DISPATCH(N, CALL);
#undef CALL
  CUDA_OK(cudaGetLastError());
  return FB200_OK;
}

}  // extern "C"
