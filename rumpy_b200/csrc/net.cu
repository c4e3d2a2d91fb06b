// Whole-network executor for the RCAN / EDSR trunk: one C call enqueues every kernel of a forward pass, one
// more enqueues the whole backward pass.
//
// Mirrors the reference's module graph (/root/reference/rumpy/SISR/models/advanced/architectures.py):
//   RCAN.forward :171-176, ResidualGroup.forward :121-124, RCAB.forward :81-84, CALayer.forward :41-44,
//   EDSR.forward :236-241, ResBlock.forward common.py:71-75, Upsampler common.py:29-44,
// and, for backward, what autograd derives from them (`loss.backward()`, base_architecture.py:432; SURVEY 8a').
// The layer program is built once per (shape, workspace, packed-weights, gradient buffers) and cached: building
// encodes every TMA tensor map on the host, so a steady-state step is just kernel launches (graph capturable).
#include "../../include/rumpy_b200.h"
#include "host_util.cuh"
#include "wgrad_tc.cuh"
#include "bwd_kernels.cuh"
#include <cmath>
#include <algorithm>
#include <memory>
#include <vector>

namespace rb {

// implemented in wgrad.cu
int wgrad_launch(const WgradJob* jobs_dev, int njobs, const WgradReduceJob* rjobs_dev, int nrjobs, cudaStream_t s);

struct ConvW {            // one 3x3 conv of the network
  int w_idx, b_idx;       // indices into the state_dict-ordered parameter list
  int cout, cin;
  int rows_padded;        // packed rows (16 for the thin tail)
  int r;                  // pixel-shuffle factor folded into the packing (1 = none)
  size_t off_fwd;         // byte offsets into the packed buffer
  size_t off_bias;        // packed bias (only when r > 1 or padded), else SIZE_MAX
  size_t off_dgrad;       // dgrad-packed weights (training), else SIZE_MAX
};

struct CAW { int w1, b1, w2, b2; };
using QScaleJob = QScaleJobHost;   // q_scale_kernel job (w1 == nullptr: no q-layer)

enum OpType { OP_HEAD, OP_QSCALE, OP_CONV, OP_CA, OP_CONV_CA, OP_TRUNK, OP_TRUNK_BWD, OP_TAIL_BWD, OP_CA_BWD, OP_ADD, OP_HEAD_WGRAD, OP_LAM, OP_CSAM, OP_DQ, OP_QGRAD, OP_CSAM_BWD, OP_LAM_BWD, OP_HAN_PG };


struct Op {
  OpType type;
  ConvPlan conv;          // OP_CONV / OP_CONV_CA
  CaFusedArgs cafused;    // OP_CONV_CA
  int bias_param;         // param index whose pointer is patched into conv.args.bias (-1: packed / none)
  bool writes_output;     // thin tail conv: out_nchw patched with the caller's y
  // OP_HEAD
  int head_w, head_b;
  float* yf; void* yb;
  // OP_CA / OP_CA_BWD
  CAW ca;
  const float* pool; const void* u; const float* x_in; float* x_out; void* x_out_b;
  float *save_mean, *save_hid, *save_y;
  float* s_partial; void* du; float* du_colsum; int ca_chunks;
  float* pool_compact;
  const float* q_scale;   // OP_CA (Q-RCAN): [N][C] multipliers of the CA vector, or nullptr
  // OP_LAM / OP_CSAM (HAN)
  const float* stack[kLamLayers]; float* lam_scratch; void* lam_out; const float* csam_x; const float* csam_out2; void* cat;
  float* han_dx[kLamLayers]; float* han_f[4]; void* han_v; int han_i;   // OP_CSAM_BWD / OP_LAM_BWD / OP_HAN_PG
  // OP_ADD / OP_HEAD_WGRAD / OP_TAIL_BWD
  const float *a, *b; float* dst_f; void* dst_b; size_t n4;
  const void* tail_in; void* g_hr; float* thin_partial;
  int h, w;
};

struct BlockRec { const void* in_b; void* t; void* u; float* sv; int conv1, conv2, ca; };
struct GroupRec { std::vector<BlockRec> blocks; const void* tail_in_b; int conv_tail; };

struct Net {
  Options opt;                // per-handle execution options (rumpy_net_set_option)
  int arch, C, n_groups, n_blocks, reduction, scale;
  float res_scale;
  int in_feats, out_feats, u_f32;
  int plan_u_f32 = 1;         // dtype of the saved pre-attention activation of the cached plan (bf16 with OP_TRUNK)
  std::vector<ConvW> convs;   // network order
  std::vector<CAW> cas;
  // Q-RCAN (meta-attention, attention_manipulators/architectures.py:154-246, q_layer.py:5-45): per-RCAB q-layer
  // parameters (w1 < 0: the RCAB has none), metadata vector length, q-layer hidden width, 'modulate' style flag
  // HAN (advanced/architectures.py:331-394): RCAN groups + layer attention + channel-spatial attention; parameter
  // indices of csa.gamma, csa.conv.weight, csa.conv.bias, la.gamma and the two extra convs (last_conv, last)
  bool han = false;
  int p_csa_gamma = -1, p_csa_w = -1, p_csa_b = -1, p_la_gamma = -1, conv_lastconv = -1, conv_last = -1;
  bool qrcan = false;
  std::vector<CAW> qs;
  int num_meta = 0, q_hidden = 0, modulate = 0, q_relu = 1;
  const float* meta_dev = nullptr;
  int meta_n = 0, meta_m = 0;
  float* q_dq = nullptr;                    // training: [n_rcab][N][64] d(loss)/dq from the backward kernel
  QGradJobHost* qg_jobs_dev = nullptr;
  std::vector<QGradJobHost> qg_jobs_uploaded;
  float* q_scale = nullptr;                 // [n_rcab][N][64] per-(image, channel) multipliers of the CA vector
  QScaleJob* q_jobs_dev = nullptr;
  std::vector<QScaleJob> q_jobs_uploaded;
  int n_params = 0;
  size_t packed_bytes = 0, packed_bytes_train = 0;
  int conv_body = 0, conv_up0 = 0, conv_tail = 0;
  // cached plan
  std::vector<Op> ops, bops;
  const void* plan_packed = nullptr;
  void* plan_ws = nullptr;
  int pN = 0, pH = 0, pW = 0, p_training = -1, p_trunk = -1;
  // backward job lists (host copies; device copies live in the workspace)
  std::vector<float*> plan_grads;
  std::vector<WgradJob> wg_jobs;
  std::vector<WgradReduceJob> wg_rjobs;   // .accumulate temporarily carries the weight's param index
  std::vector<ColsumJob> cs_jobs;
  std::vector<int> cs_bias_param;
  std::vector<PartialSumJob> ps_jobs;
  std::vector<int> ps_bias_param;
  PartialSumJob* ps_jobs_dev = nullptr;
  WgradJob* wg_jobs_dev = nullptr;
  WgradReduceJob* wg_rjobs_dev = nullptr;
  ColsumJob* cs_jobs_dev = nullptr;
  bool jobs_uploaded = false;
  // the batched wgrad runs in chunks ordered from the LAST parameters to the first; after chunk k every gradient with
  // parameter index >= wg_chunk_first_param[k] is final (an event is recorded there so that the caller's NCCL
  // all-reduce of that range overlaps the remaining chunks)
  std::vector<int> wg_chunk_job_end, wg_chunk_rjob_end, wg_chunk_first_param;
  cudaEvent_t bwd_events[8] = {};
  int n_bwd_events = 0;
  float* pg_scratch = nullptr;
  int* pg_counter = nullptr;
  float* ca_coef = nullptr;
  std::unique_ptr<TrunkPlan> trunk;            // persistent trunk kernel plan (OP_TRUNK), or null
  std::unique_ptr<TrunkBwdPlan> trunk_bwd;     // backward program of the body (OP_TRUNK_BWD), or null
  unsigned long long* ca_counters = nullptr;   // grid-barrier counters of the fused conv2+CA ops
  size_t ca_counters_bytes = 0;
  bool ca_counters_dirty = false;
  // batched weight packing
  size_t pack_jobs_bytes = 0;
  std::vector<PackJobHost> pack_jobs;
  const void* pack_sig_packed = nullptr;
  const float* pack_sig_p0 = nullptr;
  int pack_sig_training = -1;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static void add_conv(Net* n, int cout, int cin, int r = 1, int rows_padded = 0) {
  ConvW c{};
  c.w_idx = n->n_params++;
  c.b_idx = n->n_params++;
  c.cout = cout; c.cin = cin; c.r = r;
  c.rows_padded = rows_padded ? rows_padded : cout;
  c.off_fwd = n->packed_bytes;
  n->packed_bytes = align_up(n->packed_bytes + size_t(9) * c.rows_padded * cin * 2, 256);
  if (r > 1 || c.rows_padded != cout) {
    c.off_bias = n->packed_bytes;
    n->packed_bytes = align_up(n->packed_bytes + size_t(c.rows_padded) * 4, 256);
  } else {
    c.off_bias = SIZE_MAX;
  }
  c.off_dgrad = SIZE_MAX;
  n->convs.push_back(c);
}

static int upsampler_stages(int scale, int* r) {
  if (scale == 3) { *r = 3; return 1; }
  *r = 2;
  int s = 0;
  while ((1 << s) < scale) ++s;
  return ((1 << s) == scale) ? s : -1;
}

// Parameter order == reference state_dict order: head, body..., tail.
static int net_init(Net* n) {
  const int C = n->C;
  add_conv(n, C, n->in_feats);  // head (CUDA-core kernel; packed slot unused but keeps indexing uniform)
  if (n->arch == 0) {
    for (int g = 0; g < n->n_groups; ++g) {
      for (int b = 0; b < n->n_blocks; ++b) {
        add_conv(n, C, C);
        add_conv(n, C, C);
        CAW ca{n->n_params, n->n_params + 1, n->n_params + 2, n->n_params + 3};
        n->n_params += 4;
        n->cas.push_back(ca);
      }
      add_conv(n, C, C);
    }
  } else {
    for (int b = 0; b < n->n_blocks; ++b) { add_conv(n, C, C); add_conv(n, C, C); }
  }
  n->conv_body = int(n->convs.size());
  add_conv(n, C, C);  // body tail
  int r = 0;
  const int st = upsampler_stages(n->scale, &r);
  if (st < 0) return set_error(RUMPY_ERR_ARG, "scale %d unsupported (2^n or 3)", n->scale);
  n->conv_up0 = int(n->convs.size());
  for (int s = 0; s < st; ++s) add_conv(n, C * r * r, C, r);
  n->conv_tail = int(n->convs.size());
  add_conv(n, n->out_feats, C, 1, 16);  // thin tail
  if (n->qrcan && n->arch == 0) {
    // QRCAN registers its modules in a different order than RCAN (attention_manipulators/architectures.py:313-433,
    // 154-196, 249-294): final_body | head | per group: final_body, per block: QCALayer, [q_node], conv1, conv2 | tail
    // (QHAN, :643-760: head | groups | body conv -- the body conv is the last entry of `body` there, not a separate
    // final_body registered first)
    int p = 0;
    auto set = [&](int ci) { n->convs[ci].w_idx = p++; n->convs[ci].b_idx = p++; };
    if (!n->han) set(n->conv_body);
    set(0);
    const int B = n->n_blocks;
    for (int g = 0; g < n->n_groups; ++g) {
      set(1 + g * (2 * B + 1) + 2 * B);
      for (int b = 0; b < B; ++b) {
        const int rc = g * B + b;
        n->cas[rc] = CAW{p, p + 1, p + 2, p + 3};
        p += 4;
        if (n->qs[rc].w1 >= 0) { n->qs[rc] = CAW{p, p + 1, p + 2, p + 3}; p += 4; }
        set(1 + g * (2 * B + 1) + 2 * b);
        set(1 + g * (2 * B + 1) + 2 * b + 1);
      }
    }
    if (n->han) set(n->conv_body);
    for (int s2 = 0; s2 < st; ++s2) set(n->conv_up0 + s2);
    set(n->conv_tail);
    n->n_params = p;
  }
  if (n->han) {
    // registration order (architectures.py:360-366): head | body | csa.gamma, csa.conv.{weight,bias} | la.gamma |
    // last_conv | last | tail
    n->conv_lastconv = int(n->convs.size());
    add_conv(n, C, C * (n->n_groups + 1));
    n->conv_last = int(n->convs.size());
    add_conv(n, C, 2 * C);
    int p = n->convs[n->conv_body].b_idx + 1;
    auto set = [&](int ci) { n->convs[ci].w_idx = p++; n->convs[ci].b_idx = p++; };
    n->p_csa_gamma = p++; n->p_csa_w = p++; n->p_csa_b = p++; n->p_la_gamma = p++;
    set(n->conv_lastconv);
    set(n->conv_last);
    for (int s2 = 0; s2 < st; ++s2) set(n->conv_up0 + s2);
    set(n->conv_tail);
    n->n_params = p;
  }
  if (n->qrcan && n->arch == 1) {
    // QEDSR registration order (attention_manipulators/architectures.py:501-548, 463-482):
    // head | final_body | per block: body.0, body.2, [attention_layer] | tail
    int p = 0;
    auto set = [&](int ci) { n->convs[ci].w_idx = p++; n->convs[ci].b_idx = p++; };
    set(0);
    set(n->conv_body);
    for (int b = 0; b < n->n_blocks; ++b) {
      set(1 + 2 * b);
      set(2 + 2 * b);
      if (n->qs[b].w1 >= 0) { n->qs[b] = CAW{p, p + 1, p + 2, p + 3}; p += 4; }
    }
    for (int s2 = 0; s2 < st; ++s2) set(n->conv_up0 + s2);
    set(n->conv_tail);
    n->n_params = p;
  }
  // dgrad operands (training): every tensor-core conv except the head (no dX needed) and the thin tail
  n->packed_bytes_train = n->packed_bytes;
  for (int i = 1; i < n->conv_tail; ++i) {
    n->convs[i].off_dgrad = n->packed_bytes_train;
    n->packed_bytes_train = align_up(n->packed_bytes_train + size_t(9) * n->convs[i].cout * n->convs[i].cin * 2, 256);
  }
  if (n->han)
    for (int i : {n->conv_lastconv, n->conv_last}) {
      n->convs[i].off_dgrad = n->packed_bytes_train;
      n->packed_bytes_train = align_up(n->packed_bytes_train + size_t(9) * n->convs[i].cout * n->convs[i].cin * 2, 256);
    }
  n->pack_jobs_bytes = align_up(2 * n->convs.size() * sizeof(PackJobHost), 256);
  n->packed_bytes += n->pack_jobs_bytes;
  n->packed_bytes_train += n->pack_jobs_bytes;
  return RUMPY_OK;
}

struct Bump {
  char* base; size_t off = 0;
  void* take(size_t bytes) { void* p = base ? base + off : nullptr; off = align_up(off + bytes, 1024); return p; }
};

static int wgrad_splits(int m_tiles) {
  int s = (m_tiles + opt().wgrad_tiles_per_split - 1) / opt().wgrad_tiles_per_split;
  if (s > 64) s = 64;
  if (s < 1) s = 1;
  return s;
}

constexpr int kThinBlocks = 1184;
constexpr int kPlaneSlices = 64;
constexpr int kCaBwdChunks = 32;

// Lays out the workspace and (when build) builds the op lists.  Forward buffers first, then (training) the
// backward buffers, so `bytes_out` covers a whole train step.
static int build_plan(Net* n, const void* packed, void* ws, int N, int H, int W, int training, size_t* bytes_out,
                      bool build) {
  const int C = n->C;
  const size_t px = size_t(N) * H * W;
  Bump bp{static_cast<char*>(ws)};
  const char* pk = static_cast<const char*>(packed);
  std::vector<Op> ops;
  int err = 0;
  auto conv_op = [&](std::vector<Op>& list, const ConvW& cw, ConvDesc d, bool dgrad) {
    Op op{};
    op.type = OP_CONV;
    if (dgrad) {
      d.w = pk + cw.off_dgrad;
      d.bias = nullptr;
      op.bias_param = -1;
    } else {
      d.w = pk + cw.off_fwd;
      if (cw.off_bias != SIZE_MAX) { d.bias = reinterpret_cast<const float*>(pk + cw.off_bias); op.bias_param = -1; }
      else { d.bias = reinterpret_cast<const float*>(16); op.bias_param = cw.b_idx; }  // patched per call
    }
    if (build) { if (int e = conv_plan_build(&op.conv, d)) err = e; }
    list.push_back(op);
  };
  // ---- buffers
  float* head_f = static_cast<float*>(bp.take(px * C * 4));
  void* head_b = bp.take(px * C * 2);
  float* S_f = static_cast<float*>(bp.take(px * C * 4));
  // fp32 group outputs: two ping-pong buffers (RCAN) or one per group (HAN stacks them for the layer attention)
  const bool han = n->han;
  const float* han_stack[kLamLayers] = {};   // kept for the backward plan
  float* han_lam_scratch = nullptr;
  void* han_lam_out = nullptr;
  void* han_cat = nullptr;
  std::vector<float*> G_bufs(han ? n->n_groups : 2);
  for (auto& g : G_bufs) g = static_cast<float*>(bp.take(px * C * 4));
  struct { std::vector<float*>* v; bool han; float* operator[](int g) const { return (*v)[han ? g : (g & 1)]; } } G_f{&G_bufs, han};
  float* han_body_f = han ? static_cast<float*>(bp.take(px * C * 4)) : nullptr;   // body conv output (no skip in HAN)
  const int tiles = ((H + kTileH - 1) / kTileH) * ((W + kTileW - 1) / kTileW);   // dataflow trunk kernels
  const int pool_rows = conv_pool_rows(H, W, C);   // per image, written by the pooled per-layer convs
  int ci = 0;  // conv cursor
  // ---- head
  {
    Op op{};
    op.type = OP_HEAD;
    op.head_w = n->convs[0].w_idx; op.head_b = n->convs[0].b_idx;
    op.yf = head_f; op.yb = head_b;
    ops.push_back(op);
    ci = 1;
  }
  // ---- Q-RCAN: per-(RCAB, image, channel) meta-attention multipliers, evaluated once per forward
  float* q_scale = nullptr;
  if (n->qrcan) {
    if (training && n->modulate) {
      // attributes * sigmoid(q-layer) would need the two factors apart in the backward; the reference's handler never
      // builds that combination (its 'modulate' attributes are n_feats wide, the q-layers take num_metadata inputs)
      for (const CAW& q : n->qs)
        if (q.w1 >= 0)
          return set_error(RUMPY_ERR_ARG, "Q-RCAN style 'modulate' combined with q-layers: inference only");
    }
    q_scale = static_cast<float*>(bp.take(n->qs.size() * size_t(N) * C * sizeof(float)));
    QScaleJob* qj = static_cast<QScaleJob*>(bp.take(n->qs.size() * sizeof(QScaleJob)));
    if (build) {
      n->q_scale = q_scale; n->q_jobs_dev = qj; n->q_jobs_uploaded.clear();
      Op op{};
      op.type = OP_QSCALE;
      ops.push_back(op);
    }
  }
  auto q_of = [&](int cai) -> const float* {
    if (!n->qrcan || (n->qs[cai].w1 < 0 && !n->modulate)) return nullptr;
    return q_scale + size_t(cai) * N * C;
  };
  const void* cur_b = head_b;     // bf16 operand of the running activation
  const float* cur_f = head_f;    // its fp32 residual-stream copy
  const int Cr = n->arch == 0 ? C / n->reduction : 1;
  // dtype of the pre-attention activation u = conv2(t) of the per-layer path.  Training keeps the handle's choice
  // (fp32 by default: the backward re-reads it); inference stores it in bf16 unless option "infer_u_bf16" is 0:
  // measured on a full-depth RCAN, 270x480 input: max-abs vs the CPU oracle 0.00874 (bf16) / 0.00876 (fp32) -- the
  // rounding of u (it is added to the fp32 stream scaled by y) is invisible next to the operands' -- and a 1080p
  // frame takes 158.6 instead of 168.1 ms (the HBM-bound CA pass moves 12 instead of 14 bytes per element).
  const bool u32 = n->u_f32 && (training || !opt().infer_u_bf16);
  // ---- the whole body (every 64->64 conv, CA, skips) as ONE persistent dataflow kernel when the shape fits
  const bool use_trunk = opt().use_trunk && C == 64 && trunk_supported(N, H, W, C, Cr);
  std::unique_ptr<TrunkPlan> trunk;
  std::vector<GroupRec> groups;
  const void* trunk_body_in = nullptr;
  if (use_trunk) {
    // Inference: two ping-pong operand buffers.  Training: every layer writes its own buffer (backward needs the
    // RCAB inputs, the post-ReLU t, the pre-attention u (kept in bf16) and the CA vectors).
    const int n_body = n->conv_body;                            // convs 1 .. conv_body are the 64->64 body convs
    void* pp[2] = {nullptr, nullptr};
    if (!training) { pp[0] = bp.take(px * C * 2); pp[1] = bp.take(px * C * 2); }
    void* dev = bp.take(trunk_device_bytes(N, H, W, n_body, 2 * n_body + 2, 2 * n_body + 2));
    if (build) trunk.reset(new TrunkPlan());
    int L = 0;
    auto in_of = [&](const void* p) {
      for (size_t i = 0; i < trunk->in_bufs.size(); ++i) if (trunk->in_bufs[i] == p) return int(i);
      trunk->in_bufs.push_back(p);
      return int(trunk->in_bufs.size()) - 1;
    };
    auto out_of = [&](void* p) {
      for (size_t i = 0; i < trunk->out_bufs.size(); ++i) if (trunk->out_bufs[i] == p) return int(i);
      trunk->out_bufs.push_back(p);
      return int(trunk->out_bufs.size()) - 1;
    };
    auto next_out = [&]() -> void* { return training ? bp.take(px * C * 2) : pp[L & 1]; };
    auto add_layer = [&](int kind, int conv, int ca, const void* in, void* out) -> TrunkLayer* {
      ++L;
      if (!build) return nullptr;
      TrunkLayer l{};
      l.kind = kind;
      l.in_map = in_of(in);
      l.out_map = out_of(out);
      l.ca_slot = -1; l.u_map = -1; l.alpha = 1.f;
      TrunkLayerParams lp{n->convs[conv].b_idx, -1, -1, -1, -1};
      if (ca >= 0) { const CAW& c = n->cas[ca]; lp.w1 = c.w1; lp.b1 = c.b1; lp.w2 = c.w2; lp.b2 = c.b2; l.ca_slot = ca; }
      trunk->layers.push_back(l);
      trunk->lparams.push_back(lp);
      return &trunk->layers.back();
    };
    int c2 = 1, cai = 0;
    if (n->arch == 0) {
      for (int g = 0; g < n->n_groups; ++g) {
        GroupRec gr;
        for (int b = 0; b < n->n_blocks; ++b) {
          void* t = next_out();
          add_layer(kTrunkRelu, c2, -1, cur_b, t);
          void* xb = next_out();
          void* u = training ? bp.take(px * C * 2) : nullptr;
          float* sv = training ? static_cast<float*>(bp.take(size_t(N) * (2 * C + Cr) * 4)) : nullptr;
          if (TrunkLayer* l = add_layer(kTrunkCA, c2 + 1, cai, t, xb)) {
            l->q_scale = q_of(cai);
            if (training) {
              l->u_map = out_of(u);
              l->save_mean = sv; l->save_y = sv + size_t(N) * C; l->save_hid = sv + size_t(N) * 2 * C;
            }
          }
          gr.blocks.push_back(BlockRec{cur_b, t, u, sv, c2, c2 + 1, cai});
          c2 += 2; ++cai;
          cur_b = xb;
        }
        void* gout = next_out();
        gr.tail_in_b = cur_b; gr.conv_tail = c2;
        if (TrunkLayer* l = add_layer(kTrunkRes, c2++, -1, cur_b, gout)) {   // group tail conv + group skip (:121-124)
          l->res_f32 = g == 0 ? head_f : G_f[g - 1];
          l->out_f32 = G_f[g];
          l->update_s = 1;
        }
        groups.push_back(gr);
        cur_b = gout;
      }
    } else {
      GroupRec gr;
      for (int b = 0; b < n->n_blocks; ++b) {
        void* t = next_out();
        add_layer(kTrunkRelu, c2, -1, cur_b, t);
        void* xb = next_out();
        if (TrunkLayer* l = add_layer(kTrunkRes, c2 + 1, -1, t, xb)) {       // conv2(.)*res_scale + x (common.py:72-73)
          l->alpha = n->res_scale; l->update_s = 1;
          l->q_scale = q_of(b);          // Q-EDSR: (conv2(.)*res_scale) * q + x   (ParamResBlock.forward :484-493)
        }
        gr.blocks.push_back(BlockRec{cur_b, t, nullptr, nullptr, c2, c2 + 1, -1});
        c2 += 2;
        cur_b = xb;
      }
      groups.push_back(gr);
    }
    trunk_body_in = cur_b;
    void* body_b = bp.take(px * C * 2);
    if (TrunkLayer* l = add_layer(kTrunkRes, c2++, -1, cur_b, body_b)) {     // body tail conv + global skip
      if (han) { l->no_res = 1; l->out_f32 = han_body_f; }                   // HAN: plain conv, fp32 copy for LAM / CSAM
      else l->res_f32 = head_f;
    }
    if (build) {
      if (c2 != n_body + 1) err = set_error(RUMPY_ERR_ARG, "trunk: layer count mismatch");
      if (int e = trunk_plan_finish(trunk.get(), N, H, W, Cr, pk + n->convs[1].off_fwd, head_f, dev, !training)) err = e;
      Op op{};
      op.type = OP_TRUNK;
      ops.push_back(op);
    }
    ci = n_body + 1;
    cur_b = body_b;
  } else if (n->arch == 0) {
    const size_t u_bytes = px * C * (u32 ? 4 : 2);
    void* xb_shared = training ? nullptr : bp.take(px * C * 2);
    void* t_shared = training ? nullptr : bp.take(px * C * 2);
    void* u_shared = training ? nullptr : bp.take(u_bytes);
    float* pool_shared = training ? nullptr : static_cast<float*>(bp.take(size_t(N) * pool_rows * C * 4));
    void* gb[2] = {bp.take(px * C * 2), bp.take(px * C * 2)};
    float* pool_compact = static_cast<float*>(bp.take(size_t(N) * 64 * C * 4));
    const size_t n_rcab = size_t(n->n_groups) * n->n_blocks;
    unsigned long long* ca_counters = static_cast<unsigned long long*>(bp.take(n_rcab * sizeof(unsigned long long)));
    const bool fuse_ca = opt().use_fused_ca && !n->qrcan && u32 && build && conv_ca_supported(N, H, W, C, C);
    if (build) { n->ca_counters = ca_counters; n->ca_counters_bytes = n_rcab * sizeof(unsigned long long); n->ca_counters_dirty = true; }
    int cai = 0;
    for (int g = 0; g < n->n_groups; ++g) {
      GroupRec gr;
      const float* gin_f = cur_f;
      for (int b = 0; b < n->n_blocks; ++b) {
        void* t = training ? bp.take(px * C * 2) : t_shared;
        void* u = training ? bp.take(u_bytes) : u_shared;
        void* xb = training ? bp.take(px * C * 2) : xb_shared;
        float* pool = training ? static_cast<float*>(bp.take(size_t(N) * pool_rows * C * 4)) : pool_shared;
        float* sv = training ? static_cast<float*>(bp.take(size_t(N) * (2 * C + Cr) * 4)) : nullptr;
        BlockRec br{cur_b, t, u, sv, ci, ci + 1, cai};
        ConvDesc d1{};
        d1.x = cur_b; d1.y_bf16 = t; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C; d1.Cout = C; d1.flags = kConvRelu;
        d1.alpha = 1.f;
        conv_op(ops, n->convs[ci++], d1, false);
        const float* ca_x_in = (b == 0) ? gin_f : S_f;
        if (fuse_ca) {
          // conv2 + CALayer + RCAB skip in one kernel: accumulators wait in TMEM across a grid barrier
          const ConvW& cw = n->convs[ci++];
          Op op{};
          op.type = OP_CONV_CA;
          op.ca = n->cas[cai];
          op.bias_param = cw.b_idx;
          ConvDesc d2{};
          d2.x = t; d2.w = pk + cw.off_fwd; d2.bias = reinterpret_cast<const float*>(16); d2.residual = ca_x_in;
          d2.y_f32 = S_f; d2.y_bf16 = xb; d2.pool_partial = pool; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C; d2.Cout = C;
          float* sm_ = sv, *sy_ = sv ? sv + size_t(N) * C : nullptr, *sh_ = sv ? sv + size_t(N) * 2 * C : nullptr;
          if (int e = conv_ca_plan_build(&op.conv, &op.cafused, d2, training ? static_cast<float*>(u) : nullptr, Cr,
                                         ca_counters + cai, sm_, sh_, sy_))
            err = e;
          ops.push_back(op);
          ++cai;
        } else {
        ConvDesc d2{};
        d2.x = t; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C; d2.Cout = C; d2.flags = kConvPool; d2.alpha = 1.f;
        if (u32) d2.y_f32 = static_cast<float*>(u); else d2.y_bf16 = u;
        d2.pool_partial = pool;
        conv_op(ops, n->convs[ci++], d2, false);
        Op ca{};
        ca.type = OP_CA;
        ca.ca = n->cas[cai++];
        ca.pool = pool; ca.u = u; ca.x_in = ca_x_in; ca.x_out = S_f; ca.x_out_b = xb;
        ca.pool_compact = pool_compact;
        ca.q_scale = q_of(cai - 1);
        if (sv) { ca.save_mean = sv; ca.save_y = sv + size_t(N) * C; ca.save_hid = sv + size_t(N) * 2 * C; }
        ops.push_back(ca);
        }
        gr.blocks.push_back(br);
        cur_b = xb; cur_f = S_f;
      }
      ConvDesc dg{};
      dg.x = cur_b; dg.residual = gin_f; dg.y_f32 = G_f[g]; dg.y_bf16 = training ? bp.take(px * C * 2) : gb[g & 1];
      dg.N = N; dg.H = H; dg.W = W; dg.Cin = C; dg.Cout = C; dg.alpha = 1.f;
      void* gout_b = dg.y_bf16;
      gr.tail_in_b = cur_b; gr.conv_tail = ci;
      conv_op(ops, n->convs[ci++], dg, false);
      groups.push_back(gr);
      cur_b = gout_b; cur_f = G_f[g];
    }
  } else {
    void* t_shared = training ? nullptr : bp.take(px * C * 2);
    void* xb_shared = training ? nullptr : bp.take(px * C * 2);
    GroupRec gr;
    for (int b = 0; b < n->n_blocks; ++b) {
      void* t = training ? bp.take(px * C * 2) : t_shared;
      void* xb = training ? bp.take(px * C * 2) : xb_shared;
      BlockRec br{cur_b, t, nullptr, nullptr, ci, ci + 1, -1};
      ConvDesc d1{};
      d1.x = cur_b; d1.y_bf16 = t; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C; d1.Cout = C; d1.flags = kConvRelu;
      d1.alpha = 1.f;
      conv_op(ops, n->convs[ci++], d1, false);
      ConvDesc d2{};
      d2.x = t; d2.residual = cur_f; d2.y_f32 = S_f; d2.y_bf16 = xb; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C;
      d2.Cout = C; d2.alpha = n->res_scale;   // conv2(.)*res_scale + x   (common.py:72-73)
      d2.ch_scale = q_of(b);
      conv_op(ops, n->convs[ci++], d2, false);
      gr.blocks.push_back(br);
      cur_b = xb; cur_f = S_f;
    }
    groups.push_back(gr);
  }
  // ---- body tail conv + global skip (architectures.py:173-174 / :238-239): operand for the upsampler only
  const void* body_in_b = use_trunk ? trunk_body_in : cur_b;
  if (!use_trunk) {
    void* body_b = bp.take(px * C * 2);
    ConvDesc d{};
    d.x = cur_b; d.residual = head_f; d.y_bf16 = body_b; d.N = N; d.H = H; d.W = W; d.Cin = C; d.Cout = C;
    d.alpha = 1.f;
    if (han) { d.residual = nullptr; d.y_f32 = han_body_f; }     // HAN: plain conv, fp32 copy for LAM / CSAM
    conv_op(ops, n->convs[ci++], d, false);
    cur_b = body_b;
  }
  // ---- HAN: layer attention over the stacked group outputs -> last_conv; channel-spatial attention of the body
  // output; last(cat) + head skip  (architectures.py:368-392)
  if (han) {
    const int L = n->n_groups + 1;
    if (L != kLamLayers) return set_error(RUMPY_ERR_ARG, "HAN: %d residual groups (the reference fixes 10)", n->n_groups);
    void* lam_out = bp.take(px * C * L * 2);
    float* lam_scratch = static_cast<float*>(bp.take(size_t(lam_workspace_floats(N)) * 4));
    float* out2_f = static_cast<float*>(bp.take(px * C * 4));
    void* cat = bp.take(px * 2 * C * 2);
    void* last_b = bp.take(px * C * 2);
    Op lam{};
    lam.type = OP_LAM;
    lam.stack[0] = han_body_f;                                   // res1 is stacked newest first (:373-378)
    for (int k = 1; k < L; ++k) lam.stack[k] = G_f[L - 1 - k];
    lam.lam_scratch = lam_scratch; lam.lam_out = lam_out;
    ops.push_back(lam);
    for (int k = 0; k < L; ++k) han_stack[k] = lam.stack[k];
    han_lam_scratch = lam_scratch; han_lam_out = lam_out; han_cat = cat;
    ConvDesc d1{};
    d1.x = lam_out; d1.y_f32 = out2_f; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C * L; d1.Cout = C; d1.alpha = 1.f;
    conv_op(ops, n->convs[n->conv_lastconv], d1, false);
    Op cs{};
    cs.type = OP_CSAM;
    cs.csam_x = han_body_f; cs.csam_out2 = out2_f; cs.cat = cat;
    ops.push_back(cs);
    ConvDesc d2{};
    d2.x = cat; d2.residual = head_f; d2.y_bf16 = last_b; d2.N = N; d2.H = H; d2.W = W; d2.Cin = 2 * C; d2.Cout = C;
    d2.alpha = 1.f;
    conv_op(ops, n->convs[n->conv_last], d2, false);
    cur_b = last_b;
  }
  // ---- upsampler: conv C -> C*r*r with the PixelShuffle folded into the store
  int r = 0;
  const int st = upsampler_stages(n->scale, &r);
  int h = H, w = W;
  std::vector<const void*> up_in;
  for (int s = 0; s < st; ++s) {
    void* up = bp.take(size_t(N) * (h * r) * (w * r) * C * 2);
    ConvDesc d{};
    d.x = cur_b; d.y_bf16 = up; d.N = N; d.H = h; d.W = w; d.Cin = C; d.Cout = C * r * r; d.out_r = r; d.alpha = 1.f;
    up_in.push_back(cur_b);
    conv_op(ops, n->convs[ci++], d, false);
    cur_b = up; h *= r; w *= r;
  }
  // ---- thin tail conv -> caller's fp32 NCHW output
  const void* tail_in_b = cur_b;
  {
    ConvDesc d{};
    d.x = cur_b; d.N = N; d.H = h; d.W = w; d.Cin = C; d.Cout = 16; d.cout_real = n->out_feats; d.alpha = 1.f;
    d.out_nchw = reinterpret_cast<float*>(16);  // patched per call
    conv_op(ops, n->convs[ci++], d, false);
    ops.back().writes_output = true;
  }

  // =========================================================================================== backward
  std::vector<Op> bops;
  std::unique_ptr<TrunkBwdPlan> trunk_bwd;
  std::vector<WgradJob> wg_jobs;
  std::vector<WgradReduceJob> wg_rjobs;
  std::vector<ColsumJob> cs_jobs;
  std::vector<int> cs_bias_param;
  std::vector<PartialSumJob> ps_jobs;
  std::vector<int> ps_bias_param;
  struct Site { int conv; const void* g; const void* x; int h, w; float alpha; const float* part; int part_count;
                bool thin = false; /* tail conv: g is zero-padded to 64 channels, bias gradient handled elsewhere */ };
  std::vector<Site> sites;
  WgradJob* jobs_dev = nullptr;
  WgradReduceJob* rjobs_dev = nullptr;
  ColsumJob* cs_dev = nullptr;
  if (training) {
    const int Hh = h, Wh = w;  // final resolution
    // ---- tail conv (thin): dgrad on CUDA cores from the fp32 NCHW upstream gradient; wgrad + bias on CUDA cores
    void* g_hr = bp.take(size_t(N) * Hh * Wh * C * 2);
    float* thin_partial = static_cast<float*>(bp.take(size_t(kThinBlocks) * 37 * C * 4));
    {
      Op op{};
      op.type = OP_TAIL_BWD;
      op.g_hr = g_hr; op.tail_in = tail_in_b; op.thin_partial = thin_partial; op.h = Hh; op.w = Wh;
      // weight gradient on tensor cores: dy zero-padded to a 64-channel bf16 NHWC operand of the batched wgrad kernel
      op.dst_b = bp.take(size_t(N) * Hh * Wh * 64 * 2);
      bops.push_back(op);
      Site ts{n->conv_tail, op.dst_b, tail_in_b, Hh, Wh, 1.f, nullptr, 0};
      ts.thin = true;
      sites.push_back(ts);
    }
    // ---- upsampler stages reversed
    const void* g_cur = g_hr;
    float* GF_body = nullptr;
    void* GB_body = nullptr;
    int hs = Hh, wsz = Wh;
    for (int s = st - 1; s >= 0; --s) {
      hs /= r; wsz /= r;
      const ConvW& cw = n->convs[n->conv_up0 + s];
      ConvDesc d{};
      d.x = g_cur; d.in_r = r; d.N = N; d.H = hs; d.W = wsz; d.Cin = C * r * r; d.Cout = C; d.alpha = 1.f;
      void* g_prev = bp.take(size_t(N) * hs * wsz * C * 2);
      d.y_bf16 = g_prev;
      if (s == 0) { GF_body = static_cast<float*>(bp.take(px * C * 4)); d.y_f32 = GF_body; GB_body = g_prev; }
      conv_op(bops, cw, d, true);
      sites.push_back({n->conv_up0 + s, g_cur, up_in[s], hs, wsz, 1.f, nullptr, 0});
      g_cur = g_prev;
    }
    // ---- body tail conv
    float* P = static_cast<float*>(bp.take(px * C * 4));
    float* Q = static_cast<float*>(bp.take(px * C * 4));
    void* GB_cur = bp.take(px * C * 2);
    float* han_dx[kLamLayers] = {};   // HAN: gradients w.r.t. the stacked maps (body conv output, groups 9 .. 0)
    if (han) {
      // ---- HAN head of the backward (architectures.py:380-392 in reverse): last -> {CSAM, last_conv -> LAM}
      const int L = kLamLayers;
      float* dcat_f = static_cast<float*>(bp.take(px * 2 * C * 4));
      float* dc_f = static_cast<float*>(bp.take(px * C * 4));
      float* sig_f = static_cast<float*>(bp.take(px * C * 4));
      void* d2_b = bp.take(px * C * 2);
      float* glam_f = static_cast<float*>(bp.take(px * C * L * 4));
      for (int k = 0; k < L; ++k) han_dx[k] = static_cast<float*>(bp.take(px * C * 4));
      void* gb2 = bp.take(px * C * 2);
      const int cblocks = han_csam_blocks(N, H, W);
      float* hscr = static_cast<float*>(bp.take(size_t(han_bwd_scratch_floats(N, cblocks)) * 4));
      ConvDesc dl{};
      dl.x = GB_body; dl.N = N; dl.H = H; dl.W = W; dl.Cin = C; dl.Cout = 2 * C; dl.alpha = 1.f; dl.y_f32 = dcat_f;
      conv_op(bops, n->convs[n->conv_last], dl, true);
      sites.push_back({n->conv_last, GB_body, han_cat, H, W, 1.f, nullptr, 0});
      Op cb{};
      cb.type = OP_CSAM_BWD;
      cb.csam_x = han_stack[0]; cb.han_f[0] = dcat_f; cb.han_f[1] = dc_f; cb.han_f[2] = sig_f; cb.han_f[3] = hscr;
      cb.han_v = d2_b; cb.han_dx[0] = han_dx[0]; cb.han_i = cblocks;
      bops.push_back(cb);
      ConvDesc dlc{};
      dlc.x = d2_b; dlc.N = N; dlc.H = H; dlc.W = W; dlc.Cin = C; dlc.Cout = C * L; dlc.alpha = 1.f; dlc.y_f32 = glam_f;
      conv_op(bops, n->convs[n->conv_lastconv], dlc, true);
      sites.push_back({n->conv_lastconv, d2_b, han_lam_out, H, W, 1.f, nullptr, 0});
      Op lb{};
      lb.type = OP_LAM_BWD;
      for (int k = 0; k < L; ++k) { lb.stack[k] = han_stack[k]; lb.han_dx[k] = han_dx[k]; }
      lb.han_f[0] = glam_f; lb.han_f[3] = hscr; lb.lam_scratch = han_lam_scratch; lb.han_v = gb2; lb.han_i = cblocks;
      bops.push_back(lb);
      Op pg{};
      pg.type = OP_HAN_PG;
      pg.han_f[3] = hscr; pg.han_i = cblocks;
      bops.push_back(pg);
      // body conv (no skip in HAN): its input gradient + the layer-attention gradient of group 9's output
      ConvDesc d{};
      d.x = gb2; d.N = N; d.H = H; d.W = W; d.Cin = C; d.Cout = C; d.alpha = 1.f; d.residual = han_dx[1]; d.y_f32 = P;
      d.y_bf16 = GB_cur;
      conv_op(bops, n->convs[n->conv_body], d, true);
      sites.push_back({n->conv_body, gb2, body_in_b, H, W, 1.f, nullptr, 0});
    } else {
      ConvDesc d{};
      d.x = GB_body; d.N = N; d.H = H; d.W = W; d.Cin = C; d.Cout = C; d.alpha = 1.f; d.y_f32 = P; d.y_bf16 = GB_cur;
      // Q-EDSR: the bf16 gradient operand of block b's conv2 dgrad / wgrad is g * q_b; the fp32 skip stream P stays g
      if (n->qrcan && n->arch == 1) d.bf16_scale = q_of(n->n_blocks - 1);
      conv_op(bops, n->convs[n->conv_body], d, true);
      sites.push_back({n->conv_body, GB_body, body_in_b, H, W, 1.f, nullptr, 0});
    }
    const bool use_trunk_bwd = opt().use_trunk_bwd && use_trunk && n->arch == 0;
    if (use_trunk_bwd) {
      // ---- the whole backward body as ONE persistent dataflow kernel (trunk_bwd.cuh); the gradient stream Q lives
      // in tensor memory, P (gradient w.r.t. the group input) is updated in place once per group
      const int T_tiles = N * tiles;
      const int n_ca = n->n_groups * n->n_blocks;
      const int n_lay = n->n_groups * (1 + 3 * n->n_blocks);
      const size_t per = size_t(2) * C * Cr + C + Cr;
      void* dev = bp.take(trunk_bwd_device_bytes(N, H, W, n_lay, n_lay + 2, n_lay + 2, n_ca));
      float* pg_all = static_cast<float*>(bp.take(size_t(n_ca) * N * per * 4));
      float* dq_all = n->qrcan ? static_cast<float*>(bp.take(size_t(n_ca) * N * C * 4)) : nullptr;
      QGradJobHost* qgj = n->qrcan ? static_cast<QGradJobHost*>(bp.take(size_t(n_ca) * sizeof(QGradJobHost))) : nullptr;
      if (build) { n->q_dq = dq_all; n->qg_jobs_dev = qgj; n->qg_jobs_uploaded.clear(); }
      if (build) trunk_bwd.reset(new TrunkBwdPlan());
      auto in_of = [&](const void* p) {
        for (size_t i = 0; i < trunk_bwd->in_bufs.size(); ++i) if (trunk_bwd->in_bufs[i] == p) return int(i);
        trunk_bwd->in_bufs.push_back(p);
        return int(trunk_bwd->in_bufs.size()) - 1;
      };
      auto out_of = [&](void* p) {
        for (size_t i = 0; i < trunk_bwd->out_bufs.size(); ++i) if (trunk_bwd->out_bufs[i] == p) return int(i);
        trunk_bwd->out_bufs.push_back(p);
        return int(trunk_bwd->out_bufs.size()) - 1;
      };
      int L = 0, ca_ord = 0;
      const void* gb = GB_cur;   // bf16 gradient w.r.t. the current group's output ...
      int gb_epoch = 0;          // ... complete at this tile epoch (0: written before the kernel)
      auto push = [&](const TrunkBwdLayer& l, int w1 = -1, int w2 = -1) {
        if (build) { trunk_bwd->layers.push_back(l); trunk_bwd->lparams.push_back(TrunkBwdLayerParams{w1, w2}); }
        ++L;
      };
      for (int g = n->n_groups - 1; g >= 0; --g) {
        const GroupRec& gr = groups[g];
        {
          TrunkBwdLayer l{};
          l.kind = kBwdFresh; l.out_map = -1; l.ca_slot = -1;
          l.w_idx = gr.conv_tail - 1; l.wait_epoch = gb_epoch;
          if (build) l.in_map = in_of(gb);
          push(l);
          sites.push_back({gr.conv_tail, gb, gr.tail_in_b, H, W, 1.f, nullptr, 0});
        }
        for (int b = n->n_blocks - 1; b >= 0; --b) {
          const BlockRec& br = gr.blocks[b];
          void* du = bp.take(px * C * 2);
          void* dt = bp.take(px * C * 2);
          float* du_cs = static_cast<float*>(bp.take(size_t(T_tiles) * C * 4));
          float* dt_pool = static_cast<float*>(bp.take(size_t(T_tiles) * C * 4));
          const int L_ca = L;
          {
            TrunkBwdLayer l{};
            l.kind = kBwdCA; l.in_map = -1; l.w_idx = -1; l.ca_slot = ca_ord;
            l.aux = static_cast<const __nv_bfloat16*>(br.u); l.colsum = du_cs;
            l.save_mean = br.sv; l.save_y = br.sv + size_t(N) * C; l.save_hid = br.sv + size_t(N) * 2 * C;
            l.pg = pg_all + size_t(ca_ord) * N * per;
            if (n->qrcan) {
              l.q_scale = q_of(br.ca);       // q-layer output, or the 'modulate' attributes (no gradient of their own)
              if (n->qs[br.ca].w1 >= 0) l.dq = dq_all + size_t(br.ca) * N * C;
            }
            if (build) {
              l.out_map = out_of(du);
              const CAW& cw = n->cas[br.ca];
              trunk_bwd->pg_binds.push_back(TrunkBwdPgBind{L, cw.w1, cw.b1, cw.w2, cw.b2});
              push(l, cw.w1, cw.w2);
            } else {
              push(l);
            }
            ++ca_ord;
          }
          const int L_mask = L;
          {
            TrunkBwdLayer l{};
            l.kind = kBwdMask; l.ca_slot = -1; l.w_idx = br.conv2 - 1; l.wait_epoch = L_ca + 1;
            l.aux = static_cast<const __nv_bfloat16*>(br.t); l.colsum = dt_pool;
            if (build) { l.in_map = in_of(du); l.out_map = out_of(dt); }
            push(l);
          }
          {
            TrunkBwdLayer l{};
            l.kind = kBwdAcc; l.ca_slot = -1; l.out_map = -1; l.w_idx = br.conv1 - 1; l.wait_epoch = L_mask + 1;
            void* GB_new = b == 0 ? bp.take(px * C * 2) : nullptr;
            if (b == 0) { l.res_f32 = P; l.out_f32 = P; if (han && g > 0) l.add_f32 = han_dx[kLamLayers - g]; }
            if (build) { l.in_map = in_of(dt); if (b == 0) l.out_map = out_of(GB_new); }
            if (b == 0) { gb = GB_new; gb_epoch = L + 1; }
            push(l);
          }
          sites.push_back({br.conv2, du, br.t, H, W, 1.f, du_cs, T_tiles});
          sites.push_back({br.conv1, dt, br.in_b, H, W, 1.f, dt_pool, T_tiles});
        }
      }
      if (build) {
        if (int e = trunk_bwd_plan_finish(trunk_bwd.get(), N, H, W, Cr, pk + n->convs[1].off_dgrad, n->conv_body, dev))
          err = e;
        Op op{};
        op.type = OP_TRUNK_BWD;
        bops.push_back(op);
      }
    } else if (n->arch == 0) {
      float* dq_all = n->qrcan ? static_cast<float*>(bp.take(size_t(n->cas.size()) * N * C * 4)) : nullptr;
      QGradJobHost* qgj = n->qrcan ? static_cast<QGradJobHost*>(bp.take(n->cas.size() * sizeof(QGradJobHost))) : nullptr;
      if (build && n->qrcan) { n->q_dq = dq_all; n->qg_jobs_dev = qgj; n->qg_jobs_uploaded.clear(); }
      for (int g = n->n_groups - 1; g >= 0; --g) {
        const GroupRec& gr = groups[g];
        {  // group tail conv: grad wrt the last block's output, fp32 only
          ConvDesc d{};
          d.x = GB_cur; d.N = N; d.H = H; d.W = W; d.Cin = C; d.Cout = C; d.alpha = 1.f; d.y_f32 = Q;
          conv_op(bops, n->convs[gr.conv_tail], d, true);
          sites.push_back({gr.conv_tail, GB_cur, gr.tail_in_b, H, W, 1.f, nullptr, 0});
        }
        for (int b = n->n_blocks - 1; b >= 0; --b) {
          const BlockRec& br = gr.blocks[b];
          float* s_partial = static_cast<float*>(bp.take(size_t(N) * kCaBwdChunks * C * 4));
          void* du = bp.take(px * C * 2);
          void* dt = bp.take(px * C * 2);
          const size_t vec = size_t(H) * W * (C / 4);
          int ca_chunks = int((vec + 4095) / 4096);     // >= 16 vectors per thread: pure streaming kernel
          { const int cap = (148 * 4 + N - 1) / N; if (ca_chunks > cap) ca_chunks = cap; if (ca_chunks < 1) ca_chunks = 1; }
          float* du_cs = static_cast<float*>(bp.take(size_t(N) * ca_chunks * C * 4));
          float* dt_pool = static_cast<float*>(bp.take(size_t(N) * pool_rows * C * 4));
          Op cb{};
          cb.type = OP_CA_BWD;
          cb.ca = n->cas[br.ca];
          cb.a = Q; cb.u = br.u; cb.s_partial = s_partial; cb.du = du; cb.du_colsum = du_cs; cb.ca_chunks = ca_chunks;
          cb.save_mean = br.sv; cb.save_y = br.sv + size_t(N) * C; cb.save_hid = br.sv + size_t(N) * 2 * C;
          if (n->qrcan) { cb.q_scale = q_of(br.ca); cb.dst_f = n->qs[br.ca].w1 >= 0 ? dq_all + size_t(br.ca) * N * C : nullptr; }
          bops.push_back(cb);
          ConvDesc d2{};
          d2.x = du; d2.mask = br.t; d2.y_bf16 = dt; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C; d2.Cout = C; d2.alpha = 1.f;
          d2.flags = kConvPool; d2.pool_partial = dt_pool;   // per-tile column sums of dt = conv1's bias gradient
          conv_op(bops, n->convs[br.conv2], d2, true);
          ConvDesc d1{};
          d1.x = dt; d1.residual = Q; d1.y_f32 = Q; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C; d1.Cout = C; d1.alpha = 1.f;
          conv_op(bops, n->convs[br.conv1], d1, true);
          sites.push_back({br.conv2, du, br.t, H, W, 1.f, du_cs, N * ca_chunks});
          sites.push_back({br.conv1, dt, br.in_b, H, W, 1.f, dt_pool, N * pool_rows});
        }
        void* GB_new = bp.take(px * C * 2);
        Op add{};
        add.type = OP_ADD;
        add.a = P; add.b = Q; add.dst_f = P; add.dst_b = GB_new; add.n4 = px * C / 4;
        bops.push_back(add);
        if (han && g > 0) {   // + the layer-attention gradient of group g-1's output (stack index L - g)
          Op add2 = add;
          add2.b = han_dx[kLamLayers - g];
          bops.push_back(add2);
        }
        GB_cur = GB_new;
      }
      if (n->qrcan) { Op qg{}; qg.type = OP_QGRAD; qg.ca_chunks = 0; bops.push_back(qg); }   // dq [N][C] per RCAB
    } else {
      const GroupRec& gr = groups[0];
      const bool qed = n->qrcan;
      float* dq_all = qed ? static_cast<float*>(bp.take(size_t(n->n_blocks) * N * kDqSlices * C * 4)) : nullptr;
      QGradJobHost* qgj = qed ? static_cast<QGradJobHost*>(bp.take(size_t(n->n_blocks) * sizeof(QGradJobHost))) : nullptr;
      if (build && qed) { n->q_dq = dq_all; n->qg_jobs_dev = qgj; n->qg_jobs_uploaded.clear(); }
      for (int b = n->n_blocks - 1; b >= 0; --b) {
        const BlockRec& br = gr.blocks[b];
        if (qed && n->qs[b].w1 >= 0) {
          // q * dq[n][c] = sum_hw g * (out - x), taken while P still holds the gradient w.r.t. this block's output
          Op dq{};
          dq.type = OP_DQ;
          dq.a = P; dq.u = b + 1 < n->n_blocks ? gr.blocks[b + 1].in_b : body_in_b; dq.tail_in = br.in_b;
          dq.dst_f = dq_all + size_t(b) * N * kDqSlices * C;
          bops.push_back(dq);
        }
        void* dt = bp.take(px * C * 2);
        void* GB_new = bp.take(px * C * 2);
        ConvDesc d2{};
        d2.x = GB_cur; d2.mask = br.t; d2.y_bf16 = dt; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C; d2.Cout = C;
        d2.alpha = n->res_scale;
        float* dt_pool = static_cast<float*>(bp.take(size_t(N) * pool_rows * C * 4));
        d2.flags = kConvPool; d2.pool_partial = dt_pool;
        conv_op(bops, n->convs[br.conv2], d2, true);
        ConvDesc d1{};
        d1.x = dt; d1.residual = P; d1.y_f32 = P; d1.y_bf16 = GB_new; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C;
        d1.Cout = C; d1.alpha = 1.f;
        if (qed && b > 0) d1.bf16_scale = q_of(b - 1);
        conv_op(bops, n->convs[br.conv1], d1, true);
        sites.push_back({br.conv2, GB_cur, br.t, H, W, n->res_scale, nullptr, 0});
        sites.push_back({br.conv1, dt, br.in_b, H, W, 1.f, dt_pool, N * pool_rows});
        GB_cur = GB_new;
      }
      if (qed) { Op qg{}; qg.type = OP_QGRAD; qg.ca_chunks = kDqSlices; bops.push_back(qg); }   // q*dq slices
    }
    {  // head conv: weight / bias gradient only (no dX), upstream = trunk gradient + global skip
      Op op{};
      op.type = OP_HEAD_WGRAD;
      op.a = P; op.b = GF_body; op.thin_partial = thin_partial;
      bops.push_back(op);
    }
    // ---- batched tensor-core wgrad + bias gradients over every recorded site, last parameters first
    std::stable_sort(sites.begin(), sites.end(), [&](const Site& a, const Site& b) {
      return n->convs[a.conv].w_idx > n->convs[b.conv].w_idx;
    });
    size_t njobs = 0, nrjobs = 0, cs_floats = 0;
    for (const Site& s : sites) {
      const ConvW& cw = n->convs[s.conv];
      const int mt = N * ((s.h + kTileH - 1) / kTileH) * ((s.w + kTileW - 1) / kTileW);
      const int blocks = (s.thin ? 1 : cw.cout / 64) * (cw.cin / 64);
      njobs += size_t(blocks) * wgrad_splits(mt);
      nrjobs += blocks;
      if (!s.part && !s.thin) cs_floats += size_t(kColsumSlices) * cw.r * cw.cout;
    }
    jobs_dev = static_cast<WgradJob*>(bp.take(njobs * sizeof(WgradJob)));
    rjobs_dev = static_cast<WgradReduceJob*>(bp.take(nrjobs * sizeof(WgradReduceJob)));
    cs_dev = static_cast<ColsumJob*>(bp.take(sites.size() * sizeof(ColsumJob)));
    PartialSumJob* ps_dev = static_cast<PartialSumJob*>(bp.take(sites.size() * sizeof(PartialSumJob)));
    if (build) n->ps_jobs_dev = ps_dev;
    float* partials = static_cast<float*>(bp.take(njobs * 9 * 64 * 64 * sizeof(float)));
    float* cs_partials = static_cast<float*>(bp.take(cs_floats * sizeof(float)));
    float* pg_scratch = static_cast<float*>(bp.take(size_t(N) * (2 * C * Cr + C + Cr) * sizeof(float)));
    int* pg_counter = static_cast<int*>(bp.take(size_t(N + 1) * sizeof(int)));
    float* ca_coef = static_cast<float*>(bp.take(size_t(N) * C * sizeof(float)));
    if (build) { n->pg_scratch = pg_scratch; n->pg_counter = pg_counter; n->ca_coef = ca_coef; }
    if (build) {
      size_t job_cursor = 0, cs_cursor = 0;
      int last_site_param = 0;
      std::vector<int> chunk_job_end, chunk_rjob_end, chunk_first_param;
      const size_t per_chunk = (njobs + opt().wgrad_chunks - 1) / size_t(opt().wgrad_chunks);
      for (const Site& s : sites) {
        const ConvW& cw = n->convs[s.conv];
        if (!wg_jobs.empty() && wg_jobs.size() >= per_chunk * (chunk_job_end.size() + 1)) {
          chunk_job_end.push_back(int(wg_jobs.size()));     // close the chunk before this site: everything from the
          chunk_rjob_end.push_back(int(wg_rjobs.size()));   // previous site's weight index upwards is then final
          chunk_first_param.push_back(last_site_param);
        }
        last_site_param = cw.w_idx;
        const int cout_eff = s.thin ? 64 : cw.cout;
        const int rr = cw.r * cw.r, cout_sub = cout_eff / rr, chunks_per_q = cout_sub / 64;
        const int tiles_x = (s.w + kTileW - 1) / kTileW, tiles_y = (s.h + kTileH - 1) / kTileH;
        const int mt = N * tiles_x * tiles_y;
        const int splits = wgrad_splits(mt);
        std::vector<CUtensorMap> gmaps(rr);
        for (int q = 0; q < rr; ++q)
          if (int e = make_map_nhwc_sub(&gmaps[q], false, s.g, cout_sub, s.w, s.h, N, cw.r, q, kABoxH)) err = e;
        CUtensorMap xmap;
        if (int e = make_map_nhwc_sub(&xmap, false, s.x, cw.cin, s.w, s.h, N, 1, 0, kTileH)) err = e;
        for (int cb = 0; cb < cout_eff / 64; ++cb) {
          for (int ib = 0; ib < cw.cin / 64; ++ib) {
            float* pbase = partials + job_cursor * 9 * 64 * 64;
            for (int k = 0; k < splits; ++k) {
              WgradJob j{};
              j.g = gmaps[cb / chunks_per_q]; j.x = xmap;
              j.gc0 = (cb % chunks_per_q) * 64; j.xc0 = ib * 64;
              j.tile_begin = int((long long)mt * k / splits); j.tile_end = int((long long)mt * (k + 1) / splits);
              j.tiles_x = tiles_x; j.tiles_y = tiles_y;
              j.out = pbase + size_t(k) * 9 * 64 * 64;
              wg_jobs.push_back(j);
            }
            job_cursor += splits;
            WgradReduceJob rj{};
            rj.partial = pbase; rj.dw = nullptr; rj.splits = splits; rj.cout = cw.cout; rj.cin = cw.cin;
            rj.co0 = cb * 64; rj.ci0 = ib * 64; rj.r = cw.r; rj.alpha = s.alpha;
            rj.accumulate = cw.w_idx;  // carries the param index until net_backward binds the gradient pointer
            wg_rjobs.push_back(rj);
          }
        }
        if (s.thin) continue;   // bias gradient of the tail conv: plane sums of dy (OP_TAIL_BWD)
        if (s.part) {
          PartialSumJob pj{};
          pj.partial = s.part; pj.db = nullptr; pj.count = s.part_count; pj.C = cw.cout; pj.alpha = s.alpha;
          ps_jobs.push_back(pj);
          ps_bias_param.push_back(cw.b_idx);
        } else {
          ColsumJob cj{};
          cj.g = static_cast<const __nv_bfloat16*>(s.g); cj.db = nullptr; cj.partial = cs_partials + cs_cursor;
          cj.outer = N * s.h; cj.r = cw.r; cj.inner = s.w; cj.C = cout_sub; cj.alpha = s.alpha;
          cs_cursor += size_t(kColsumSlices) * cw.r * cw.cout;
          cs_jobs.push_back(cj);
          cs_bias_param.push_back(cw.b_idx);
        }
      }
      chunk_job_end.push_back(int(wg_jobs.size()));
      chunk_rjob_end.push_back(int(wg_rjobs.size()));
      chunk_first_param.push_back(0);
      n->wg_chunk_job_end.swap(chunk_job_end);
      n->wg_chunk_rjob_end.swap(chunk_rjob_end);
      n->wg_chunk_first_param.swap(chunk_first_param);
    }
  }
  *bytes_out = bp.off;
  if (err) return err;
  if (build) {
    n->ops.swap(ops);
    n->bops.swap(bops);
    n->trunk = std::move(trunk);
    n->trunk_bwd = std::move(trunk_bwd);
    n->plan_u_f32 = use_trunk ? 0 : int(u32);
    n->wg_jobs.swap(wg_jobs);
    n->wg_rjobs.swap(wg_rjobs);
    n->cs_jobs.swap(cs_jobs);
    n->cs_bias_param.swap(cs_bias_param);
    n->ps_jobs.swap(ps_jobs);
    n->ps_bias_param.swap(ps_bias_param);
    n->wg_jobs_dev = jobs_dev; n->wg_rjobs_dev = rjobs_dev; n->cs_jobs_dev = cs_dev;
    n->plan_grads.clear();
    n->jobs_uploaded = false;
  }
  return RUMPY_OK;
}

static int ensure_plan(Net* n, const void* packed, void* workspace, int N, int H, int W, int training) {
  if (n->plan_packed != packed || n->plan_ws != workspace || n->pN != N || n->pH != H || n->pW != W ||
      n->p_training != training || n->p_trunk != int(opt().plan_sig())) {
    size_t bytes = 0;
    n->plan_packed = nullptr;
    if (int e = build_plan(n, packed, workspace, N, H, W, training, &bytes, true)) return e;
    n->plan_packed = packed; n->plan_ws = workspace; n->pN = N; n->pH = H; n->pW = W; n->p_training = training;
    n->p_trunk = int(opt().plan_sig());
  }
  return RUMPY_OK;
}

static int launch_conv_op(Op& op, const float* const* params, float* y_nchw, cudaStream_t stream) {
  if (op.bias_param >= 0) op.conv.args.bias = params[op.bias_param];
  if (op.writes_output) op.conv.args.out_nchw = y_nchw;
  return conv_plan_launch(op.conv, stream);
}

}  // namespace rb

using namespace rb;

extern "C" {

int rumpy_net_create(void** out, int arch, int n_feats, int n_groups, int n_blocks, int reduction, int scale,
                     float res_scale, int in_feats, int out_feats, int u_f32) {
  if (!out) return set_error(RUMPY_ERR_ARG, "net_create: null out");
  if (arch != 0 && arch != 1 && arch != 2) return set_error(RUMPY_ERR_ARG, "net_create: arch %d", arch);
  const bool han = arch == 2;
  if (han) {
    if (n_feats != 64 || n_groups != kLamLayers - 1)
      return set_error(RUMPY_ERR_ARG, "net_create: HAN needs n_feats=64 and %d residual groups", kLamLayers - 1);
    arch = 0;
  }
  if (n_feats % 64 != 0 || n_feats <= 0 || n_feats > 256)
    return set_error(RUMPY_ERR_ARG, "net_create: n_feats=%d must be 64, 128, 192 or 256", n_feats);
  if (in_feats < 1 || in_feats > 4 || out_feats < 1 || out_feats > 4)
    return set_error(RUMPY_ERR_ARG, "net_create: in_feats=%d out_feats=%d (1..4 supported)", in_feats, out_feats);
  if (arch == 0 && (reduction < 1 || n_feats % reduction != 0 || n_feats / reduction > 16))
    return set_error(RUMPY_ERR_ARG, "net_create: reduction=%d (n_feats/reduction must be 1..16)", reduction);
  Net* n = new Net();
  n->arch = arch; n->C = n_feats; n->n_groups = n_groups; n->n_blocks = n_blocks; n->reduction = reduction;
  n->scale = scale; n->res_scale = res_scale; n->in_feats = in_feats; n->out_feats = out_feats; n->u_f32 = u_f32;
  n->han = han;
  if (int e = net_init(n)) { delete n; return e; }
  *out = n;
  return RUMPY_OK;
}

/* Meta-attention networks (reference SISR/models/attention_manipulators/architectures.py):
 *   arch 0  Q-RCAN (:313-462), QCALayer style 'standard' or 'modulate', optional q-nodes in the RCABs;
 *   arch 1  Q-EDSR (:496-556), ParamResBlocks with optional q-layers (n_groups ignored).
 * block_has_q[i] != 0: block i (RCAB g*n_blocks+b / ResBlock b) owns a 2-layer ParaCALayer (q_layer.py:5-45:
 * num_metadata -> q_hidden -> n_feats, ReLU in between iff q_relu).  Parameter order = the reference's state_dict. */
int rumpy_net_create_q(void** out, int arch, int n_feats, int n_groups, int n_blocks, int reduction, int scale,
                       float res_scale, int in_feats, int out_feats, int num_metadata, int q_hidden,
                       const unsigned char* block_has_q, int modulate, int q_relu) {
  if (!out || !block_has_q) return set_error(RUMPY_ERR_ARG, "net_create_q: null pointer");
  if (arch != 0 && arch != 1 && arch != 2) return set_error(RUMPY_ERR_ARG, "net_create_q: arch %d", arch);
  const bool han = arch == 2;
  if (han) {
    if (n_groups != kLamLayers - 1)
      return set_error(RUMPY_ERR_ARG, "net_create_q: Q-HAN needs %d residual groups", kLamLayers - 1);
    arch = 0;
  }
  if (arch == 0 && n_feats != 64) return set_error(RUMPY_ERR_ARG, "net_create_q: Q-RCAN n_feats=%d (64 supported)", n_feats);
  if (n_feats % 64 != 0 || n_feats <= 0 || n_feats > 256)
    return set_error(RUMPY_ERR_ARG, "net_create_q: n_feats=%d must be 64, 128, 192 or 256", n_feats);
  if (in_feats < 1 || in_feats > 4 || out_feats < 1 || out_feats > 4)
    return set_error(RUMPY_ERR_ARG, "net_create_q: in_feats=%d out_feats=%d (1..4 supported)", in_feats, out_feats);
  if (arch == 0 && (reduction < 1 || n_feats % reduction != 0 || n_feats / reduction > 16))
    return set_error(RUMPY_ERR_ARG, "net_create_q: reduction=%d", reduction);
  if (num_metadata < 1 || num_metadata > 1024 || q_hidden < 1 || q_hidden > 1024)
    return set_error(RUMPY_ERR_ARG, "net_create_q: num_metadata=%d q_hidden=%d", num_metadata, q_hidden);
  if (arch == 1 && modulate) return set_error(RUMPY_ERR_ARG, "net_create_q: 'modulate' is a Q-RCAN style");
  Net* n = new Net();
  n->arch = arch; n->qrcan = true; n->han = han;
  n->C = n_feats; n->n_groups = arch == 0 ? n_groups : 1; n->n_blocks = n_blocks; n->reduction = reduction;
  n->scale = scale; n->res_scale = arch == 0 ? 1.f : res_scale; n->in_feats = in_feats; n->out_feats = out_feats;
  n->u_f32 = 1;
  n->num_meta = num_metadata; n->q_hidden = q_hidden; n->modulate = modulate; n->q_relu = q_relu;
  n->qs.resize(size_t(n->n_groups) * n_blocks);
  for (size_t i = 0; i < n->qs.size(); ++i) n->qs[i] = CAW{block_has_q[i] ? 0 : -1, -1, -1, -1};
  if (int e = net_init(n)) { delete n; return e; }
  *out = n;
  return RUMPY_OK;
}

/* metadata: device fp32 [N][num_metadata] (what QRCAN.forward receives as `metadata`, squeezed); read by the next
 * rumpy_net_forward of a Q-RCAN handle on its stream.  M = num_metadata; style 'modulate' without q-layers
 * also takes M = n_feats (the handler's scale_qpi vector, reference attention_manipulators/handlers.py:65-73). */
int rumpy_net_set_metadata(void* net, const float* metadata, int N, int M) {
  Net* n = static_cast<Net*>(net);
  if (!n || !n->qrcan) return set_error(RUMPY_ERR_ARG, "net_set_metadata: not a meta-attention (rumpy_net_create_q) handle");
  bool any_q = false;
  for (const CAW& q : n->qs) any_q |= q.w1 >= 0;
  if (any_q && M != n->num_meta)
    return set_error(RUMPY_ERR_ARG, "net_set_metadata: %d attributes per image, the q-layers take %d", M, n->num_meta);
  if (n->modulate && M != 1 && M != n->C)
    return set_error(RUMPY_ERR_ARG, "net_set_metadata: style 'modulate' takes 1 or %d attributes per image, got %d", n->C, M);
  if (M < 1 || M > 1024) return set_error(RUMPY_ERR_ARG, "net_set_metadata: M=%d", M);
  n->meta_dev = metadata;
  n->meta_n = N;
  n->meta_m = M;
  return RUMPY_OK;
}


/* Per-handle execution options (no process-global switches).  Takes effect for plans built afterwards: the cached
 * plan is rebuilt by the next forward when an option it depends on changed. */
int rumpy_net_set_option(void* net, const char* name, long long value) {
  if (!net || !name) return set_error(RUMPY_ERR_ARG, "net_set_option: null");
  Options& o = static_cast<Net*>(net)->opt;
  const std::string k(name);
  const int v = int(value);
  if (k == "trunk") o.use_trunk = v != 0;
  else if (k == "cluster") o.use_cluster = v != 0;
  else if (k == "cluster_groups") o.cluster_groups = v == 4 ? 4 : 2;
  else if (k == "cluster_split") o.cluster_split = v != 0;
  else if (k == "cluster_dbg") o.cluster_dbg = v;
  else if (k == "infer_u_bf16") o.infer_u_bf16 = v != 0;
  else if (k == "band") o.use_band = v != 0;
  else if (k == "trunk_bwd") o.use_trunk_bwd = v != 0;
  else if (k == "fused_ca") o.use_fused_ca = v != 0;
  else if (k == "wgrad_chunks") o.wgrad_chunks = v < 1 ? 1 : (v > 8 ? 8 : v);
  else if (k == "wgrad_tiles_per_split") o.wgrad_tiles_per_split = v < 1 ? 1 : v;
  else if (k == "pdl") o.use_pdl = v != 0;
  else if (k == "conv_dbg") o.conv_dbg = v;
  else if (k == "trunk_sync_mode") o.trunk_sync_mode = v;
  else return set_error(RUMPY_ERR_ARG, "net_set_option: unknown option '%s'", name);
  return RUMPY_OK;
}

long long rumpy_net_get_option(void* net, const char* name) {
  if (!net || !name) return -1;
  const Options& o = static_cast<Net*>(net)->opt;
  const std::string k(name);
  if (k == "trunk") return o.use_trunk;
  if (k == "cluster") return o.use_cluster;
  if (k == "cluster_groups") return o.cluster_groups;
  if (k == "cluster_split") return o.cluster_split;
  if (k == "infer_u_bf16") return o.infer_u_bf16;
  if (k == "band") return o.use_band;
  if (k == "trunk_bwd") return o.use_trunk_bwd;
  if (k == "fused_ca") return o.use_fused_ca;
  if (k == "wgrad_chunks") return o.wgrad_chunks;
  if (k == "wgrad_tiles_per_split") return o.wgrad_tiles_per_split;
  if (k == "pdl") return o.use_pdl;
  if (k == "trunk_sync_mode") return o.trunk_sync_mode;
  if (k == "cluster_dbg") return o.cluster_dbg;
  if (k == "conv_dbg") return o.conv_dbg;
  return -1;
}

/* Measurement hook: CUDA events (cudaEvent_t, caller-owned) recorded on the launching stream right before / after
 * the trunk kernel of every forward of THIS handle; NULL, NULL switches it off. */
int rumpy_net_set_trunk_events(void* net, void* ev_start, void* ev_stop) {
  if (!net) return set_error(RUMPY_ERR_ARG, "net_set_trunk_events: null");
  Options& o = static_cast<Net*>(net)->opt;
  o.trunk_ev0 = static_cast<cudaEvent_t>(ev_start);
  o.trunk_ev1 = static_cast<cudaEvent_t>(ev_stop);
  return RUMPY_OK;
}

/* Diagnostics: device buffer (caller-owned, int64) that the kernels of THIS handle fill with clock64 stamps; `layers`
 * > 0 also asks the trunk kernels for a per-layer timeline of that many layers.  NULL / 0 switches it off. */
int rumpy_net_set_timeline(void* net, void* buf, int layers) {
  if (!net) return set_error(RUMPY_ERR_ARG, "net_set_timeline: null");
  Options& o = static_cast<Net*>(net)->opt;
  o.timeline = static_cast<long long*>(buf);
  o.trunk_dbg_layers = buf ? layers : 0;
  return RUMPY_OK;
}

int rumpy_net_destroy(void* net) {
  delete static_cast<Net*>(net);
  return RUMPY_OK;
}

/* kernels enqueued by one forward of the cached plan (0 before the first forward) */
int rumpy_net_num_launches(void* net) { return net ? int(static_cast<Net*>(net)->ops.size()) : -1; }

/* how the 64-channel body of the cached plan runs: 0 one kernel per layer, 1 persistent dataflow kernel
 * (trunk_pipe.cuh), 2 one thread-block cluster per image (trunk_cluster.cuh) */
int rumpy_net_trunk_mode(void* net) {
  if (!net) return -1;
  Net* n = static_cast<Net*>(net);
  return n->trunk ? (n->trunk->band ? 3 : (n->trunk->cluster ? 2 : 1)) : 0;
}

/* kernels enqueued by one backward of the cached plan (0 when the plan is inference-only) */
int rumpy_net_num_launches_backward(void* net) {
  if (!net) return -1;
  Net* n = static_cast<Net*>(net);
  if (n->bops.empty()) return 0;
  int chunks = 0, j0 = 0;                     // wgrad chunks that actually hold jobs: wgrad kernel + its reduce each
  for (int j1 : n->wg_chunk_job_end) { chunks += j1 > j0; j0 = j1; }
  int c = 2 * chunks + (n->cs_jobs.empty() ? 0 : 2) + (n->ps_jobs.empty() ? 0 : 1);   // + colsum, its reduce, partial sums
  for (const Op& op : n->bops) {
    switch (op.type) {
      case OP_TAIL_BWD: c += 4; break;        // dgrad, padded dy operand, plane sums (2)
      case OP_CA_BWD: c += 2; break;
      case OP_TRUNK_BWD: c += n->qrcan ? 3 : 2; break;   // dataflow kernel + CA parameter-gradient finalize [+ q-layer grads]
      case OP_HEAD_WGRAD: c += 2; break;
      default: c += 1;
    }
  }
  return c;
}

int rumpy_net_num_params(void* net) { return net ? static_cast<Net*>(net)->n_params : -1; }

long long rumpy_net_packed_bytes(void* net, int training) {
  if (!net) return -1;
  Net* n = static_cast<Net*>(net);
  return (long long)(training ? n->packed_bytes_train : n->packed_bytes);
}

long long rumpy_net_workspace_bytes(void* net, int N, int H, int W, int training) {
  if (!net) return -1;
  size_t bytes = 0;
  OptScope scope(&static_cast<Net*>(net)->opt);
  if (build_plan(static_cast<Net*>(net), nullptr, nullptr, N, H, W, training, &bytes, false)) return -1;
  return (long long)bytes;
}

// fp32 OIHW parameters (device pointers, state_dict order) -> packed bf16 operands.  Call after every
// optimiser step (weights changed) and before the first forward.  training != 0 also packs the dgrad operands.
int rumpy_net_pack(void* net_, const float* const* params, void* packed, int training, void* stream) {
  Net* n = static_cast<Net*>(net_);
  if (!n || !params || !packed) return set_error(RUMPY_ERR_ARG, "net_pack: null");
  OptScope scope(&n->opt);
  if (int e = device_info(nullptr)) return e;
  char* pk = static_cast<char*>(packed);
  const size_t total = training ? n->packed_bytes_train : n->packed_bytes;
  PackJobHost* jobs_dev = reinterpret_cast<PackJobHost*>(pk + total - n->pack_jobs_bytes);
  // the job list (pointers into params / packed) is rebuilt and re-uploaded only when those pointers change
  if (n->pack_sig_packed != packed || n->pack_sig_p0 != params[0] || n->pack_sig_training != training ||
      n->pack_jobs.empty()) {
    n->pack_jobs.clear();
    for (size_t i = 1; i < n->convs.size(); ++i) {  // conv 0 is the fp32 head
      const ConvW& c = n->convs[i];
      PackJobHost j{};
      j.w = params[c.w_idx]; j.p = pk + c.off_fwd; j.cout = c.cout; j.cin = c.cin; j.rows_padded = c.rows_padded;
      j.r = c.r; j.dgrad = 0;
      if (c.off_bias != SIZE_MAX) { j.b = params[c.b_idx]; j.bp = reinterpret_cast<float*>(pk + c.off_bias); }
      n->pack_jobs.push_back(j);
      if (training && c.off_dgrad != SIZE_MAX) {
        PackJobHost d{};
        d.w = params[c.w_idx]; d.p = pk + c.off_dgrad; d.cout = c.cout; d.cin = c.cin; d.rows_padded = c.cout;
        d.r = c.r; d.dgrad = 1;
        n->pack_jobs.push_back(d);
      }
    }
    if (cudaMemcpyAsync(jobs_dev, n->pack_jobs.data(), n->pack_jobs.size() * sizeof(PackJobHost),
                        cudaMemcpyHostToDevice, cudaStream_t(stream)) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "net_pack: job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    n->pack_sig_packed = packed; n->pack_sig_p0 = params[0]; n->pack_sig_training = training;
  }
  return pack_batched_launch(jobs_dev, int(n->pack_jobs.size()), cudaStream_t(stream));
}

int rumpy_net_forward(void* net_, const float* const* params, const void* packed, const float* x_nchw,
                      float* y_nchw, void* workspace, int N, int H, int W, int training, void* stream_) {
  Net* n = static_cast<Net*>(net_);
  if (!n || !params || !packed || !x_nchw || !y_nchw || !workspace)
    return set_error(RUMPY_ERR_ARG, "net_forward: null pointer");
  OptScope scope(&n->opt);
  if (int e = device_info(nullptr)) return e;
  cudaStream_t stream = cudaStream_t(stream_);
  if (int e = ensure_plan(n, packed, workspace, N, H, W, training)) return e;
  if (n->ca_counters_dirty && n->ca_counters) {
    if (cudaMemsetAsync(n->ca_counters, 0, n->ca_counters_bytes, stream) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "net_forward: counter reset failed");
    n->ca_counters_dirty = false;
  }
  for (Op& op : n->ops) {
    switch (op.type) {
      case OP_HEAD:
        if (int e = rumpy_head_conv(x_nchw, params[op.head_w], params[op.head_b], op.yf, op.yb, N, H, W, n->in_feats,
                                    n->C, stream))
          return e;
        break;
      case OP_LAM:
        if (int e = lam_launch(op.stack, op.lam_scratch, params[n->p_la_gamma], op.lam_out, N, H * W, stream)) return e;
        break;
      case OP_CSAM:
        if (int e = csam_cat_launch(op.csam_x, op.csam_out2, params[n->p_csa_w], params[n->p_csa_b],
                                    params[n->p_csa_gamma], op.cat, N, H, W, stream))
          return e;
        break;
      case OP_QSCALE: {
        if (!n->meta_dev || n->meta_n != N)
          return set_error(RUMPY_ERR_ARG, "meta-attention forward: call rumpy_net_set_metadata with a [%d][%d] tensor first", N,
                           n->num_meta);
        std::vector<QScaleJob> jobs(n->qs.size());
        for (size_t i = 0; i < jobs.size(); ++i) {
          const CAW& q = n->qs[i];
          jobs[i] = q.w1 >= 0 ? QScaleJob{params[q.w1], params[q.b1], params[q.w2], params[q.b2], nullptr}
                              : QScaleJob{nullptr, nullptr, nullptr, nullptr, nullptr};
          jobs[i].out = n->q_scale + i * size_t(N) * n->C;
        }
        if (jobs.size() != n->q_jobs_uploaded.size() ||
            memcmp(jobs.data(), n->q_jobs_uploaded.data(), jobs.size() * sizeof(QScaleJob)) != 0) {
          n->q_jobs_uploaded = jobs;
          if (cudaMemcpyAsync(n->q_jobs_dev, n->q_jobs_uploaded.data(), jobs.size() * sizeof(QScaleJob),
                              cudaMemcpyHostToDevice, stream) != cudaSuccess)
            return set_error(RUMPY_ERR_CUDA, "Q-RCAN: job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
          cudaStreamSynchronize(stream);
        }
        if (int e = q_scale_launch(n->q_jobs_dev, int(jobs.size()), n->meta_dev,
                                   N, n->meta_m, n->q_hidden, n->C, n->modulate, n->q_relu, stream))
          return e;
        break;
      }
      case OP_CONV:
        if (int e = launch_conv_op(op, params, y_nchw, stream)) return e;
        break;
      case OP_TRUNK:
        if (int e = trunk_launch(n->trunk.get(), params, stream)) return e;
        break;
      case OP_CONV_CA:
        op.conv.args.bias = params[op.bias_param];
        op.cafused.w1 = params[op.ca.w1]; op.cafused.b1 = params[op.ca.b1];
        op.cafused.w2 = params[op.ca.w2]; op.cafused.b2 = params[op.ca.b2];
        if (int e = conv_ca_launch(op.conv, op.cafused, stream)) return e;
        break;
      case OP_CA:
        if (int e = ca_apply_launch(op.pool, conv_pool_rows(H, W, n->C),
                                    op.pool_compact, op.u, n->plan_u_f32, op.x_in, params[op.ca.w1], params[op.ca.b1],
                                    params[op.ca.w2], params[op.ca.b2], op.x_out, op.x_out_b, op.save_mean,
                                    op.save_hid, op.save_y, N, H, W, n->C, n->C / n->reduction, stream, op.q_scale))
          return e;
        break;
      default:
        return set_error(RUMPY_ERR_ARG, "net_forward: unexpected op");
    }
  }
  return RUMPY_OK;
}

// Backward of the last training forward on the same (packed, workspace, shape): dy is the upstream gradient in
// the reference's fp32 NCHW layout; grads[i] receives d(loss)/d(params[i]) (overwritten, not accumulated).
int rumpy_net_backward(void* net_, const float* const* params, const void* packed, const float* x_nchw,
                       const float* dy_nchw, float* const* grads, void* workspace, int N, int H, int W,
                       void* stream_) {
  Net* n = static_cast<Net*>(net_);
  if (!n || !params || !packed || !x_nchw || !dy_nchw || !grads || !workspace)
    return set_error(RUMPY_ERR_ARG, "net_backward: null pointer");
  OptScope scope(&n->opt);
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  cudaStream_t stream = cudaStream_t(stream_);
  if (int e = ensure_plan(n, packed, workspace, N, H, W, 1)) return e;
  const int C = n->C, Cr = C / n->reduction;
  // (re)bind gradient pointers into the reduce / colsum job lists and upload them when anything changed
  bool rebind = !n->jobs_uploaded || n->plan_grads.size() != size_t(n->n_params);
  if (!rebind)
    for (int i = 0; i < n->n_params; ++i)
      if (n->plan_grads[i] != grads[i]) { rebind = true; break; }
  if (rebind) {
    n->plan_grads.assign(grads, grads + n->n_params);
    std::vector<WgradReduceJob> rj = n->wg_rjobs;
    for (WgradReduceJob& j : rj) { j.dw = grads[j.accumulate]; j.accumulate = 0; }
    std::vector<ColsumJob> cj = n->cs_jobs;
    for (size_t i = 0; i < cj.size(); ++i) cj[i].db = grads[n->cs_bias_param[i]];
    cudaError_t e1 = cudaMemcpyAsync(n->wg_jobs_dev, n->wg_jobs.data(), n->wg_jobs.size() * sizeof(WgradJob),
                                     cudaMemcpyHostToDevice, stream);
    cudaError_t e2 = cudaMemcpyAsync(n->wg_rjobs_dev, rj.data(), rj.size() * sizeof(WgradReduceJob),
                                     cudaMemcpyHostToDevice, stream);
    cudaError_t e3 = cudaMemcpyAsync(n->cs_jobs_dev, cj.data(), cj.size() * sizeof(ColsumJob), cudaMemcpyHostToDevice,
                                     stream);
    std::vector<PartialSumJob> pj = n->ps_jobs;
    for (size_t i = 0; i < pj.size(); ++i) pj[i].db = grads[n->ps_bias_param[i]];
    cudaError_t e4 = pj.empty() ? cudaSuccess : cudaMemcpyAsync(n->ps_jobs_dev, pj.data(), pj.size() * sizeof(PartialSumJob), cudaMemcpyHostToDevice, stream);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "net_backward: job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    if (n->pg_counter) cudaMemsetAsync(n->pg_counter, 0, size_t(N + 1) * sizeof(int), stream);
    cudaStreamSynchronize(stream);  // the temporaries above die at scope exit; happens once per plan
    n->jobs_uploaded = true;
  }
  const ConvW& tail = n->convs[n->conv_tail];
  const ConvW& head = n->convs[0];
  const int thin_block = (256 / C > 0 ? 256 / C : 1) * C, thin_lanes = thin_block / C;
  for (Op& op : n->bops) {
    switch (op.type) {
      case OP_TAIL_BWD: {
        const int Hh = op.h, Wh = op.w, M = n->out_feats;
        const size_t items = size_t(N) * Hh * ((Wh + 3) / 4) * (C / 8);
        tail_dgrad_kernel<<<grid_for(items, 256, 8), 256, size_t(M) * 9 * C * sizeof(float), stream>>>(
            dy_nchw, params[tail.w_idx], static_cast<__nv_bfloat16*>(op.g_hr), N, Hh, Wh, C, M);
        if (int e = check_launch("tail_dgrad")) return e;
        pad_thin_grad_kernel<<<grid_for(size_t(N) * Hh * Wh * 8, 256, 8), 256, 0, stream>>>(
            dy_nchw, static_cast<__nv_bfloat16*>(op.dst_b), N, M, Hh * Wh);
        if (int e = check_launch("tail_grad_pad")) return e;
        plane_sum_kernel<<<dim3(kPlaneSlices, M), 512, 0, stream>>>(dy_nchw, op.thin_partial, N, M, Hh * Wh);
        if (int e = check_launch("tail_bias_grad")) return e;
        plane_sum_finalize_kernel<<<1, 32, 0, stream>>>(op.thin_partial, kPlaneSlices, grads[tail.b_idx], M);
        if (int e = check_launch("tail_bias_grad_finalize")) return e;
        break;
      }
      case OP_CONV:
        if (int e = launch_conv_op(op, params, nullptr, stream)) return e;
        break;
      case OP_CSAM_BWD:
        if (int e = csam_bwd_launch(op.csam_x, op.han_f[0], params[n->p_csa_w], params[n->p_csa_b], params[n->p_csa_gamma],
                                    op.han_f[1], op.han_f[2], op.han_v, op.han_dx[0], op.han_f[3], N, H, W, stream))
          return e;
        break;
      case OP_LAM_BWD:
        if (int e = lam_bwd_launch(op.stack, op.han_f[0], lam_att_ptr(op.lam_scratch, N), params[n->p_la_gamma], op.han_dx,
                                   op.han_v, op.han_f[3], op.han_i, N, H * W, stream))
          return e;
        break;
      case OP_HAN_PG:
        if (int e = han_param_grad_launch(op.han_f[3], op.han_i, N, grads[n->p_csa_w], grads[n->p_csa_b],
                                          grads[n->p_csa_gamma], grads[n->p_la_gamma], stream))
          return e;
        break;
      case OP_DQ:
        if (int e = dq_reduce_launch(op.a, op.u, op.tail_in, op.dst_f, N, H * W, C, stream)) return e;
        break;
      case OP_QGRAD: {   // q-layer parameter gradients: Q-EDSR from the q * dq slices of the OP_DQ passes, Q-RCAN
                         // (per-layer path) from the dq the CA backward wrote
        std::vector<QGradJobHost> jobs;
        for (size_t i = 0; i < n->qs.size(); ++i) {
          const CAW& q = n->qs[i];
          if (q.w1 < 0) continue;
          jobs.push_back(QGradJobHost{params[q.w1], params[q.b1], params[q.w2], params[q.b2],
                                      n->q_scale + i * size_t(N) * C,
                                      n->q_dq + i * size_t(N) * (op.ca_chunks > 0 ? op.ca_chunks : 1) * C,
                                      grads[q.w1], grads[q.b1], grads[q.w2], grads[q.b2]});
        }
        if (!jobs.empty() && (jobs.size() != n->qg_jobs_uploaded.size() ||
                              memcmp(jobs.data(), n->qg_jobs_uploaded.data(), jobs.size() * sizeof(QGradJobHost)) != 0)) {
          n->qg_jobs_uploaded = jobs;
          if (cudaMemcpyAsync(n->qg_jobs_dev, n->qg_jobs_uploaded.data(), jobs.size() * sizeof(QGradJobHost),
                              cudaMemcpyHostToDevice, stream) != cudaSuccess)
            return set_error(RUMPY_ERR_CUDA, "Q-EDSR: gradient job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
          cudaStreamSynchronize(stream);
        }
        if (int e = q_grad_launch(n->qg_jobs_dev, int(jobs.size()), n->meta_dev, N, n->meta_m, n->q_hidden, C, n->q_relu,
                                  op.ca_chunks, stream))
          return e;
        break;
      }
      case OP_TRUNK_BWD:
        if (int e = trunk_bwd_launch(n->trunk_bwd.get(), params, grads, stream)) return e;
        if (n->qrcan) {   // q-layer parameter gradients from the dq = s*y terms the kernel left behind
          std::vector<QGradJobHost> jobs;
          for (size_t i = 0; i < n->qs.size(); ++i) {
            const CAW& q = n->qs[i];
            if (q.w1 < 0) continue;
            jobs.push_back(QGradJobHost{params[q.w1], params[q.b1], params[q.w2], params[q.b2],
                                        n->q_scale + i * size_t(N) * C, n->q_dq + i * size_t(N) * C,
                                        grads[q.w1], grads[q.b1], grads[q.w2], grads[q.b2]});
          }
          if (!jobs.empty() && (jobs.size() != n->qg_jobs_uploaded.size() ||
                                memcmp(jobs.data(), n->qg_jobs_uploaded.data(), jobs.size() * sizeof(QGradJobHost)) != 0)) {
            n->qg_jobs_uploaded = jobs;
            if (cudaMemcpyAsync(n->qg_jobs_dev, n->qg_jobs_uploaded.data(), jobs.size() * sizeof(QGradJobHost),
                                cudaMemcpyHostToDevice, stream) != cudaSuccess)
              return set_error(RUMPY_ERR_CUDA, "Q-RCAN: gradient job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
            cudaStreamSynchronize(stream);
          }
          if (int e = q_grad_launch(n->qg_jobs_dev, int(jobs.size()), n->meta_dev, N, n->meta_m, n->q_hidden, C, n->q_relu,
                                    0, stream))
            return e;
        }
        break;
      case OP_CA_BWD: {
        const int HW = H * W;
        CaBwdArgs a{};
        a.G = op.a; a.u = op.u; a.save_mean = op.save_mean; a.save_hid = op.save_hid; a.save_y = op.save_y;
        a.w1 = params[op.ca.w1]; a.w2 = params[op.ca.w2];
        a.dw1 = grads[op.ca.w1]; a.db1 = grads[op.ca.b1]; a.dw2 = grads[op.ca.w2]; a.db2 = grads[op.ca.b2];
        a.s_partial = op.s_partial; a.coef = n->ca_coef; a.pg_scratch = n->pg_scratch; a.counters = n->pg_counter;
        a.N = N; a.HW = HW; a.C = C; a.Cr = Cr;
        a.q_scale = op.q_scale; a.dq = op.dst_f;
        dim3 g1(kCaBwdChunks, N);
        const size_t smem = size_t(256 / (C / 4)) * C * sizeof(float);
        if (n->plan_u_f32) ca_bwd_reduce_kernel<true><<<g1, 256, smem, stream>>>(a);
        else ca_bwd_reduce_kernel<false><<<g1, 256, smem, stream>>>(a);
        if (int e = check_launch("ca_bwd_reduce")) return e;
        ca_bwd_apply_kernel<<<dim3(op.ca_chunks, N), 256, 0, stream>>>(
            op.a, op.save_y, n->ca_coef, static_cast<__nv_bfloat16*>(op.du), op.du_colsum, HW, C, op.q_scale);
        if (int e = check_launch("ca_bwd_apply")) return e;
        break;
      }
      case OP_ADD:
        add_f32_bf16_kernel<<<grid_for(op.n4, 256, 8), 256, 0, stream>>>(op.a, op.b, op.dst_f,
                                                                        static_cast<__nv_bfloat16*>(op.dst_b), op.n4);
        if (int e = check_launch("add_f32_bf16")) return e;
        break;
      case OP_HEAD_WGRAD: {
        const int M = n->in_feats, rows = M * 9 + 1;
        if (M <= 3)
          thin_wgrad_kernel<false, 3><<<kThinBlocks, thin_block, size_t(thin_lanes) * rows * C * sizeof(float), stream>>>(
              x_nchw, op.a, op.b, op.thin_partial, N, H, W, C, M, +1);
        else
          thin_wgrad_kernel<false, 4><<<kThinBlocks, thin_block, size_t(thin_lanes) * rows * C * sizeof(float), stream>>>(
              x_nchw, op.a, op.b, op.thin_partial, N, H, W, C, M, +1);
        if (int e = check_launch("head_wgrad")) return e;
        thin_wgrad_reduce_kernel<<<((M * 9 + 1) * C + 31) / 32, 256, 0, stream>>>(op.thin_partial, kThinBlocks, grads[head.w_idx],
                                                        grads[head.b_idx], C, M, 1);
        if (int e = check_launch("head_wgrad_reduce")) return e;
        break;
      }
      default:
        return set_error(RUMPY_ERR_ARG, "net_backward: unexpected op");
    }
  }
  const int ncs = int(n->cs_jobs.size());
  if (ncs > 0) {
    colsum_kernel<<<dim3(kColsumSlices, ncs), 768, 768 * sizeof(float), stream>>>(n->cs_jobs_dev);
    if (int e = check_launch("colsum")) return e;
    colsum_reduce_kernel<<<ncs, 256, 0, stream>>>(n->cs_jobs_dev);
    if (int e = check_launch("colsum_reduce")) return e;
  }
  if (!n->ps_jobs.empty()) {
    const int block = (256 / C > 0 ? 256 / C : 1) * C;
    partial_sum_kernel<<<int(n->ps_jobs.size()), block, block * sizeof(float), stream>>>(n->ps_jobs_dev);
    if (int e = check_launch("partial_sum")) return e;
  }
  // weight gradients last, in chunks from the last parameters to the first: after chunk k every gradient with
  // parameter index >= wg_chunk_first_param[k] is final
  int j0 = 0, r0 = 0;
  for (size_t k = 0; k < n->wg_chunk_job_end.size(); ++k) {
    const int j1 = n->wg_chunk_job_end[k], r1 = n->wg_chunk_rjob_end[k];
    if (j1 > j0)
      if (int e = wgrad_launch(n->wg_jobs_dev + j0, j1 - j0, n->wg_rjobs_dev + r0, r1 - r0, stream)) return e;
    if (int(k) < n->n_bwd_events && n->bwd_events[k] != nullptr) cudaEventRecord(n->bwd_events[k], stream);
    j0 = j1; r0 = r1;
  }
  return RUMPY_OK;
}

/* Gradient ranges of the chunked backward: after chunk k (and the event registered for it) every gradient with
 * parameter index >= first_param[k] is final; first_param is decreasing and ends with 0.  Valid after the training
 * plan exists (any rumpy_net_forward(training=1) with the current shape).  Returns the number of chunks. */
int rumpy_net_backward_chunks(void* net, int* first_param, int max_chunks) {
  Net* n = static_cast<Net*>(net);
  if (!n || !first_param) return set_error(RUMPY_ERR_ARG, "net_backward_chunks: null");
  const int k = int(n->wg_chunk_first_param.size());
  for (int i = 0; i < k && i < max_chunks; ++i) first_param[i] = n->wg_chunk_first_param[i];
  return k;
}

/* events: cudaEvent_t handles (at most 8) recorded by rumpy_net_backward on its stream after chunk k; NULL / 0: none */
int rumpy_net_set_backward_events(void* net, void* const* events, int n_events) {
  Net* n = static_cast<Net*>(net);
  if (!n || n_events < 0 || n_events > 8) return set_error(RUMPY_ERR_ARG, "net_set_backward_events: bad arguments");
  for (int i = 0; i < 8; ++i) n->bwd_events[i] = (events && i < n_events) ? static_cast<cudaEvent_t>(events[i]) : nullptr;
  n->n_bwd_events = events ? n_events : 0;
  return RUMPY_OK;
}

// ------------------------------------------------------------------ loss / optimiser kernels
long long rumpy_l1_workspace_floats(void) { return 1024 + 4; }

// loss = mean |out - y| (nn.L1Loss, base_architecture.py:40) and dy = gscale * sign(out - y) / numel.
// ws: rumpy_l1_workspace_floats() floats.  loss_out: one float (device).  dy may be NULL (loss only).
int rumpy_l1_loss_grad(const float* out, const float* y, float* dy, float* loss_out, float* ws, long long numel,
                       float gscale, void* stream_) {
  if (int e = device_info(nullptr)) return e;
  if (!out || !y || !loss_out || !ws || numel <= 0) return set_error(RUMPY_ERR_ARG, "l1_loss_grad: bad args");
  cudaStream_t stream = cudaStream_t(stream_);
  size_t blocks = (size_t(numel) / 4 + 255) / 256;
  if (blocks > 1024) blocks = 1024;
  if (blocks < 1) blocks = 1;
  l1_loss_grad_kernel<<<int(blocks), 256, 0, stream>>>(out, y, dy, ws, size_t(numel), gscale);
  if (int e = check_launch("l1_loss_grad")) return e;
  l1_loss_finalize_kernel<<<1, 32, 0, stream>>>(ws, int(blocks), loss_out, size_t(numel));
  return check_launch("l1_loss_finalize");
}

// clip_grad_norm_ coefficient (base_architecture.py:434-435): coef_out[0] = min(1, max_norm / (||g|| + 1e-6)),
// coef_out[1] = ||g||.  ws: 1024 floats.
int rumpy_grad_clip_coef(const float* grad_flat, long long n, float max_norm, float* coef_out, float* ws,
                         void* stream_) {
  if (int e = device_info(nullptr)) return e;
  if (!grad_flat || !coef_out || !ws || n <= 0) return set_error(RUMPY_ERR_ARG, "grad_clip_coef: bad args");
  cudaStream_t stream = cudaStream_t(stream_);
  sumsq_kernel<<<1024, 256, 0, stream>>>(grad_flat, size_t(n), ws);
  if (int e = check_launch("sumsq")) return e;
  clip_coef_kernel<<<1, 32, 0, stream>>>(ws, 1024, max_norm, coef_out);
  return check_launch("clip_coef");
}

// One fused Adam step over flat fp32 buffers (torch.optim.Adam defaults, base_architecture.py:93-95).
// step is 1-based.  grad_scale_dev (device float, may be NULL) and grad_scale multiply the gradient first.
int rumpy_adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, int step, const float* grad_scale_dev, float grad_scale, void* stream_) {
  if (int e = device_info(nullptr)) return e;
  if (!p || !g || !m || !v || n <= 0 || step < 1) return set_error(RUMPY_ERR_ARG, "adam_step: bad args");
  const double bc1 = 1.0 - pow(double(beta1), step), bc2 = 1.0 - pow(double(beta2), step);
  adam_kernel<<<grid_for(size_t(n), 256, 8), 256, 0, cudaStream_t(stream_)>>>(
      p, g, m, v, size_t(n), lr, beta1, beta2, eps, float(bc1), float(sqrt(bc2)), grad_scale_dev, grad_scale);
  return check_launch("adam_step");
}

}  // extern "C"
