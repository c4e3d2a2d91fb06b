// Whole-network executor for the RCAN / EDSR trunk: one C call enqueues every kernel of a forward pass.
//
// Mirrors the reference's module graph (/root/reference/rumpy/SISR/models/advanced/architectures.py):
//   RCAN.forward :171-176, ResidualGroup.forward :121-124, RCAB.forward :81-84, CALayer.forward :41-44,
//   EDSR.forward :236-241, ResBlock.forward common.py:71-75, Upsampler common.py:29-44.
// The layer program is built once per (shape, workspace, packed-weights) and cached: building encodes every
// TMA tensor map on the host, so a steady-state forward is just kernel launches (CUDA-graph capturable).
#include "../../include/rumpy_b200.h"
#include "host_util.cuh"
#include <vector>

namespace rb {

struct ConvW {            // one 3x3 conv of the network
  int w_idx, b_idx;       // indices into the state_dict-ordered parameter list
  int cout, cin;
  int rows_padded;        // packed rows (16 for the thin tail)
  int r;                  // pixel-shuffle factor folded into the packing (1 = none)
  size_t off_fwd;         // byte offsets into the packed buffer
  size_t off_bias;        // packed bias (only when r > 1 or padded), else SIZE_MAX
};

struct CAW { int w1, b1, w2, b2; };

enum OpType { OP_HEAD, OP_CONV, OP_CA };

struct Op {
  OpType type;
  ConvPlan conv;          // OP_CONV
  int bias_param;         // param index whose pointer is patched into conv.args.bias (-1: packed / none)
  bool writes_output;     // thin tail conv: out_nchw patched with the caller's y
  // OP_HEAD
  int head_w, head_b;
  float* yf; void* yb;
  // OP_CA
  CAW ca;
  const float* pool; const void* u; const float* x_in; float* x_out; void* x_out_b;
  float *save_mean, *save_hid, *save_y;
};

struct Net {
  int arch, C, n_groups, n_blocks, reduction, scale;
  float res_scale;
  int in_feats, out_feats, u_f32;
  std::vector<ConvW> convs;   // network order
  std::vector<CAW> cas;
  int n_params = 0;
  size_t packed_bytes = 0;
  // cached plan
  std::vector<Op> ops;
  const void* plan_packed = nullptr;
  void* plan_ws = nullptr;
  int pN = 0, pH = 0, pW = 0, p_training = -1;
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
  add_conv(n, C, C);  // body tail
  int r = 0;
  const int st = upsampler_stages(n->scale, &r);
  if (st < 0) return set_error(RUMPY_ERR_ARG, "scale %d unsupported (2^n or 3)", n->scale);
  for (int s = 0; s < st; ++s) add_conv(n, C * r * r, C, r);
  add_conv(n, n->out_feats, C, 1, 16);  // thin tail
  return RUMPY_OK;
}

struct Bump {
  char* base; size_t off = 0;
  void* take(size_t bytes) { void* p = base ? base + off : nullptr; off = align_up(off + bytes, 1024); return p; }
};

// Lays out the workspace and (when ws != nullptr) builds the op list.  Returns bytes needed.
static int build_plan(Net* n, const void* packed, void* ws, int N, int H, int W, int training, size_t* bytes_out,
                      bool build) {
  const int C = n->C;
  const size_t px = size_t(N) * H * W;
  Bump bp{static_cast<char*>(ws)};
  const char* pk = static_cast<const char*>(packed);
  std::vector<Op> ops;
  int ci = 0;  // conv cursor
  auto conv_op = [&](const ConvW& cw, ConvDesc d, int* err) {
    Op op{};
    op.type = OP_CONV;
    d.w = pk + cw.off_fwd;
    if (cw.off_bias != SIZE_MAX) { d.bias = reinterpret_cast<const float*>(pk + cw.off_bias); op.bias_param = -1; }
    else { d.bias = reinterpret_cast<const float*>(16); op.bias_param = cw.b_idx; }  // patched per call
    if (build) { if (int e = conv_plan_build(&op.conv, d)) *err = e; }
    ops.push_back(op);
  };
  int err = 0;
  // ---- buffers
  float* head_f = static_cast<float*>(bp.take(px * C * 4));
  void* head_b = bp.take(px * C * 2);
  float* S_f = static_cast<float*>(bp.take(px * C * 4));
  float* G_f[2] = {static_cast<float*>(bp.take(px * C * 4)), static_cast<float*>(bp.take(px * C * 4))};
  const int tiles = ((H + kTileH - 1) / kTileH) * ((W + kTileW - 1) / kTileW);
  // ---- head
  {
    Op op{};
    op.type = OP_HEAD;
    op.head_w = n->convs[0].w_idx; op.head_b = n->convs[0].b_idx;
    op.yf = head_f; op.yb = head_b;
    ops.push_back(op);
    ci = 1;
  }
  const void* cur_b = head_b;     // bf16 operand of the running activation
  const float* cur_f = head_f;    // its fp32 residual-stream copy
  if (n->arch == 0) {
    const int Cr = C / n->reduction;
    const size_t u_bytes = px * C * (n->u_f32 ? 4 : 2);
    void* xb_shared = training ? nullptr : bp.take(px * C * 2);
    void* t_shared = training ? nullptr : bp.take(px * C * 2);
    void* u_shared = training ? nullptr : bp.take(u_bytes);
    float* pool_shared = training ? nullptr : static_cast<float*>(bp.take(size_t(N) * tiles * 2 * C * 4));
    void* gb[2] = {bp.take(px * C * 2), bp.take(px * C * 2)};
    int cai = 0;
    for (int g = 0; g < n->n_groups; ++g) {
      const float* gin_f = cur_f;
      for (int b = 0; b < n->n_blocks; ++b) {
        void* t = training ? bp.take(px * C * 2) : t_shared;
        void* u = training ? bp.take(u_bytes) : u_shared;
        void* xb = training ? bp.take(px * C * 2) : xb_shared;
        float* pool = training ? static_cast<float*>(bp.take(size_t(N) * tiles * 2 * C * 4)) : pool_shared;
        float* sv = training ? static_cast<float*>(bp.take(size_t(N) * (2 * C + Cr) * 4)) : nullptr;
        ConvDesc d1{};
        d1.x = cur_b; d1.y_bf16 = t; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C; d1.Cout = C; d1.flags = kConvRelu;
        d1.alpha = 1.f;
        conv_op(n->convs[ci++], d1, &err);
        ConvDesc d2{};
        d2.x = t; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C; d2.Cout = C; d2.flags = kConvPool; d2.alpha = 1.f;
        if (n->u_f32) d2.y_f32 = static_cast<float*>(u); else d2.y_bf16 = u;
        d2.pool_partial = pool;
        conv_op(n->convs[ci++], d2, &err);
        Op ca{};
        ca.type = OP_CA;
        ca.ca = n->cas[cai++];
        ca.pool = pool; ca.u = u; ca.x_in = (b == 0) ? gin_f : S_f; ca.x_out = S_f; ca.x_out_b = xb;
        if (sv) { ca.save_mean = sv; ca.save_y = sv + size_t(N) * C; ca.save_hid = sv + size_t(N) * 2 * C; }
        ops.push_back(ca);
        cur_b = xb; cur_f = S_f;
      }
      ConvDesc dg{};
      dg.x = cur_b; dg.residual = gin_f; dg.y_f32 = G_f[g & 1]; dg.y_bf16 = training ? bp.take(px * C * 2) : gb[g & 1];
      dg.N = N; dg.H = H; dg.W = W; dg.Cin = C; dg.Cout = C; dg.alpha = 1.f;
      void* gout_b = dg.y_bf16;
      conv_op(n->convs[ci++], dg, &err);
      cur_b = gout_b; cur_f = G_f[g & 1];
    }
  } else {
    void* t_shared = training ? nullptr : bp.take(px * C * 2);
    void* xb_shared = training ? nullptr : bp.take(px * C * 2);
    for (int b = 0; b < n->n_blocks; ++b) {
      void* t = training ? bp.take(px * C * 2) : t_shared;
      void* xb = training ? bp.take(px * C * 2) : xb_shared;
      ConvDesc d1{};
      d1.x = cur_b; d1.y_bf16 = t; d1.N = N; d1.H = H; d1.W = W; d1.Cin = C; d1.Cout = C; d1.flags = kConvRelu;
      d1.alpha = 1.f;
      conv_op(n->convs[ci++], d1, &err);
      ConvDesc d2{};
      d2.x = t; d2.residual = cur_f; d2.y_f32 = S_f; d2.y_bf16 = xb; d2.N = N; d2.H = H; d2.W = W; d2.Cin = C;
      d2.Cout = C; d2.alpha = n->res_scale;   // conv2(.)*res_scale + x   (common.py:72-73)
      conv_op(n->convs[ci++], d2, &err);
      cur_b = xb; cur_f = S_f;
    }
  }
  // ---- body tail conv + global skip (architectures.py:173-174 / :238-239): operand for the upsampler only
  void* body_b = bp.take(px * C * 2);
  {
    ConvDesc d{};
    d.x = cur_b; d.residual = head_f; d.y_bf16 = body_b; d.N = N; d.H = H; d.W = W; d.Cin = C; d.Cout = C;
    d.alpha = 1.f;
    conv_op(n->convs[ci++], d, &err);
    cur_b = body_b;
  }
  // ---- upsampler: conv C -> C*r*r with the PixelShuffle folded into the store
  int r = 0;
  const int st = upsampler_stages(n->scale, &r);
  int h = H, w = W;
  for (int s = 0; s < st; ++s) {
    void* up = bp.take(size_t(N) * (h * r) * (w * r) * C * 2);
    ConvDesc d{};
    d.x = cur_b; d.y_bf16 = up; d.N = N; d.H = h; d.W = w; d.Cin = C; d.Cout = C * r * r; d.out_r = r; d.alpha = 1.f;
    conv_op(n->convs[ci++], d, &err);
    cur_b = up; h *= r; w *= r;
  }
  // ---- thin tail conv -> caller's fp32 NCHW output
  {
    ConvDesc d{};
    d.x = cur_b; d.N = N; d.H = h; d.W = w; d.Cin = C; d.Cout = 16; d.cout_real = n->out_feats; d.alpha = 1.f;
    d.out_nchw = reinterpret_cast<float*>(16);  // patched per call
    conv_op(n->convs[ci++], d, &err);
    ops.back().writes_output = true;
  }
  *bytes_out = bp.off;
  if (err) return err;
  if (build) n->ops.swap(ops);
  return RUMPY_OK;
}

}  // namespace rb

using namespace rb;

extern "C" {

int rumpy_net_create(void** out, int arch, int n_feats, int n_groups, int n_blocks, int reduction, int scale,
                     float res_scale, int in_feats, int out_feats, int u_f32) {
  if (!out) return set_error(RUMPY_ERR_ARG, "net_create: null out");
  if (arch != 0 && arch != 1) return set_error(RUMPY_ERR_ARG, "net_create: arch %d", arch);
  if (n_feats % 64 != 0 || n_feats <= 0 || n_feats > 256)
    return set_error(RUMPY_ERR_ARG, "net_create: n_feats=%d must be 64, 128, 192 or 256", n_feats);
  if (in_feats < 1 || in_feats > 4 || out_feats < 1 || out_feats > 16)
    return set_error(RUMPY_ERR_ARG, "net_create: in_feats=%d out_feats=%d", in_feats, out_feats);
  if (arch == 0 && (reduction < 1 || n_feats % reduction != 0 || n_feats / reduction > 64))
    return set_error(RUMPY_ERR_ARG, "net_create: reduction=%d", reduction);
  Net* n = new Net();
  n->arch = arch; n->C = n_feats; n->n_groups = n_groups; n->n_blocks = n_blocks; n->reduction = reduction;
  n->scale = scale; n->res_scale = res_scale; n->in_feats = in_feats; n->out_feats = out_feats; n->u_f32 = u_f32;
  if (int e = net_init(n)) { delete n; return e; }
  *out = n;
  return RUMPY_OK;
}

int rumpy_net_destroy(void* net) {
  delete static_cast<Net*>(net);
  return RUMPY_OK;
}

/* kernels enqueued by one forward of the cached plan (0 before the first forward) */
int rumpy_net_num_launches(void* net) { return net ? int(static_cast<Net*>(net)->ops.size()) : -1; }

int rumpy_net_num_params(void* net) { return net ? static_cast<Net*>(net)->n_params : -1; }

long long rumpy_net_packed_bytes(void* net) { return net ? (long long)static_cast<Net*>(net)->packed_bytes : -1; }

long long rumpy_net_workspace_bytes(void* net, int N, int H, int W, int training) {
  if (!net) return -1;
  size_t bytes = 0;
  if (build_plan(static_cast<Net*>(net), nullptr, nullptr, N, H, W, training, &bytes, false)) return -1;
  return (long long)bytes;
}

// fp32 OIHW parameters (device pointers, state_dict order) -> packed bf16 operands.  Call after every
// optimiser step (weights changed) and before the first forward.
int rumpy_net_pack(void* net_, const float* const* params, void* packed, void* stream) {
  Net* n = static_cast<Net*>(net_);
  if (!n || !params || !packed) return set_error(RUMPY_ERR_ARG, "net_pack: null");
  char* pk = static_cast<char*>(packed);
  for (size_t i = 1; i < n->convs.size(); ++i) {  // conv 0 is the fp32 head
    const ConvW& c = n->convs[i];
    if (int e = rumpy_pack_conv3x3(params[c.w_idx], pk + c.off_fwd, c.cout, c.cin, c.rows_padded, c.r, 0, stream))
      return e;
    if (c.off_bias != SIZE_MAX)
      if (int e = rumpy_pack_bias(params[c.b_idx], reinterpret_cast<float*>(pk + c.off_bias), c.cout, c.rows_padded,
                                  c.r, stream))
        return e;
  }
  return RUMPY_OK;
}

int rumpy_net_forward(void* net_, const float* const* params, const void* packed, const float* x_nchw,
                      float* y_nchw, void* workspace, int N, int H, int W, int training, void* stream_) {
  Net* n = static_cast<Net*>(net_);
  if (!n || !params || !packed || !x_nchw || !y_nchw || !workspace)
    return set_error(RUMPY_ERR_ARG, "net_forward: null pointer");
  if (int e = device_info(nullptr)) return e;
  cudaStream_t stream = cudaStream_t(stream_);
  if (n->plan_packed != packed || n->plan_ws != workspace || n->pN != N || n->pH != H || n->pW != W ||
      n->p_training != training) {
    size_t bytes = 0;
    n->plan_packed = nullptr;
    if (int e = build_plan(n, packed, workspace, N, H, W, training, &bytes, true)) return e;
    n->plan_packed = packed; n->plan_ws = workspace; n->pN = N; n->pH = H; n->pW = W; n->p_training = training;
  }
  for (Op& op : n->ops) {
    switch (op.type) {
      case OP_HEAD:
        if (int e = rumpy_head_conv(x_nchw, params[op.head_w], params[op.head_b], op.yf, op.yb, N, H, W, n->in_feats,
                                    n->C, stream))
          return e;
        break;
      case OP_CONV:
        if (op.bias_param >= 0) op.conv.args.bias = params[op.bias_param];
        if (op.writes_output) op.conv.args.out_nchw = y_nchw;
        if (int e = conv_plan_launch(op.conv, stream)) return e;
        break;
      case OP_CA:
        if (int e = rumpy_ca_apply(op.pool, op.u, n->u_f32, op.x_in, params[op.ca.w1], params[op.ca.b1],
                                   params[op.ca.w2], params[op.ca.b2], op.x_out, op.x_out_b, op.save_mean,
                                   op.save_hid, op.save_y, N, H, W, n->C, n->C / n->reduction, stream))
          return e;
        break;
    }
  }
  return RUMPY_OK;
}

}  // extern "C"
