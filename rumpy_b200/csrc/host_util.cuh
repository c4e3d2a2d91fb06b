// Host-side declarations shared by api.cu / net.cu.
#pragma once
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <cuda.h>
#include <cuda_runtime.h>
#include "conv3x3_tc.cuh"
#include "conv3x3_ca.cuh"
#include "trunk_pipe.cuh"
#include "trunk_cluster.cuh"
#include "trunk_band.cuh"
#include "trunk_bwd.cuh"
#include <vector>

namespace rb {
// Per-handle execution options (rumpy_net_set_option / rumpy_net_set_trunk_events / rumpy_net_set_timeline in
// include/rumpy_b200.h).  There is NO process-global switch: every rumpy_net_* entry point installs its handle's
// options for the duration of the call on the calling thread (OptScope); code that runs outside a net handle (the
// stand-alone conv / wgrad entry points) sees the defaults.
struct Options {
  int use_trunk = 1;          // whole 64-channel body in ONE kernel when the shape fits (0: one kernel per layer)
  int use_cluster = 1;        // prefer the one-cluster-per-image kernel (trunk_cluster.cuh) when it fits
  int cluster_groups = 2;     // epilogue groups of the cluster kernel (2 or 4)
  int cluster_dbg = 0;        // cluster kernel timing experiments (garbage results): see ClusterArgs::dbg_flags
  int cluster_split = 1;      // cluster kernel: two-phase layer hand-over on vertical strips (trunk_cluster.cuh)
  int use_band = 0;           // role-swapped band kernel (trunk_band.cuh): experiment, slower than the cluster kernel
  int use_trunk_bwd = 1;      // backward of the RCAN body in the persistent dataflow kernel (trunk_bwd.cuh)
  int use_fused_ca = 0;       // conv2 + CALayer in one kernel (conv3x3_ca.cuh): correct, not faster
  int wgrad_chunks = 4;       // chunks of the batched wgrad (gradient ranges handed to the all-reduce one by one)
  int wgrad_tiles_per_split = 64;   // pixel tiles per split-K job (measured: 32 -> 14.97, 64 -> 14.76, 128 -> 14.73 ms)
  int use_pdl = 1;            // programmatic dependent launch between per-layer kernels
  int infer_u_bf16 = 1;       // per-layer inference: pre-attention activation u in bf16 (net.cu build_plan)
  int conv_dbg = 0;           // per-layer conv kernel timing experiments (garbage results): see ConvArgs::dbg_mode
  int trunk_sync_mode = 8;    // dataflow kernel: release store of the tile epoch (needed, DESIGN.md trunk protocol)
  int trunk_dbg_layers = 0;   // > 0: trunk kernels write a clock64 timeline of this many layers to `timeline`
  long long* timeline = nullptr;
  cudaEvent_t trunk_ev0 = nullptr, trunk_ev1 = nullptr;   // recorded right before / after the trunk kernel
  // everything a cached plan depends on
  unsigned plan_sig() const {
    return unsigned(use_trunk) | unsigned(use_cluster) << 1 | unsigned(use_trunk_bwd) << 2 | unsigned(use_band) << 3 |
           unsigned(use_fused_ca) << 4 | unsigned(cluster_groups == 4) << 5 | unsigned(use_pdl) << 6 |
           unsigned(wgrad_chunks) << 8 | unsigned(wgrad_tiles_per_split) << 12 |
           unsigned(timeline != nullptr) << 28 | unsigned(cluster_split) << 29 | unsigned(infer_u_bf16) << 30 | unsigned(cluster_dbg != 0 || conv_dbg != 0) << 31;
  }
};
const Options& opt();
struct OptScope {
  const Options* prev;
  explicit OptScope(const Options* o);
  ~OptScope();
};
}  // namespace rb

namespace rb {

extern thread_local std::string g_last_error;
int set_error(int code, const char* fmt, ...);
int device_info(int* num_sms);
int grid_for(size_t work_items, int block, int per_sm);

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(-3, "%s launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

// One convolution call, host view.
struct ConvDesc {
  const void* x;          // bf16 NHWC input (see in_r)
  const void* w;          // packed bf16 weights
  const float* bias;      // packed-row order or null
  const float* residual;  // fp32 NHWC or null
  const void* mask;       // bf16 NHWC or null
  void* y_bf16;           // bf16 out or null (see out_r)
  float* y_f32;           // fp32 out or null
  float* pool_partial;    // CA pool partials or null
  float* out_nchw;        // thin tail variant: fp32 NCHW output
  int cout_real;          // thin tail variant
  int N, H, W, Cin, Cout;
  int in_r, out_r;        // pixel-unshuffle on load / pixel-shuffle on store
  int force_bn;           // 0 = auto
  unsigned flags;
  float alpha;
  const float* ch_scale;  // [N][Cout] per-(image, channel) multiplier of alpha * (acc + bias) (Q-EDSR), or null
  const float* bf16_scale;  // [N][Cout] multiplier of the bf16 output only (Q-EDSR backward), or null
};

// Everything a launch needs: tensor maps (host copy, passed by value as __grid_constant__) + args.
struct ConvPlan {
  ConvMaps maps;
  ConvArgs args;
  int bn;
  bool resident;
  int grid;
  size_t smem;
};

int make_map_nhwc_sub(CUtensorMap* m, bool f32, const void* base, int C, int W, int H, int N, int r, int q,
                      int box_h, int box_w = kTileW);
// Rows per image of the pool-partial buffer a pooled C-channel conv writes (two 64-pixel halves per tile; the tile
// geometry of conv3x3_tc.cuh depends on whether the weights stay resident in shared memory, i.e. on C).
inline int conv_pool_rows(int H, int W, int C) {
  const bool tall = size_t(9) * (C / 64) * 64 * 128 <= 96 * 1024;
  const int th = tall ? kTallH : kTileH, tw = tall ? kTallW : kTileW;
  return 2 * ((H + th - 1) / th) * ((W + tw - 1) / tw);
}
int conv_plan_build(ConvPlan* p, const ConvDesc& d);
// RCAB tail fused kernel (conv2 + channel attention + skip), conv3x3_ca.cuh.  d: x = conv2 input (bf16), w/bias,
// residual = x_in (fp32), y_f32 / y_bf16 = x_out, pool_partial; u_store (fp32 NHWC) may be null (inference).
bool conv_ca_supported(int N, int H, int W, int Cin, int Cout);
int conv_ca_plan_build(ConvPlan* p, CaFusedArgs* ca, const ConvDesc& d, float* u_store, int Cr,
                       unsigned long long* grid_bar, float* save_mean, float* save_hid, float* save_y);
int conv_ca_launch(const ConvPlan& p, const CaFusedArgs& ca, cudaStream_t s);
int ca_apply_launch(const float* pool_partial, int partials_per_img, float* compact_scratch, const void* u,
                    int u_is_f32, const float* x_in, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* x_out, void* x_out_bf16, float* save_mean, float* save_hid,
                    float* save_y, int N, int H, int W, int C, int Cr, cudaStream_t stream,
                    const float* q_scale = nullptr);
// Q-RCAN meta-attention multipliers (misc_kernels.cuh q_scale_kernel): one job per RCAB
struct QScaleJobHost { const float *w1, *b1, *w2, *b2; float* out; };
int q_scale_launch(const QScaleJobHost* jobs_dev, int njobs, const float* meta, int N, int M, int hidden, int C,
                   int modulate, int relu, cudaStream_t s);

// HAN attention modules (han.cu): layer attention over the kLamLayers stacked fp32 NHWC feature maps -> bf16 NHWC with
// kLamLayers*64 channels; channel-spatial attention + concat -> bf16 NHWC with 128 channels
constexpr int kLamLayers = 11;   // the reference hard-codes n_feats*11 (10 residual groups + the body conv)
int lam_workspace_floats(int N);
int lam_launch(const float* const* stack, float* scratch, const float* gamma, void* out_bf16, int N, int HW, cudaStream_t s);
int csam_cat_launch(const float* x, const float* out2, const float* w, const float* b, const float* gamma, void* cat_bf16,
                    int N, int H, int W, cudaStream_t s);

// HAN training (han.cu): backward of CSAM / LAM and their parameter gradients
int han_csam_blocks(int N, int H, int W);
int han_bwd_scratch_floats(int N, int csam_blocks);
const float* lam_att_ptr(const float* lam_scratch, int N);
int csam_bwd_launch(const float* x, const float* dcat, const float* w, const float* b, const float* gamma, float* dc,
                    float* sig, void* d2_b, float* dx0, float* scratch, int N, int H, int W, cudaStream_t s);
int lam_bwd_launch(const float* const* stack, const float* G, const float* att, const float* gamma, float* const* dx,
                   void* dx0_b, float* scratch, int csam_blocks, int N, int HW, cudaStream_t s);
int han_param_grad_launch(const float* scratch, int csam_blocks, int N, float* dw, float* db, float* dg_csa, float* dg_la,
                          cudaStream_t s);

// one q-layer's parameter-gradient job (trunk_bwd.cuh: QGradJob has the same layout)
struct QGradJobHost { const float *w1, *b1, *w2, *b2, *q, *dq; float *dw1, *db1, *dw2, *db2; };
// dq_slices > 0: dq holds [N][dq_slices][C] partial sums of dq * q (Q-EDSR); 0: dq holds [N][C] (Q-RCAN)
int q_grad_launch(const QGradJobHost* jobs_dev, int njobs, const float* meta, int N, int M, int hidden, int C, int relu,
                  int dq_slices, cudaStream_t s);
// Q-EDSR: partial[n][slice][c] = sum over a pixel slice of g[n,p,c] * (out[n,p,c] - x[n,p,c])  (= dq * q)
constexpr int kDqSlices = 32;
int dq_reduce_launch(const float* g_f32, const void* out_bf16, const void* x_bf16, float* partial, int N, int HW, int C,
                     cudaStream_t s);

// one conv's packing job for the batched pack kernel (misc_kernels.cuh: PackJob has the same layout)
struct PackJobHost {
  const float* w; void* p; const float* b; float* bp;
  int cout, cin, rows_padded, r, dgrad;
};
int pack_batched_launch(const PackJobHost* jobs_dev, int njobs, cudaStream_t s);
int conv_plan_launch(const ConvPlan& p, cudaStream_t s);

// ------------------------------------------------------------------ persistent trunk kernel (trunk_pipe.cuh)
// packed weights of `n_layers` consecutive 64->64 convs ([layer][tap][64][64] bf16) -> 4-D map, box = one kx third
int make_map_weight_layers(CUtensorMap* m, const void* base, int n_layers, int box_taps = 3);

struct TrunkLayerParams { int bias, w1, b1, w2, b2; };   // indices into the caller's parameter list (-1: none)

struct TrunkPlan {
  CUtensorMap w_map;
  TrunkArgs args;
  int grid = 0;
  std::vector<TrunkLayer> layers;          // host copy; parameter pointers are patched per call
  std::vector<TrunkLayerParams> lparams;
  std::vector<TrunkLayer> uploaded;        // what the device table currently holds
  std::vector<const void*> in_bufs;        // bf16 NHWC tensors read through in_maps[i]
  std::vector<void*> out_bufs;             // bf16 NHWC tensors written through out_maps[i]
  // device areas (carved from the caller's workspace)
  TrunkLayer* layers_dev = nullptr;
  CUtensorMap *in_maps_dev = nullptr, *out_maps_dev = nullptr;
  void* flags_dev = nullptr;
  size_t flags_bytes = 0;
  bool maps_uploaded = false;
  // one-cluster-per-image variant (trunk_cluster.cuh), chosen when the image fits a cluster's shared memory
  bool cluster = false;
  ClusterArgs cargs;
  int cluster_size = 0;
  int cluster_groups = 2;   // epilogue groups (template parameter of the cluster kernel) the plan was built for
  size_t cluster_smem = 0;
  // role-swapped band kernel (trunk_band.cuh): weights in tensor memory, one cluster of row bands per image
  bool band = false;
  BandArgs bargs;
  CUtensorMap w_tap_map;    // same packed weights, box = one tap (8 KB)
  size_t band_smem = 0;
};
// backward program of the body (trunk_bwd.cuh)
struct TrunkBwdLayerParams { int w1, w2; };              // parameter indices of a kBwdCA layer (-1: none)
struct TrunkBwdPgBind { int layer, dw1, db1, dw2, db2; };   // kBwdCA layer -> gradient parameter indices
struct CaPgJobHost { const float* pg; float *dw1, *db1, *dw2, *db2; };
struct TrunkBwdPlan {
  CUtensorMap w_map;
  TrunkBwdArgs args;
  int grid = 0;
  std::vector<TrunkBwdLayer> layers, uploaded;
  std::vector<TrunkBwdLayerParams> lparams;
  std::vector<TrunkBwdPgBind> pg_binds;
  std::vector<const void*> in_bufs;
  std::vector<void*> out_bufs;
  TrunkBwdLayer* layers_dev = nullptr;
  CUtensorMap *in_maps_dev = nullptr, *out_maps_dev = nullptr;
  void* flags_dev = nullptr;
  size_t flags_bytes = 0;
  CaPgJobHost* pg_jobs_dev = nullptr;
  std::vector<CaPgJobHost> pg_jobs_uploaded;
  bool maps_uploaded = false;
};
size_t trunk_bwd_device_bytes(int N, int H, int W, int n_layers, int n_in_maps, int n_out_maps, int n_ca);
int trunk_bwd_plan_finish(TrunkBwdPlan* plan, int N, int H, int W, int Cr, const void* w_base, int n_w_layers,
                          void* dev);
int trunk_bwd_launch(TrunkBwdPlan* plan, const float* const* params, float* const* grads, cudaStream_t s);

bool trunk_supported(int N, int H, int W, int C, int Cr);
// device bytes needed next to the activations: layer table, tensor maps, flags, pool partials
size_t trunk_device_bytes(int N, int H, int W, int n_layers, int n_in_maps, int n_out_maps);
// fills args / maps from plan->layers, in_bufs, out_bufs; `dev` = trunk_device_bytes() bytes of device memory
// allow_cluster: the cluster-per-image kernel ping-pongs two resident buffers, so it serves inference plans only
int trunk_plan_finish(TrunkPlan* plan, int N, int H, int W, int Cr, const void* w_base, const float* s_init,
                      void* dev, bool allow_cluster);
int trunk_launch(TrunkPlan* plan, const float* const* params, cudaStream_t s);

}  // namespace rb
