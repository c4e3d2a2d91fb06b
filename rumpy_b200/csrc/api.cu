// C-ABI entry points (include/rumpy_b200.h): argument checks, TMA tensor-map encoding, kernel launches.
#define RB_CONV_CA_KERNEL_IMPL
#include "../../include/rumpy_b200.h"
#include "host_util.cuh"
#include "conv3x3_tc.cuh"
#include "misc_kernels.cuh"

namespace rb {

thread_local std::string g_last_error;
static const Options g_default_options{};
static thread_local const Options* t_options = nullptr;   // options of the net whose entry point runs on this thread
const Options& opt() { return t_options ? *t_options : g_default_options; }
OptScope::OptScope(const Options* o) : prev(t_options) { t_options = o; }
OptScope::~OptScope() { t_options = prev; }

int set_error(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

// ------------------------------------------------------------------ driver entry point for TMA descriptors
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    (void)cudaGetLastError();
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  return fn;
}

int device_info(int* num_sms) {
  static int cached_sms = -1;
  if (cached_sms < 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
      (void)cudaGetLastError();
      return set_error(RUMPY_ERR_DEVICE, "no CUDA device (rumpy_b200 has no CPU fallback)");
    }
    int major = 0, sms = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (major != 10) return set_error(RUMPY_ERR_DEVICE, "device compute capability %d.x is not sm_100", major);
    if (!get_encode_fn()) return set_error(RUMPY_ERR_DEVICE, "driver lacks cuTensorMapEncodeTiled");
    cached_sms = sms;
  }
  if (num_sms) *num_sms = cached_sms;
  return RUMPY_OK;
}

// 4-D NHWC map {C, W, H, N} with explicit strides (bytes) and box {box_c, 16, 8, 1}, 128B swizzle, zero OOB fill.
int make_map_nhwc(CUtensorMap* m, bool f32, const void* base, int C, int W, int H, int N, uint64_t stride_w,
                  uint64_t stride_h, uint64_t stride_n, int box_h, int box_w) {
  const cuuint64_t dims[4] = {cuuint64_t(C), cuuint64_t(W), cuuint64_t(H), cuuint64_t(N)};
  const cuuint64_t strides[3] = {stride_w, stride_h, stride_n};
  const cuuint32_t box[4] = {cuuint32_t(f32 ? 32 : 64), cuuint32_t(box_w), cuuint32_t(box_h), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(RUMPY_ERR_ARG, "tensor base not 16B aligned");
  CUresult r = get_encode_fn()(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                               const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RUMPY_ERR_CUDA, "cuTensorMapEncodeTiled(NHWC) failed: %d", int(r));
  return RUMPY_OK;
}

// Dense NHWC tensor, optionally viewed through a pixel (un)shuffle of factor r: sub-pixel q=(i,j) of the
// [N, H*r, W*r, C] tensor is the strided [N,H,W,C] view starting at (i, j).
int make_map_nhwc_sub(CUtensorMap* m, bool f32, const void* base, int C, int W, int H, int N, int r, int q,
                      int box_h, int box_w) {
  const uint64_t es = f32 ? 4 : 2;
  const int i = q / r, j = q % r;
  const uint64_t Wf = uint64_t(W) * r, Hf = uint64_t(H) * r;
  const char* b = static_cast<const char*>(base) + (uint64_t(i) * Wf + j) * C * es;
  return make_map_nhwc(m, f32, b, C, W, H, N, uint64_t(r) * C * es, uint64_t(r) * Wf * C * es, Hf * Wf * C * es,
                       box_h, box_w);
}

// packed weights [9][rows][k] bf16 -> 3-D map {k, rows, 9}, box {64, bn, taps} (9 taps resident, 3 streaming)
int make_map_weights(CUtensorMap* m, const void* base, int k, int rows, int bn, int box_taps) {
  const cuuint64_t dims[3] = {cuuint64_t(k), cuuint64_t(rows), 9};
  const cuuint64_t strides[2] = {cuuint64_t(k) * 2, cuuint64_t(k) * 2 * rows};
  const cuuint32_t box[3] = {64, cuuint32_t(bn), cuuint32_t(box_taps)};
  const cuuint32_t estr[3] = {1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(RUMPY_ERR_ARG, "weights not 16B aligned");
  CUresult r = get_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RUMPY_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", int(r));
  return RUMPY_OK;
}

int make_map_weight_layers(CUtensorMap* m, const void* base, int n_layers, int box_taps) {
  const cuuint64_t dims[4] = {64, 64, 9, cuuint64_t(n_layers)};
  const cuuint64_t strides[3] = {128, 64 * 128, 9 * 64 * 128};
  const cuuint32_t box[4] = {64, 64, cuuint32_t(box_taps), 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(RUMPY_ERR_ARG, "weights not 16B aligned");
  CUresult r = get_encode_fn()(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(RUMPY_ERR_CUDA, "cuTensorMapEncodeTiled(weight layers) failed: %d", int(r));
  return RUMPY_OK;
}

// ------------------------------------------------------------------ conv launch plan
int conv_plan_build(ConvPlan* p, const ConvDesc& d) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  if (d.N <= 0 || d.H <= 0 || d.W <= 0) return set_error(RUMPY_ERR_ARG, "conv3x3: bad shape %dx%dx%d", d.N, d.H, d.W);
  if (d.Cin % 64 != 0 || d.Cin <= 0) return set_error(RUMPY_ERR_ARG, "conv3x3: Cin=%d must be a multiple of 64", d.Cin);
  if (!d.x || !d.w) return set_error(RUMPY_ERR_ARG, "conv3x3: null input / weights");
  const int rin = d.in_r < 1 ? 1 : d.in_r, rout = d.out_r < 1 ? 1 : d.out_r;
  if (rin > 3 || rout > 3) return set_error(RUMPY_ERR_ARG, "conv3x3: shuffle factor > 3 unsupported");
  ConvArgs& a = p->args;
  memset(p, 0, sizeof(*p));
  const bool thin = d.out_nchw != nullptr;
  int bn;
  if (thin) {
    if (d.cout_real > 16 || d.cout_real < 1) return set_error(RUMPY_ERR_ARG, "tail conv: cout_real=%d", d.cout_real);
    bn = 16;
  } else {
    if (d.Cout % 64 != 0 || d.Cout <= 0)
      return set_error(RUMPY_ERR_ARG, "conv3x3: Cout=%d must be a multiple of 64", d.Cout);
    if (d.Cin % (64 * rin * rin) != 0 || d.Cout % (64 * rout * rout) != 0)
      return set_error(RUMPY_ERR_ARG, "conv3x3: channels/shuffle mismatch");
    bn = 64;
    if (d.force_bn) bn = d.force_bn;
    if (d.Cout % bn != 0) return set_error(RUMPY_ERR_ARG, "conv3x3: Cout %% BN != 0");
  }
  const int cin_chunks = d.Cin / 64;
  const bool resident = size_t(9) * cin_chunks * bn * 128 <= 96 * 1024;
  p->bn = bn;
  p->resident = resident;
  a.N = d.N; a.H = d.H; a.W = d.W;
  a.n_tiles = thin ? 1 : d.Cout / bn;
  a.cin_chunks = cin_chunks;
  a.a_chunks_per_map = cin_chunks / (rin * rin);
  a.o_chunks_per_map = thin ? 1 : (d.Cout / 64) / (rout * rout);
  a.cout = thin ? 16 : d.Cout;
  a.alpha = d.alpha;
  a.ch_scale = d.ch_scale;
  a.bf16_scale = d.bf16_scale;
  a.bias = d.bias;
  a.pool_partial = d.pool_partial;
  a.out_nchw = d.out_nchw;
  a.cout_real = d.cout_real;
  a.dbg = opt().timeline;
  a.dbg_mode = opt().conv_dbg;
  uint32_t flags = d.flags & (kConvRelu | kConvPool);
  if (d.y_bf16) flags |= kConvOutBf16;
  if (d.y_f32) flags |= kConvOutF32;
  if (d.residual) flags |= kConvResF32;
  if (d.mask) flags |= kConvMask;
  if (!thin && !(flags & (kConvOutBf16 | kConvOutF32))) return set_error(RUMPY_ERR_ARG, "conv3x3: no output");
  if ((flags & kConvPool) && !d.pool_partial) return set_error(RUMPY_ERR_ARG, "conv3x3: POOL needs pool_partial");
  if ((flags & kConvPool) && (d.Cin != d.Cout || resident != (size_t(9) * cin_chunks * 64 * 128 <= 96 * 1024)))
    return set_error(RUMPY_ERR_ARG, "conv3x3: POOL needs Cin == Cout and the default BN (pool rows: conv_pool_rows)");
  if (rout > 1 && (flags & (kConvOutF32 | kConvResF32 | kConvMask | kConvPool)))
    return set_error(RUMPY_ERR_ARG, "conv3x3: shuffle store supports bf16 output only");
  a.flags = flags;
  // pipeline depth / staging slots from the 227 KB budget: two staging slots when >= 4 stages still fit
  const size_t budget = kConvSmemBudget;
  const bool has_in = (flags & (kConvResF32 | kConvMask)) != 0;
  int bufs = 2, stages = kMaxStages;
  auto fit = [&](int b) {
    int st = kMaxStages;
    while (st > 2 && conv_smem_bytes(bn, resident, cin_chunks, st, flags, b) > budget) --st;
    return st;
  };
  stages = fit(2);
  if (conv_smem_bytes(bn, resident, cin_chunks, stages, flags, 2) > budget || (has_in && stages < 4)) {
    bufs = has_in ? 1 : 2;
    stages = fit(bufs);
  }
  if (conv_smem_bytes(bn, resident, cin_chunks, stages, flags, bufs) > budget)
    return set_error(RUMPY_ERR_ARG, "conv3x3: configuration does not fit shared memory");
  a.stages = stages;
  a.stg_bufs = bufs;
  p->smem = conv_smem_bytes(bn, resident, cin_chunks, stages, flags, bufs);
  const int ctas_per_sm = 1;
  // resident weights: 16 x 8 tiles fed by one halo box; streamed weights: 8 x 16 tiles, one box per kx (conv3x3_tc.cuh)
  const int th = p->resident ? kTallH : kTileH, tw = p->resident ? kTallW : kTileW;
  a.tiles_x = (d.W + tw - 1) / tw;
  a.tiles_y = (d.H + th - 1) / th;
  a.m_tiles = d.N * a.tiles_x * a.tiles_y;
  int grid = sms * ctas_per_sm < a.m_tiles * a.n_tiles ? sms * ctas_per_sm : a.m_tiles * a.n_tiles;
  grid -= grid % a.n_tiles;
  if (grid < a.n_tiles) grid = a.n_tiles;
  p->grid = grid;
  // tensor maps
  const int cin_sub = d.Cin / (rin * rin);
  for (int q = 0; q < rin * rin; ++q)
    if (int e = make_map_nhwc_sub(&p->maps.a[q], false, d.x, cin_sub, d.W, d.H, d.N, rin, q,
                                  p->resident ? kTallBoxH : kABoxH, p->resident ? kTallBoxW : kTileW)) return e;
  if (int e = make_map_weights(&p->maps.w, d.w, d.Cin, thin ? 16 : d.Cout, bn, p->resident ? 9 : 3)) return e;
  if (d.y_bf16) {
    const int cout_sub = d.Cout / (rout * rout);
    for (int q = 0; q < rout * rout; ++q)
      if (int e = make_map_nhwc_sub(&p->maps.ob[q], false, d.y_bf16, cout_sub, d.W, d.H, d.N, rout, q, th, tw)) return e;
  }
  if (d.y_f32)
    if (int e = make_map_nhwc_sub(&p->maps.of, true, d.y_f32, d.Cout, d.W, d.H, d.N, 1, 0, th, tw)) return e;
  if (d.residual)
    if (int e = make_map_nhwc_sub(&p->maps.rf, true, d.residual, d.Cout, d.W, d.H, d.N, 1, 0, th, tw)) return e;
  if (d.mask)
    if (int e = make_map_nhwc_sub(&p->maps.mb, false, d.mask, d.Cout, d.W, d.H, d.N, 1, 0, th, tw)) return e;
  return RUMPY_OK;
}

template <int BN, bool RES>
static int launch_conv_t(const ConvPlan& p, cudaStream_t s) {
  auto kern = conv3x3_tc_kernel<BN, RES>;
  static bool attr_set = false;
  static int max_dyn = 0;
  if (!attr_set) {
    // dynamic + this instantiation's static shared memory (barriers, bias, pool sums) must fit the 227 KB of a CTA
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, kern) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "cudaFuncGetAttributes: %s", cudaGetErrorString(cudaGetLastError()));
    max_dyn = 227 * 1024 - int(fa.sharedSizeBytes);
    if (max_dyn > int(kConvSmemBudget)) max_dyn = int(kConvSmemBudget);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, max_dyn) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  if (int(p.smem) > max_dyn)
    return set_error(RUMPY_ERR_ARG, "conv3x3: %zu B of dynamic shared memory do not fit next to the kernel's static part",
                     p.smem);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.grid); cfg.blockDim = dim3(conv_threads(RES)); cfg.dynamicSmemBytes = p.smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = opt().use_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, p.maps, p.args);
  if (e != cudaSuccess) return set_error(RUMPY_ERR_CUDA, "conv3x3 launch: %s", cudaGetErrorString(e));
  return RUMPY_OK;
}

int conv_plan_launch(const ConvPlan& p, cudaStream_t s) {
  switch (p.bn) {
    case 16: return p.resident ? launch_conv_t<16, true>(p, s) : launch_conv_t<16, false>(p, s);
    case 64: return p.resident ? launch_conv_t<64, true>(p, s) : launch_conv_t<64, false>(p, s);
    case 128: return p.resident ? launch_conv_t<128, true>(p, s) : launch_conv_t<128, false>(p, s);
    case 256: return p.resident ? launch_conv_t<256, true>(p, s) : launch_conv_t<256, false>(p, s);
  }
  return set_error(RUMPY_ERR_ARG, "conv3x3: unsupported BN %d", p.bn);
}

static_assert(sizeof(PackJobHost) == sizeof(PackJob), "PackJobHost / PackJob layout mismatch");
int pack_batched_launch(const PackJobHost* jobs_dev, int njobs, cudaStream_t s) {
  if (njobs <= 0) return RUMPY_OK;
  pack_conv3x3_batched_kernel<<<dim3(8, njobs), 256, 0, s>>>(reinterpret_cast<const PackJob*>(jobs_dev));
  return check_launch("pack_conv3x3_batched");
}

static_assert(sizeof(QScaleJobHost) == sizeof(QScaleJobDev), "QScaleJobHost layout mismatch");
int q_scale_launch(const QScaleJobHost* jobs_dev, int njobs, const float* meta, int N, int M, int hidden, int C,
                   int modulate, int relu, cudaStream_t s) {
  if (njobs <= 0) return RUMPY_OK;
  q_scale_kernel<<<dim3(njobs, N), C, size_t(M + hidden) * sizeof(float), s>>>(
      reinterpret_cast<const QScaleJobDev*>(jobs_dev), meta, M, hidden, modulate, relu);
  return check_launch("q_scale");
}

bool conv_ca_supported(int N, int H, int W, int Cin, int Cout) {
  int sms = 0;
  if (device_info(&sms)) return false;
  const int m_tiles = N * ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  return Cin == 64 && Cout == 64 && m_tiles <= kCaMaxTiles * sms;
}

int conv_ca_plan_build(ConvPlan* p, CaFusedArgs* ca, const ConvDesc& d, float* u_store, int Cr,
                       unsigned long long* grid_bar, float* save_mean, float* save_hid, float* save_y) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  if (!conv_ca_supported(d.N, d.H, d.W, d.Cin, d.Cout)) return set_error(RUMPY_ERR_ARG, "conv_ca: unsupported shape");
  if (!d.x || !d.w || !d.residual || !d.y_f32 || !d.y_bf16 || !d.pool_partial || !grid_bar)
    return set_error(RUMPY_ERR_ARG, "conv_ca: null pointer");
  if (Cr < 1 || Cr > 16) return set_error(RUMPY_ERR_ARG, "conv_ca: Cr=%d", Cr);
  memset(p, 0, sizeof(*p));
  memset(ca, 0, sizeof(*ca));
  ConvArgs& a = p->args;
  a.N = d.N; a.H = d.H; a.W = d.W;
  a.tiles_x = (d.W + kTileW - 1) / kTileW;
  a.tiles_y = (d.H + kTileH - 1) / kTileH;
  a.m_tiles = d.N * a.tiles_x * a.tiles_y;
  a.n_tiles = 1; a.cin_chunks = 1; a.a_chunks_per_map = 1; a.o_chunks_per_map = 1; a.cout = 64;
  a.alpha = 1.f; a.bias = d.bias; a.pool_partial = d.pool_partial; a.dbg = opt().timeline;
  a.stages = 4; a.stg_bufs = 2;
  p->bn = 64; p->resident = true;
  p->smem = conv_ca_smem_bytes(a.stages);
  const int T = (a.m_tiles + sms - 1) / sms;
  p->grid = (a.m_tiles + T - 1) / T;
  ca->tiles_per_cta = T;
  ca->cr = Cr; ca->hw = d.H * d.W; ca->partials_per_img = 2 * a.tiles_x * a.tiles_y;
  ca->store_u = u_store != nullptr;
  ca->grid_bar = grid_bar;
  ca->save_mean = save_mean; ca->save_hid = save_hid; ca->save_y = save_y;
  if (int e = make_map_nhwc_sub(&p->maps.a[0], false, d.x, 64, d.W, d.H, d.N, 1, 0, kABoxH)) return e;
  if (int e = make_map_weights(&p->maps.w, d.w, 64, 64, 64, 9)) return e;
  if (int e = make_map_nhwc_sub(&p->maps.rf, true, d.residual, 64, d.W, d.H, d.N, 1, 0, kTileH)) return e;
  if (int e = make_map_nhwc_sub(&p->maps.of, true, d.y_f32, 64, d.W, d.H, d.N, 1, 0, kTileH)) return e;
  if (int e = make_map_nhwc_sub(&p->maps.ob[0], false, d.y_bf16, 64, d.W, d.H, d.N, 1, 0, kTileH)) return e;
  if (u_store)
    if (int e = make_map_nhwc_sub(&p->maps.mb, true, u_store, 64, d.W, d.H, d.N, 1, 0, kTileH)) return e;
  return RUMPY_OK;
}

int conv_ca_launch(const ConvPlan& p, const CaFusedArgs& ca, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(conv3x3_ca_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kConvSmemBudget)) !=
        cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "conv_ca cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.grid); cfg.blockDim = dim3(kConvCaThreads); cfg.dynamicSmemBytes = p.smem; cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = opt().use_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv3x3_ca_kernel, p.maps, p.args, ca);
  if (e != cudaSuccess) return set_error(RUMPY_ERR_CUDA, "conv_ca launch: %s", cudaGetErrorString(e));
  return RUMPY_OK;
}

int grid_for(size_t work_items, int block, int per_sm = 8) {
  int sms = 148;
  device_info(&sms);
  size_t blocks = (work_items + block - 1) / block;
  size_t cap = size_t(sms) * per_sm;
  return int(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace rb

using namespace rb;

extern "C" {

int rumpy_version(void) { return RUMPY_B200_VERSION; }
/* debug hook, not part of the public header: per-CTA clock64 timeline (16 slots per CTA) for conv kernels */
const char* rumpy_last_error(void) { return g_last_error.c_str(); }
int rumpy_device_check(void) { return device_info(nullptr); }

int rumpy_pack_conv3x3(const float* w, void* p, int cout, int cin, int rows_padded, int r, int dgrad, void* stream) {
  if (int e = device_info(nullptr)) return e;
  if (!w || !p || cout <= 0 || cin <= 0) return set_error(RUMPY_ERR_ARG, "pack_conv3x3: bad args");
  if (r < 1) r = 1;
  if (rows_padded < cout) rows_padded = cout;
  if (cout % (r * r) != 0) return set_error(RUMPY_ERR_ARG, "pack_conv3x3: cout %% r^2 != 0");
  const size_t total = size_t(9) * (dgrad ? size_t(cin) * cout : size_t(rows_padded) * cin);
  pack_conv3x3_kernel<<<grid_for(total, 256), 256, 0, cudaStream_t(stream)>>>(
      w, static_cast<__nv_bfloat16*>(p), cout, cin, rows_padded, r, dgrad);
  return check_launch("pack_conv3x3");
}

int rumpy_pack_bias(const float* b, float* p, int cout, int rows_padded, int r, void* stream) {
  if (int e = device_info(nullptr)) return e;
  if (!b || !p) return set_error(RUMPY_ERR_ARG, "pack_bias: null");
  if (r < 1) r = 1;
  if (rows_padded < cout) rows_padded = cout;
  pack_bias_kernel<<<grid_for(rows_padded, 128), 128, 0, cudaStream_t(stream)>>>(b, p, cout, rows_padded, r);
  return check_launch("pack_bias");
}

int rumpy_conv3x3(const void* x, const void* w, const float* bias, const float* residual, const void* mask,
                  void* y_bf16, float* y_f32, float* pool_partial, int N, int H, int W, int Cin, int Cout,
                  int in_unshuffle_r, int out_shuffle_r, unsigned flags, float alpha, void* stream) {
  ConvDesc d{};
  d.x = x; d.w = w; d.bias = bias; d.residual = residual; d.mask = mask; d.y_bf16 = y_bf16; d.y_f32 = y_f32;
  d.pool_partial = pool_partial; d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.Cout = Cout; d.in_r = in_unshuffle_r;
  d.out_r = out_shuffle_r; d.flags = flags; d.alpha = alpha;
  ConvPlan p;
  if (int e = conv_plan_build(&p, d)) return e;
  return conv_plan_launch(p, cudaStream_t(stream));
}

int rumpy_conv3x3_tail(const void* x, const void* w, const float* bias16, float* y_nchw, int N, int H, int W,
                       int Cin, int cout_real, void* stream) {
  if (!y_nchw) return set_error(RUMPY_ERR_ARG, "conv3x3_tail: null output");
  ConvDesc d{};
  d.x = x; d.w = w; d.bias = bias16; d.N = N; d.H = H; d.W = W; d.Cin = Cin; d.Cout = 16; d.out_nchw = y_nchw;
  d.cout_real = cout_real; d.alpha = 1.f; d.in_r = 1; d.out_r = 1;
  ConvPlan p;
  if (int e = conv_plan_build(&p, d)) return e;
  return conv_plan_launch(p, cudaStream_t(stream));
}

int rumpy_head_conv(const float* x, const float* w, const float* bias, float* yf, void* yb, int N, int H, int W,
                    int Cin, int C, void* stream) {
  if (int e = device_info(nullptr)) return e;
  if (!x || !w || !bias || !yf || !yb) return set_error(RUMPY_ERR_ARG, "head_conv: null pointer");
  if (Cin < 1 || Cin > 4 || C % 8 != 0 || C > 512) return set_error(RUMPY_ERR_ARG, "head_conv: Cin=%d C=%d", Cin, C);
  const size_t smem = (size_t(Cin) * 9 * C + C) * sizeof(float);
  const size_t items = size_t(N) * H * W * (C / 8);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(head_conv_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  head_conv_kernel<4><<<grid_for(items, 256, 4), 256, smem, cudaStream_t(stream)>>>(
      x, w, bias, yf, static_cast<__nv_bfloat16*>(yb), N, H, W, Cin, C);
  return check_launch("head_conv");
}

}  // extern "C"

namespace rb {
// pool_partial: [N][partials_per_img][C].  compact_scratch (N*64*C floats, may be NULL): when an image has more
// than 256 partial rows they are first compacted to 64 rows per image.
int ca_apply_launch(const float* pool_partial, int partials_per_img, float* compact_scratch, const void* u,
                    int u_is_f32, const float* x_in, const float* w1, const float* b1, const float* w2,
                    const float* b2, float* x_out, void* x_out_bf16, float* save_mean, float* save_hid,
                    float* save_y, int N, int H, int W, int C, int Cr, cudaStream_t stream,
                    const float* q_scale) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  if (!pool_partial || !u || !x_in || !w1 || !b1 || !w2 || !b2 || !x_out || !x_out_bf16)
    return set_error(RUMPY_ERR_ARG, "ca_apply: null pointer");
  if (C % 4 != 0 || C > 256 || Cr < 1 || Cr > 64) return set_error(RUMPY_ERR_ARG, "ca_apply: C=%d Cr=%d", C, Cr);
  if (save_y && (!save_mean || !save_hid)) return set_error(RUMPY_ERR_ARG, "ca_apply: save_* must come together");
  int partials = partials_per_img;
  if (compact_scratch && partials > 256) {
    const int block = (256 / C > 0 ? 256 / C : 1) * C;
    pool_compact_kernel<<<dim3(64, N), block, block * sizeof(float), stream>>>(pool_partial, partials,
                                                                               compact_scratch, C);
    if (int e = check_launch("pool_compact")) return e;
    pool_partial = compact_scratch;
    partials = 64;
  }
  const size_t vec = size_t(H) * W * (C / 4);
  // 4 vectors per thread (all in flight) for small images, 8 once an image is >= 64K vectors: the per-CTA FC
  // prologue is amortised over more streaming work
  const int per_thread = vec >= 60000 ? 8 : 4;
  int chunks = int((vec + 256 * per_thread - 1) / (256 * per_thread));
  const int cap = (sms * 8 + N - 1) / N;
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(chunks, N); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = opt().use_pdl ? 1 : 0;
  const int HW = H * W;
  __nv_bfloat16* xob = static_cast<__nv_bfloat16*>(x_out_bf16);
  cudaError_t le;
  if (u_is_f32)
    le = cudaLaunchKernelEx(&cfg, ca_apply_kernel<true>, pool_partial, partials, u, x_in, w1, b1, w2, b2, x_out, xob,
                            save_mean, save_hid, save_y, HW, C, Cr, q_scale);
  else
    le = cudaLaunchKernelEx(&cfg, ca_apply_kernel<false>, pool_partial, partials, u, x_in, w1, b1, w2, b2, x_out, xob,
                            save_mean, save_hid, save_y, HW, C, Cr, q_scale);
  if (le != cudaSuccess) return set_error(RUMPY_ERR_CUDA, "ca_apply launch: %s", cudaGetErrorString(le));
  return RUMPY_OK;
}
}  // namespace rb

extern "C" {

int rumpy_pool_rows(int H, int W, int C) { return (H > 0 && W > 0 && C >= 64) ? conv_pool_rows(H, W, C) : 0; }

int rumpy_ca_apply(const float* pool_partial, const void* u, int u_is_f32, const float* x_in, const float* w1,
                   const float* b1, const float* w2, const float* b2, float* x_out, void* x_out_bf16,
                   float* save_mean, float* save_hid, float* save_y, int N, int H, int W, int C, int Cr,
                   void* stream) {
  return ca_apply_launch(pool_partial, conv_pool_rows(H, W, C), nullptr, u, u_is_f32, x_in, w1, b1, w2, b2, x_out, x_out_bf16,
                         save_mean, save_hid, save_y, N, H, W, C, Cr, cudaStream_t(stream), nullptr);
}

int rumpy_nchw_to_nhwc(const float* x, float* y_f32, void* y_bf16, int N, int C, int H, int W, void* stream) {
  if (int e = device_info(nullptr)) return e;
  if (!x || (!y_f32 && !y_bf16) || N <= 0 || C <= 0 || H <= 0 || W <= 0)
    return set_error(RUMPY_ERR_ARG, "nchw_to_nhwc: bad args");
  const int P = H * W;
  dim3 grid((P + 31) / 32, (C + 31) / 32, N), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, cudaStream_t(stream)>>>(x, y_f32, static_cast<__nv_bfloat16*>(y_bf16), C, P);
  return check_launch("nchw_to_nhwc");
}

int rumpy_nhwc_to_nchw(const void* x, int x_is_bf16, float* y, int N, int C, int H, int W, void* stream) {
  if (int e = device_info(nullptr)) return e;
  if (!x || !y || N <= 0 || C <= 0 || H <= 0 || W <= 0) return set_error(RUMPY_ERR_ARG, "nhwc_to_nchw: bad args");
  const int P = H * W;
  dim3 grid((P + 31) / 32, (C + 31) / 32, N), block(32, 8);
  if (x_is_bf16) nhwc_to_nchw_kernel<true><<<grid, block, 0, cudaStream_t(stream)>>>(x, y, C, P);
  else nhwc_to_nchw_kernel<false><<<grid, block, 0, cudaStream_t(stream)>>>(x, y, C, P);
  return check_launch("nhwc_to_nchw");
}

int rumpy_pool_sum(const float* x_nhwc, float* pool_partial, int N, int H, int W, int C, void* stream) {
  if (int e = device_info(nullptr)) return e;
  if (!x_nhwc || !pool_partial || C <= 0 || C > 256) return set_error(RUMPY_ERR_ARG, "pool_sum: bad args");
  const int block = (1024 / C) * C;
  pool_sum_kernel<<<N, block, block * sizeof(float), cudaStream_t(stream)>>>(x_nhwc, pool_partial,
                                                                              conv_pool_rows(H, W, C), H * W, C);
  return check_launch("pool_sum");
}

}  // extern "C"
