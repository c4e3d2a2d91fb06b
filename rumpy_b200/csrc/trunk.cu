// Host side of the persistent trunk kernel (trunk_pipe.cuh): plan construction, device tables, launch.
#define RB_TRUNK_KERNEL_IMPL
#include "../../include/rumpy_b200.h"
#include "host_util.cuh"

namespace rb {

constexpr int kClusterMaxDyn = 227 * 1024 - 8192;
constexpr int kBandMaxDyn = 227 * 1024 - 2048;   // the band kernel has < 2 KB of static shared memory

static size_t al(size_t v) { return (v + 1023) / 1024 * 1024; }

bool trunk_supported(int N, int H, int W, int C, int Cr) {
  int sms = 0;
  if (device_info(&sms)) return false;
  const int P = ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  const long long T = (long long)N * P;
  return C == 64 && Cr >= 1 && Cr <= 16 && T <= (long long)kTrunkMaxK * sms && P <= 512;
}

size_t trunk_device_bytes(int N, int H, int W, int n_layers, int n_in_maps, int n_out_maps) {
  const size_t P = size_t((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  const size_t T = size_t(N) * P;
  return al(size_t(n_layers) * sizeof(TrunkLayer)) + al(size_t(n_in_maps) * sizeof(CUtensorMap)) +
         al(size_t(n_out_maps) * sizeof(CUtensorMap)) + al(T * sizeof(int) + 2 * T * 64 * sizeof(unsigned long long));
}

int trunk_plan_finish(TrunkPlan* plan, int N, int H, int W, int Cr, const void* w_base, const float* s_init,
                      void* dev, bool allow_cluster) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  if (!trunk_supported(N, H, W, 64, Cr)) return set_error(RUMPY_ERR_ARG, "trunk: unsupported shape");
  TrunkArgs& a = plan->args;
  memset(&a, 0, sizeof(a));
  a.n_layers = int(plan->layers.size());
  a.N = N; a.H = H; a.W = W;
  a.tiles_x = (W + kTileW - 1) / kTileW;
  a.tiles_y = (H + kTileH - 1) / kTileH;
  a.tiles_per_img = a.tiles_x * a.tiles_y;
  a.T = N * a.tiles_per_img;
  a.K = (a.T + sms - 1) / sms;
  plan->grid = (a.T + a.K - 1) / a.K;
  // whole images per slot row => every image lives in one tile slot => pool / apply may be interleaved per tile
  if (a.K > 1 && a.tiles_per_img <= sms) {
    const int g = (sms / a.tiles_per_img) * a.tiles_per_img;
    if ((a.T + g - 1) / g <= a.K) plan->grid = g;
  }
  // pool(j), apply(j) back to back is legal when every image lives in one tile slot (grid % tiles_per_img == 0),
  // but measured slower than all-pools-then-applies at 16x64x64 (5.10 vs 4.93 ms): the group stalls in apply(j)
  // while the next tile's accumulator is already waiting.  Kept as a kernel option, off.
  a.interleave = 0;
  a.w_layer0 = 0;
  a.cr = Cr;
  a.inv_hw = 1.f / float(H * W);
  a.s_init = s_init;
  char* p = static_cast<char*>(dev);
  plan->layers_dev = reinterpret_cast<TrunkLayer*>(p); p += al(plan->layers.size() * sizeof(TrunkLayer));
  plan->in_maps_dev = reinterpret_cast<CUtensorMap*>(p); p += al(plan->in_bufs.size() * sizeof(CUtensorMap));
  plan->out_maps_dev = reinterpret_cast<CUtensorMap*>(p); p += al(plan->out_bufs.size() * sizeof(CUtensorMap));
  plan->flags_dev = p;   // [T] tile epochs, then the tagged pool partials: one memset per launch clears both
  const size_t ready_bytes = (size_t(a.T) * sizeof(int) + 15) / 16 * 16;
  plan->flags_bytes = ready_bytes + size_t(2) * a.T * 64 * sizeof(unsigned long long);
  a.ready = reinterpret_cast<int*>(p);
  a.pool_partial = reinterpret_cast<unsigned long long*>(p + ready_bytes);
  a.layers = plan->layers_dev;
  a.in_maps = plan->in_maps_dev;
  a.out_maps = plan->out_maps_dev;
  if (int e = make_map_weight_layers(&plan->w_map, w_base, a.n_layers)) return e;
  // ---- one cluster per image?  pick the decomposition with the fewest tiles per CTA that keeps all N clusters
  // co-resident (clusters are independent, so this is a performance condition, not a correctness one)
  plan->cluster = false;
  if (opt().use_cluster && allow_cluster) {
    static bool attr_set = false;
    if (!attr_set) {
      // dynamic + static (3.2 KB with 2 epilogue groups, 5.8 KB with 4) must stay within 227 KB
      if (cudaFuncSetAttribute(trunk_cluster_kernel_t<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterMaxDyn) != cudaSuccess ||
          cudaFuncSetAttribute(trunk_cluster_kernel_t<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterMaxDyn) != cudaSuccess)
        return set_error(RUMPY_ERR_CUDA, "trunk_cluster cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
      attr_set = true;
    }
    int best_tiles = 99, best_perim = 1 << 30;
    for (int th = 1; th <= 4; ++th)
      for (int tw = 1; th * tw <= kTrunkMaxK; ++tw) {
        const int cy = (H + kClusterTileH * th - 1) / (kClusterTileH * th);
        const int cx = (W + kClusterTileW * tw - 1) / (kClusterTileW * tw);
        const int C = cx * cy;
        if (C > 8) continue;
        const size_t smem = cluster_smem_bytes(th, tw, C);
        if (smem > size_t(kClusterMaxDyn)) continue;
        // vertical strips (one tile wide, the whole image height) allow the two-phase hand-over (trunk_cluster.cuh
        // `split`): preferred among decompositions with the same tile count, then the smaller perimeter
        const bool can_split = opt().cluster_split && tw == 1 && cy == 1 && th >= 3;
        const int perim = kClusterTileH * th + kClusterTileW * tw - (can_split ? 1000 : 0);
        if (th * tw > best_tiles || (th * tw == best_tiles && perim >= best_perim)) continue;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(N * C); cfg.blockDim = dim3(cluster_threads(opt().cluster_groups)); cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int active = 0;
        if ((opt().cluster_groups == 4 ? cudaOccupancyMaxActiveClusters(&active, trunk_cluster_kernel_t<4>, &cfg)
                                   : cudaOccupancyMaxActiveClusters(&active, trunk_cluster_kernel_t<2>, &cfg)) != cudaSuccess) {
          (void)cudaGetLastError();
          continue;
        }
        if (active < N) continue;
        best_tiles = th * tw; best_perim = perim;
        plan->cluster = true;
        plan->cluster_size = C;
        plan->cluster_groups = opt().cluster_groups;
        plan->cluster_smem = smem;
        ClusterArgs& c = plan->cargs;
        memset(&c, 0, sizeof(c));
        c.layers = plan->layers_dev;
        c.s_init = s_init;
        c.x_init = static_cast<const __nv_bfloat16*>(plan->in_bufs.front());
        c.out_bf16 = static_cast<__nv_bfloat16*>(plan->out_bufs.back());
        c.n_layers = a.n_layers; c.N = N; c.H = H; c.W = W; c.cx = cx; c.cy = cy; c.th = th; c.tw = tw; c.cr = Cr;
        c.inv_hw = a.inv_hw;
        c.split = can_split ? th - 1 : 0;
        c.dbg_flags = opt().cluster_dbg;
        for (const TrunkLayer& l : plan->layers) c.n_ca += l.kind == kTrunkCA ? 1 : 0;
      }
  }
  // ---- role-swapped band kernel: one cluster of row bands per image, weights in tensor memory.  Picks the
  // largest cluster (most SMs) whose bands fit (<= 3 chunks of <= 143 linear pixels, shared memory) with all N
  // clusters co-resident.
  plan->band = false;
  if (opt().use_band && allow_cluster && W + 1 >= 4) {
    static bool attr_set = false;
    if (!attr_set) {
      if (cudaFuncSetAttribute(trunk_band_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBandMaxDyn) != cudaSuccess)
        return set_error(RUMPY_ERR_CUDA, "trunk_band cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
      attr_set = true;
    }
    const int P = W + 1;
    for (int C = kBandMaxCluster; C >= 1 && !plan->band; --C) {
      const int R = (H + C - 1) / C;
      if ((C - 1) * R >= H) continue;                       // every CTA owns at least one row
      const int count = R * P - 1;                          // linear output pixels of a full band
      const int n_chunks = (count + kBandNQMax - 1) / kBandNQMax;
      if (n_chunks > kBandMaxChunks) continue;
      const size_t smem = band_smem_bytes(R, P, C);
      if (smem > size_t(kBandMaxDyn)) continue;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(N * C); cfg.blockDim = dim3(kBandThreads); cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int active = 0;
      if (cudaOccupancyMaxActiveClusters(&active, trunk_band_kernel, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        continue;
      }
      if (active < N) continue;
      plan->band = true;
      plan->band_smem = smem;
      BandArgs& b = plan->bargs;
      memset(&b, 0, sizeof(b));
      b.layers = plan->layers_dev;
      b.s_init = s_init;
      b.out_bf16 = static_cast<__nv_bfloat16*>(plan->out_bufs.back());
      b.n_layers = a.n_layers; b.N = N; b.H = H; b.W = W; b.C = C; b.R = R; b.P = P; b.n_chunks = n_chunks; b.cr = Cr;
      b.inv_hw = a.inv_hw;
      b.plane_bytes = band_plane_bytes(R, P);
      b.pmagic = ((1u << 24) + uint32_t(P) - 1) / uint32_t(P);
      int left = count;
      for (int ch = 0; ch < n_chunks; ++ch) {
        b.nq[ch] = left < kBandNQMax ? left : kBandNQMax;
        b.nn[ch] = (b.nq[ch] + 1 + 15) / 16 * 16;
        left -= b.nq[ch];
      }
      for (const TrunkLayer& l : plan->layers) b.n_ca += l.kind == kTrunkCA ? 1 : 0;
      if (int e = make_map_weight_layers(&plan->w_tap_map, w_base, a.n_layers, 1)) return e;
    }
  }
  plan->uploaded.clear();
  plan->maps_uploaded = false;
  return RUMPY_OK;
}

int trunk_launch(TrunkPlan* plan, const float* const* params, cudaStream_t s) {
  TrunkArgs& a = plan->args;
  if (!plan->maps_uploaded) {
    std::vector<CUtensorMap> im(plan->in_bufs.size()), om(plan->out_bufs.size());
    for (size_t i = 0; i < im.size(); ++i)
      if (int e = make_map_nhwc_sub(&im[i], false, plan->in_bufs[i], 64, a.W, a.H, a.N, 1, 0, kABoxH)) return e;
    for (size_t i = 0; i < om.size(); ++i)
      if (int e = make_map_nhwc_sub(&om[i], false, plan->out_bufs[i], 64, a.W, a.H, a.N, 1, 0, kTileH)) return e;
    if (cudaMemcpyAsync(plan->in_maps_dev, im.data(), im.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s) !=
            cudaSuccess ||
        cudaMemcpyAsync(plan->out_maps_dev, om.data(), om.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s) !=
            cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk: map upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);   // `im` / `om` die at scope exit; once per plan
    plan->maps_uploaded = true;
  }
  for (size_t i = 0; i < plan->layers.size(); ++i) {
    TrunkLayer& l = plan->layers[i];
    const TrunkLayerParams& lp = plan->lparams[i];
    l.bias = params[lp.bias];
    if (lp.w1 >= 0) { l.w1 = params[lp.w1]; l.b1 = params[lp.b1]; l.w2 = params[lp.w2]; l.b2 = params[lp.b2]; }
  }
  if (plan->uploaded.size() != plan->layers.size() ||
      memcmp(plan->uploaded.data(), plan->layers.data(), plan->layers.size() * sizeof(TrunkLayer)) != 0) {
    plan->uploaded = plan->layers;   // persistent host copy: the async copy may read it after we return
    if (cudaMemcpyAsync(plan->layers_dev, plan->uploaded.data(), plan->uploaded.size() * sizeof(TrunkLayer),
                        cudaMemcpyHostToDevice, s) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk: layer table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);
  }
  if (plan->band) {
    BandArgs& b = plan->bargs;
    b.dbg = opt().trunk_dbg_layers > 0 ? opt().timeline : nullptr;
    b.dbg_layers = opt().trunk_dbg_layers;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(b.N * b.C); cfg.blockDim = dim3(kBandThreads);
    cfg.dynamicSmemBytes = plan->band_smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = b.C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (opt().trunk_ev0) cudaEventRecord(opt().trunk_ev0, s);
    cudaError_t e = cudaLaunchKernelEx(&cfg, trunk_band_kernel, plan->w_tap_map, b);
    if (e != cudaSuccess) return set_error(RUMPY_ERR_CUDA, "trunk_band launch: %s", cudaGetErrorString(e));
    if (opt().trunk_ev1) cudaEventRecord(opt().trunk_ev1, s);
    return RUMPY_OK;
  }
  if (plan->cluster) {
    ClusterArgs& c = plan->cargs;
    c.dbg = opt().trunk_dbg_layers > 0 ? opt().timeline : nullptr;
    c.dbg_layers = opt().trunk_dbg_layers;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c.N * plan->cluster_size); cfg.blockDim = dim3(cluster_threads(plan->cluster_groups));
    cfg.dynamicSmemBytes = plan->cluster_smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = plan->cluster_size; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    if (opt().trunk_ev0) cudaEventRecord(opt().trunk_ev0, s);
    cudaError_t e = plan->cluster_groups == 4 ? cudaLaunchKernelEx(&cfg, trunk_cluster_kernel_t<4>, plan->w_map, c)
                                              : cudaLaunchKernelEx(&cfg, trunk_cluster_kernel_t<2>, plan->w_map, c);
    if (e != cudaSuccess) return set_error(RUMPY_ERR_CUDA, "trunk_cluster launch: %s", cudaGetErrorString(e));
    if (opt().trunk_ev1) cudaEventRecord(opt().trunk_ev1, s);
    return RUMPY_OK;
  }
  if (cudaMemsetAsync(plan->flags_dev, 0, plan->flags_bytes, s) != cudaSuccess)
    return set_error(RUMPY_ERR_CUDA, "trunk: flag reset failed: %s", cudaGetErrorString(cudaGetLastError()));
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(trunk_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTrunkSmemBytes)) !=
        cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  a.dbg = opt().trunk_dbg_layers > 0 ? opt().timeline : nullptr;
  a.dbg_layers = opt().trunk_dbg_layers;
  a.sync_mode = opt().trunk_sync_mode;
  if (opt().trunk_ev0) cudaEventRecord(opt().trunk_ev0, s);
  trunk_pipe_kernel<<<plan->grid, kTrunkThreads, kTrunkSmemBytes, s>>>(plan->w_map, a);
  if (int e = check_launch("trunk_pipe")) return e;
  if (opt().trunk_ev1) cudaEventRecord(opt().trunk_ev1, s);
  return RUMPY_OK;
}

// ------------------------------------------------------------------ backward program (trunk_bwd.cuh)
size_t trunk_bwd_device_bytes(int N, int H, int W, int n_layers, int n_in_maps, int n_out_maps, int n_ca) {
  const size_t P = size_t((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  const size_t T = size_t(N) * P;
  return al(size_t(n_layers) * sizeof(TrunkBwdLayer)) + al(size_t(n_in_maps) * sizeof(CUtensorMap)) +
         al(size_t(n_out_maps) * sizeof(CUtensorMap)) + al(T * sizeof(int) + 2 * T * 64 * sizeof(unsigned long long)) +
         al(size_t(n_ca) * sizeof(CaPgJobHost));
}

int trunk_bwd_plan_finish(TrunkBwdPlan* plan, int N, int H, int W, int Cr, const void* w_base, int n_w_layers,
                          void* dev) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  if (!trunk_supported(N, H, W, 64, Cr)) return set_error(RUMPY_ERR_ARG, "trunk_bwd: unsupported shape");
  TrunkBwdArgs& a = plan->args;
  memset(&a, 0, sizeof(a));
  a.n_layers = int(plan->layers.size());
  a.N = N; a.H = H; a.W = W;
  a.tiles_x = (W + kTileW - 1) / kTileW;
  a.tiles_y = (H + kTileH - 1) / kTileH;
  a.tiles_per_img = a.tiles_x * a.tiles_y;
  a.T = N * a.tiles_per_img;
  a.K = (a.T + sms - 1) / sms;
  plan->grid = (a.T + a.K - 1) / a.K;
  if (a.K > 1 && a.tiles_per_img <= sms) {
    const int g = (sms / a.tiles_per_img) * a.tiles_per_img;
    if ((a.T + g - 1) / g <= a.K) plan->grid = g;
  }
  a.cr = Cr;
  a.inv_hw = 1.f / float(H * W);
  char* p = static_cast<char*>(dev);
  plan->layers_dev = reinterpret_cast<TrunkBwdLayer*>(p); p += al(plan->layers.size() * sizeof(TrunkBwdLayer));
  plan->in_maps_dev = reinterpret_cast<CUtensorMap*>(p); p += al(plan->in_bufs.size() * sizeof(CUtensorMap));
  plan->out_maps_dev = reinterpret_cast<CUtensorMap*>(p); p += al(plan->out_bufs.size() * sizeof(CUtensorMap));
  plan->flags_dev = p;
  const size_t ready_bytes = (size_t(a.T) * sizeof(int) + 15) / 16 * 16;
  plan->flags_bytes = ready_bytes + size_t(2) * a.T * 64 * sizeof(unsigned long long);
  a.ready = reinterpret_cast<int*>(p);
  a.pool_partial = reinterpret_cast<unsigned long long*>(p + ready_bytes);
  p += al(size_t(a.T) * sizeof(int) + 2 * size_t(a.T) * 64 * sizeof(unsigned long long));
  plan->pg_jobs_dev = reinterpret_cast<CaPgJobHost*>(p);
  a.layers = plan->layers_dev;
  a.in_maps = plan->in_maps_dev;
  a.out_maps = plan->out_maps_dev;
  if (int e = make_map_weight_layers(&plan->w_map, w_base, n_w_layers)) return e;
  plan->uploaded.clear();
  plan->pg_jobs_uploaded.clear();
  plan->maps_uploaded = false;
  return RUMPY_OK;
}

static_assert(sizeof(CaPgJobHost) == sizeof(CaPgJob), "CaPgJobHost / CaPgJob layout mismatch");

int trunk_bwd_launch(TrunkBwdPlan* plan, const float* const* params, float* const* grads, cudaStream_t s) {
  TrunkBwdArgs& a = plan->args;
  if (!plan->maps_uploaded) {
    std::vector<CUtensorMap> im(plan->in_bufs.size()), om(plan->out_bufs.size());
    for (size_t i = 0; i < im.size(); ++i)
      if (int e = make_map_nhwc_sub(&im[i], false, plan->in_bufs[i], 64, a.W, a.H, a.N, 1, 0, kABoxH)) return e;
    for (size_t i = 0; i < om.size(); ++i)
      if (int e = make_map_nhwc_sub(&om[i], false, plan->out_bufs[i], 64, a.W, a.H, a.N, 1, 0, kTileH)) return e;
    if (cudaMemcpyAsync(plan->in_maps_dev, im.data(), im.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s) !=
            cudaSuccess ||
        cudaMemcpyAsync(plan->out_maps_dev, om.data(), om.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice, s) !=
            cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk_bwd: map upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);
    plan->maps_uploaded = true;
  }
  for (size_t i = 0; i < plan->layers.size(); ++i) {
    const TrunkBwdLayerParams& lp = plan->lparams[i];
    if (lp.w1 >= 0) { plan->layers[i].w1 = params[lp.w1]; plan->layers[i].w2 = params[lp.w2]; }
  }
  if (plan->uploaded.size() != plan->layers.size() ||
      memcmp(plan->uploaded.data(), plan->layers.data(), plan->layers.size() * sizeof(TrunkBwdLayer)) != 0) {
    plan->uploaded = plan->layers;
    if (cudaMemcpyAsync(plan->layers_dev, plan->uploaded.data(), plan->uploaded.size() * sizeof(TrunkBwdLayer),
                        cudaMemcpyHostToDevice, s) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk_bwd: layer table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);
  }
  std::vector<CaPgJobHost> jobs;
  for (const TrunkBwdPgBind& b : plan->pg_binds)
    jobs.push_back(CaPgJobHost{plan->layers[b.layer].pg, grads[b.dw1], grads[b.db1], grads[b.dw2], grads[b.db2]});
  if (!jobs.empty() && (jobs.size() != plan->pg_jobs_uploaded.size() ||
                        memcmp(jobs.data(), plan->pg_jobs_uploaded.data(), jobs.size() * sizeof(CaPgJobHost)) != 0)) {
    plan->pg_jobs_uploaded = jobs;
    if (cudaMemcpyAsync(plan->pg_jobs_dev, plan->pg_jobs_uploaded.data(), jobs.size() * sizeof(CaPgJobHost),
                        cudaMemcpyHostToDevice, s) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk_bwd: job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    cudaStreamSynchronize(s);
  }
  if (cudaMemsetAsync(plan->flags_dev, 0, plan->flags_bytes, s) != cudaSuccess)
    return set_error(RUMPY_ERR_CUDA, "trunk_bwd: flag reset failed: %s", cudaGetErrorString(cudaGetLastError()));
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(trunk_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTrunkSmemBytes)) !=
        cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "trunk_bwd cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  trunk_bwd_kernel<<<plan->grid, kTrunkThreads, kTrunkSmemBytes, s>>>(plan->w_map, a);
  if (int e = check_launch("trunk_bwd")) return e;
  if (!jobs.empty()) {
    ca_pg_finalize_kernel<<<int(jobs.size()), 256, 0, s>>>(reinterpret_cast<const CaPgJob*>(plan->pg_jobs_dev), a.N, a.cr);
    if (int e = check_launch("ca_pg_finalize")) return e;
  }
  return RUMPY_OK;
}

static_assert(sizeof(QGradJobHost) == sizeof(QGradJob), "QGradJobHost layout mismatch");
int q_grad_launch(const QGradJobHost* jobs_dev, int njobs, const float* meta, int N, int M, int hidden, int C, int relu,
                  int dq_slices, cudaStream_t s) {
  if (njobs <= 0) return RUMPY_OK;
  const size_t smem = size_t(N) * (M + 2 * hidden + C) * sizeof(float);
  if (smem > 48 * 1024)
    return set_error(RUMPY_ERR_ARG, "q_grad: batch %d x (metadata %d + hidden %d) does not fit 48 KB of shared memory", N, M,
                     hidden);
  q_grad_kernel<<<njobs, 256, smem, s>>>(reinterpret_cast<const QGradJob*>(jobs_dev), meta, N, M, hidden, C, relu,
                                         dq_slices);
  return check_launch("q_grad");
}

int dq_reduce_launch(const float* g_f32, const void* out_bf16, const void* x_bf16, float* partial, int N, int HW, int C,
                     cudaStream_t s) {
  if (C < 1 || C > 256) return set_error(RUMPY_ERR_ARG, "dq_reduce: C=%d", C);
  const int threads = (256 / C) * C;
  dq_reduce_kernel<<<dim3(kDqSlices, N), threads, threads * sizeof(float), s>>>(
      g_f32, static_cast<const __nv_bfloat16*>(out_bf16), static_cast<const __nv_bfloat16*>(x_bf16), partial, HW, C);
  return check_launch("dq_reduce");
}

}  // namespace rb
