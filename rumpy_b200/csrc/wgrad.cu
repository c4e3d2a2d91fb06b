// Host side of the tensor-core weight gradient: job lists (TMA descriptors in device memory) + launches.
#include "../../include/rumpy_b200.h"
#include "host_util.cuh"
#define RB_WGRAD_KERNELS_IMPL
#include "wgrad_tc.cuh"
#include <vector>

namespace rb {

static size_t wg_align(size_t v) { return (v + 255) / 256 * 256; }

// Layout of the caller-provided workspace for one wgrad call.
struct WgradLayout {
  int blocks_co, blocks_ci, splits, m_tiles, njobs;
  size_t off_jobs, off_rjobs, off_partials, bytes;
};

WgradLayout wgrad_layout(int N, int H, int W, int Cin, int Cout, int sms, int splits_override) {
  WgradLayout L{};
  L.blocks_co = Cout / 64; L.blocks_ci = Cin / 64;
  const int tiles = ((W + kTileW - 1) / kTileW) * ((H + kTileH - 1) / kTileH);
  L.m_tiles = N * tiles;
  const int blocks = L.blocks_co * L.blocks_ci;
  int splits = splits_override > 0 ? splits_override : (sms + blocks - 1) / blocks;
  if (splits > L.m_tiles) splits = L.m_tiles;
  if (splits < 1) splits = 1;
  L.splits = splits;
  L.njobs = blocks * splits;
  L.off_jobs = 0;
  L.off_rjobs = wg_align(size_t(L.njobs) * sizeof(WgradJob));
  L.off_partials = L.off_rjobs + wg_align(size_t(blocks) * sizeof(WgradReduceJob));
  L.bytes = L.off_partials + size_t(L.njobs) * 9 * 64 * 64 * sizeof(float);
  return L;
}

// Fills host-side job arrays for one conv layer.  g: [N,H*r,W*r,Cout/r^2] bf16 (r = pixel-unshuffle factor of
// the gradient operand), x: [N,H,W,Cin] bf16, partial/jobs addresses are DEVICE addresses inside `ws`.
int wgrad_build_jobs(std::vector<WgradJob>& jobs, std::vector<WgradReduceJob>& rjobs, const WgradLayout& L,
                     const void* g, const void* x, float* dw, char* ws_dev, int N, int H, int W, int Cin, int Cout,
                     int r, float alpha, int accumulate) {
  const int rr = r * r;
  if (Cout % (64 * rr) != 0 || Cin % 64 != 0) return set_error(RUMPY_ERR_ARG, "wgrad: channels must be multiples of 64");
  const int cout_sub = Cout / rr, chunks_per_q = cout_sub / 64;
  const int tiles_x = (W + kTileW - 1) / kTileW, tiles_y = (H + kTileH - 1) / kTileH;
  std::vector<CUtensorMap> gmaps(rr);
  for (int q = 0; q < rr; ++q)
    if (int e = make_map_nhwc_sub(&gmaps[q], false, g, cout_sub, W, H, N, r, q, kABoxH)) return e;
  CUtensorMap xmap;
  if (int e = make_map_nhwc_sub(&xmap, false, x, Cin, W, H, N, 1, 0, kTileH)) return e;
  float* partials = reinterpret_cast<float*>(ws_dev + L.off_partials);
  for (int cb = 0; cb < L.blocks_co; ++cb) {
    for (int ib = 0; ib < L.blocks_ci; ++ib) {
      const int block = cb * L.blocks_ci + ib;
      float* pbase = partials + size_t(block) * L.splits * 9 * 64 * 64;
      for (int s = 0; s < L.splits; ++s) {
        WgradJob j{};
        j.g = gmaps[cb / chunks_per_q];
        j.x = xmap;
        j.gc0 = (cb % chunks_per_q) * 64;
        j.xc0 = ib * 64;
        j.tile_begin = int((long long)L.m_tiles * s / L.splits);
        j.tile_end = int((long long)L.m_tiles * (s + 1) / L.splits);
        j.tiles_x = tiles_x; j.tiles_y = tiles_y;
        j.out = pbase + size_t(s) * 9 * 64 * 64;
        jobs.push_back(j);
      }
      WgradReduceJob rj{};
      rj.partial = pbase; rj.dw = dw; rj.splits = L.splits; rj.cout = Cout; rj.cin = Cin; rj.co0 = cb * 64;
      rj.ci0 = ib * 64; rj.r = r; rj.accumulate = accumulate; rj.alpha = alpha;
      rjobs.push_back(rj);
    }
  }
  return RUMPY_OK;
}

int wgrad_launch(const WgradJob* jobs_dev, int njobs, const WgradReduceJob* rjobs_dev, int nrjobs, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes) != cudaSuccess)
      return set_error(RUMPY_ERR_CUDA, "wgrad cudaFuncSetAttribute: %s", cudaGetErrorString(cudaGetLastError()));
    attr_set = true;
  }
  wgrad_tc_kernel<<<njobs, kWgThreads, kWgSmemBytes, s>>>(jobs_dev);
  if (int e = check_launch("wgrad_tc")) return e;
  wgrad_reduce_kernel<<<dim3(36, nrjobs), 256, 0, s>>>(rjobs_dev);
  return check_launch("wgrad_reduce");
}

}  // namespace rb

using namespace rb;

extern "C" {

long long rumpy_conv3x3_wgrad_workspace(int N, int H, int W, int Cin, int Cout) {
  int sms = 148;
  if (device_info(&sms)) return -1;
  if (Cin % 64 || Cout % 64 || N <= 0 || H <= 0 || W <= 0) return -1;
  return (long long)wgrad_layout(N, H, W, Cin, Cout, sms, 0).bytes;
}

int rumpy_conv3x3_wgrad(const void* g_bf16, const void* x_bf16, float* dw_oihw, void* workspace, int N, int H, int W,
                        int Cin, int Cout, int g_unshuffle_r, float alpha, int accumulate, void* stream) {
  int sms = 0;
  if (int e = device_info(&sms)) return e;
  if (!g_bf16 || !x_bf16 || !dw_oihw || !workspace) return set_error(RUMPY_ERR_ARG, "wgrad: null pointer");
  if (Cin % 64 || Cout % 64 || N <= 0 || H <= 0 || W <= 0) return set_error(RUMPY_ERR_ARG, "wgrad: bad shape");
  const int r = g_unshuffle_r < 1 ? 1 : g_unshuffle_r;
  const WgradLayout L = wgrad_layout(N, H, W, Cin, Cout, sms, 0);
  std::vector<WgradJob> jobs;
  std::vector<WgradReduceJob> rjobs;
  char* ws = static_cast<char*>(workspace);
  if (int e = wgrad_build_jobs(jobs, rjobs, L, g_bf16, x_bf16, dw_oihw, ws, N, H, W, Cin, Cout, r, alpha, accumulate))
    return e;
  cudaStream_t s = cudaStream_t(stream);
  // op-level path: descriptors are uploaded per call (the whole-network executor uploads them once per plan)
  if (cudaMemcpyAsync(ws + L.off_jobs, jobs.data(), jobs.size() * sizeof(WgradJob), cudaMemcpyHostToDevice, s) !=
          cudaSuccess ||
      cudaMemcpyAsync(ws + L.off_rjobs, rjobs.data(), rjobs.size() * sizeof(WgradReduceJob), cudaMemcpyHostToDevice,
                      s) != cudaSuccess)
    return set_error(RUMPY_ERR_CUDA, "wgrad: job upload failed: %s", cudaGetErrorString(cudaGetLastError()));
  cudaStreamSynchronize(s);  // host vectors go out of scope; op-level path only
  return wgrad_launch(reinterpret_cast<const WgradJob*>(ws + L.off_jobs), int(jobs.size()),
                      reinterpret_cast<const WgradReduceJob*>(ws + L.off_rjobs), int(rjobs.size()), s);
}

}  // extern "C"
