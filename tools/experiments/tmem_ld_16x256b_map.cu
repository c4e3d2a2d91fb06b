// Experiment (GPU box): which (TMEM lane, column) does each register of tcgen05.ld.16x256b.x4 hold?  Rows are written with
// the 32x32b shape (thread = lane, register = column) as lane*100 + column, read back with the 16x256b shape from lane
// offsets 0 and 16 of the warp's quadrant.  If a thread holds the SAME few columns of several rows, per-column sums over
// the 128 rows of an accumulator need 7 shuffles per warp instead of the 31 of a full 32x32 lane transpose.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I rumpy_b200/csrc -o tools/experiments/tmem_ld_16x256b_map tools/experiments/tmem_ld_16x256b_map.cu
#include "ptx.cuh"
#include <vector>
using namespace rb;

__global__ void __launch_bounds__(128, 1) map_kernel(uint32_t* out) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) tmem_alloc<64>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  const uint32_t quad = tmem + (uint32_t(warp * 32) << 16);
  uint32_t v[32];
  for (int c = 0; c < 32; ++c) v[c] = uint32_t((warp * 32 + lane) * 100 + c);
  tmem_st(quad, v);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  for (int half = 0; half < 2; ++half) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(quad + (uint32_t(half * 16) << 16)));
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[((warp * 2 + half) * 32 + lane) * 16 + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(tmem); }
}

int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
  map_kernel<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<uint32_t> h(4 * 2 * 32 * 16);
  cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
  for (int warp : {0, 1})
    for (int half = 0; half < 2; ++half)
      for (int lane : {0, 1, 2, 3, 4, 5, 8, 31}) {
        printf("warp %d half %d lane %2d:", warp, half, lane);
        for (int i = 0; i < 16; ++i) { const uint32_t x = h[((warp * 2 + half) * 32 + lane) * 16 + i]; printf(" r%d=(%u,%u)", i, x / 100, x % 100); }
        printf("\n");
      }
  return 0;
}
