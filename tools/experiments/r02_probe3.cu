// Round-2 probe 3: tcgen05.ld 32x32b at ODD column addresses (x8 / x16), needed by the role-swapped epilogue
// (upper accumulator half is shifted by one column).
#include "ptx.cuh"
#include <vector>
using namespace rb;
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr) : "memory");
}
__global__ void __launch_bounds__(128, 1) k(int* bad) {
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t base = tmem_base_s + (uint32_t(warp * 32) << 16);
  for (int c0 = 0; c0 < 512; c0 += 32) {
    uint32_t v[32];
    for (int i = 0; i < 32; ++i) v[i] = (threadIdx.x << 16) | (c0 + i);
    tmem_st32(base + c0, v);
  }
  tmem_st_wait();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  int nbad = 0;
  for (int col : {1, 3, 5, 7, 25, 49, 97, 145, 151}) {
    uint32_t a[16], b[8];
    tmem_ld16(base + col, a);
    tmem_ld8(base + col + 16, b);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) nbad += a[i] != ((threadIdx.x << 16) | (col + i));
    for (int i = 0; i < 8; ++i) nbad += b[i] != ((threadIdx.x << 16) | (col + 16 + i));
  }
  // unaligned tcgen05.st x8-equivalent: st16 at odd column
  {
    uint32_t v[16];
    for (int i = 0; i < 16; ++i) v[i] = 0xabc00000u | (threadIdx.x << 8) | i;
    tmem_st16(base + 201, v);
    tmem_st_wait();
    uint32_t a[16];
    tmem_ld16(base + 201, a);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) nbad += a[i] != v[i];
    uint32_t c[16];
    tmem_ld16(base + 192, c);   // neighbours untouched?
    tmem_ld_wait();
    for (int i = 0; i < 9; ++i) nbad += c[i] != ((threadIdx.x << 16) | (192 + i));
  }
  atomicAdd(bad, nbad);
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem_base_s); }
}
int main() {
  int* d; cudaMalloc(&d, 4); cudaMemset(d, 0, 4);
  k<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  int h = -1; cudaMemcpy(&h, d, 4, cudaMemcpyDeviceToHost);
  printf("odd-column tcgen05.ld/st x8/x16: %d mismatches [%s]\n", h, cudaGetErrorString(e));
  return 0;
}
