// Micro-benchmark: cost of tcgen05.mma chains on B200 as seen from issue to mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_latency tools/tc_latency.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../wakeword_detection_b200/csrc/tc_common.cuh"
using namespace wwb::tc;

// n_mma MMAs of shape 128 x N x 16; accumulator index cycles through n_acc independent TMEM tiles.
__global__ void lat_kernel(int N, int n_mma, int n_acc, int rows_a, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 60000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  if (tid < 32) {   // whole warp runs the loop with warp-uniform operands; one elected lane issues
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 2 * rows_a * 16;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      int acc = 0;
      for (int j = 0; j < n_mma; ++j) {
        uint64_t da = make_desc(a + (j & 3) * 16, rows_a * 16, 128);
        uint64_t db = make_desc(b, N * 16, 128);
        if (elect_one()) mma_f16_ss(tmem + acc * 64, da, db, idesc, j >= n_acc);
        acc = (acc + 1 == n_acc) ? 0 : acc + 1;
      }
      long long t1 = clock64();
      if (elect_one()) mma_commit(&bar);
      mbar_wait(&bar, rep & 1);
      long long t2 = clock64();
      if (tid == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  fence_before_sync(); __syncthreads();
  if (tid < 32) tmem_dealloc(tmem, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(lat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  int cfgs[][4] = {{32,1,1,656},{32,9,1,656},{32,9,3,656},{32,45,5,656},{32,45,1,656},{48,3,1,640},{48,15,5,640},{48,15,1,640},
                   {96,3,1,656},{192,4,1,128},{192,12,1,128},{256,12,1,128},{32,9,1,128},{64,9,1,128},{128,9,1,128}};
  for (auto& c : cfgs) {
    lat_kernel<<<1, 128, 61440>>>(c[0], c[1], c[2], c[3], d);
    long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
    printf("N=%3d n_mma=%2d n_acc=%d rowsA=%3d : issue %5lld clk, done %5lld clk  (%.1f clk/MMA)\n", c[0], c[1], c[2], c[3], h[4], h[5], (double)h[5] / c[1]);
  }
  return 0;
}
