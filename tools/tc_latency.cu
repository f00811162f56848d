// Micro-benchmark: cost of tcgen05.mma chains on B200 as seen from issue to mbarrier completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_latency tools/tc_latency.cu
// mode 0: every MMA issued under its own elect_one() by a converged warp (descriptors warp-uniform)
// mode 1: one elect_one() branch around the whole chain (CUTLASS style), descriptors built in-thread
// mode 2: like 1 but descriptors advanced by integer adds on a precomputed base
#include <cstdio>
#include <cuda_runtime.h>
#include "../wakeword_detection_b200/csrc/tc_common.cuh"
using namespace wwb::tc;

__global__ void lat_kernel(int N, int n_mma, int n_acc, int rows_a, int mode, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 60000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  if (tid < 32) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a = smem_u32(smem), b = smem_u32(smem) + 2 * rows_a * 16;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      long long t1;
      if (mode == 0) {
        int acc = 0;
        for (int j = 0; j < n_mma; ++j) {
          uint64_t da = make_desc(a + (j & 3) * 16, rows_a * 16, 128);
          uint64_t db = make_desc(b, N * 16, 128);
          if (elect_one()) mma_f16_ss(tmem + acc * 64, da, db, idesc, j >= n_acc);
          acc = (acc + 1 == n_acc) ? 0 : acc + 1;
        }
        t1 = clock64();
        if (elect_one()) mma_commit(&bar);
      } else if (mode == 1) {
        if (elect_one()) {
          int acc = 0;
          for (int j = 0; j < n_mma; ++j) {
            uint64_t da = make_desc(a + (j & 3) * 16, rows_a * 16, 128);
            uint64_t db = make_desc(b, N * 16, 128);
            mma_f16_ss(tmem + acc * 64, da, db, idesc, j >= n_acc);
            acc = (acc + 1 == n_acc) ? 0 : acc + 1;
          }
          mma_commit(&bar);
        }
        t1 = clock64();
      } else {
        if (elect_one()) {
          int acc = 0;
          const uint64_t da0 = make_desc(a, rows_a * 16, 128);
          const uint64_t db = make_desc(b, N * 16, 128);
#pragma unroll 4
          for (int j = 0; j < n_mma; ++j) {
            mma_f16_ss(tmem + acc * 64, da0 + (uint64_t)(j & 3), db, idesc, j >= n_acc);
            acc = (acc + 1 == n_acc) ? 0 : acc + 1;
          }
          mma_commit(&bar);
        }
        t1 = clock64();
      }
      __syncwarp();
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
  int cfgs[][4] = {{32,1,1,656},{32,9,1,656},{32,45,5,656},{32,45,1,656},{48,15,5,640},{48,15,1,640},{64,32,4,128},
                   {96,32,4,128},{128,32,4,128},{192,32,2,128},{256,32,2,128},{32,64,8,128},{16,64,8,128},{8,64,8,128}};
  for (int mode = 0; mode < 3; ++mode)
    for (auto& c : cfgs) {
      lat_kernel<<<1, 128, 61440>>>(c[0], c[1], c[2], c[3], mode, d);
      long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("mode %d N=%3d n_mma=%2d n_acc=%d rowsA=%3d : issue %5lld clk, done %5lld clk  (%.1f clk/MMA)\n", mode, c[0], c[1], c[2], c[3], h[4], h[5], (double)h[5] / c[1]);
    }
  return 0;
}
