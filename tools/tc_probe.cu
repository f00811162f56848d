// Probe of the tcgen05 primitives in csrc/tc_common.cuh on a real B200:
// D[128,N] = A[rows off..off+127, K] * B[N,K]^T with fp16 operands in the no-swizzle
// chunk-panel layout, fp32 accumulation in TMEM.  Verifies descriptor semantics (LBO/SBO,
// arbitrary 16-byte row offsets of the start address) against a CPU product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_probe tools/tc_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../wakeword_detection_b200/csrc/tc_common.cuh"

using namespace wwb::tc;

__global__ void probe_kernel(const __half* A, const __half* B, float* D, int RA, int N, int K, int off, int reps) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  unsigned char* sA = smem;                         // K/8 panels of RA*16 bytes
  unsigned char* sB = smem + (size_t)(K / 8) * RA * 16;   // K/8 panels of N*16 bytes
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&tmem_slot, 64);
  for (int i = tid; i < RA * (K / 8); i += blockDim.x) {
    int r = i % RA, c = i / RA;
    *reinterpret_cast<uint4*>(sA + (size_t)c * RA * 16 + r * 16) = *reinterpret_cast<const uint4*>(A + (size_t)r * K + c * 8);
  }
  for (int i = tid; i < N * (K / 8); i += blockDim.x) {
    int r = i % N, c = i / N;
    *reinterpret_cast<uint4*>(sB + (size_t)c * N * 16 + r * 16) = *reinterpret_cast<const uint4*>(B + (size_t)r * K + c * 8);
  }
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    const uint32_t idesc = make_idesc_f16(128, N);
    for (int rep = 0; rep < reps; ++rep)
      for (int kk = 0; kk < K / 16; ++kk) {
        uint64_t da = make_desc(smem_u32(sA) + off * 16 + kk * 2 * RA * 16, RA * 16, 128);
        uint64_t db = make_desc(smem_u32(sB) + kk * 2 * N * 16, N * 16, 128);
        mma_f16_ss(tmem, da, db, idesc, (kk | rep) != 0);
      }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[(size_t)(warp * 32 + lane) * N + c0 + i] = v[i];
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static int run(int RA, int N, int K, int off, int reps) {
  std::vector<__half> hA((size_t)RA * K), hB((size_t)N * K);
  std::vector<float> fA(hA.size()), fB(hB.size());
  srand(1234 + N + K + off);
  for (size_t i = 0; i < hA.size(); ++i) { float x = (rand() % 2001 - 1000) / 500.0f; hA[i] = __float2half(x); fA[i] = __half2float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float x = (rand() % 2001 - 1000) / 800.0f; hB[i] = __float2half(x); fB[i] = __half2float(hB[i]); }
  __half *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, 128 * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, 128 * N * 4);
  size_t smem = (size_t)(K / 8) * (RA + N) * 16;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(dA, dB, dD, RA, N, K, off, reps);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("RA=%d N=%d K=%d off=%d: CUDA error %s\n", RA, N, K, off, cudaGetErrorString(e)); return 2; }
  std::vector<float> hD(128 * N);
  cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += (double)fA[(size_t)(m + off) * K + k] * fB[(size_t)n * K + k];
      ref *= reps;
      maxerr = fmax(maxerr, fabs(ref - hD[m * N + n]));
      maxref = fmax(maxref, fabs(ref));
    }
  printf("RA=%3d N=%3d K=%3d off=%2d reps=%d: max|err|=%.3e (max|ref|=%.2f) %s\n", RA, N, K, off, reps, maxerr, maxref,
         maxerr < 1e-3 * maxref ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-3 * maxref ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run(128, 32, 16, 0, 1);
  bad += run(128, 32, 48, 0, 1);
  bad += run(128, 48, 16, 0, 1);
  bad += run(128, 32, 112, 0, 1);
  bad += run(128, 192, 64, 0, 1);
  bad += run(128, 96, 32, 0, 3);
  bad += run(160, 32, 48, 8, 1);
  bad += run(160, 32, 48, 16, 1);
  bad += run(160, 32, 48, 1, 1);
  bad += run(160, 32, 48, 2, 1);
  bad += run(160, 32, 48, 4, 1);
  bad += run(160, 32, 48, 13, 1);
  bad += run(1600, 32, 48, 1011, 1);
  bad += run(128, 16, 16, 0, 1);
  bad += run(128, 8, 16, 0, 1);
  printf(bad ? "PROBE FAILED (%d)\n" : "PROBE PASSED\n", bad);
  return bad ? 1 : 0;
}
