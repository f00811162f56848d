// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (".ts" form) on B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/tc_ts_probe tools/tc_ts_probe.cu
// 1. layout check: A[128,16] fp16 written by tcgen05.st.32x32b.x8 (lane = row, column j = k 2j,2j+1),
//    D = A * B^T compared with the host product;
// 2. cost of MMA chains: A from TMEM vs A from shared memory (aligned / row-shifted start), N = 32/48/64.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../wakeword_detection_b200/csrc/tc_common.cuh"
using namespace wwb::tc;

__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
  uint32_t acc = accumulate ? 1u : 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- 1. layout / correctness ----
__global__ void ts_check_kernel(const __half* A /*[128][16]*/, const __half* B /*[N][16]*/, int N, float* D /*[128][N]*/) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // B as chunk panels: chunk c (8 halves) of row n at c*N*16 + n*16
  for (int i = tid; i < N * 16; i += blockDim.x) {
    const int n = i / 16, k = i % 16;
    reinterpret_cast<__half*>(smem + (k / 8) * N * 16 + n * 16)[k % 8] = B[n * 16 + k];
  }
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  // A -> TMEM columns 256..263
  uint32_t r[8];
  for (int j = 0; j < 8; ++j) r[j] = pack_h2(A[tid * 16 + 2 * j], A[tid * 16 + 2 * j + 1]);
  tmem_st8(tmem + ((uint32_t)(warp * 32) << 16) + 256, r);
  tmem_st_wait();
  fence_before_sync(); __syncthreads(); fence_after_sync();
  if (warp == 0) {
    if (elect_one()) {
      mma_f16_ts(tmem, tmem + 256, make_desc(smem_u32(smem), N * 16, 128), make_idesc_f16(128, N), false);
      mma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  fence_after_sync();
  for (int c = 0; c < N; c += 16) {
    float v[16];
    tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) D[tid * N + c + i] = v[i];
  }
  (void)lane;
  fence_before_sync(); __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- 2. cost ----
// mode 0: A from smem, aligned; 1: A from smem, start shifted by `shift` rows; 2: A from TMEM
__global__ void ts_cost_kernel(int N, int n_mma, int n_acc, int mode, int shift, int rows_a, long long* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x;
  for (int i = tid; i < 60000 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (tid < 32) tmem_alloc(&slot, 512);
  fence_async_smem(); fence_before_sync(); __syncthreads(); fence_after_sync();
  const uint32_t tmem = slot;
  {
    uint32_t r[8];
    for (int j = 0; j < 8; ++j) r[j] = 0x3c003c00u;
    for (int c = 0; c < 4; ++c) tmem_st8(tmem + ((uint32_t)((tid >> 5) * 32) << 16) + 448 + c * 8, r);
    tmem_st_wait();
  }
  fence_before_sync(); __syncthreads(); fence_after_sync();
  if (tid < 32) {
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t a = smem_u32(smem) + (mode == 1 ? shift * 16 : 0), b = smem_u32(smem) + 2 * rows_a * 16;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      if (elect_one()) {
        int acc = 0;
        const uint64_t da0 = make_desc(a, rows_a * 16, 128);
        const int astep = mode == 0 ? shift : 8;   // A start advance between MMAs (16-byte units)
        const uint64_t db = make_desc(b, N * 16, 128);
#pragma unroll 4
        for (int j = 0; j < n_mma; ++j) {
          if (mode == 2) mma_f16_ts(tmem + acc * 64, tmem + 448 + (j & 3) * 8, db, idesc, j >= n_acc);
          else mma_f16_ss(tmem + acc * 64, da0 + (uint64_t)((j & 3) * astep), db, idesc, j >= n_acc);
          acc = (acc + 1 == n_acc) ? 0 : acc + 1;
        }
        mma_commit(&bar);
      }
      long long t1 = clock64();
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
  // ---- layout check ----
  for (int N : {32, 48}) {
    __half hA[128 * 16], hB[64 * 16];
    for (int i = 0; i < 128 * 16; ++i) hA[i] = __float2half((float)((i * 7 + i / 16) % 13 - 6));
    for (int i = 0; i < N * 16; ++i) hB[i] = __float2half((float)((i * 5 + i / 16) % 11 - 5));
    __half *dA, *dB; float* dD;
    cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dB, sizeof(hB)); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice); cudaMemcpy(dB, hB, sizeof(hB), cudaMemcpyHostToDevice);
    ts_check_kernel<<<1, 128, 4096>>>(dA, dB, N, dD);
    float* hD = (float*)malloc(128 * N * 4);
    cudaError_t e = cudaMemcpy(hD, dD, 128 * N * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("check error %s\n", cudaGetErrorString(e)); return 1; }
    int bad = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        float ref = 0;
        for (int k = 0; k < 16; ++k) ref += __half2float(hA[m * 16 + k]) * __half2float(hB[n * 16 + k]);
        if (ref != hD[m * N + n]) { if (bad < 5) printf("  mismatch m=%d n=%d got %g want %g\n", m, n, hD[m * N + n], ref); ++bad; }
      }
    printf("TS layout check N=%d: %s (%d mismatches)\n", N, bad ? "FAIL" : "OK", bad);
  }
  // ---- cost ----
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(ts_cost_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 61440);
  int Ns[] = {16, 32, 48, 64, 96, 128};
  for (int N : Ns)
    for (int mode = 0; mode < 3; ++mode)
      for (int shift : {1, 3, 8}) {
        if (mode != 1 && shift != 1) continue;
        for (int n_acc : {1, 4}) {
          const int n_mma = 48;
          ts_cost_kernel<<<1, 128, 61440>>>(N, n_mma, n_acc, mode, mode == 0 ? 8 : shift, 656, d);
          long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          printf("N=%3d mode=%s shift=%d n_acc=%d : issue %5lld clk, done %5lld clk (%.1f clk/MMA)\n", N,
                 mode == 0 ? "ss-aligned" : mode == 1 ? "ss-shifted" : "ts        ", mode == 1 ? shift : 0, n_acc, h[4], h[5],
                 (double)h[5] / n_mma);
        }
      }
  // ---- LBO (panel stride) sensitivity, A from smem, N = 32 ----
  for (int rows_a : {128, 136, 192, 256, 384, 512, 640, 648, 656, 664, 672, 704, 768, 1024, 1280})
    for (int astep : {0, 1, 8}) {
      const int n_mma = 48, N = 32;
      ts_cost_kernel<<<1, 128, 61440>>>(N, n_mma, 4, 0, astep, rows_a, d);
      long long h[6]; cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("LBO rows_a=%4d (%6d B) N=%d astep=%d : %.1f clk/MMA\n", rows_a, rows_a * 16, N, astep, (double)h[5] / n_mma);
    }
  return 0;
}
