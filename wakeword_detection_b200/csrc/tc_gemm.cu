// Tensor-core GEMM with bias for the CRNN input projections:
//     C[M, 192] = A[M, K] * W^T + b      (K = 640 for GRU layer 1, 64 for layer 2)
// A is fp32 in global memory; it is split on the fly into fp16 hi + lo planes, W is
// pre-split on the host, and three tcgen05 MMAs per k-step (hi*hi + lo*hi + hi*lo) with
// fp32 accumulation in TMEM reproduce the fp32 product to ~1e-6 (precision).
//
// Warp-specialised persistent kernel, one CTA per SM:
//   warps 0-7  : A producers  (global fp32 -> hi/lo fp16 -> shared chunk panels)
//   warp  8    : MMA issuer   (one elected thread, tcgen05.mma, accumulators in TMEM)
//   warp  9    : B loader     (cp.async.bulk of the pre-packed weight stage, 48 KB each)
//   warps 10-13: epilogue     (tcgen05.ld -> +bias -> global), double-buffered accumulator
// Stages of 64 k-values are double-buffered in shared memory behind full/empty mbarriers.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wwb {

using namespace tc;

constexpr int G_N = 192;
constexpr int G_KS = 64;                          // k-values per stage
constexpr int G_CH = G_KS / 8;                    // 16-byte chunks per row per stage
constexpr int G_STAGES = 2;
constexpr int G_A_BYTES = 2 * G_CH * 128 * 16;    // hi+lo planes: 32 KB
constexpr int G_B_BYTES = 2 * G_CH * G_N * 16;    // 48 KB
constexpr int G_PRODUCERS = 8;
constexpr int G_THREADS = (G_PRODUCERS + 2 + 4) * 32;   // 448
constexpr int G_TMEM_COLS = 512;                  // two 192-column accumulators

struct GemmSmem {
  unsigned char a[G_STAGES][G_A_BYTES];
  unsigned char b[G_STAGES][G_B_BYTES];
  uint64_t full[G_STAGES], empty[G_STAGES], accfull[2], accempty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(G_THREADS, 1)
tc_gemm_bias_kernel(const float* __restrict__ A, const unsigned char* __restrict__ Bpacked,
                    const float* __restrict__ bias, float* __restrict__ C, int64_t M, int K, int nsplit,
                    int xw_layout) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_ks = K / G_KS;
  // xw_layout: M = B*19 rows (window b, step t).  A tile is then 128 consecutive WINDOWS at ONE step t, so that a
  // warp's stores into [b / 128][t][column][b % 128] are 512 contiguous bytes.  (With 128 consecutive (b, t) rows per
  // tile every lane wrote its own 16 bytes 98 KB apart: 33 M scattered sector writes per launch, ~3/4 of what the L2
  // accepts - that, not HBM, bounded the kernel.)
  const int64_t n_win = xw_layout ? M / 19 : 0;
  const int64_t n_tiles = xw_layout ? ((n_win + 127) / 128) * 19 : (M + 127) / 128;
  auto row_of = [&](int64_t tile, int r) -> int64_t {     // global row of tile row r, or -1
    if (!xw_layout) { const int64_t g = tile * 128 + r; return g < M ? g : -1; }
    const int64_t bt = tile / 19, b = bt * 128 + r;
    return b < n_win ? b * 19 + (tile - bt * 19) : -1;
  };

  if (tid == 0) {
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(&sm.full[s], G_PRODUCERS + 1); mbar_init(&sm.empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&sm.accfull[b], 1); mbar_init(&sm.accempty[b], 4); }
    mbar_fence_init();
  }
  if (warp == G_PRODUCERS) tmem_alloc(&sm.tmem_base, G_TMEM_COLS);
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;

  if (warp < G_PRODUCERS) {
    // ---------------- A producers ----------------
    // Software-pipelined: the global loads of step it+1 are in flight while step it is converted and stored,
    // and they are issued BEFORE waiting for the shared-memory slot (the kernel is HBM-bound for K = 64:
    // 128 KB of traffic against 1260 clk of tensor work per tile; exposed load latency was 71 % of its stalls).
    uint32_t it = 0;
    const int r8 = lane & 7, cq = lane >> 3;          // row within an 8-row group, chunk 0..3
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int64_t total = my_tiles * n_ks;
    float4 cur[4][2], nxt[4][2];
    auto load = [&](int64_t step, float4 (&v)[4][2]) {
      const int64_t tile = blockIdx.x + (step / n_ks) * gridDim.x;
      const int ks = (int)(step % n_ks);
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int r = warp * 16 + rr * 8 + r8;
          const int c = cc * 4 + cq;
          const int64_t grow = row_of(tile, r);
          v[rr * 2 + cc][0] = make_float4(0.f, 0.f, 0.f, 0.f);
          v[rr * 2 + cc][1] = v[rr * 2 + cc][0];
          if (grow >= 0) {
            const float4* src = reinterpret_cast<const float4*>(A + grow * K + ks * G_KS + c * 8);
            v[rr * 2 + cc][0] = __ldg(src);
            v[rr * 2 + cc][1] = __ldg(src + 1);
          }
        }
    };
    if (total > 0) load(0, cur);
    for (int64_t step = 0; step < total; ++step, ++it) {
      if (step + 1 < total) load(step + 1, nxt);
      const int s = it % G_STAGES;
      mbar_wait(&sm.empty[s], ((it / G_STAGES) & 1) ^ 1);
      unsigned char* hi = sm.a[s];
      unsigned char* lo = sm.a[s] + G_CH * 128 * 16;
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int r = warp * 16 + rr * 8 + r8;     // row within the tile
          const int c = cc * 4 + cq;
          const float4 v0 = cur[rr * 2 + cc][0], v1 = cur[rr * 2 + cc][1];
          const float x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
          __half h[8], l[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) split_f16(x[i], h[i], l[i]);
          uint4 ph = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
          uint4 pl = make_uint4(pack_h2(l[0], l[1]), pack_h2(l[2], l[3]), pack_h2(l[4], l[5]), pack_h2(l[6], l[7]));
          *reinterpret_cast<uint4*>(hi + (c * 128 + r) * 16) = ph;
          *reinterpret_cast<uint4*>(lo + (c * 128 + r) * 16) = pl;
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.full[s]);
#pragma unroll
      for (int i = 0; i < 4; ++i) { cur[i][0] = nxt[i][0]; cur[i][1] = nxt[i][1]; }
    }
  } else if (warp == G_PRODUCERS) {
    // ---------------- MMA issuer ----------------
    const uint32_t idesc = make_idesc_f16(128, G_N);
    uint32_t it = 0, tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int ab = tcount & 1;
      mbar_wait(&sm.accempty[ab], ((tcount >> 1) & 1) ^ 1);
      fence_after_sync();
      const uint32_t d = tmem + ab * 256;
      for (int ks = 0; ks < n_ks; ++ks, ++it) {
        const int s = it % G_STAGES;
        mbar_wait(&sm.full[s], (it / G_STAGES) & 1);
        fence_after_sync();
        {
          const uint32_t a_hi = smem_u32(sm.a[s]), a_lo = a_hi + G_CH * 128 * 16;
          const uint32_t b_hi = smem_u32(sm.b[s]), b_lo = b_hi + G_CH * G_N * 16;
#pragma unroll
          for (int kk = 0; kk < G_KS / 16; ++kk) {
            const uint32_t ao = kk * 2 * 128 * 16, bo = kk * 2 * G_N * 16;
            const uint64_t dah = make_desc(a_hi + ao, 128 * 16, 128), dal = make_desc(a_lo + ao, 128 * 16, 128);
            const uint64_t dbh = make_desc(b_hi + bo, G_N * 16, 128), dbl = make_desc(b_lo + bo, G_N * 16, 128);
            mma_f16_ss_w(d, dah, dbh, idesc, (ks | kk) != 0);
            if (nsplit == 3) {
              mma_f16_ss_w(d, dal, dbh, idesc, true);
              mma_f16_ss_w(d, dah, dbl, idesc, true);
            }
          }
          if (elect_one()) {
            mma_commit(&sm.empty[s]);
            if (ks == n_ks - 1) mma_commit(&sm.accfull[ab]);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == G_PRODUCERS + 1) {
    // ---------------- B loader ----------------
    if (lane == 0) {
      uint32_t it = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int ks = 0; ks < n_ks; ++ks, ++it) {
          const int s = it % G_STAGES;
          mbar_wait(&sm.empty[s], ((it / G_STAGES) & 1) ^ 1);
          mbar_arrive_expect_tx(&sm.full[s], G_B_BYTES);
          bulk_g2s(sm.b[s], Bpacked + (size_t)ks * G_B_BYTES, G_B_BYTES, &sm.full[s]);
        }
      }
    }
  } else {
    // ---------------- epilogue ----------------
    const int q = warp & 3;                      // TMEM lane quadrant of this warp
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int ab = tcount & 1;
      mbar_wait(&sm.accfull[ab], (tcount >> 1) & 1);
      fence_after_sync();
      const int64_t grow = row_of(tile, q * 32 + lane);
      const uint32_t taddr = tmem + ab * 256 + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c0 = 0; c0 < G_N; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + c0, v);
        tmem_ld_wait();
        if (grow >= 0) {
          // xw_layout: rows are (window b, t) pairs and the output goes to the coalesced layout of the
          // recurrence kernel, [b / 128][t][48 float4 columns][b % 128] (crnn_tc.cu)
          const int64_t b = grow / 19;
          const int t = (int)(grow - b * 19);
          float4* dst = xw_layout ? reinterpret_cast<float4*>(C) + (((b >> 7) * 19 + t) * 48 + c0 / 4) * 128 + (b & 127)
                                  : reinterpret_cast<float4*>(C + grow * G_N + c0);
          const int64_t step = xw_layout ? 128 : 1;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            dst[i * step] = make_float4(v[4 * i] + __ldg(bias + c0 + 4 * i), v[4 * i + 1] + __ldg(bias + c0 + 4 * i + 1),
                                        v[4 * i + 2] + __ldg(bias + c0 + 4 * i + 2), v[4 * i + 3] + __ldg(bias + c0 + 4 * i + 3));
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.accempty[ab]);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == G_PRODUCERS) tmem_dealloc(tmem, G_TMEM_COLS);
}

// Host: W [N=192 rows][K] fp32 (row n = output unit) -> per-stage packed hi/lo chunk panels
std::vector<unsigned char> pack_gemm_b(const float* w_nk, int K, bool split) {
  const int n_ks = K / G_KS;
  std::vector<unsigned char> out((size_t)n_ks * G_B_BYTES, 0);
  for (int ks = 0; ks < n_ks; ++ks)
    for (int c = 0; c < G_CH; ++c)
      for (int n = 0; n < G_N; ++n)
        for (int e = 0; e < 8; ++e) {
          float x = w_nk[(size_t)n * K + ks * G_KS + c * 8 + e];
          __half h = __float2half_rn(x);
          __half l = split ? __float2half_rn(x - __half2float(h)) : __float2half_rn(0.f);
          size_t off = (size_t)ks * G_B_BYTES + ((size_t)c * G_N + n) * 16 + e * 2;
          memcpy(&out[off], &h, 2);
          memcpy(&out[off + (size_t)G_CH * G_N * 16], &l, 2);
        }
  return out;
}

int tc_gemm_bias(wwb_ctx* ctx, const float* A, const unsigned char* Bpacked, const float* bias, float* C, int64_t M,
                 int K, int nsplit, int xw_layout, cudaStream_t st) {
  if (M == 0) return WWB_OK;
  if (K % G_KS) return fail(ctx, WWB_ERR_ARG, "tc_gemm: K must be a multiple of %d", G_KS);
  const size_t smem = sizeof(GemmSmem) + 128;
  WWB_CUDA(ctx, cudaFuncSetAttribute(tc_gemm_bias_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = xw_layout ? ((M / 19 + 127) / 128) * 19 : (M + 127) / 128;
  const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, ctx->sm_count);
  tc_gemm_bias_kernel<<<grid, G_THREADS, smem, st>>>(A, Bpacked, bias, C, M, K, nsplit, xw_layout);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

}  // namespace wwb
