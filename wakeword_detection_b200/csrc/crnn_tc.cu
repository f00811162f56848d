// K2/K3 — CRNN conv + GRU-1 input projection fused on the tensor cores
// (precision WWB_PREC_TC / TC_FAST).  Replaces CRNN/encode.tflite op 0 (CONV_2D 5x20,
// stride (2,8), SAME, ReLU) and the x-side FULLY_CONNECTED of the two layer-1 WHILE
// bodies for a batch of windows:
//     mel window [151,40]  ->  xw1[b, t, 0:192] = conv_out[b, t, 0:640] . W1^T + b_in
// (fwd gates z|r|h then bwd gates z|r|h).  The 48.6 KB/window conv output never leaves the SM.
//
// One persistent CTA per SM works on tiles of 6 windows.  Row r = wl*21 + t of a tile
// (wl = window in the tile, t = conv time step; 19 of every 21 rows are real) is the M index
// of every GEMM, so one thread owns one (window, t) pair.
//
//  * The conv is an implicit GEMM per output frequency f: D[128,32] = A_f[128,128] . Wc^T with
//    k = (freq tap kf, time tap kt).  A_f is never materialised: the window is kept in shared
//    memory TRANSPOSED and fp16-split, XP[freq row p][element e] = 8 consecutive time samples
//    (16 bytes), e = wl*21 + chunk.  Because the conv's time stride is 8 = one element, the
//    operand "row r, k-chunk (kf, j)" is XP[2f+kf][r + j]: linear in r (16 bytes per row) — the
//    tcgen05 shared-memory descriptor reads it in place (SBO = 128, LBO = 16 or, for the chunk
//    pair that straddles two freq taps, row pitch - 32).
//  * XP streams through a ring of three 8-row groups (+ one mirror row so that "next row" is
//    always physically adjacent), filled by four producer warps from the mel windows in global
//    memory (sector-aligned 128-bit loads, register transpose, conflict-free 128-bit stores).
//  * The conv accumulator of frequency f is read by the epilogue warps (bias, ReLU, fp16
//    hi/lo split) and written to shared memory as the k-slice [32f, 32f+32) of the GRU input
//    projection, whose 128x192 accumulator stays in TMEM over the 20 slices; W1 slices
//    (24 KB, pre-packed on the host) stream through a 3-slot cp.async.bulk ring.
//  * Everything is fp16 hi/lo split with three MMAs per product (hi*hi + lo*hi + hi*lo), fp32
//    accumulation in TMEM: fp32-equivalent results (precision).
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wwb {

using namespace tc;

constexpr int CA_WPT = 6;                         // windows per tile
constexpr int CA_TP = 21;                         // row slots per window
constexpr int CA_T = 19, CA_F = 20;
constexpr int CA_PITCH = 131 * 16;                // bytes per XP row (130 elements used; odd multiple of 16)
constexpr int CA_RING_ROWS = 25;                  // 3 groups of 8 + mirror of physical row 0
constexpr int CA_XP_PLANE = CA_RING_ROWS * CA_PITCH;
constexpr int CA_GROUPS = 7;                      // 8-row groups per tile (padded rows p = freq + 8, 0..55)
constexpr int CA_A1_PLANE = 4 * 128 * 16;         // one k-slice (32 values) of 128 rows
constexpr int CA_W1_SLICE = 2 * 4 * 192 * 16;     // hi + lo planes of one k-slice of W1
constexpr int CA_W1_SLOTS = 3;
constexpr int CA_CW_PLANE = 16 * 32 * 16;         // conv weights: 16 k-chunks x 32 channels
constexpr int CA_EPI_WARPS = 4, CA_PROD_WARPS = 4;
constexpr int CA_THREADS = (CA_EPI_WARPS + 2 + CA_PROD_WARPS) * 32;   // 320

struct CaSmem {
  unsigned char xp[2 * CA_XP_PLANE];
  unsigned char a1[2][2 * CA_A1_PLANE];
  unsigned char w1[CA_W1_SLOTS][CA_W1_SLICE];
  unsigned char cw[2 * CA_CW_PLANE];
  float conv_b[32];
  float b_in[192];
  uint64_t xp_full[3], xp_empty[3], cacc_full[2], cacc_empty[2], a1_full[2], a1_empty[2];
  uint64_t w1_full[CA_W1_SLOTS], w1_empty[CA_W1_SLOTS], pacc_full[2], pacc_empty[2];
  uint32_t tmem_base;
};

struct CaParams {
  WinMap wm;
  const unsigned char* cw;      // packed conv weights (2 planes)
  const unsigned char* w1;      // [20][CA_W1_SLICE]
  const float* conv_b;
  const float* b_in;            // [192]
  float* xw1;                   // [n_win, 19, 192]
  int L;
  int nsplit;
};

__device__ __forceinline__ void split8v(const float (&x)[8], uint4& hi, uint4& lo) {
  split_pair(x[0], x[1], hi.x, lo.x);
  split_pair(x[2], x[3], hi.y, lo.y);
  split_pair(x[4], x[5], hi.z, lo.z);
  split_pair(x[6], x[7], hi.w, lo.w);
}

__global__ void __launch_bounds__(CA_THREADS, 1) crnn_front_tc_kernel(const CaParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  CaSmem& sm = *reinterpret_cast<CaSmem*>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_win = P.wm.n_win_dev ? (int64_t)*P.wm.n_win_dev : P.wm.n_win;
  const int64_t n_tiles = (n_win + CA_WPT - 1) / CA_WPT;

  // ---- one-time setup ----
  for (int i = tid; i < (int)(sizeof(sm.xp) / 16); i += CA_THREADS) reinterpret_cast<uint4*>(sm.xp)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (int)(sizeof(sm.cw) / 16); i += CA_THREADS)
    reinterpret_cast<uint4*>(sm.cw)[i] = reinterpret_cast<const uint4*>(P.cw)[i];
  if (tid < 32) sm.conv_b[tid] = P.conv_b[tid];
  if (tid < 192) sm.b_in[tid] = P.b_in[tid];
  if (tid == 0) {
    for (int i = 0; i < 3; ++i) { mbar_init(&sm.xp_full[i], CA_PROD_WARPS); mbar_init(&sm.xp_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sm.cacc_full[i], 1); mbar_init(&sm.cacc_empty[i], CA_EPI_WARPS);
      mbar_init(&sm.a1_full[i], CA_EPI_WARPS); mbar_init(&sm.a1_empty[i], 1);
      mbar_init(&sm.pacc_full[i], 1); mbar_init(&sm.pacc_empty[i], CA_EPI_WARPS);
    }
    for (int i = 0; i < CA_W1_SLOTS; ++i) { mbar_init(&sm.w1_full[i], 1); mbar_init(&sm.w1_empty[i], 1); }
    mbar_fence_init();
  }
  if (warp == CA_EPI_WARPS) tmem_alloc(&sm.tmem_base, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t TM_CACC = 0, TM_PACC = 64;   // conv acc: 2 x 32 cols; projection acc: 2 x 192 cols

  if (warp < CA_EPI_WARPS) {
    // =========================== epilogue warps: one row each ===========================
    const int q = warp;
    const int r = q * 32 + lane;
    const int wl = r / CA_TP, t = r - wl * CA_TP;
    const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int64_t b = tile * CA_WPT + wl;
      const bool valid = (wl < CA_WPT) && (t < CA_T) && (b < n_win);
      for (int f = 0; f < CA_F; ++f) {
        const uint32_t ci = tcount * CA_F + f;
        const int cb = ci & 1;
        mbar_wait(&sm.cacc_full[cb], (ci >> 1) & 1);
        fence_after_sync();
        float v0[16], v1[16];
        tmem_ld16(tlane + TM_CACC + cb * 32, v0);
        tmem_ld16(tlane + TM_CACC + cb * 32 + 16, v1);
        tmem_ld_wait();
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.cacc_empty[cb]);
        uint4 hi[4], lo[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float a = (c < 2) ? v0[c * 8 + i] : v1[(c - 2) * 8 + i];
            x[i] = fmaxf(a + sm.conv_b[c * 8 + i], 0.f);
          }
          split8v(x, hi[c], lo[c]);
        }
        mbar_wait(&sm.a1_empty[cb], ((ci >> 1) & 1) ^ 1);
        unsigned char* a1 = sm.a1[cb] + r * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          *reinterpret_cast<uint4*>(a1 + c * 2048) = hi[c];
          *reinterpret_cast<uint4*>(a1 + CA_A1_PLANE + c * 2048) = lo[c];
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.a1_full[cb]);
      }
      // ---- projection epilogue: + b_in -> xw1 ----
      const int pb = tcount & 1;
      mbar_wait(&sm.pacc_full[pb], (tcount >> 1) & 1);
      fence_after_sync();
      float* dst = P.xw1 + (b * CA_T + t) * 192;
#pragma unroll 1
      for (int c0 = 0; c0 < 192; c0 += 16) {
        float v[16];
        tmem_ld16(tlane + TM_PACC + pb * 192 + c0, v);
        tmem_ld_wait();
        if (valid) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            reinterpret_cast<float4*>(dst + c0)[i] =
                make_float4(v[4 * i] + sm.b_in[c0 + 4 * i], v[4 * i + 1] + sm.b_in[c0 + 4 * i + 1],
                            v[4 * i + 2] + sm.b_in[c0 + 4 * i + 2], v[4 * i + 3] + sm.b_in[c0 + 4 * i + 3]);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.pacc_empty[pb]);
    }
  } else if (warp == CA_EPI_WARPS) {
    // =========================== MMA issuer ===========================
    const uint32_t idesc_c = make_idesc_f16(128, 32), idesc_p = make_idesc_f16(128, 192);
    const uint32_t uXP = smem_u32(sm.xp), uCW = smem_u32(sm.cw);
    const int nsplit = P.nsplit;
    uint32_t tcount = 0, w1cnt = 0;
    auto inproj = [&](uint32_t ci, int f, uint32_t pacc) {
      const int cb = ci & 1;
      const int sl = w1cnt % CA_W1_SLOTS;
      mbar_wait(&sm.a1_full[cb], (ci >> 1) & 1);
      mbar_wait(&sm.w1_full[sl], (w1cnt / CA_W1_SLOTS) & 1);
      fence_after_sync();
      const uint32_t a_hi = smem_u32(sm.a1[cb]), a_lo = a_hi + CA_A1_PLANE;
      const uint32_t b_hi = smem_u32(sm.w1[sl]), b_lo = b_hi + CA_W1_SLICE / 2;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const uint64_t dah = make_desc(a_hi + kk * 4096, 2048, 128), dal = make_desc(a_lo + kk * 4096, 2048, 128);
        const uint64_t dbh = make_desc(b_hi + kk * 6144, 3072, 128), dbl = make_desc(b_lo + kk * 6144, 3072, 128);
        mma_f16_ss_w(pacc, dah, dbh, idesc_p, (f | kk) != 0);
        if (nsplit == 3) {
          mma_f16_ss_w(pacc, dal, dbh, idesc_p, true);
          mma_f16_ss_w(pacc, dah, dbl, idesc_p, true);
        }
      }
      if (elect_one()) {
        mma_commit(&sm.a1_empty[cb]);
        mma_commit(&sm.w1_empty[sl]);
      }
      __syncwarp();
      ++w1cnt;
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const int pb = tcount & 1;
      const uint32_t pacc = tmem + TM_PACC + pb * 192;
      const uint32_t gbase = tcount * CA_GROUPS;
      int groups_ready = 0;
      for (int f = 0; f < CA_F; ++f) {
        const uint32_t ci = tcount * CA_F + f;
        const int cb = ci & 1;
        const int g_hi = (2 * f + 12) >> 3;
        for (; groups_ready <= g_hi; ++groups_ready) {
          const uint32_t gg = gbase + groups_ready;
          mbar_wait(&sm.xp_full[gg % 3], (gg / 3) & 1);
        }
        mbar_wait(&sm.cacc_empty[cb], ((ci >> 1) & 1) ^ 1);
        fence_after_sync();
        const uint32_t cacc = tmem + TM_CACC + cb * 32;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const int q0 = 2 * m, kf0 = q0 / 3, j0 = q0 % 3;
          const int rowp = 2 * f + 7 + kf0;
          const uint32_t phys = ((gbase + (rowp >> 3)) % 3) * 8 + (rowp & 7);
          const uint32_t lbo = (j0 == 2 && m != 7) ? (uint32_t)(CA_PITCH - 32) : 16u;
          const uint32_t a0 = uXP + phys * CA_PITCH + j0 * 16;
          const uint64_t dah = make_desc(a0, lbo, 128), dal = make_desc(a0 + CA_XP_PLANE, lbo, 128);
          const uint64_t dbh = make_desc(uCW + q0 * 512, 512, 128), dbl = make_desc(uCW + CA_CW_PLANE + q0 * 512, 512, 128);
          mma_f16_ss_w(cacc, dah, dbh, idesc_c, m != 0);
          if (nsplit == 3) {
            mma_f16_ss_w(cacc, dal, dbh, idesc_c, true);
            mma_f16_ss_w(cacc, dah, dbl, idesc_c, true);
          }
        }
        if (elect_one()) {
          mma_commit(&sm.cacc_full[cb]);
          // group f/4 was last read by conv(f) when f = 4*(f/4)
          if ((f & 3) == 0) mma_commit(&sm.xp_empty[(gbase + (f >> 2)) % 3]);
          if (f == CA_F - 1) { mma_commit(&sm.xp_empty[(gbase + 5) % 3]); mma_commit(&sm.xp_empty[(gbase + 6) % 3]); }
        }
        __syncwarp();
        if (f == 0) {
          mbar_wait(&sm.pacc_empty[pb], ((tcount >> 1) & 1) ^ 1);
          fence_after_sync();
        } else {
          inproj(ci - 1, f - 1, pacc);
        }
      }
      inproj(tcount * CA_F + CA_F - 1, CA_F - 1, pacc);
      if (elect_one()) mma_commit(&sm.pacc_full[pb]);
      __syncwarp();
    }
  } else if (warp == CA_EPI_WARPS + 1) {
    // =========================== W1 slice loader ===========================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int f = 0; f < CA_F; ++f, ++cnt) {
          const int sl = cnt % CA_W1_SLOTS;
          mbar_wait(&sm.w1_empty[sl], ((cnt / CA_W1_SLOTS) & 1) ^ 1);
          mbar_arrive_expect_tx(&sm.w1_full[sl], CA_W1_SLICE);
          bulk_g2s(sm.w1[sl], P.w1 + (size_t)f * CA_W1_SLICE, CA_W1_SLICE, &sm.w1_full[sl]);
        }
    }
  } else {
    // =========================== XP producers ===========================
    const int task = tid - (CA_EPI_WARPS + 2) * 32;      // 0..127; element e = task
    const int wl = task / CA_TP, c = task - wl * CA_TP;  // window in tile, time chunk (frames 8c-6 .. 8c+1)
    const bool has_task = task < CA_WPT * CA_TP;
    const int L = P.L;
    uint32_t gg = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t b = tile * CA_WPT + wl;
      const bool live = has_task && b < n_win;
      for (int g = 0; g < CA_GROUPS; ++g, ++gg) {
        const int slot = gg % 3;
        // issue the loads before waiting for the slot
        float4 v[8][2];
        const bool data = live && g >= 1 && g <= 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int frame = 8 * c - 6 + i;
          if (data && frame >= 0 && frame < L) {
            const float4* src = reinterpret_cast<const float4*>(win_row(P.wm, b, frame) + 8 * (g - 1));
            v[i][0] = __ldg(src);
            v[i][1] = __ldg(src + 1);
          } else {
            v[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            v[i][1] = v[i][0];
          }
        }
        mbar_wait(&sm.xp_empty[slot], ((gg / 3) & 1) ^ 1);
        if (has_task) {
          unsigned char* base = sm.xp + (slot * 8) * CA_PITCH + task * 16;
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float x[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 s = v[i][k >> 2];
              x[i] = (k & 3) == 0 ? s.x : (k & 3) == 1 ? s.y : (k & 3) == 2 ? s.z : s.w;
            }
            uint4 hi, lo;
            split8v(x, hi, lo);
            *reinterpret_cast<uint4*>(base + k * CA_PITCH) = hi;
            *reinterpret_cast<uint4*>(base + k * CA_PITCH + CA_XP_PLANE) = lo;
            if (slot == 0 && k == 0) {   // mirror of physical row 0 after the last ring row
              *reinterpret_cast<uint4*>(sm.xp + 24 * CA_PITCH + task * 16) = hi;
              *reinterpret_cast<uint4*>(sm.xp + 24 * CA_PITCH + task * 16 + CA_XP_PLANE) = lo;
            }
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.xp_full[slot]);
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == CA_EPI_WARPS) tmem_dealloc(tmem, 512);
}

// ---- host side ------------------------------------------------------------------------------
static void put_split16(std::vector<unsigned char>& buf, size_t hi_off, size_t lo_off, float x) {
  __half h = __float2half_rn(x);
  __half l = __float2half_rn(x - __half2float(h));
  memcpy(&buf[hi_off], &h, 2);
  memcpy(&buf[lo_off], &l, 2);
}

// conv_w [32][5][20] -> [plane][16 chunks][32 channels][8 halves]; chunk q = kf*3 + j holds time taps 8j..8j+7
std::vector<unsigned char> crnn_pack_conv(const float* conv_w) {
  std::vector<unsigned char> out(2 * CA_CW_PLANE, 0);
  for (int q = 0; q < 15; ++q)
    for (int n = 0; n < 32; ++n)
      for (int e = 0; e < 8; ++e) {
        const int kf = q / 3, kt = (q % 3) * 8 + e;
        if (kt >= 20) continue;
        const size_t off = ((size_t)q * 32 + n) * 16 + e * 2;
        put_split16(out, off, off + CA_CW_PLANE, conv_w[(n * 5 + kf) * 20 + kt]);
      }
  return out;
}

// W1 [192][640] (row = fwd gates then bwd gates) -> 20 slices [plane][4 chunks][192][8 halves]
std::vector<unsigned char> crnn_pack_w1(const float* w_nk) {
  std::vector<unsigned char> out((size_t)20 * CA_W1_SLICE, 0);
  for (int f = 0; f < 20; ++f)
    for (int c = 0; c < 4; ++c)
      for (int n = 0; n < 192; ++n)
        for (int e = 0; e < 8; ++e) {
          const size_t off = (size_t)f * CA_W1_SLICE + ((size_t)c * 192 + n) * 16 + e * 2;
          put_split16(out, off, off + CA_W1_SLICE / 2, w_nk[(size_t)n * 640 + f * 32 + c * 8 + e]);
        }
  return out;
}

int crnn_front_tc(wwb_ctx* ctx, const WinMap& wm, float* xw1, cudaStream_t st) {
  if (wm.n_win == 0) return WWB_OK;
  CaParams P;
  P.wm = wm;
  P.cw = ctx->crnn.tc_conv;
  P.w1 = ctx->crnn.tc_w1;
  P.conv_b = ctx->crnn.conv_b;
  P.b_in = ctx->crnn.gru_bi[0];
  P.xw1 = xw1;
  P.L = ctx->L;
  P.nsplit = ctx->precision == WWB_PREC_TC ? 3 : 1;
  const size_t smem = sizeof(CaSmem) + 128;
  WWB_CUDA(ctx, cudaFuncSetAttribute(crnn_front_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_tiles = (wm.n_win + CA_WPT - 1) / CA_WPT;
  const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, ctx->sm_count);
  crnn_front_tc_kernel<<<grid, CA_THREADS, smem, st>>>(P);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

}  // namespace wwb
