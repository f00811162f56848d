// K5/K6 — WaveNet encode + detect on the tensor cores (precision WWB_PREC_TC / TC_FAST).
//
// One persistent CTA per SM processes groups of WN_G = 3 windows.  The rows of a group
// (window w, time t) -> o = w*198 + t form the M dimension of every GEMM (5 tiles of 128
// rows; the 16 rows after each window are the next window's causal zero padding).  Each of
// the 640 epilogue threads owns ONE row for the whole 24-block stack: its 16-channel
// residual stream and 32-channel skip sum never leave registers.  Per block:
//   gate GEMM   D[128,32] = sum_tap U[rows - (2-tap)*d, 16] * Wg_tap      (tcgen05, TMEM)
//               the dilated taps are row-shifted views of one shared-memory buffer: the
//               operand layout is linear in the row index, so a tap is a start-address offset
//   epilogue 1  g = tanh(.)*sigmoid(.) (ex2/rcp), fp16 hi/lo -> shared memory
//   res/skip    D[128,48] = g[128,16] * [Wres | Wskip]                   (tcgen05, TMEM)
//   epilogue 2  x += ReLU(res); skip += ReLU(.); u = BN_next(x) -> hi/lo -> shared memory
// then the detect head (32->32 on the tensor core, 32->2, max over time, softmax).
// fp16 hi/lo operand split (3 MMAs per product) keeps the result at fp32 accuracy
// (DESIGN.md §precision).  Per-block weights (9.5 KB) stream through a 2-stage
// cp.async.bulk ring; tiles of a block are pipelined through per-tile mbarriers.
#include <string.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace wwb {

using namespace tc;

constexpr int WN_G = 3;                       // windows per group
constexpr int WN_SLOT = 198;                  // rows per window slot (182 + 16 pad)
constexpr int WN_NT = 5;                      // M tiles per group
constexpr int WN_ROWS = WN_NT * 128;          // 640
constexpr int WN_UROWS = WN_ROWS + 16;        // U buffer has 16 leading zero rows
constexpr int WN_PU = WN_UROWS * 16;          // bytes per U chunk panel
constexpr int WN_PG = WN_ROWS * 16;           // bytes per G chunk panel
constexpr int WN_EPI_WARPS = WN_NT * 4;       // 20
constexpr int WN_EPI_THREADS = WN_EPI_WARPS * 32;
constexpr int WN_THREADS = (WN_EPI_WARPS + 1) * 32;   // + MMA/loader warp = 672 (leaves 96 registers per thread)
constexpr int WN_WBLK = 9728;                 // bytes of one block's weight blob
constexpr int WN_GATE_B = 6144, WN_RS_B = 3072;
constexpr int WN_WST = 4;                     // weight ring stages
constexpr int WN_TMEM_TILE = 96;              // columns per tile: gate 32 @0, res/skip 48 @32

// resident head blob (floats unless noted)
struct WnHead {
  float in_w[40 * 16];      // [k][c]
  float in_b[16];
  float bn0_mul[16], bn0_add[16];
  float det1_b[32];
  float det2_w[2 * 32];
  float det2_b[2];
  float pad_[2];
  unsigned char det1_B[2 * 4 * 32 * 16];   // hi/lo planes, 4 chunks x 32 rows x 16 B
};

struct WnSmem {
  unsigned char U[2 * 2 * WN_PU];          // [plane][chunk][row]
  unsigned char Gb[2 * 2 * WN_PG];
  unsigned char W[WN_WST][WN_WBLK];
  WnHead head;
  uint64_t bar_u[WN_NT], bar_gate[WN_NT], bar_g[WN_NT], bar_rs[WN_NT];
  uint64_t wfull[WN_WST];
  uint32_t tmem_base;
  int zmax[WN_G][2];
};

struct WnTcParams {
  WinMap wm;
  const unsigned char* wblob;   // [24][WN_WBLK]
  const WnHead* head;
  int L;
  int nsplit;
  int dil[24];                  // dilation per block (kernel-parameter space keeps it in uniform registers)
  float* enc_out;
  float* det_out;
  float* post;
  long long* dbg;   // optional timeline dump (block 0, first group): [6 roles][24 blocks][4 events]
};

#define WN_DBG(role, k, ev) do { if (P.dbg && blockIdx.x == 0 && grp == (int64_t)gridDim.x) P.dbg[((role) * 24 + (k)) * 4 + (ev)] = clock64(); } while (0)

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh(a) * sigmoid(b) with two ex2 and one rcp
// `a2` = -2*log2(e)*a and `b2` = -log2(e)*b arrive pre-scaled (bias folded in by an FFMA).
// Only e^(-2a) can make the quotient inf/inf, so only it is clamped (tanh is +-1 to 1e-13 there).
__device__ __forceinline__ float gate_fn(float a2, float b2) {
  const float ea = ex2_approx(fminf(a2, 43.f));   // e^(-2a), a >= -14.9
  const float eb = ex2_approx(b2);                // e^(-b); inf -> rcp(inf) = 0 -> g = 0
  return (1.f - ea) * rcp_approx((1.f + ea) * (1.f + eb));
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 fp32 -> one 16-byte chunk of hi halves and one of lo halves
__device__ __forceinline__ void split8(const float (&x)[8], uint4& hi, uint4& lo) {
  split_pair(x[0], x[1], hi.x, lo.x);
  split_pair(x[2], x[3], hi.y, lo.y);
  split_pair(x[4], x[5], hi.z, lo.z);
  split_pair(x[6], x[7], hi.w, lo.w);
}
__device__ __forceinline__ void ld8f(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ void atomic_max_float(int* addr, float v) {
  if (v >= 0.f) atomicMax(addr, __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WN_EPI_THREADS) : "memory"); }

__global__ void __launch_bounds__(WN_THREADS, 1) wavenet_tc_kernel(const WnTcParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  WnSmem& sm = *reinterpret_cast<WnSmem*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
  const int64_t n_win = P.wm.n_win_dev ? (int64_t)*P.wm.n_win_dev : P.wm.n_win;
  const int64_t n_groups = (n_win + WN_G - 1) / WN_G;
  const int L = P.L;

  // ---- one-time setup ----
  for (int i = tid; i < (int)(sizeof(sm.U) / 16); i += WN_THREADS) reinterpret_cast<uint4*>(sm.U)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (int)(sizeof(sm.Gb) / 16); i += WN_THREADS) reinterpret_cast<uint4*>(sm.Gb)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < (int)(sizeof(WnHead) / 16); i += WN_THREADS)
    reinterpret_cast<uint4*>(&sm.head)[i] = reinterpret_cast<const uint4*>(P.head)[i];
  if (tid < WN_G * 2) sm.zmax[tid >> 1][tid & 1] = (int)0xff800000;   // -inf
  if (tid == 0) {
    for (int i = 0; i < WN_NT; ++i) {
      mbar_init(&sm.bar_u[i], 4); mbar_init(&sm.bar_gate[i], 1); mbar_init(&sm.bar_g[i], 4); mbar_init(&sm.bar_rs[i], 1);
    }
    for (int s = 0; s < WN_WST; ++s) mbar_init(&sm.wfull[s], 1);
    mbar_fence_init();
  }
  if (warp == WN_EPI_WARPS) tmem_alloc(&sm.tmem_base, 512);
  fence_async_smem();
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem = sm.tmem_base;
  const uint32_t uU = smem_u32(sm.U), uG = smem_u32(sm.Gb);
  const int nsplit = P.nsplit;

  if (warp < WN_EPI_WARPS) {
    // =========================== epilogue threads: one row each ===========================
    const int tile = warp >> 2, q = warp & 3;
    const int o = tile * 128 + q * 32 + lane;       // row within the group
    const int w = o / WN_SLOT, t = o - w * WN_SLOT;
    const uint32_t tbase = tmem + ((uint32_t)(q * 32) << 16) + tile * WN_TMEM_TILE;
    uint32_t n_gate = 0, n_rs = 0, n_w = 0;          // completed phases of bar_gate / bar_rs / weight ring
    unsigned char* const Urow = sm.U + (16 + o) * 16;
    unsigned char* const Grow = sm.Gb + o * 16;

    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      const int64_t b = grp * WN_G + w;
      const bool valid = (w < WN_G) && (t < L) && (b < n_win);
      float x[16], skip[32];
      // ---- input layer: x = ReLU(in_w * mel + in_b); u0 = BN_0(x) ----
      {
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = sm.head.in_b[c];
        if (valid) {
          const float4* row = reinterpret_cast<const float4*>(win_row(P.wm, b, t));
#pragma unroll
          for (int k4 = 0; k4 < 10; ++k4) {
            const float4 m = __ldg(row + k4);
            const float mm[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float4* wr = reinterpret_cast<const float4*>(sm.head.in_w + (k4 * 4 + j) * 16);
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const float4 wv = wr[c4];
                x[c4 * 4] = fmaf(wv.x, mm[j], x[c4 * 4]);
                x[c4 * 4 + 1] = fmaf(wv.y, mm[j], x[c4 * 4 + 1]);
                x[c4 * 4 + 2] = fmaf(wv.z, mm[j], x[c4 * 4 + 2]);
                x[c4 * 4 + 3] = fmaf(wv.w, mm[j], x[c4 * 4 + 3]);
              }
            }
          }
        }
#pragma unroll
        for (int c = 0; c < 16; ++c) x[c] = fmaxf(x[c], 0.f);
#pragma unroll
        for (int n = 0; n < 32; ++n) skip[n] = 0.f;
        if (valid) {
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
              u[i] = __fadd_rn(__fmul_rn(x[h8 * 8 + i], sm.head.bn0_mul[h8 * 8 + i]), sm.head.bn0_add[h8 * 8 + i]);
            uint4 hi, lo;
            split8(u, hi, lo);
            *reinterpret_cast<uint4*>(Urow + h8 * WN_PU) = hi;
            *reinterpret_cast<uint4*>(Urow + 2 * WN_PU + h8 * WN_PU) = lo;
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_u[tile]);
      }

      for (int k = 0; k < 24; ++k, ++n_w) {
        const int ws = n_w % WN_WST;
        mbar_wait(&sm.wfull[ws], (n_w / WN_WST) & 1);
        const float* wf = reinterpret_cast<const float*>(sm.W[ws] + WN_GATE_B + WN_RS_B);   // gate_b[32] rs_b[48] bn_mul[16] bn_add[16]
        // ---- epilogue 1: gated activation ----
        mbar_wait(&sm.bar_gate[tile], n_gate & 1);
        ++n_gate;
        fence_after_sync();
        if (q == 0 && lane == 0) WN_DBG(tile, k, 0);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          float at[8], as[8];
          tmem_ld8(tbase + h8 * 8, at);
          tmem_ld8(tbase + 16 + h8 * 8, as);
          tmem_ld_wait();
          float g[8], bt[8], bs[8];
          ld8f(wf + h8 * 8, bt);
          ld8f(wf + 16 + h8 * 8, bs);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            g[i] = gate_fn(fmaf(at[i], -2.8853900817779268f, bt[i]), fmaf(as[i], -1.4426950408889634f, bs[i]));
          uint4 hi, lo;
          split8(g, hi, lo);
          *reinterpret_cast<uint4*>(Grow + h8 * WN_PG) = hi;
          *reinterpret_cast<uint4*>(Grow + 2 * WN_PG + h8 * WN_PG) = lo;
        }
        fence_before_sync();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_g[tile]);
        if (q == 0 && lane == 0) WN_DBG(tile, k, 1);

        // ---- epilogue 2: residual + skip, next block's BN ----
        mbar_wait(&sm.bar_rs[tile], n_rs & 1);
        ++n_rs;
        fence_after_sync();
        if (q == 0 && lane == 0) WN_DBG(tile, k, 2);
        const bool last = (k == 23);
#pragma unroll
        for (int h8 = 0; h8 < 2; ++h8) {
          float r[8];
          tmem_ld8(tbase + 32 + h8 * 8, r);
          tmem_ld_wait();
          if (!last) {
            float u[8], br[8], bm[8], ba[8];
            ld8f(wf + 32 + h8 * 8, br);
            ld8f(wf + 80 + h8 * 8, bm);
            ld8f(wf + 96 + h8 * 8, ba);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = h8 * 8 + i;
              x[c] = fmaxf(r[i] + br[i], 0.f) + x[c];
              u[i] = fmaf(x[c], bm[i], ba[i]);
            }
            if (valid) {
              uint4 hi, lo;
              split8(u, hi, lo);
              *reinterpret_cast<uint4*>(Urow + h8 * WN_PU) = hi;
              *reinterpret_cast<uint4*>(Urow + 2 * WN_PU + h8 * WN_PU) = lo;
            }
          }
        }
#pragma unroll
        for (int h8 = 0; h8 < 4; ++h8) {
          float s[8], bk[8];
          tmem_ld8(tbase + 48 + h8 * 8, s);
          ld8f(wf + 48 + h8 * 8, bk);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) skip[h8 * 8 + i] += fmaxf(s[i] + bk[i], 0.f);
        }
        if (last) {
          // detect input: ReLU(skip) hi/lo; channels 0-15 -> G panels, 16-31 -> U panels
#pragma unroll
          for (int h8 = 0; h8 < 4; ++h8) {
            float e[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) e[i] = fmaxf(skip[h8 * 8 + i], 0.f);
            uint4 hi, lo;
            split8(e, hi, lo);
            if (h8 < 2) {
              *reinterpret_cast<uint4*>(Grow + h8 * WN_PG) = hi;
              *reinterpret_cast<uint4*>(Grow + 2 * WN_PG + h8 * WN_PG) = lo;
            } else if (valid) {
              *reinterpret_cast<uint4*>(Urow + (h8 - 2) * WN_PU) = hi;
              *reinterpret_cast<uint4*>(Urow + 2 * WN_PU + (h8 - 2) * WN_PU) = lo;
            }
          }
        }
        fence_before_sync();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_u[tile]);
        if (q == 0 && lane == 0) WN_DBG(tile, k, 3);
      }

      if (P.enc_out && valid) {
        float4* dst = reinterpret_cast<float4*>(P.enc_out + (b * L + t) * 32);
#pragma unroll
        for (int n = 0; n < 32; n += 4) dst[n / 4] = make_float4(skip[n], skip[n + 1], skip[n + 2], skip[n + 3]);
      }
      // ---- detect head epilogue: ReLU(D + b1) -> 32->2 -> max over time ----
      mbar_wait(&sm.bar_gate[tile], n_gate & 1);
      ++n_gate;
      fence_after_sync();
      float z0 = sm.head.det2_b[0], z1 = sm.head.det2_b[1];
#pragma unroll
      for (int h8 = 0; h8 < 4; ++h8) {
        float d[8];
        tmem_ld8(tbase + h8 * 8, d);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e = fmaxf(d[i] + sm.head.det1_b[h8 * 8 + i], 0.f);
          z0 = fmaf(sm.head.det2_w[h8 * 8 + i], e, z0);
          z1 = fmaf(sm.head.det2_w[32 + h8 * 8 + i], e, z1);
        }
      }
      fence_before_sync();
      if (valid) {
        atomic_max_float(&sm.zmax[w][0], z0);
        atomic_max_float(&sm.zmax[w][1], z1);
      }
      epi_bar_sync();
      if (tid < WN_G) {
        const int64_t bb = grp * WN_G + tid;
        if (bb < n_win) {
          const float a0 = __int_as_float(sm.zmax[tid][0]), a1 = __int_as_float(sm.zmax[tid][1]);
          const float m = fmaxf(a0, a1);
          const float e0 = expf(a0 - m), e1 = expf(a1 - m), s = e0 + e1;
          if (P.det_out) { P.det_out[bb * 2] = e0 / s; P.det_out[bb * 2 + 1] = e1 / s; }
          if (P.post) P.post[bb] = e1 / s;
        }
        sm.zmax[tid][0] = (int)0xff800000;   // -inf for the next group
        sm.zmax[tid][1] = (int)0xff800000;
      }
      epi_bar_sync();
    }
  } else {
    // =========================== MMA issuer + weight loader ===========================
    // Per block: all five gate GEMMs first (each as soon as its tile's U is ready), then the
    // five res/skip GEMMs (each as soon as its tile's g is ready).  rs(k,i) is therefore always
    // issued after gate(k,i+1), which reads the last rows of tile i's U through the row-shifted
    // taps and must finish before epilogue 2 of (k,i) overwrites them (tcgen05 ops of one thread
    // complete in order).  Tiles run staggered: while tile 4 is still in epilogue 1 of block k,
    // tile 0 is already in epilogue 2, which keeps the SFU and FMA pipes both busy.
    // The weight ring is refilled right after the gate GEMMs of block k are issued: by then
    // every tile has finished block k-1, so the other stage is free.
    const uint32_t idesc_gate = make_idesc_f16(128, 32), idesc_rs = make_idesc_f16(128, 48);
    const uint32_t hb = smem_u32(sm.head.det1_B);
    uint32_t n_u = 0, n_g = 0, n_w = 0, n_load = 0;
    uint32_t my_groups = 0;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) ++my_groups;
    const uint32_t total_loads = my_groups * 24;
    // prologue: fill WN_WST-1 stages
    for (; n_load < (uint32_t)(WN_WST - 1) && n_load < total_loads; ++n_load)
      if (lane == 0) {
        mbar_arrive_expect_tx(&sm.wfull[n_load % WN_WST], WN_WBLK);
        bulk_g2s(sm.W[n_load % WN_WST], P.wblob + (size_t)(n_load % 24) * WN_WBLK, WN_WBLK, &sm.wfull[n_load % WN_WST]);
      }
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      for (int k = 0; k < 24; ++k, ++n_w) {
        const int ws = n_w % WN_WST;
        mbar_wait(&sm.wfull[ws], (n_w / WN_WST) & 1);
        const int d = P.dil[k];
        const uint32_t wb = smem_u32(sm.W[ws]);
        for (int i = 0; i < WN_NT; ++i) {
          mbar_wait(&sm.bar_u[i], n_u & 1);
          fence_after_sync();
          if (elect_one()) {   // one lane issues the whole tile (uniform descriptors, no per-MMA election)
            const uint32_t dst = tmem + i * WN_TMEM_TILE;
            const uint64_t db = make_desc(wb, 512, 128);
#pragma unroll
            for (int tap = 0; tap < 3; ++tap) {
              const uint32_t arow = (uint32_t)(16 + i * 128 - (2 - tap) * d) * 16;
              const uint64_t ah = make_desc(uU + arow, WN_PU, 128), al = ah + (uint64_t)((2 * WN_PU) >> 4);
              const uint64_t bh = db + (uint64_t)((tap * 2 * 512) >> 4), bl = bh + (uint64_t)(3072 >> 4);
              mma_f16_ss(dst, ah, bh, idesc_gate, tap != 0);
              if (nsplit == 3) {
                mma_f16_ss(dst, al, bh, idesc_gate, true);
                mma_f16_ss(dst, ah, bl, idesc_gate, true);
              }
            }
            mma_commit(&sm.bar_gate[i]);
            if (i == 0) WN_DBG(5, k, 0);
            if (i == WN_NT - 1) WN_DBG(5, k, 1);
          }
          __syncwarp();
        }
        ++n_u;
        // all tiles have finished block k-1, so the stage that held its weights is free:
        // refill it with the block WN_WST-1 ahead (bulk copies take ~2.4 us, i.e. > one block)
        if (n_load < total_loads) {
          if (lane == 0) {
            mbar_arrive_expect_tx(&sm.wfull[n_load % WN_WST], WN_WBLK);
            bulk_g2s(sm.W[n_load % WN_WST], P.wblob + (size_t)(n_load % 24) * WN_WBLK, WN_WBLK, &sm.wfull[n_load % WN_WST]);
          }
          ++n_load;
        }
        for (int i = 0; i < WN_NT; ++i) {
          mbar_wait(&sm.bar_g[i], n_g & 1);
          fence_after_sync();
          if (elect_one()) {
            const uint32_t dst = tmem + i * WN_TMEM_TILE + 32;
            const uint32_t arow = (uint32_t)(i * 128) * 16;
            const uint64_t ah = make_desc(uG + arow, WN_PG, 128), al = ah + (uint64_t)((2 * WN_PG) >> 4);
            const uint64_t bh = make_desc(wb + WN_GATE_B, 768, 128), bl = bh + (uint64_t)(1536 >> 4);
            mma_f16_ss(dst, ah, bh, idesc_rs, false);
            if (nsplit == 3) {
              mma_f16_ss(dst, al, bh, idesc_rs, true);
              mma_f16_ss(dst, ah, bl, idesc_rs, true);
            }
            mma_commit(&sm.bar_rs[i]);
            if (i == 0) WN_DBG(5, k, 2);
            if (i == WN_NT - 1) WN_DBG(5, k, 3);
          }
          __syncwarp();
        }
        ++n_g;
      }
      // detect head: D[128,32] = ReLU(skip)[128,32] * W1^T ; k-step 0 from G panels, k-step 1 from U panels
      for (int i = 0; i < WN_NT; ++i) {
        mbar_wait(&sm.bar_u[i], n_u & 1);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t dst = tmem + i * WN_TMEM_TILE;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t abase = kk == 0 ? uG + (uint32_t)(i * 128) * 16 : uU + (uint32_t)(16 + i * 128) * 16;
            const uint32_t pl = kk == 0 ? WN_PG : WN_PU;
            const uint64_t ah = make_desc(abase, pl, 128), al = make_desc(abase + 2 * pl, pl, 128);
            const uint64_t bh = make_desc(hb + kk * 2 * 512, 512, 128), bl = make_desc(hb + 2048 + kk * 2 * 512, 512, 128);
            mma_f16_ss(dst, ah, bh, idesc_gate, kk != 0);
            if (nsplit == 3) {
              mma_f16_ss(dst, al, bh, idesc_gate, true);
              mma_f16_ss(dst, ah, bl, idesc_gate, true);
            }
          }
          mma_commit(&sm.bar_gate[i]);
        }
        __syncwarp();
      }
      ++n_u;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == WN_EPI_WARPS) tmem_dealloc(tmem, 512);
}

// ---- host side: pack the weights ------------------------------------------------------------
static void put_split(std::vector<unsigned char>& buf, size_t hi_off, size_t lo_off, float x, bool split) {
  __half h = __float2half_rn(x);
  __half l = split ? __float2half_rn(x - __half2float(h)) : __float2half_rn(0.f);
  memcpy(&buf[hi_off], &h, 2);
  memcpy(&buf[lo_off], &l, 2);
}

// gate_w [24][48][32] (k = tap*16+ch ; n), rs_w [24][16][48], biases, bn, dilation (fp32 device-layout copies
// made by api.cu on the host before upload)
std::vector<unsigned char> wavenet_pack_blocks(const float* gate_w, const float* gate_b, const float* rs_w,
                                               const float* rs_b, const float* bn_mul, const float* bn_add,
                                               const int* dilation) {
  std::vector<unsigned char> out((size_t)24 * WN_WBLK, 0);
  for (int b = 0; b < 24; ++b) {
    const size_t base = (size_t)b * WN_WBLK;
    for (int k = 0; k < 48; ++k)
      for (int n = 0; n < 32; ++n) {
        const int c = k / 8, e = k % 8;
        const size_t off = base + ((size_t)c * 32 + n) * 16 + e * 2;
        put_split(out, off, off + 3072, gate_w[((size_t)b * 48 + k) * 32 + n], true);
      }
    for (int k = 0; k < 16; ++k)
      for (int n = 0; n < 48; ++n) {
        const int c = k / 8, e = k % 8;
        const size_t off = base + WN_GATE_B + ((size_t)c * 48 + n) * 16 + e * 2;
        put_split(out, off, off + 1536, rs_w[((size_t)b * 16 + k) * 48 + n], true);
      }
    float f[113] = {0};
    for (int n = 0; n < 32; ++n)   // pre-scaled for gate_fn: tanh half by -2*log2(e), sigmoid half by -log2(e)
      f[n] = (float)((double)gate_b[b * 32 + n] * (n < 16 ? -2.8853900817779268 : -1.4426950408889634));
    for (int n = 0; n < 48; ++n) f[32 + n] = rs_b[b * 48 + n];
    if (b < 23)
      for (int c = 0; c < 16; ++c) { f[80 + c] = bn_mul[(b + 1) * 16 + c]; f[96 + c] = bn_add[(b + 1) * 16 + c]; }
    f[112] = (float)dilation[b];
    memcpy(&out[base + WN_GATE_B + WN_RS_B], f, sizeof(f));
  }
  return out;
}

std::vector<unsigned char> wavenet_pack_head(const float* in_w_kc, const float* in_b, const float* bn_mul0,
                                             const float* bn_add0, const float* det1_w_nk, const float* det1_b,
                                             const float* det2_w, const float* det2_b) {
  std::vector<unsigned char> out(sizeof(WnHead), 0);
  WnHead* h = reinterpret_cast<WnHead*>(out.data());
  memcpy(h->in_w, in_w_kc, sizeof(h->in_w));
  memcpy(h->in_b, in_b, sizeof(h->in_b));
  memcpy(h->bn0_mul, bn_mul0, sizeof(h->bn0_mul));
  memcpy(h->bn0_add, bn_add0, sizeof(h->bn0_add));
  memcpy(h->det1_b, det1_b, sizeof(h->det1_b));
  memcpy(h->det2_w, det2_w, sizeof(h->det2_w));
  memcpy(h->det2_b, det2_b, sizeof(h->det2_b));
  const size_t boff = offsetof(WnHead, det1_B);
  for (int n = 0; n < 32; ++n)
    for (int k = 0; k < 32; ++k) {
      const int c = k / 8, e = k % 8;
      const size_t off = boff + ((size_t)c * 32 + n) * 16 + e * 2;
      put_split(out, off, off + 2048, det1_w_nk[n * 32 + k], true);
    }
  return out;
}

int wavenet_tc_posteriors(wwb_ctx* ctx, const WinMap& wm, float* enc_out, float* det_out, float* post,
                          cudaStream_t st) {
  if (wm.n_win == 0) return WWB_OK;
  if (ctx->L > 182) return fail(ctx, WWB_ERR_ARG, "tensor-core WaveNet path supports windows up to 182 frames");
  WnTcParams P;
  P.wm = wm;
  P.wblob = ctx->wn.tc_blocks;
  P.head = reinterpret_cast<const WnHead*>(ctx->wn.tc_head);
  P.L = ctx->L;
  P.nsplit = ctx->precision == WWB_PREC_TC ? 3 : 1;
  for (int b = 0; b < 24; ++b) P.dil[b] = ctx->wn.dilation[b];
  P.enc_out = enc_out; P.det_out = det_out; P.post = post;
  P.dbg = reinterpret_cast<long long*>(ctx->debug_buf);
  const size_t smem = sizeof(WnSmem) + 128;
  WWB_CUDA(ctx, cudaFuncSetAttribute(wavenet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t n_groups = (wm.n_win + WN_G - 1) / WN_G;
  const unsigned grid = (unsigned)std::min<int64_t>(n_groups, ctx->sm_count);
  wavenet_tc_kernel<<<grid, WN_THREADS, smem, st>>>(P);
  WWB_CHECK_LAUNCH(ctx);
  return WWB_OK;
}

}  // namespace wwb
