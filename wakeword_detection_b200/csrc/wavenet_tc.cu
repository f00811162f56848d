// K5/K6 — WaveNet encode + detect on the tensor cores (precision WWB_PREC_TC / TC_FAST).
//
// One persistent CTA per SM processes groups of WN_G = 3 windows.  The rows of a group
// (window w, time t) -> o = w*198 + t form the M dimension of every GEMM (5 tiles of 128
// rows; the 16 rows after each window are the next window's causal zero padding).  Each of
// the 640 epilogue threads owns ONE row for the whole 24-block stack: its 16-channel
// residual stream and 32-channel skip sum never leave registers.  Per block and tile:
//   gate GEMM   D[128,32] = sum_tap U[rows - (2-tap)*d, 16] * Wg_tap          (tcgen05, TMEM)
//               taps 0/1 (row-shifted) read A from shared memory: the operand layout is linear
//               in the row index, so a dilated tap is a start-address offset;
//               tap 2 (no shift) reads A straight from TENSOR MEMORY, where the row's owner
//               thread stored it (tcgen05.st): 15.5 clk per MMA instead of 39.4 (tools/tc_rate_probe.cu)
//   epilogue 1  g = tanh(.)*sigmoid(.) (2 ex2 per channel, 1 rcp per channel pair), fp16 hi/lo -> TMEM, over the
//               consumed gate accumulator (A operand of the next GEMM)
//   res/skip    D[128,48] = g[128,16] * [Wres | Wskip]       (A from TMEM, own accumulator columns)
//   epilogue 2a x += ReLU(res); u = BN_next(x) -> hi/lo -> shared memory + TMEM   (releases the next gate GEMM)
//   epilogue 2b skip += ReLU(.)                                                   (overlaps that GEMM)
// The input 1x1 conv depends only on the mel row, which ~91 overlapping windows share: wn_input_kernel computes it once
// per row (fp32) and a group starts by loading its rows' 16 values.  The detect head (32->32 with A from TMEM; then
// 32->2, max over time, softmax) is a tensor-core GEMM of the same kind as the blocks' at the end of a group.
// Biases are added by the tensor core too: one extra k-step whose A chunk is the constant (1, 1, 0, ...)
// and whose B rows hold (bias_hi, bias_lo, 0, ...) - broadcast loads of per-block constants from shared
// memory cost one wavefront per 4 bytes and were the largest shared-memory consumer (profiles/).
// The gate weights and biases are pre-scaled by -2*log2(e) (tanh half) / -log2(e) (sigmoid half).
// The per-block BatchNorm affine u = bn_mul * x + bn_add is FOLDED into the gate GEMM (weights * bn_mul, bias +
// W * bn_add), so the A operand is the residual stream x itself and the epilogue neither loads the 32 constants
// (8 broadcast LDS.128 = 32 shared-memory wavefronts per warp and block, 17 % of the shared-memory pipe) nor
// applies them.  The causal zero padding is of u, not x: the 16 padding rows in front of every window therefore
// hold x_pad = -bn_add / bn_mul (so that u_pad = 0), rewritten for each block by the threads that own those rows.
// fp16 hi/lo operand split (3 MMAs per product) keeps the result at fp32 accuracy
// (DESIGN.md §precision).  Epilogue arithmetic uses the packed fp32x2 instructions (FFMA2/FADD2).
// Per-block weights (12 KB incl. the bias operands) stream through a 4-stage cp.async.bulk ring.
//
// GEMM issue: one extra warp issues the gate GEMMs tile after tile (in order, so they run back to back and
// stagger the tiles: the tensor pipe, the SFU and the FMA/ALU pipes then work on different tiles at the
// same time); a second extra warp issues the small res/skip GEMMs, so they never queue behind a wait of the
// gate warp.  Epilogue 2 is split: 2a (residual -> u) releases the next gate GEMM, 2b (skip sum) overlaps it.
// Ordering rules:
//   gate(k,i) needs epilogue 2 of (k-1,i) and (k-1,i-1)      (its taps reach 16 rows into tile i-1)
//   rs(k,i)   needs epilogue 1 of (k,i) and gate(k,i+1) COMPLETE: epilogue 2 of (k,i) overwrites rows that
//             GEMM reads, and GEMMs issued by different threads have no implicit order.
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
constexpr int WN_EPI_WARPS = WN_NT * 4;       // 20
constexpr int WN_EPI_THREADS = WN_EPI_WARPS * 32;
constexpr int WN_THREADS = (WN_EPI_WARPS + 2) * 32;   // + the gate-GEMM / loader warp + the res/skip-GEMM warp = 704
constexpr int WN_GATE_B = 6144, WN_RS_B = 3072;   // gate / res+skip B operands (hi and lo planes)
constexpr int WN_F32_B = 512;                     // misc: bytes [320, 384) = the NEXT block's padding rows (hi0, hi1, lo0, lo1)
constexpr int WN_GBIAS_B = 1024, WN_RBIAS_B = 1536;   // bias B operands for the 'ones' GEMM (k0 = hi, k1 = lo)
constexpr int WN_OFF_F32 = WN_GATE_B + WN_RS_B, WN_OFF_GBIAS = WN_OFF_F32 + WN_F32_B, WN_OFF_RBIAS = WN_OFF_GBIAS + WN_GBIAS_B;
constexpr int WN_WBLK = WN_OFF_RBIAS + WN_RBIAS_B;    // 12288 bytes of one block's weight blob
constexpr int WN_WST = 4;                     // weight ring stages
constexpr int WN_TMEM_TILE = 96;              // TMEM columns per tile:
constexpr int WN_C_G = 0;                     //   0..31  gate accumulator; g hi/lo (A of res/skip) overwrites 0..15 once epilogue 1 has read it
constexpr int WN_C_R = 32;                    //   32..47 res accumulator, 48..79 skip accumulator (detect: 32..63)
constexpr int WN_C_U = 80;                    //   80..95 u hi/lo (A of the gate's unshifted tap)
constexpr int WN_C_ONE = WN_NT * WN_TMEM_TILE;   // 480..487: constant A chunk (k0 = k1 = 1, rest 0), shared by all tiles: adds the biases

// resident head blob (floats unless noted)
struct WnHead {
  unsigned char pad0[64];   // block 0's padding rows (see the BN note in the header): hi chunk 0, hi chunk 1, lo chunk 0, lo chunk 1
  unsigned char rsv_[64];
  float det1_b[32];
  float det2_w[2 * 32];
  float det2_b[2];
  float pad_[2];
  unsigned char det1_B[2 * 4 * 32 * 16];   // hi/lo planes, 4 chunks x 32 rows x 16 B
};

struct WnSmem {
  unsigned char U[2 * 2 * WN_PU];          // [plane][chunk][row]
  unsigned char W[WN_WST][WN_WBLK];
  WnHead head;
  uint64_t bar_gate[WN_NT], bar_rs[WN_NT], bar_det[WN_NT];   // GEMM completion (tcgen05.commit) -> the tile's four warps
  uint64_t bar_u[WN_NT];                     // the tile's four warps finished epilogue 2 -> gate warp
  uint64_t bar_g[WN_NT];                     // the tile's four warps finished epilogue 1 -> res/skip warp
  uint64_t wfull[WN_WST];
  uint32_t tmem_base;
  int zmax[2][WN_G][2];                      // per-window max of the two logits, double-buffered by group parity
};

struct WnTcParams {
  WinMap wm;
  const unsigned char* wblob;   // [24][WN_WBLK]
  const WnHead* head;
  const float* x0;              // [n_streams * ring][16]: ReLU(in_w * mel + in_b) of every mel row (wn_input_kernel)
  int L;
  int nsplit;
  int dil[24];                  // dilation per block (kernel-parameter space keeps it in uniform registers)
  float* enc_out;
  float* det_out;
  float* post;
  long long* dbg;   // optional timeline dump (block 0, second group): [8 roles][2 groups x 24 blocks][4 events]
};

#define WN_DBG(role, k, ev) do { if (P.dbg && blockIdx.x == 0 && (grp == (int64_t)gridDim.x || grp == 2 * (int64_t)gridDim.x)) P.dbg[((role) * 48 + (k) + (grp == (int64_t)gridDim.x ? 0 : 24)) * 4 + (ev)] = clock64(); } while (0)

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
__device__ __forceinline__ u64 relu2(u64 v) { float a, b; upk(v, a, b); return pk(fmaxf(a, 0.f), fmaxf(b, 0.f)); }
// acc + ReLU(v) for a pair in two packed instructions: v + |v| = 2 max(v, 0) exactly (FADD2 takes |.| as an operand
// modifier), and fma(., 0.5, acc) rounds once - bit-identical to acc + max(v, 0), one instruction less than
// 2 FMNMX + FADD2, and on the FMA pipe instead of the busier ALU pipe.
__device__ __forceinline__ u64 add_relu2(u64 acc, u64 v) {
  float a, b;
  upk(v, a, b);
  return ffma2(fadd2(v, pk(fabsf(a), fabsf(b))), pk(0.5f, 0.5f), acc);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// order-preserving float <-> int key (signed compare): the per-window maxima are kept as keys, so one warp-wide
// integer max-reduction (REDUX) and one atomicMax per warp replace 32 serialised shared-memory atomics
__device__ __forceinline__ int f2key(float v) { const int i = __float_as_int(v); return i ^ ((i >> 31) & 0x7fffffff); }
__device__ __forceinline__ float key2f(int k) { return __int_as_float(k ^ ((k >> 31) & 0x7fffffff)); }
constexpr int WN_KEY_NEG_INF = (int)0x807fffff;   // f2key(-inf)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(WN_EPI_THREADS) : "memory"); }

// ---- hang diagnosis (compile with -DWWB_HANG_DEBUG and pass a debug buffer): instead of trapping, the first
// thread whose wait times out dumps its position and the shared progress counters behind the timeline
// area of the debug buffer, and every timed-out thread exits, so the kernel ends and the host can read it.
#ifdef WWB_HANG_DEBUG
#define WN_SPIN_LIMIT (1u << 17)
__device__ __noinline__ void wn_hang(long long* dbg, const uint32_t* cnt_u, const uint32_t* cnt_g, int id, int tile, int q,
                                     uint32_t n_gate, uint32_t n_rs, uint32_t n_u, uint32_t n_w) {
  if (dbg) {
    long long* base = dbg + 8 * 48 * 4;
    if (atomicCAS(reinterpret_cast<unsigned long long*>(base), 0ull, 1ull) == 0ull) {
      base[1] = id; base[2] = blockIdx.x; base[3] = tile; base[4] = q; base[5] = n_gate; base[6] = n_rs;
      base[7] = n_u; base[8] = n_w;
      if (cnt_u) for (int i = 0; i < 5; ++i) { base[9 + i] = cnt_u[i]; base[14 + i] = cnt_g[i]; }
    }
  }
  asm volatile("exit;");
}
#define WN_HANG(id) wn_hang(P.dbg, nullptr, nullptr, id, tile, q, n_gate, n_rs, n_u, n_w)
#define WN_MBAR_WAIT(bar, par, id) do { if (!mbar_try_wait(bar, par)) { uint32_t sp_ = 0; while (!mbar_try_wait(bar, par)) if (++sp_ > WN_SPIN_LIMIT) WN_HANG(id); } } while (0)
#else
#define WN_MBAR_WAIT(bar, par, id) mbar_wait(bar, par)
#endif

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
  for (int i = tid; i < (int)(sizeof(WnHead) / 16); i += WN_THREADS)
    reinterpret_cast<uint4*>(&sm.head)[i] = reinterpret_cast<const uint4*>(P.head)[i];
  if (tid < 2 * WN_G * 2) (&sm.zmax[0][0][0])[tid] = WN_KEY_NEG_INF;
  if (tid == 0) {
    for (int i = 0; i < WN_NT; ++i) {
      mbar_init(&sm.bar_u[i], 4); mbar_init(&sm.bar_g[i], 4); mbar_init(&sm.bar_gate[i], 1); mbar_init(&sm.bar_rs[i], 1);
      mbar_init(&sm.bar_det[i], 1);
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
  const int nsplit = P.nsplit;
  const uint32_t idesc_gate = make_idesc_f16(128, 32), idesc_rs = make_idesc_f16(128, 48);

  if (warp < WN_EPI_WARPS) {
    // =========================== epilogue threads: one row each ===========================
    const int tile = warp >> 2, q = warp & 3;
    const int o = tile * 128 + q * 32 + lane;       // row within the group
    const int w = o / WN_SLOT, t = o - w * WN_SLOT;
    const uint32_t tacc = tmem + tile * WN_TMEM_TILE;                 // this tile's TMEM columns (lane 0)
    const uint32_t tbase = tacc + ((uint32_t)(q * 32) << 16);         // ... seen from this warp's lane quadrant
    uint32_t n_gate = 0, n_rs = 0, n_w = 0, n_u = 0;  // completed phases of bar_gate / bar_rs ; global block index ; groups done
    unsigned char* const Urow = sm.U + (16 + o) * 16;
    // padding rows: the 16 rows after each window (they precede the next one; owned by otherwise idle threads) and,
    // for window 0, U rows 0..15, which the threads of rows 0..15 write in addition to their own row (they belong to
    // tile 0, whose gate GEMM is the reader, so the tile's arrival barrier orders the write)
    const bool pad_row = (w < WN_G && t >= 182) || (o < 16);
    unsigned char* const Upad = (o < 16) ? sm.U + o * 16 : Urow;
    auto store_pad = [&](const unsigned char* chunk) {   // chunk: 64 bytes (hi0, hi1, lo0, lo1)
      const uint4* c4 = reinterpret_cast<const uint4*>(chunk);
      *reinterpret_cast<uint4*>(Upad) = c4[0];
      *reinterpret_cast<uint4*>(Upad + WN_PU) = c4[1];
      *reinterpret_cast<uint4*>(Upad + 2 * WN_PU) = c4[2];
      *reinterpret_cast<uint4*>(Upad + 3 * WN_PU) = c4[3];
    };
    {
      uint32_t one[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) one[i] = 0u;
      one[0] = 0x3c003c00u;   // (1.0h, 1.0h)
      tmem_st16(tmem + ((uint32_t)(q * 32) << 16) + WN_C_ONE, one);
      tmem_st_wait();
    }
    const u64 ZERO2 = pk(0.f, 0.f), ONE2 = pk(1.f, 1.f);

    // The input layer x0 = ReLU(in_w * mel + in_b) depends only on the mel row, and every mel row is shared by ~91
    // overlapping windows: it is computed once per row by wn_input_kernel (64 B per row instead of 160 B of mel per row
    // and window), which also takes a tcgen05.st / GEMM / tcgen05.ld round trip off the group-boundary chain.  The row
    // of the NEXT group is fetched into registers before this group's detect epilogue (x / skip are dead by then), so
    // its global-memory latency is not on that chain either.
    float4 xrow[4];
    auto fetch_x0 = [&](int64_t grp_next) {
      const int64_t bn = grp_next * WN_G + w;
      const bool vn = (grp_next < n_groups) && (w < WN_G) && (t < L) && (bn < n_win);
      int64_t s0 = 0;
      int start = 0;
      if (vn) win_origin(P.wm, bn, s0, start);
      int rr = start + (vn ? t : 0);
      if (rr >= P.wm.ring) rr -= P.wm.ring;
      const float4* row = reinterpret_cast<const float4*>(P.x0 + (s0 * P.wm.ring + rr) * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) xrow[i] = vn ? __ldg(row + i) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    int64_t fin_grp = -1;
    int fin_zp = 0;
    auto finalise = [&]() {
      if (fin_grp >= 0 && tid < WN_G) {
        const int64_t bb = fin_grp * WN_G + tid;
        if (bb < n_win) {
          const float a0 = key2f(sm.zmax[fin_zp][tid][0]), a1 = key2f(sm.zmax[fin_zp][tid][1]);
          const float m = fmaxf(a0, a1);
          const float e0 = expf(a0 - m), e1 = expf(a1 - m), s = e0 + e1;
          if (P.det_out) { P.det_out[bb * 2] = e0 / s; P.det_out[bb * 2 + 1] = e1 / s; }
          if (P.post) P.post[bb] = e1 / s;
        }
        sm.zmax[fin_zp][tid][0] = WN_KEY_NEG_INF;   // for the group after next
        sm.zmax[fin_zp][tid][1] = WN_KEY_NEG_INF;
      }
      fin_grp = -1;
    };
    if ((int64_t)blockIdx.x < n_groups) fetch_x0(blockIdx.x);
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ++n_u) {
      const int64_t b = grp * WN_G + w;
      const bool valid = (w < WN_G) && (t < L) && (b < n_win);
      u64 x[8], skip[16];     // channel pairs
      // ---- input layer: x = x0 row (prefetched); block 0's BatchNorm is folded into its gate weights ----
      {
        if (tile == 0 && q == 0 && lane == 0) WN_DBG(6, 21, 0);
        finalise();   // previous group's posteriors
#pragma unroll
        for (int n = 0; n < 16; ++n) skip[n] = ZERO2;
        uint32_t ur[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          x[2 * i] = pk(xrow[i].x, xrow[i].y);
          x[2 * i + 1] = pk(xrow[i].z, xrow[i].w);
          split2(x[2 * i], ur[2 * i], ur[8 + 2 * i]);
          split2(x[2 * i + 1], ur[2 * i + 1], ur[8 + 2 * i + 1]);
        }
        if (tile == 0 && q == 0 && lane == 0) WN_DBG(6, 21, 1);
        if (valid) {
          *reinterpret_cast<uint4*>(Urow) = make_uint4(ur[0], ur[1], ur[2], ur[3]);
          *reinterpret_cast<uint4*>(Urow + WN_PU) = make_uint4(ur[4], ur[5], ur[6], ur[7]);
          *reinterpret_cast<uint4*>(Urow + 2 * WN_PU) = make_uint4(ur[8], ur[9], ur[10], ur[11]);
          *reinterpret_cast<uint4*>(Urow + 3 * WN_PU) = make_uint4(ur[12], ur[13], ur[14], ur[15]);
        }
        if (pad_row) store_pad(sm.head.pad0);
        tmem_st16(tbase + WN_C_U, ur);
        tmem_st_wait();
        fence_before_sync();
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_u[tile]);
      }

      for (int k = 0; k < 24; ++k, ++n_w) {
        const int ws = n_w % WN_WST;
        const float* wf = reinterpret_cast<const float*>(sm.W[ws] + WN_OFF_F32);
        WN_MBAR_WAIT(&sm.wfull[ws], (n_w / WN_WST) & 1, 4);   // BN constants travel with the block's weights (off the critical path: the gate GEMM is running)
        // ---- epilogue 1: gated activation ----
        WN_MBAR_WAIT(&sm.bar_gate[tile], n_gate & 1, 5);
        ++n_gate;
        fence_after_sync();
        if (q == 0 && lane == 0) WN_DBG(tile, k, 0);
        {
          uint32_t gr[16];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float at[8], as[8];
            tmem_ld8(tbase + h8 * 8, at);
            tmem_ld8(tbase + 16 + h8 * 8, as);
            tmem_ld_wait();
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              // accumulators arrive as a2 = -2*log2(e)*(a + b_t), b2 = -log2(e)*(b + b_s) (scaled weights, bias GEMM)
              const float a0 = at[2 * p], a1 = at[2 * p + 1], b0 = as[2 * p], b1 = as[2 * p + 1];
              // exponents clamped at 30: tanh is -1 and sigmoid 0 to 1e-9 beyond, and both denominators stay below
              // 2^61, so ONE reciprocal serves the pair of channels (1/d0 = d1/(d0 d1)): 5 SFU ops per pair instead
              // of 6 - the SFU queue is what stretches this epilogue (profiles/r1_notes.md)
              const float ea0 = ex2_approx(fminf(a0, 30.f)), ea1 = ex2_approx(fminf(a1, 30.f));
              const float eb0 = ex2_approx(fminf(b0, 30.f)), eb1 = ex2_approx(fminf(b1, 30.f));
              const u64 tt = fadd2(pk(eb0, eb1), ONE2);
              float d0, d1;
              upk(ffma2(pk(ea0, ea1), tt, tt), d0, d1);                 // (1 + e^-2a)(1 + e^-b)
              const float rp = rcp_approx(d0 * d1);
              const float r0 = rp * d1, r1 = rp * d0;
              const u64 g = pk(fmaf(-ea0, r0, r0), fmaf(-ea1, r1, r1)); // tanh(a) * sigmoid(b)
              split2(g, gr[h8 * 4 + p], gr[8 + h8 * 4 + p]);
            }
          }
          tmem_st16(tbase + WN_C_G, gr);
          tmem_st_wait();
        }
        fence_before_sync();
        if (q == 0 && lane == 0) WN_DBG(tile, k, 1);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.bar_g[tile]);

        // ---- epilogue 2a: residual, next block's BN -> u (releases the next gate GEMM) ----
        WN_MBAR_WAIT(&sm.bar_rs[tile], n_rs & 1, 6);
        ++n_rs;
        fence_after_sync();
        if (q == 0 && lane == 0) WN_DBG(tile, k, 2);
        const bool last = (k == 23);
        if (!last) {
          uint32_t ur[16];
#pragma unroll
          for (int h8 = 0; h8 < 2; ++h8) {
            float r[8];
            tmem_ld8(tbase + WN_C_R + h8 * 8, r);
            tmem_ld_wait();
#pragma unroll
            for (int p = 0; p < 4; ++p) {
              const int c = h8 * 4 + p;
              x[c] = add_relu2(x[c], pk(r[2 * p], r[2 * p + 1]));
              split2(x[c], ur[c], ur[8 + c]);   // the next block's BN is folded into its gate weights
            }
          }
          if (valid) {
            *reinterpret_cast<uint4*>(Urow) = make_uint4(ur[0], ur[1], ur[2], ur[3]);
            *reinterpret_cast<uint4*>(Urow + WN_PU) = make_uint4(ur[4], ur[5], ur[6], ur[7]);
            *reinterpret_cast<uint4*>(Urow + 2 * WN_PU) = make_uint4(ur[8], ur[9], ur[10], ur[11]);
            *reinterpret_cast<uint4*>(Urow + 3 * WN_PU) = make_uint4(ur[12], ur[13], ur[14], ur[15]);
          }
          if (pad_row) store_pad(reinterpret_cast<const unsigned char*>(wf) + 320);
          tmem_st16(tbase + WN_C_U, ur);
          tmem_st_wait();
          fence_before_sync();
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.bar_u[tile]);
          if (q == 0 && lane == 0) WN_DBG(tile, k, 3);
        }
        // ---- epilogue 2b: skip sum (off the critical path: the next gate GEMM is already running) ----
#pragma unroll
        for (int h8 = 0; h8 < 4; ++h8) {
          float s[8];
          tmem_ld8(tbase + WN_C_R + 16 + h8 * 8, s);
          tmem_ld_wait();
#pragma unroll
          for (int p = 0; p < 4; ++p) skip[h8 * 4 + p] = add_relu2(skip[h8 * 4 + p], pk(s[2 * p], s[2 * p + 1]));
        }
        fence_before_sync();
        if (last) {
          // detect input: ReLU(skip) hi/lo; channels 0-15 -> the u columns, 16-31 -> the g columns
          uint32_t er[16];
#pragma unroll
          for (int c = 0; c < 8; ++c) split2(relu2(skip[c]), er[c], er[8 + c]);
          tmem_st16(tbase + WN_C_U, er);
#pragma unroll
          for (int c = 0; c < 8; ++c) split2(relu2(skip[8 + c]), er[c], er[8 + c]);
          tmem_st16(tbase + WN_C_G, er);
          tmem_st_wait();
          fence_before_sync();
          __syncwarp();
          if (lane == 0) mbar_arrive(&sm.bar_u[tile]);
          if (q == 0 && lane == 0) WN_DBG(tile, k, 3);
        }
      }

      if (P.enc_out && valid) {
        float4* dst = reinterpret_cast<float4*>(P.enc_out + (b * L + t) * 32);
#pragma unroll
        for (int n = 0; n < 8; ++n) {
          float s0, s1, s2, s3;
          upk(skip[2 * n], s0, s1);
          upk(skip[2 * n + 1], s2, s3);
          dst[n] = make_float4(s0, s1, s2, s3);
        }
      }
      if (tile == 0 && q == 0 && lane == 0) WN_DBG(6, 20, 0);   // boundary timeline: e2b(23) done
      fetch_x0(grp + gridDim.x);   // next group's input row: in flight during the detect epilogue
      // ---- detect head epilogue: ReLU(D + b1) -> 32->2 -> max over time ----
      WN_MBAR_WAIT(&sm.bar_det[tile], n_u & 1, 8);
      fence_after_sync();
      if (tile == 0 && q == 0 && lane == 0) WN_DBG(6, 20, 1);   // detect GEMM done
      // 32 -> 2 with 128-bit loads of the constants and four independent partial sums per logit (as 96 scalar loads
      // feeding two 32-long dependent FMA chains this epilogue took ~2500 clk on the group-boundary critical path)
      float z0, z1;
      {
        u64 acc0[2] = {pk(0.f, 0.f), pk(0.f, 0.f)}, acc1[2] = {pk(0.f, 0.f), pk(0.f, 0.f)};
#pragma unroll
        for (int h8 = 0; h8 < 4; ++h8) {
          float d[8];
          tmem_ld8(tbase + WN_C_R + h8 * 8, d);
          const ulonglong2* b1 = reinterpret_cast<const ulonglong2*>(sm.head.det1_b + h8 * 8);
          const ulonglong2* w0 = reinterpret_cast<const ulonglong2*>(sm.head.det2_w + h8 * 8);
          const ulonglong2* w1 = reinterpret_cast<const ulonglong2*>(sm.head.det2_w + 32 + h8 * 8);
          const ulonglong2 b1a = b1[0], b1b = b1[1], w0a = w0[0], w0b = w0[1], w1a = w1[0], w1b = w1[1];
          tmem_ld_wait();
          const u64 e0 = relu2(fadd2(pk(d[0], d[1]), b1a.x)), e1 = relu2(fadd2(pk(d[2], d[3]), b1a.y));
          const u64 e2 = relu2(fadd2(pk(d[4], d[5]), b1b.x)), e3 = relu2(fadd2(pk(d[6], d[7]), b1b.y));
          acc0[0] = ffma2(w0a.x, e0, acc0[0]); acc0[1] = ffma2(w0a.y, e1, acc0[1]);
          acc0[0] = ffma2(w0b.x, e2, acc0[0]); acc0[1] = ffma2(w0b.y, e3, acc0[1]);
          acc1[0] = ffma2(w1a.x, e0, acc1[0]); acc1[1] = ffma2(w1a.y, e1, acc1[1]);
          acc1[0] = ffma2(w1b.x, e2, acc1[0]); acc1[1] = ffma2(w1b.y, e3, acc1[1]);
        }
        float a, b;
        upk(fadd2(acc0[0], acc0[1]), a, b);
        z0 = (a + b) + sm.head.det2_b[0];
        upk(fadd2(acc1[0], acc1[1]), a, b);
        z1 = (a + b) + sm.head.det2_b[1];
      }
      fence_before_sync();
      const int zp = (int)(n_u & 1);
      {
        // a warp's 32 rows touch at most two windows (slot = 198 rows): one reduction + one atomic per window
        const int w_first = __shfl_sync(0xffffffffu, w, 0);
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
          const bool mine = valid && (w == w_first + dw);
          const unsigned m = __ballot_sync(0xffffffffu, mine);
          if (m) {
            const int k0 = __reduce_max_sync(0xffffffffu, mine ? f2key(z0) : WN_KEY_NEG_INF);
            const int k1 = __reduce_max_sync(0xffffffffu, mine ? f2key(z1) : WN_KEY_NEG_INF);
            if (lane == 0) {
              atomicMax(&sm.zmax[zp][w_first + dw][0], k0);
              atomicMax(&sm.zmax[zp][w_first + dw][1], k1);
            }
          }
        }
      }
      if (tile == 0 && q == 0 && lane == 0) WN_DBG(6, 20, 2);   // detect epilogue done
      epi_bar_sync();   // the only barrier per group: zmax is double-buffered, and its reset (in finalise, which runs
                        // before the barrier of group g+1) is ordered before the atomics of group g+2
      fin_grp = grp;    // softmax + stores of this group's posteriors: deferred into the next group's input-GEMM wait
      fin_zp = zp;
      if (tile == 0 && q == 0 && lane == 0) WN_DBG(6, 20, 3);   // barrier done
    }
    finalise();
  } else if (warp == WN_EPI_WARPS) {
    // =========================== gate-GEMM warp + weight loader ===========================
    // Issues the gate GEMMs tile after tile, each as soon as the tile's epilogue 2 has arrived.  Because one
    // thread issues them, they execute back to back in tile order, which staggers the tiles: while tile i+1's
    // gate GEMM runs, tile i is in epilogue 1 (SFU), tile i-1 in its res/skip GEMM or epilogue 2 (FMA/ALU).
    // gate(k,i) also needs epilogue 2 of (k-1,i-1) (its taps reach 16 rows into tile i-1): awaited one step earlier.
    const int tile = 0, q = 0;   // (for the hang report)
    uint32_t n_gate = 0, n_rs = 0, n_u = 0;
    const uint64_t dU = make_desc(smem_u32(sm.U), WN_PU, 128);                   // U row 0, hi plane
    const uint64_t dWg = make_desc(smem_u32(sm.W[0]), 512, 128);                 // gate B: stage 0, tap 0, hi plane
    const uint64_t dBg = make_desc(smem_u32(sm.W[0]) + WN_OFF_GBIAS, 512, 128);  // gate bias B: stage 0
    const uint64_t dH = make_desc(smem_u32(sm.head.det1_B), 512, 128);           // detect B: k-step 0, hi plane
    uint32_t my_groups = 0;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) ++my_groups;
    const uint32_t total_loads = my_groups * 24;
    for (uint32_t n = 0; n < (uint32_t)WN_WST && n < total_loads; ++n)
      if (lane == 0) {
        mbar_arrive_expect_tx(&sm.wfull[n], WN_WBLK);
        bulk_g2s(sm.W[n], P.wblob + (size_t)n * WN_WBLK, WN_WBLK, &sm.wfull[n]);
      }
    uint32_t n_w = 0;         // global block index (24 per group)
    uint32_t ubase = 0;       // bar_u phases before this group (25 per group)
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x, ubase += 25, ++n_u) {
      for (int k = 0; k < 24; ++k, ++n_w) {
        WN_MBAR_WAIT(&sm.wfull[n_w % WN_WST], (n_w / WN_WST) & 1, 1);
        const uint32_t d = (uint32_t)P.dil[k];
        const uint64_t wofs = (uint64_t)((n_w % WN_WST) * (WN_WBLK >> 4));
        const uint64_t b0 = dWg + wofs, bb = dBg + wofs;
#pragma unroll
        for (int i = 0; i < WN_NT; ++i) {
          if (i == 1) WN_DBG(6, k, 0);
          WN_MBAR_WAIT(&sm.bar_u[i], (ubase + k) & 1, 2);
          if (i == 1) WN_DBG(6, k, 1);
          fence_after_sync();
          if (elect_one()) {   // one lane issues the whole tile (uniform descriptors, no per-MMA election)
            const uint32_t tacc = tmem + i * WN_TMEM_TILE;
            const uint64_t a0 = dU + (uint64_t)(16 + i * 128 - 2 * d), a1 = dU + (uint64_t)(16 + i * 128 - d);
            const uint64_t lo_a = (uint64_t)((2 * WN_PU) >> 4), lo_b = (uint64_t)(3072 >> 4);
            mma_f16_ss(tacc, a0, b0, idesc_gate, false);
            if (nsplit == 3) {
              mma_f16_ss(tacc, a0 + lo_a, b0, idesc_gate, true);
              mma_f16_ss(tacc, a0, b0 + lo_b, idesc_gate, true);
            }
            mma_f16_ss(tacc, a1, b0 + 64, idesc_gate, true);
            if (nsplit == 3) {
              mma_f16_ss(tacc, a1 + lo_a, b0 + 64, idesc_gate, true);
              mma_f16_ss(tacc, a1, b0 + 64 + lo_b, idesc_gate, true);
            }
            mma_f16_ts(tacc, tacc + WN_C_U, b0 + 128, idesc_gate, true);
            if (nsplit == 3) {
              mma_f16_ts(tacc, tacc + WN_C_U + 8, b0 + 128, idesc_gate, true);
              mma_f16_ts(tacc, tacc + WN_C_U, b0 + 128 + lo_b, idesc_gate, true);
            }
            mma_f16_ts(tacc, tmem + WN_C_ONE, bb, idesc_gate, true);
            mma_commit(&sm.bar_gate[i]);
            if (i == 0) WN_DBG(5, k, 0);
            if (i == WN_NT - 1) WN_DBG(5, k, 1);
          }
          __syncwarp();
          if (i == 1) WN_DBG(6, k, 2);
        }
        // every tile has finished block n_w-1 (its epilogue 2 was awaited above): refill that stage
        if (n_w >= 1 && n_w - 1 + WN_WST < total_loads) {
          const uint32_t nl = n_w - 1 + WN_WST;
          if (lane == 0) {
            mbar_arrive_expect_tx(&sm.wfull[nl % WN_WST], WN_WBLK);
            bulk_g2s(sm.W[nl % WN_WST], P.wblob + (size_t)(nl % 24) * WN_WBLK, WN_WBLK, &sm.wfull[nl % WN_WST]);
          }
        }
      }
      // detect head: D[128,32] = ReLU(skip)[128,32] * W1^T ; k-step 0 from the u columns, k-step 1 from the g columns
#pragma unroll
      for (int i = 0; i < WN_NT; ++i) {
        WN_MBAR_WAIT(&sm.bar_u[i], (ubase + 24) & 1, 9);
        fence_after_sync();
        if (elect_one()) {
          const uint32_t tacc = tmem + i * WN_TMEM_TILE;
#pragma unroll
          for (int kk = 0; kk < 2; ++kk) {
            const uint32_t ta = tacc + (kk == 0 ? WN_C_U : WN_C_G);
            const uint64_t bh = dH + (uint64_t)((kk * 2 * 512) >> 4), bl = bh + (uint64_t)(2048 >> 4);
            mma_f16_ts(tacc + WN_C_R, ta, bh, idesc_gate, kk != 0);
            if (nsplit == 3) {
              mma_f16_ts(tacc + WN_C_R, ta + 8, bh, idesc_gate, true);
              mma_f16_ts(tacc + WN_C_R, ta, bl, idesc_gate, true);
            }
          }
          mma_commit(&sm.bar_det[i]);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================== res/skip-GEMM warp ===========================
    // Tile after tile: when the tile's four warps have stored g (epilogue 1), issue its res and skip GEMMs
    // (A = g from TMEM, + bias k-step).  Epilogue 2a of the tile will then overwrite rows that gate(k) of
    // tile+1 reads, so that GEMM must have COMPLETED first (it is issued by another thread: no implicit order).
    const int tile = 1, q = 0;   // (for the hang report)
    uint32_t n_gate = 0, n_rs = 0, n_u = 0, n_w = 0;
    const uint64_t dWr = make_desc(smem_u32(sm.W[0]) + WN_GATE_B, 768, 128);     // res/skip B: stage 0, hi plane
    const uint64_t dBr = make_desc(smem_u32(sm.W[0]) + WN_OFF_RBIAS, 768, 128);  // res/skip bias B: stage 0
    const uint32_t ones = tmem + WN_C_ONE;
    for (int64_t grp = blockIdx.x; grp < n_groups; grp += gridDim.x) {
      for (int k = 0; k < 24; ++k, ++n_w) {
        const uint64_t wofs = (uint64_t)((n_w % WN_WST) * (WN_WBLK >> 4));
        const uint64_t bh = dWr + wofs, bb = dBr + wofs, lo_b = (uint64_t)(1536 >> 4);
        WN_MBAR_WAIT(&sm.wfull[n_w % WN_WST], (n_w / WN_WST) & 1, 11);   // (observe the weight load ourselves)
#pragma unroll
        for (int i = 0; i < WN_NT; ++i) {
          if (i == 3) WN_DBG(7, k, 0);
          WN_MBAR_WAIT(&sm.bar_g[i], n_w & 1, 10);
          if (i < WN_NT - 1) WN_MBAR_WAIT(&sm.bar_gate[i + 1], n_w & 1, 3);
          if (i == 3) WN_DBG(7, k, 1);
          fence_after_sync();
          if (elect_one()) {
            // res (columns 32..47) and skip (48..79) in one N = 48 GEMM.  One commit: the next gate GEMM overwrites the
            // g columns, so epilogue 2a must not release it before this GEMM has read them
            const uint32_t tacc = tmem + i * WN_TMEM_TILE;
            mma_f16_ts(tacc + WN_C_R, tacc + WN_C_G, bh, idesc_rs, false);
            if (nsplit == 3) {
              mma_f16_ts(tacc + WN_C_R, tacc + WN_C_G + 8, bh, idesc_rs, true);
              mma_f16_ts(tacc + WN_C_R, tacc + WN_C_G, bh + lo_b, idesc_rs, true);
            }
            mma_f16_ts(tacc + WN_C_R, ones, bb, idesc_rs, true);
            mma_commit(&sm.bar_rs[i]);
            if (i == 0) WN_DBG(5, k, 2);
            if (i == WN_NT - 1) WN_DBG(5, k, 3);
          }
          __syncwarp();
          if (i == 3) WN_DBG(7, k, 2);
        }
      }
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == WN_EPI_WARPS) tmem_dealloc(tmem, 512);
}

// the gate pre-activations are produced pre-scaled for ex2: tanh half by -2*log2(e), sigmoid half by -log2(e)
static double gate_scale(int n) { return n < 16 ? -2.8853900817779268 : -1.4426950408889634; }

static void put_split(std::vector<unsigned char>& buf, size_t hi_off, size_t lo_off, float x, bool split) {
  __half h = __float2half_rn(x);
  __half l = split ? __float2half_rn(x - __half2float(h)) : __float2half_rn(0.f);
  memcpy(&buf[hi_off], &h, 2);
  memcpy(&buf[lo_off], &l, 2);
}

// gate_w [24][48][32] (k = tap*16+ch ; n), rs_w [24][16][48], biases, bn, dilation (fp32 device-layout copies
// made by api.cu on the host before upload)
// padding rows for a block whose BatchNorm is (mul, add): x_pad = -add / mul (=> u_pad = 0), as the 64-byte
// (hi chunk 0, hi chunk 1, lo chunk 0, lo chunk 1) image of one U row.  false if a scale is 0 (cannot be folded).
static bool put_pad_rows(std::vector<unsigned char>& out, size_t off, const float* mul, const float* add) {
  for (int c = 0; c < 16; ++c) {
    if (mul[c] == 0.f) return false;
    const size_t o = off + (size_t)(c / 8) * 16 + (c % 8) * 2;
    put_split(out, o, o + 32, (float)(-(double)add[c] / (double)mul[c]), true);
  }
  return true;
}

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
        put_split(out, off, off + 3072, (float)((double)gate_w[((size_t)b * 48 + k) * 32 + n] * (double)bn_mul[b * 16 + k % 16] * gate_scale(n)), true);
      }
    for (int k = 0; k < 16; ++k)
      for (int n = 0; n < 48; ++n) {
        const int c = k / 8, e = k % 8;
        const size_t off = base + WN_GATE_B + ((size_t)c * 48 + n) * 16 + e * 2;
        put_split(out, off, off + 1536, rs_w[((size_t)b * 16 + k) * 48 + n], true);
      }
    float f[113] = {0};
    f[112] = (float)dilation[b];
    memcpy(&out[base + WN_OFF_F32], f, sizeof(f));
    if (b < 23 && !put_pad_rows(out, base + WN_OFF_F32 + 320, bn_mul + (b + 1) * 16, bn_add + (b + 1) * 16)) return {};
    // bias operands of the 'ones' GEMM: row n = (hi, lo, 0, ...); the second k-chunk stays zero
    for (int n = 0; n < 32; ++n) {
      const size_t off = base + WN_OFF_GBIAS + (size_t)n * 16;
      double bsum = gate_b[b * 32 + n];   // + W * bn_add (the folded BatchNorm shift)
      for (int k = 0; k < 48; ++k) bsum += (double)gate_w[((size_t)b * 48 + k) * 32 + n] * (double)bn_add[b * 16 + k % 16];
      put_split(out, off, off + 2, (float)(bsum * gate_scale(n)), true);
    }
    for (int n = 0; n < 48; ++n) {
      const size_t off = base + WN_OFF_RBIAS + (size_t)n * 16;
      put_split(out, off, off + 2, rs_b[b * 48 + n], true);
    }
  }
  return out;
}

std::vector<unsigned char> wavenet_pack_head(const float* bn_mul0, const float* bn_add0, const float* det1_w_nk,
                                             const float* det1_b, const float* det2_w, const float* det2_b) {
  std::vector<unsigned char> out(sizeof(WnHead), 0);
  WnHead* h = reinterpret_cast<WnHead*>(out.data());
  if (!put_pad_rows(out, offsetof(WnHead, pad0), bn_mul0, bn_add0)) return {};
  h = reinterpret_cast<WnHead*>(out.data());
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

// Input layer (Wavenet/encode.tflite op 0: 1x1 conv 40 -> 16 + ReLU) once per mel row, fp32:
// x0[row, c] = ReLU(b[c] + sum_k w[k][c] * mel[row, k]).  4 threads per row, 4 channels each.
__global__ void __launch_bounds__(256) wn_input_kernel(const float* __restrict__ mel, const float* __restrict__ w_kc,
                                                       const float* __restrict__ b, float* __restrict__ x0, int64_t n_rows) {
  __shared__ float ws[kMel * 16];
  __shared__ float bs[16];
  for (int i = threadIdx.x; i < kMel * 16; i += 256) ws[i] = w_kc[i];
  if (threadIdx.x < 16) bs[threadIdx.x] = b[threadIdx.x];
  __syncthreads();
  const int c4 = threadIdx.x & 3;
  for (int64_t row = (int64_t)blockIdx.x * 64 + (threadIdx.x >> 2); row < n_rows; row += (int64_t)gridDim.x * 64) {
    const float4* m4 = reinterpret_cast<const float4*>(mel + row * kMel);
    float a0 = bs[4 * c4], a1 = bs[4 * c4 + 1], a2 = bs[4 * c4 + 2], a3 = bs[4 * c4 + 3];
#pragma unroll
    for (int k4 = 0; k4 < kMel / 4; ++k4) {
      const float4 m = __ldg(m4 + k4);
      const float mv[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float4 wv = *reinterpret_cast<const float4*>(ws + (4 * k4 + e) * 16 + 4 * c4);
        a0 = fmaf(wv.x, mv[e], a0); a1 = fmaf(wv.y, mv[e], a1); a2 = fmaf(wv.z, mv[e], a2); a3 = fmaf(wv.w, mv[e], a3);
      }
    }
    reinterpret_cast<float4*>(x0 + row * 16)[c4] = make_float4(fmaxf(a0, 0.f), fmaxf(a1, 0.f), fmaxf(a2, 0.f), fmaxf(a3, 0.f));
  }
}

int wavenet_tc_posteriors(wwb_ctx* ctx, const WinMap& wm, float* enc_out, float* det_out, float* post,
                          cudaStream_t st) {
  if (wm.n_win == 0) return WWB_OK;
  if (ctx->L > 182) return fail(ctx, WWB_ERR_ARG, "tensor-core WaveNet path supports windows up to 182 frames");
  const int64_t n_rows = wm.n_streams * (int64_t)wm.ring;
  void* x0;
  int rc = workspace(ctx, 1, (size_t)std::max<int64_t>(n_rows, 1) * 16 * sizeof(float), &x0);
  if (rc) return rc;
  wn_input_kernel<<<(unsigned)std::min<int64_t>((n_rows + 63) / 64, (int64_t)ctx->sm_count * 8), 256, 0, st>>>(
      wm.mel, ctx->wn.in_w, ctx->wn.in_b, (float*)x0, n_rows);
  WWB_CHECK_LAUNCH(ctx);
  WnTcParams P;
  P.x0 = (const float*)x0;
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
