// Declarations of the tensor-core WaveNet kernel (wavenet_tc.cu): geometry, operand-blob layout, kernel parameters and
// the epilogue helpers.
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

namespace wwb {

using namespace tc;

constexpr int WN_G = 4;                       // windows per group
constexpr int WN_NT = 6;                      // M tiles per group (4 x 182 = 728 of 768 rows)
constexpr int WN_ROWS = WN_NT * 128;          // 768
constexpr int WN_PAD = 16 * WN_G;             // leading padding rows of the U buffer: 2 * max dilation (8) * WN_G
constexpr int WN_UROWS = WN_ROWS + WN_PAD;
constexpr int WN_MAXT = 182;                  // WN_G * 182 = 728 rows <= WN_ROWS
constexpr int WN_CHUNK_WARM = 180;            // stream mode: rows of a chunk inside some receptive field (D(23) = 180)
constexpr int WN_CHUNK_STEP = WN_ROWS - WN_CHUNK_WARM;
constexpr int WN_SNAP_F = 48;                 // floats per snapshot row: x[16], skip prefix sum[32] (stored as 12 float4 column planes)
constexpr int WN_PU = WN_UROWS * 16;          // bytes per U chunk panel
constexpr int WN_UBUF = 4 * WN_PU;            // the U buffer: hi and lo planes of two k-chunks
constexpr int WN_EPI_WARPS = WN_NT * 4;       // 24
constexpr int WN_EPI_THREADS = WN_EPI_WARPS * 32;
constexpr int WN_THREADS = (WN_EPI_WARPS + 2) * 32;   // + the gate-GEMM / loader warp + the res/skip-GEMM warp = 832
constexpr int WN_GATE_B = 6144, WN_RS_B = 3072;   // gate / res+skip B operands (hi and lo planes)
constexpr int WN_F32_B = 512;                     // misc: bytes [320, 384) = the NEXT block's padding rows (hi0, hi1, lo0, lo1)
constexpr int WN_GBIAS_B = 1024, WN_RBIAS_B = 1536;   // bias B operands for the 'ones' GEMM (k0 = hi, k1 = lo)
constexpr int WN_OFF_F32 = WN_GATE_B + WN_RS_B, WN_OFF_GBIAS = WN_OFF_F32 + WN_F32_B, WN_OFF_RBIAS = WN_OFF_GBIAS + WN_GBIAS_B;
constexpr int WN_WBLK = WN_OFF_RBIAS + WN_RBIAS_B;    // 12288 bytes of one block's weight blob
constexpr int WN_WST = 5;                     // weight ring stages
constexpr int WN_TMEM_TILE = 80;              // TMEM columns per tile:
constexpr int WN_C_G = 0;                     //   0..31  gate accumulator; g hi/lo (A of res/skip) overwrites 0..15 once epilogue 1 has read it
constexpr int WN_C_R = 16;                    //   16..31 res accumulator (over the consumed sigmoid half of the gate accumulator), 32..63 skip accumulator (detect: 16..47)
constexpr int WN_C_U = 64;                    //   64..79 u hi/lo (A of the gate's unshifted tap)
constexpr int WN_DBG_ROLES = WN_NT + 3;          // timeline roles: tiles 0..NT-1, issuing warps, group boundary, res/skip warp
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

struct WnTcParams {
  WinMap wm;
  const unsigned char* wblob;   // [24][WN_WBLK]
  const WnHead* head;
  const float* x0;              // [4 float4 columns][n_streams * ring] float4: ReLU(in_w * mel + in_b) of every mel row (wn_input_kernel)
  int L;
  int nsplit;
  int dil[24];                  // dilation per block (kernel-parameter space keeps it in uniform registers)
  int rstride;                  // rows per time step: WN_G (window groups, row = t*3 + w) or 1 (stream chunks)
  int stream_mode;              // 1: a group is a chunk of WN_ROWS consecutive frames of one stream; writes snapshots
  int chunks_per_stream;
  int join[WN_NT];              // first block tile i takes part in (0 = from the input layer)
  int src_slot[WN_NT];          // snapshot slot a joining tile starts from (-1: input layer x0, skip = 0)
  int snap_slot[24];            // stream mode: snapshot slot written after block k (-1: none)
  // detect head constants of the 32 -> 2 layer: read as kernel parameters (constant bank, uniform datapath) - as
  // broadcast shared-memory loads they were ~100 wavefronts per warp and group, with all tiles reaching the detect
  // epilogue at about the same time (1900 clk of the ~4900 clk group boundary)
  float det1_b[32], det2_w[64], det2_b[2];
  float* snap;                  // [n_slots][WN_SNAP_F / 4 float4 columns][n_rows] float4 (column planes)
  int64_t n_rows;               // rows behind x0 / snap (n_streams * ring)
  float* enc_out;
  float* det_out;
  float* post;
  long long* dbg;   // optional timeline dump (block 0, second group): [8 roles][2 groups x 24 blocks][4 events]
};

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

}  // namespace wwb
