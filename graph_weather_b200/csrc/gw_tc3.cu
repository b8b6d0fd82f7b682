// gw_tc3.cu -- the fused MLP-chain kernel on Blackwell tensor cores (tcgen05 + TMEM), precision GW_PREC_FP32_TC / BF16_TC.
//
// One persistent CTA per SM walks 128-row tiles of a gw::TcChain.  For every tile the whole chain
//     A0 = assemble(row sources)                                   (stream / gather / relu(gather+const) ...)
//     for each layer:  D = A . W^T  (tcgen05.mma, fp32 accumulate in TMEM)
//                      v = D*s + bias + gathered addends ; ReLU | LayerNorm ; + residual
//                      v -> global (fp32)  and/or  v -> split fp16 hi/lo -> shared memory = A operand of the next layer
// runs without the activations ever leaving the SM: the reference's x[row]/x[col] gathers, cat, 3 Linear + LayerNorm and
// residual (graph_net_block.py:131-135, 184-191) are one kernel per edge pass and one per node pass.
//
// fp32 fidelity on fp16 tensor cores: every fp32 operand a is split a = hi + lo (two fp16, 22 significand bits) and each
// product is hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM.  Weights are pre-scaled by a power of two so their lo
// parts stay normal; the scale is undone exactly in the epilogue.
//
// Layout of the work (third generation; the second, with mover warps and shared-memory staging -- git history, profiles/r01_c_* --
// spent 40 % of its time in mbarrier hand-offs, measured with tools/ablate.py):
//   * 16 worker warps, four per TMEM lane quadrant.  A worker reads the accumulator with tcgen05.ld.16x256b, whose register
//     fragment gives four adjacent lanes one row.  The weights are packed with output rows and K columns permuted inside groups
//     of 16 (perm16, gw_pack.cu) so that a thread's fragment is four CONSECUTIVE features of a row: in that layout the workers
//     load gathered addends / residual rows and store outputs DIRECTLY from/to global memory with 16-byte accesses (8 rows x 64 B
//     per warp instruction, every fetched sector fully used): no staging buffers, no mover warps, no hand-off barriers.  Loads
//     for the next 64-column chunk (global operands and the accumulator chunk itself) are issued while the current one is
//     converted.
//   * the A operand ring holds a full K = 256 operand (4 chunks x [128 x 64] hi|lo = 128 KB).  Because a layer's epilogue
//     starts only when that layer's MMAs have completed, every operand slot is known to be free when the epilogue refills
//     it: per tile a worker waits on 1 barrier per layer (accumulator complete) and arrives on 1 per produced chunk, one
//     elected lane per warp.
//   * the stage-0 operand of the NEXT tile is assembled before the current tile's last epilogue, slot by slot as the last
//     layer's MMAs release them (empty_a): the assembly overlaps the tail of those MMAs, and the tensor pipe then runs the next
//     tile's first layer on the other accumulator under the LayerNorm epilogue of this tile.
//   * two instantiations per precision: the lean path (every source / output 16-byte aligned and as wide as the layer; rows are a
//     warp-uniform 64-bit base + 32-bit offsets; the epilogue is specialised per feature mask) and the general path (any width /
//     alignment).  setmaxnreg gives the workers 120 registers and the auxiliary warpgroup 32; the lean path has no spills.
//   * warp 16: weight producer (cp.async.bulk of pre-swizzled 32 KB panels), warp 17: MMA issuer (one thread).
// TMEM: 512 columns = two 128x256 fp32 accumulators alternating by layer.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <type_traits>

#include "gw_internal.h"
#include "gw_ops.h"
#include "gw_tc_ptx.cuh"

namespace gw {
namespace t3 {

#ifdef GW_ABLATE
#define ABL3(bit) ((ch.ablate & (bit)) != 0)
#else
#define ABL3(bit) false
#endif
enum { ABL_FENCE = 1, ABL_LOADS = 2, ABL_CONVERT = 4, ABL_STORES = 8, ABL_LN = 16, ABL_TMEM = 32, ABL_MMA = 64, ABL_WEIGHTS = 128 };

// epilogue feature mask of a layer (TcLayer::kind); the lean path is instantiated for the masks that occur
enum { F_ADD0 = 1, F_ADD1 = 2, F_RELU = 4, F_LN = 8, F_RES = 16, F_OUT = 32, F_FEEDS = 64 };
#ifndef GW_CHUNK_UNROLL
#define GW_CHUNK_UNROLL 1  // the per-chunk loops stay rolled: unrolled, their code no longer fits the instruction cache
#endif
constexpr int CHUNK_UNROLL = GW_CHUNK_UNROLL;
constexpr int TILE_M = 128;
constexpr int A_SLOTS = 4, B_STAGES = 2;
constexpr int A_HALF_BYTES = TILE_M * 128;      // [128 rows x 64 halfs]
constexpr int A_SLOT_BYTES = 2 * A_HALF_BYTES;  // hi | lo
constexpr int B_STAGE_BYTES = 256 * 128;        // [256 rows x 64 halfs], hi OR lo panel
constexpr int WSPLIT = 4;                       // worker warps per TMEM lane quadrant; each owns 16 columns of every 64-column chunk
constexpr int WORKER_WARPS = 4 * WSPLIT, NUM_WORKERS = 32 * WORKER_WARPS;
constexpr int WARP_PRODUCER = WORKER_WARPS, WARP_MMA = WORKER_WARPS + 1;
constexpr int NUM_THREADS = NUM_WORKERS + 128;  // + one auxiliary warpgroup: producer, MMA issuer, two idle warps
constexpr int WORKER_REGS = 104, AUX_REGS = 56;  // setmaxnreg: 4 x 32 x 120 + 32 x 32 = 16384 registers per SM sub-partition
constexpr int PAR_LAYERS = 6;
constexpr int OFF_A = 0;
constexpr int OFF_B = A_SLOTS * A_SLOT_BYTES;
constexpr int OFF_BAR = OFF_B + B_STAGES * B_STAGE_BYTES;
constexpr int NUM_BARS = 2 * A_SLOTS + 2 * B_STAGES + 4;
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int OFF_PAR = OFF_TMEM + 16;                // float bias[PAR_LAYERS][256]
constexpr int OFF_LNP = OFF_PAR + PAR_LAYERS * 1024;  // float gamma_beta[2][2][256]
constexpr int OFF_LN = OFF_LNP + 4 * 1024;            // float ln_x[WSPLIT][128], ln_y[WSPLIT][128]: row statistics exchange
constexpr int OFF_PRE = OFF_LN + 2 * WSPLIT * 128 * 4;  // uint32 pre[8][NUM_WORKERS]: layer-0 gather offsets of the coming tile
constexpr int SMEM_BYTES = OFF_PRE + 8 * NUM_WORKERS * 4;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(OFF_B % 1024 == 0 && A_SLOT_BYTES % 1024 == 0 && B_STAGE_BYTES % 1024 == 0, "SWIZZLE_128B needs 1 KB alignment");

// tcgen05.ld 16 lanes x 256 bit, x2: 16 accumulator columns of 16 rows.  Lane t holds (cute SM100_TMEM_LOAD_16dp256b2x):
//   r0,r1 = (row t/4    , cols 2(t%4)+{0,1})      r2,r3 = (row t/4 + 8, same cols)
//   r4,r5 = (row t/4    , cols 8+2(t%4)+{0,1})    r6,r7 = (row t/4 + 8, same cols)
__device__ __forceinline__ void tmem_ld_16x256b_x2(uint32_t taddr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Wait for TMEM loads that were issued earlier into v (asynchronously: the compiler believes v was written by the issuing
// statement).  Passing v through the wait as in/out operands makes every later read of v depend on the wait.
__device__ __forceinline__ void tmem_wait_ld_into(float (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                 "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
               :
               : "memory");
}

// A worker thread's fragment of one 64-column chunk: 16 values, index i = 8g + 4j + 2m + e
//   tile row  r(k) = 32 q + 16 g + lane/4 + 8 m   (k = 2g + m),   accumulator column = 16 hq + 8 j + 2 (lane%4) + e
// The weights are packed with their output rows AND K columns permuted inside every group of 16 (perm16, gw_pack.cu) so that
// accumulator column 8j + 2(lane%4) + e holds LOGICAL feature 4 (lane%4) + 2j + e: a thread owns 4 consecutive features of a row
// (16 bytes of fp32) -> global loads / stores are 128-bit, four adjacent lanes cover 64 B of one row.  Operands live in shared
// memory in accumulator order on both sides of every product, so nothing else changes.
__device__ __forceinline__ bool vec2_ok(const float* base, int ld) { return ((reinterpret_cast<uintptr_t>(base) & 7) == 0) && ((ld & 1) == 0); }
__device__ __forceinline__ bool src_gathered(int kind) { return kind == SRC_GATHER || kind == SRC_BGATHER || kind == SRC_GATHER_BCAST_RELU; }
__device__ __forceinline__ bool src_per_sample(int kind) { return kind == SRC_STREAM || kind == SRC_GATHER || kind == SRC_GATHER_BCAST_RELU; }

// source rows of my four tile rows (gather indices are the only per-tile state a source needs in registers; everything
// else is re-read from the kernel parameters where it is used)
__device__ __forceinline__ void rows_of(const RowSrc& s, int i0, const int (&rl)[4], int (&ri)[4]) {
  const bool g = src_gathered(s.kind);
#pragma unroll
  for (int k = 0; k < 4; ++k) ri[k] = g ? __ldg(s.idx + i0 + rl[k]) : i0 + rl[k];
}
// my 16 values (4 rows x logical columns col .. col+3; `col` includes 16hq + 4(lane%4)).  Warp-uniform tiers: 128-bit loads when
// the warp's 16 columns lie inside the source and rows are 16-byte aligned, 64-bit when 8-byte aligned (the 102-wide
// features), else bounds-checked scalars.
__device__ __forceinline__ bool vec4_ok(const float* base, int ld) { return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ld & 3) == 0); }
__device__ __forceinline__ void load16(const float* base, int ld, int width, const int (&ri)[4], int col, int lc, float (&o)[16]) {
  const bool inside = col - 4 * lc + 16 <= width;
  if (inside && vec4_ok(base, ld)) {
    float4 t[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) t[k] = __ldg(reinterpret_cast<const float4*>(base + (size_t)ri[k] * (size_t)ld + col));
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 8 * (k >> 1) + 2 * (k & 1);
      o[i] = t[k].x, o[i + 1] = t[k].y, o[i + 4] = t[k].z, o[i + 5] = t[k].w;
    }
  } else if (inside && vec2_ok(base, ld)) {
    float2 t[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* rowp = base + (size_t)ri[k] * (size_t)ld + col;
      t[2 * k] = __ldg(reinterpret_cast<const float2*>(rowp));
      t[2 * k + 1] = __ldg(reinterpret_cast<const float2*>(rowp + 2));
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 8 * (k >> 1) + 2 * (k & 1);
      o[i] = t[2 * k].x, o[i + 1] = t[2 * k].y, o[i + 4] = t[2 * k + 1].x, o[i + 5] = t[2 * k + 1].y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float* rowp = base + (size_t)ri[k] * (size_t)ld;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int i = 8 * (k >> 1) + 4 * j + 2 * (k & 1), c = col + 2 * j;
        o[i] = (c < width) ? __ldg(rowp + c) : 0.f;
        o[i + 1] = (c + 1 < width) ? __ldg(rowp + c + 1) : 0.f;
      }
    }
  }
}
__device__ __forceinline__ void load16(const RowSrc& s, int b, const int (&ri)[4], int col, int lc, float (&o)[16]) {
  const float* base = s.base + (src_per_sample(s.kind) ? (size_t)b * (size_t)s.src_rows * (size_t)s.ld : (size_t)0) + s.col0;
  load16(base, s.ld, s.width, ri, col, lc, o);
}

// Split my 16 values into fp16 hi/lo (or bf16) and store them into the swizzled K-major operand slot: row r, logical
// 16-byte chunk c16 = 2hq + j lives at chunk position c16 ^ (r & 7); my two halfs sit at byte 4 (lane%4) of the chunk.
__device__ __forceinline__ void store_operand16(uint8_t* slot, const int (&rt)[4], int hq, int lc, const float (&v)[16], bool split,
                                                float& amax) {
#pragma unroll
  for (int g = 0; g < 2; ++g)
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int r = rt[2 * g + m];
      uint8_t* row_hi = slot + r * 128 + 4 * lc;
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float a0 = v[8 * g + 4 * j + 2 * m], a1 = v[8 * g + 4 * j + 2 * m + 1];
        amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
        const int pos = ((2 * hq + j) ^ (r & 7)) << 4;
        if (split) {
          const __half2 hh = __floats2half2_rn(a0, a1);
          const float2 hf = __half22float2(hh);
          const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
          *reinterpret_cast<__half2*>(row_hi + pos) = hh;
          *reinterpret_cast<__half2*>(row_hi + A_HALF_BYTES + pos) = ll;
        } else {
          *reinterpret_cast<__nv_bfloat162*>(row_hi + pos) = __floats2bfloat162_rn(a0, a1);
        }
      }
    }
}


// ---- lean full-width path --------------------------------------------------------------------------------------------
// When every source / output of a layer is 16-byte aligned and at least as wide as the layer (the processor and decoder edge
// and node passes: >95 % of the run time), a source is a warp-uniform 64-bit base plus four 32-bit byte offsets (one per row
// of mine), resolved once per layer; every access is base + offset + immediate.  Contiguous sources are addressed relative
// to the tile's first row, gathered ones relative to the sample (the launcher checks that this fits 32 bits).
__device__ __forceinline__ const float* row_base(const RowSrc& s, int b, int i0) {
  const float* base = s.base + (src_per_sample(s.kind) ? (size_t)b * (size_t)s.src_rows * (size_t)s.ld : (size_t)0) + s.col0;
  return src_gathered(s.kind) ? base : base + (size_t)i0 * (size_t)s.ld;
}
__device__ __forceinline__ const float* row_refs(const RowSrc& s, int b, int i0, const int (&rl)[4], int cofs, uint32_t (&off)[4]) {
  const float* base = s.base + (src_per_sample(s.kind) ? (size_t)b * (size_t)s.src_rows * (size_t)s.ld : (size_t)0) + s.col0;
  if (src_gathered(s.kind)) {
#pragma unroll
    for (int k = 0; k < 4; ++k) off[k] = ((uint32_t)__ldg(s.idx + i0 + rl[k]) * (uint32_t)s.ld + (uint32_t)cofs) * 4u;
    return base;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) off[k] = ((uint32_t)rl[k] * (uint32_t)s.ld + (uint32_t)cofs) * 4u;
  return base + (size_t)i0 * (size_t)s.ld;
}
// my 16 values at float offset `coff` from the row references: 4 LDG.128
__device__ __forceinline__ void ldfrag(const float* base, const uint32_t (&off)[4], int coff, float (&o)[16]) {
  const char* b = reinterpret_cast<const char*>(base + coff);
  float4 t[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) t[k] = __ldg(reinterpret_cast<const float4*>(b + off[k]));
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = 8 * (k >> 1) + 2 * (k & 1);
    o[i] = t[k].x, o[i + 1] = t[k].y, o[i + 4] = t[k].z, o[i + 5] = t[k].w;
  }
}
// Operand store with precomputed addressing.  `sa` = shared address of (my row 32q + lane/4, my half2, chunk j = 0) inside the
// slot; the j = 1 chunk is sa ^ 16 (SWIZZLE_128B flips bit 4), my other rows are +8 / +16 / +24 rows = immediates.
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }
template <bool SPLIT>
__device__ __forceinline__ void store_operand_fast(uint32_t sa, const float (&v)[16], float& amax) {
  const uint32_t sa1 = sa ^ 16u;
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int i = 8 * (k >> 1) + 4 * j + 2 * (k & 1);
      const float a0 = v[i], a1 = v[i + 1];
      amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
      const uint32_t addr = (j ? sa1 : sa) + (16 * (k >> 1) + 8 * (k & 1)) * 128;
      if (SPLIT) {
        const __half2 hh = __floats2half2_rn(a0, a1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(a0 - hf.x, a1 - hf.y);
        sts32(addr, *reinterpret_cast<const uint32_t*>(&hh));
        sts32(addr + A_HALF_BYTES, *reinterpret_cast<const uint32_t*>(&ll));
      } else {
        const __nv_bfloat162 bb = __floats2bfloat162_rn(a0, a1);
        sts32(addr, *reinterpret_cast<const uint32_t*>(&bb));
      }
    }
}

// MODE 0: every part takes the general path; 1: every part takes the lean full-width path.  (A third mode that chose per
// part inside one kernel was measured slower than the general path: the live state of both paths spills.)
template <bool SPLIT, int MODE>
__global__ void __launch_bounds__(NUM_THREADS, 1) gw_chain_tc3_kernel(const __grid_constant__ TcChain ch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = ch.rows_per_sample, batch = ch.batch;
  const int tiles_per_sample = (rows + TILE_M - 1) / TILE_M;
  const int num_tiles = tiles_per_sample * batch;  // batch-major: tile -> (sample = tile % batch, row block = tile / batch)
  constexpr bool split = SPLIT;
  constexpr int parts = SPLIT ? 2 : 1;

  const uint32_t bar_full_a = sbase + OFF_BAR;             // [A_SLOTS] workers -> MMA (one arrival per worker warp)
  const uint32_t bar_empty_a = bar_full_a + 8 * A_SLOTS;   // [A_SLOTS] MMA -> workers (tcgen05.commit); waited only when stage 0 wraps the ring
  const uint32_t bar_full_b = bar_empty_a + 8 * A_SLOTS;   // [B_STAGES] bulk copy -> MMA
  const uint32_t bar_empty_b = bar_full_b + 8 * B_STAGES;  // [B_STAGES] MMA -> producer
  const uint32_t bar_full_d = bar_empty_b + 8 * B_STAGES;  // [2] MMA -> workers: accumulator complete
  const uint32_t bar_empty_d = bar_full_d + 16;            // [2] workers -> MMA: accumulator drained
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {  // SWIZZLE_128B operand tiles must be 1 KB aligned
      if (ch.status) atomicOr(ch.status, 4);
      __trap();
    }
    for (int i = 0; i < A_SLOTS; ++i) mbar_init(bar_full_a + 8 * i, WORKER_WARPS), mbar_init(bar_empty_a + 8 * i, 1);
    for (int i = 0; i < B_STAGES; ++i) mbar_init(bar_full_b + 8 * i, 1), mbar_init(bar_empty_b + 8 * i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_full_d + 8 * i, 1), mbar_init(bar_empty_d + 8 * i, WORKER_WARPS);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) {  // TMEM: all 512 columns (two fp32 accumulators of 256 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + OFF_TMEM), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // per-layer column parameters -> shared memory (zero beyond n_valid)
    float* par = reinterpret_cast<float*>(smem + OFF_PAR);
    float* lnp = reinterpret_cast<float*>(smem + OFF_LNP);
    int ln_slot = 0;
    for (int l = 0; l < ch.n_layers; ++l) {
      const TcLayer& L = ch.layer[l];
      for (int c = threadIdx.x; c < 256; c += NUM_THREADS) par[l * 256 + c] = (L.bias && c < L.n_valid) ? __ldg(L.bias + c) : 0.f;
      if (L.ln_g) {
        for (int c = threadIdx.x; c < 256; c += NUM_THREADS) {
          lnp[(ln_slot * 2 + 0) * 256 + c] = (c < L.n_valid) ? __ldg(L.ln_g + c) : 0.f;
          lnp[(ln_slot * 2 + 1) * 256 + c] = (c < L.n_valid) ? __ldg(L.ln_b + c) : 0.f;
        }
        ++ln_slot;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // Register reallocation (inside each role's branch, so that ptxas budgets the roles separately): the auxiliary warpgroup
  // keeps 32 registers per thread, the workers (which hold a 64-value LayerNorm fragment per thread) grow to 120.
  if (warp >= WORKER_WARPS) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
  if (warp == WARP_PRODUCER) {
    // ===================================== weight producer =========================================================
    if (lane == 0) {
      uint32_t bi = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        for (int l = 0; l < ch.n_layers; ++l) {
          const TcLayer& L = ch.layer[l];
          const uint32_t panel = (uint32_t)L.N * 128u;
          const uint8_t* w = static_cast<const uint8_t*>(L.Wp);
          const int nk = L.K >> 6;
          for (int kc = 0; kc < nk; ++kc) {
            for (int part = 0; part < parts; ++part, ++bi) {
              const uint32_t stage = bi % B_STAGES, n = bi / B_STAGES;
              mbar_wait(bar_empty_b + 8 * stage, (n & 1) ^ 1, ch.status);
              if (ABL3(ABL_WEIGHTS)) {
                mbar_arrive(bar_full_b + 8 * stage);
                continue;
              }
              mbar_expect_tx(bar_full_b + 8 * stage, panel);
              bulk_g2s(sbase + OFF_B + stage * B_STAGE_BYTES, w + (size_t)(kc * parts + part) * panel, panel, bar_full_b + 8 * stage);
            }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================================== MMA issuer ==============================================================
    if (lane == 0) {
      uint32_t bi = 0, fi = 0, li = 0;
      Tracer tr;
      tr.init(ch.trace, 1, true);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        uint32_t prev_first = fi;
        tr.ev(1000);
        for (int l = 0; l < ch.n_layers; ++l, ++li) {
          const TcLayer& L = ch.layer[l];
          const int nk = L.K >> 6;
          const uint32_t idesc = umma_idesc(L.N, !split);
          const uint32_t acc = li & 1, use = li >> 1;
          tr.ev(100 + l);
          mbar_wait(bar_empty_d + 8 * acc, (use & 1) ^ 1, ch.status);  // epilogue of layer li-2 has drained this accumulator
          tc_fence_after();
          tr.ev(110 + l);
          const uint32_t d_tmem = tmem_base + acc * 256;
          const uint32_t first = L.reuse_a ? prev_first : fi;
          const bool last_use = !(l + 1 < ch.n_layers && ch.layer[l + 1].reuse_a);
          for (int kc = 0; kc < nk; ++kc) {
            const uint32_t f = first + kc, slot = f % A_SLOTS, n = f / A_SLOTS;
            mbar_wait(bar_full_a + 8 * slot, n & 1, ch.status);  // (already complete when the operand is re-used)
            tc_fence_after();
            tr.ev(200 + kc);
            const uint32_t a_hi = sbase + OFF_A + slot * A_SLOT_BYTES, a_lo = a_hi + A_HALF_BYTES;
            {  // hi weight panel: A_hi.B_hi (+ A_lo.B_hi)
              const uint32_t stage = bi % B_STAGES, nb = bi / B_STAGES;
              mbar_wait(bar_full_b + 8 * stage, nb & 1, ch.status);
              tc_fence_after();
              const uint32_t b = sbase + OFF_B + stage * B_STAGE_BYTES;
              if (!ABL3(ABL_MMA)) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, umma_desc(a_hi + 32 * ks), umma_desc(b + 32 * ks), idesc, (kc | ks) != 0);
                if (split) {
#pragma unroll
                  for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, umma_desc(a_lo + 32 * ks), umma_desc(b + 32 * ks), idesc, 1);
                }
              }
              tc_commit(bar_empty_b + 8 * stage);
              ++bi;
            }
            if (split) {  // lo weight panel: A_hi.B_lo
              const uint32_t stage = bi % B_STAGES, nb = bi / B_STAGES;
              mbar_wait(bar_full_b + 8 * stage, nb & 1, ch.status);
              tc_fence_after();
              const uint32_t b = sbase + OFF_B + stage * B_STAGE_BYTES;
              if (!ABL3(ABL_MMA)) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, umma_desc(a_hi + 32 * ks), umma_desc(b + 32 * ks), idesc, 1);
              }
              tc_commit(bar_empty_b + 8 * stage);
              ++bi;
            }
            if (last_use) tc_commit(bar_empty_a + 8 * slot);  // one phase per use of the slot
            tr.ev(300 + kc);
          }
          tc_commit(bar_full_d + 8 * acc);  // accumulator complete -> epilogue
          if (!L.reuse_a) {
            prev_first = fi;
            fi += nk;
          }
        }
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WORKER_REGS));
    // ===================================== workers =================================================================
    const int q = warp & 3;    // TMEM lane quadrant (a warp may only touch lanes 32 (warp % 4) ..+31)
    const int hq = warp >> 2;  // which 16 columns of every 64-column chunk
    const int lr = lane >> 2, lc = lane & 3;
    int rt[4];  // my four tile rows (= TMEM lanes)
#pragma unroll
    for (int k = 0; k < 4; ++k) rt[k] = 32 * q + 16 * (k >> 1) + lr + 8 * (k & 1);
    const int cofs = 16 * hq + 4 * lc;  // my first (logical) column inside a 64-column chunk; the operand position uses 2 * lc
    float* ln_x = reinterpret_cast<float*>(smem + OFF_LN);
    float* ln_y = ln_x + WSPLIT * 128;
    uint32_t fi = 0, li = 0;
    float amax = 0.f;
    Tracer tr;
    tr.init(ch.trace, 5 + (hq & 1), q == 0 && lane == 0 && hq < 2);

    // publish one finished operand chunk: my writes -> async proxy, then one arrival per warp
    auto publish = [&](uint32_t slot) {
      if (!ABL3(ABL_FENCE)) fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full_a + 8 * slot);
    };

    // ---- stage 0: assemble the fp16 hi/lo operand of the first layer straight from global memory -----------------------
    auto stage0 = [&](int tile) {
      const int bs = tile % batch, i0 = (tile / batch) * TILE_M;
      const int nvalid = min(TILE_M, rows - i0);
      int rl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) rl[k] = min(rt[k], nvalid - 1);
      const int nk0 = ch.K0 >> 6, w0 = ch.a0[0].width;
      int ri[4] = {0, 0, 0, 0}, ri2[4] = {0, 0, 0, 0};  // rows of the current source (and of its broadcast partner)
      int cur_src = -1;
      auto fetch = [&](int c, float (&o)[16]) {
        const int colc = 64 * c;
        const int which = colc < w0 ? 0 : 1;
        const RowSrc& src = ch.a0[which];
        const int rel = which ? colc - w0 : colc;
        if (src.kind == SRC_NONE || rel >= src.width || ABL3(ABL_LOADS)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = 0.f;
          return;
        }
        if (which != cur_src) {
          rows_of(src, i0, rl, ri);
#pragma unroll
          for (int k = 0; k < 4; ++k) ri2[k] = i0 + rl[k];
          cur_src = which;
        }
        load16(src, bs, ri, rel + cofs, lc, o);
        if (src.kind == SRC_GATHER_BCAST_RELU) {
          float t[16];
          load16(src.base2, src.ld2, src.width, ri2, rel + cofs, lc, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = fmaxf(o[i] + t[i], 0.f);
        }
      };
      float cur[16], nxt[16];
      fetch(0, cur);
      for (int c = 0; c < nk0; ++c) {
        if (c + 1 < nk0) fetch(c + 1, nxt);
        const uint32_t f = fi + c, slot = f % A_SLOTS, n = f / A_SLOTS;
        if (n > 0) mbar_wait(bar_empty_a + 8 * slot, (n - 1) & 1, ch.status);  // the MMAs that read this slot last have completed
        if (!ABL3(ABL_CONVERT)) store_operand16(smem + OFF_A + slot * A_SLOT_BYTES, rt, hq, lc, cur, split, amax);
        publish(slot);
        tr.ev(500 + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
      }
      fi += nk0;
    };

    // operand-store address of (my first row, chunk j = 0) inside a slot; see store_operand_fast
    const uint32_t sa0 = sbase + OFF_A + (32 * q + lr) * 128 + 4 * lc + (((2 * hq) ^ lr) << 4);

    // gather offsets of layer 0's addends for the tile whose stage 0 ran last (resolved there, so that the index loads do
    // not sit between two tiles)
    uint32_t* pre_s = reinterpret_cast<uint32_t*>(smem + OFF_PRE) + threadIdx.x;  // [8][NUM_WORKERS], one column per thread
    const bool l0_add0 = ch.layer[0].add[0].kind != SRC_NONE, l0_add1 = ch.layer[0].add[1].kind != SRC_NONE;

    // ---- stage 0, lean path: every 64-column chunk lies inside one aligned source -----------------------------------------
    auto stage0_fast = [&](int tile) {
      const int bs = tile % batch, i0 = (tile / batch) * TILE_M;
      const int nvalid = min(TILE_M, rows - i0);
      int rl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) rl[k] = min(rt[k], nvalid - 1);
      const int nk0 = ch.K0 >> 6, nc0 = ch.a0[0].width >> 6;
      const float* ba = nullptr;
      const float* bb = nullptr;
      uint32_t oa[4] = {0u, 0u, 0u, 0u}, ob[4] = {0u, 0u, 0u, 0u};  // (initialised: arrays assigned only on some paths end up in local memory)
      bool gbr = false;
      auto setsrc = [&](const RowSrc& src) {
        ba = row_refs(src, bs, i0, rl, cofs, oa);
        gbr = src.kind == SRC_GATHER_BCAST_RELU;
        if (gbr) {
          bb = src.base2 + (size_t)i0 * (size_t)src.ld2;
#pragma unroll
          for (int k = 0; k < 4; ++k) ob[k] = ((uint32_t)rl[k] * (uint32_t)src.ld2 + (uint32_t)cofs) * 4u;
        }
      };
      auto fetch = [&](int c, float (&o)[16]) {
        if (c == nc0) setsrc(ch.a0[1]);
        const int off = 64 * (c < nc0 ? c : c - nc0);
        if (ABL3(ABL_LOADS)) return;
        ldfrag(ba, oa, off, o);
        if (gbr) {
          float t[16];
          ldfrag(bb, ob, off, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) o[i] = fmaxf(o[i] + t[i], 0.f);
        }
      };
      float cur[16], nxt[16];
      setsrc(ch.a0[0]);
      fetch(0, cur);
      if (l0_add0) {  // kept in shared memory (one word per thread and row): registers are the scarce resource of the epilogues
        uint32_t t[4];
        (void)row_refs(ch.layer[0].add[0], bs, i0, rl, cofs, t);
#pragma unroll
        for (int k = 0; k < 4; ++k) pre_s[k * NUM_WORKERS] = t[k];
      }
      if (l0_add1) {
        uint32_t t[4];
        (void)row_refs(ch.layer[0].add[1], bs, i0, rl, cofs, t);
#pragma unroll
        for (int k = 0; k < 4; ++k) pre_s[(4 + k) * NUM_WORKERS] = t[k];
      }
      for (int c = 0; c < nk0; ++c) {
        if (c + 1 < nk0) fetch(c + 1, nxt);
        const uint32_t f = fi + c, slot = f % A_SLOTS, n = f / A_SLOTS;
        if (n > 0) mbar_wait(bar_empty_a + 8 * slot, (n - 1) & 1, ch.status);  // the MMAs that read this slot last have completed
        if (!ABL3(ABL_CONVERT)) store_operand_fast<SPLIT>(sa0 + slot * A_SLOT_BYTES, cur, amax);
        publish(slot);
        tr.ev(500 + c);
#pragma unroll
        for (int i = 0; i < 16; ++i) cur[i] = nxt[i];
      }
      fi += nk0;
    };

    // ---- one layer's epilogue, lean path: N = 64 np (LayerNorm: N = 256), full-width aligned addends / residual / output ------------------------
    // Register budget: v[16] + pf0[16] + aux[16] + two row-pointer sets.  `aux` is the add[1] prefetch or, on LayerNorm layers
    // (which have no addends), the per-row scale/shift; `p1` is the add[1] rows or the output rows (never both: launcher).
    auto layer_fast = [&](auto FLc, int l, uint32_t acc, uint32_t use, bool waited, int bs, int i0, int nvalid, int ln_slot) {
      constexpr int F = decltype(FLc)::value;  // >= 0: the layer's feature mask is a compile-time constant
      const TcLayer& L = ch.layer[l];
      const float wsi = L.wscale_inv;
      const int np = L.N >> 6;
      const uint32_t bias_o = OFF_PAR + 4 * cofs + l * 1024;
      const bool has_add0 = F >= 0 ? (F & F_ADD0) != 0 : L.add[0].kind != SRC_NONE;
      const bool has_add1 = F >= 0 ? (F & F_ADD1) != 0 : L.add[1].kind != SRC_NONE;
      const bool has_res = F >= 0 ? (F & F_RES) != 0 : L.residual.kind != SRC_NONE;
      const bool has_out = F >= 0 ? (F & F_OUT) != 0 : L.out != nullptr;
      const bool relu = F >= 0 ? (F & F_RELU) != 0 : L.relu != 0;
      const bool has_ln = F >= 0 ? (F & F_LN) != 0 : L.ln_g != nullptr;
      const bool feeds = F >= 0 ? (F & F_FEEDS) != 0 : L.feeds_next != 0;
      const uint32_t g_o = OFF_LNP + 4 * cofs + (ln_slot * 2) * 1024, b_o = g_o + 1024;
      const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256 + 16 * hq;
      const RowSrc& src0 = has_add0 ? L.add[0] : L.residual;
      const bool has0 = (has_add0 || has_res) && !ABL3(ABL_LOADS), has1 = has_add1 && !ABL3(ABL_LOADS);
      const float* b0 = nullptr;  // pf0 source
      const float* b1 = nullptr;  // add[1] source, or the output rows
      uint32_t o0[4] = {0u, 0u, 0u, 0u}, o1[4] = {0u, 0u, 0u, 0u};  // (initialised: see stage0_fast)
      float pf0[16] = {}, aux[16] = {};
      {
        int rl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) rl[k] = min(rt[k], nvalid - 1);
        if (l == 0 && has_add0) {  // resolved during this tile's stage 0
          b0 = row_base(src0, bs, i0);
#pragma unroll
          for (int k = 0; k < 4; ++k) o0[k] = pre_s[k * NUM_WORKERS];
        } else if (has0) {
          b0 = row_refs(src0, bs, i0, rl, cofs, o0);
        }
        if (l == 0 && has1) {
          b1 = row_base(L.add[1], bs, i0);
#pragma unroll
          for (int k = 0; k < 4; ++k) o1[k] = pre_s[(4 + k) * NUM_WORKERS];
        } else if (has1) {
          b1 = row_refs(L.add[1], bs, i0, rl, cofs, o1);
        }
      }
      if (has0 && !has_ln) ldfrag(b0, o0, 0, pf0);  // (LayerNorm layers: after the statistics pass, which needs the registers)
      if (has1) ldfrag(b1, o1, 0, aux);
      if (!waited) {
        tr.ev(600 + l);
        mbar_wait(bar_full_d + 8 * acc, use & 1, ch.status);
        tc_fence_after();
        tr.ev(610 + l);
      }
      if (has_ln) {  // LayerNorm as v * aux[k] + aux[4 + k] per row
#pragma unroll
        for (int k = 0; k < 4; ++k) aux[k] = 1.f, aux[4 + k] = 0.f;
      }
      if (has_ln && !ABL3(ABL_LN)) {
        float pv[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll CHUNK_UNROLL
        for (int s = 0; s < 4; ++s) {
          float v[16];
          tmem_ld_16x256b_x2(taddr + 64 * s, v);
          tmem_ld_16x256b_x2(taddr + 64 * s + (16u << 16), v + 8);
          tmem_wait_ld();
          const float4 b4 = *reinterpret_cast<const float4*>(smem + bias_o + 256 * s);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], wsi, ((i >> 2) & 1) ? ((i & 1) ? b4.w : b4.z) : ((i & 1) ? b4.y : b4.x));
          if (s == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) pv[k] = v[8 * (k >> 1) + 2 * (k & 1)];
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = 2 * (i >> 3) + ((i >> 1) & 1);
            const float d = v[i] - pv[k];
            s1[k] += d;
            s2[k] = fmaf(d, d, s2[k]);
          }
        }
        if (has0) ldfrag(b0, o0, 0, pf0);  // residual rows of chunk 0: in flight during the merge below
        float mean[4], m2[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = s1[k] * (1.0f / 16.0f);
          mean[k] = pv[k] + a;
          m2[k] = fmaxf(s2[k] - s1[k] * a, 0.f);
        }
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float mb = __shfl_xor_sync(0xffffffffu, mean[k], o), qb = __shfl_xor_sync(0xffffffffu, m2[k], o);
            const float d = mb - mean[k];
            mean[k] = 0.5f * (mean[k] + mb);
            m2[k] = (m2[k] + qb) + d * d * (o == 1 ? 8.0f : 16.0f);
          }
        }
        if (lc == 0) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ln_x[hq * 128 + rt[k]] = mean[k], ln_y[hq * 128 + rt[k]] = m2[k];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // the four warps of this lane quadrant
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          float n = 0.f, mu = 0.f, q2 = 0.f;
#pragma unroll
          for (int w = 0; w < WSPLIT; ++w) {  // same order in every thread of the row; 64 values per warp partial
            const float mw = ln_x[w * 128 + rt[k]], qw = ln_y[w * 128 + rt[k]];
            const float d = mw - mu, nt = n + 64.f;
            mu += d * (64.f / nt);
            q2 += qw + d * d * (n * 64.f / nt);
            n = nt;
          }
          const float rstd = 1.0f / sqrtf(q2 * (1.0f / 256.0f) + 1e-5f);
          aux[k] = rstd, aux[4 + k] = -mu * rstd;
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // ln_x / ln_y may be rewritten by the next LayerNorm
      }
      if (has_ln && ABL3(ABL_LN) && has0) ldfrag(b0, o0, 0, pf0);
      if (has_out) {
        b1 = L.out + ((size_t)bs * rows + i0) * (size_t)L.ldo;
#pragma unroll
        for (int k = 0; k < 4; ++k) o1[k] = ((uint32_t)rt[k] * (uint32_t)L.ldo + (uint32_t)cofs) * 4u;
      }
      // The accumulator chunk s+1 is fetched from TMEM (into vn) while chunk s is processed (in v).
      float vn[16] = {};
      if (!ABL3(ABL_TMEM)) {
        tmem_ld_16x256b_x2(taddr, vn);
        tmem_ld_16x256b_x2(taddr + (16u << 16), vn + 8);
      }
#pragma unroll CHUNK_UNROLL
      for (int s = 0; s < np; ++s) {
        float v[16];
        tmem_wait_ld_into(vn);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = vn[i];
        if (s + 1 < np && !ABL3(ABL_TMEM)) {
          tmem_ld_16x256b_x2(taddr + 64 * (s + 1), vn);
          tmem_ld_16x256b_x2(taddr + 64 * (s + 1) + (16u << 16), vn + 8);
        }
        if (s + 1 == np) {  // my last read of this accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_empty_d + 8 * acc);
        }
        const float4 b4 = *reinterpret_cast<const float4*>(smem + bias_o + 256 * s);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], wsi, ((i >> 2) & 1) ? ((i & 1) ? b4.w : b4.z) : ((i & 1) ? b4.y : b4.x));
        if (has_add0 && has0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += pf0[i];
        }
        if (has1) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += aux[i];
        }
        if (relu) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
        }
        if (has_ln) {
          const float4 g4 = *reinterpret_cast<const float4*>(smem + g_o + 256 * s);
          const float4 e4 = *reinterpret_cast<const float4*>(smem + b_o + 256 * s);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int k = 2 * (i >> 3) + ((i >> 1) & 1);
            const float gx = ((i >> 2) & 1) ? ((i & 1) ? g4.w : g4.z) : ((i & 1) ? g4.y : g4.x);
            const float ex = ((i >> 2) & 1) ? ((i & 1) ? e4.w : e4.z) : ((i & 1) ? e4.y : e4.x);
            v[i] = fmaf(fmaf(v[i], aux[k], aux[4 + k]), gx, ex);
          }
        }
        if (!has_add0 && has0) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += pf0[i];
        }
        if (s + 1 < np) {  // next chunk's global operands: in flight while this chunk is stored / converted
          if (has0) ldfrag(b0, o0, 64 * (s + 1), pf0);
          if (has1) ldfrag(b1, o1, 64 * (s + 1), aux);
        }
        if (has_out && !ABL3(ABL_STORES)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (rt[k] < nvalid) {
              const int i = 8 * (k >> 1) + 2 * (k & 1);
              char* o = reinterpret_cast<char*>(const_cast<float*>(b1) + 64 * s) + o1[k];
              *reinterpret_cast<float4*>(o) = make_float4(v[i], v[i + 1], v[i + 4], v[i + 5]);
            }
          }
        }
        if (feeds) {
          const uint32_t slot = (fi + s) % A_SLOTS;
          if (!ABL3(ABL_CONVERT)) store_operand_fast<SPLIT>(sa0 + slot * A_SLOT_BYTES, v, amax);
          publish(slot);
        }
        tr.ev(700 + 10 * l + s);
      }
      if (feeds) fi += np;
    };

    const int n_layers = ch.n_layers;
    bool first_tile = true;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, first_tile = false) {
      const int bs = tile % batch, i0 = (tile / batch) * TILE_M;
      const int nvalid = min(TILE_M, rows - i0);
      const int next_tile = tile + gridDim.x;
      int rl[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) rl[k] = min(rt[k], nvalid - 1);
      int ln_slot = 0;
      // l == -1 (first tile only): this tile's own stage 0.  Afterwards stage 0 of tile t+1 runs inside tile t's last layer.
      for (int l = first_tile ? -1 : 0; l < n_layers; ++l) {
        const uint32_t acc = li & 1, use = li >> 1;
        const bool last_layer = l + 1 == n_layers;
        if (l < 0 || last_layer) {
          // Stage 0 of the next tile is assembled here, before this tile's last epilogue: each operand slot is refilled as soon
          // as the last layer's MMAs have read it (empty_a), so the assembly overlaps the tail of those MMAs and the next
          // tile's first layer then runs on the other accumulator under this tile's last epilogue.  (Placed before any
          // per-layer state is live: the assembly needs the registers.)
          const int t = l < 0 ? tile : next_tile;
          if (t < num_tiles) {
            if constexpr (MODE == 1) {
              stage0_fast(t);
            } else {
              stage0(t);
            }
          }
          if (l < 0) continue;
          tr.ev(600 + l);
          mbar_wait(bar_full_d + 8 * acc, use & 1, ch.status);
          tc_fence_after();
          tr.ev(610 + l);
        }
        const TcLayer& L = ch.layer[l];
        const bool has_ln = L.ln_g != nullptr;
        if constexpr (MODE == 1) {
#define GW_LF(M) case (M): layer_fast(std::integral_constant<int, (M)>{}, l, acc, use, last_layer, bs, i0, nvalid, ln_slot); break
          switch (L.kind) {
            GW_LF(F_ADD0 | F_ADD1 | F_RELU | F_FEEDS);  // edge layer 1: gathered P[src] + P[dst]
            GW_LF(F_ADD0 | F_RELU | F_FEEDS);           // encoder edge layer 1: broadcast constant term
            GW_LF(F_RELU | F_FEEDS);                    // hidden layers
            GW_LF(F_LN | F_RES | F_OUT);                // last layer of an edge / node MLP
            GW_LF(F_LN | F_RES | F_OUT | F_FEEDS);      // ... whose rows are also the operand of the next block's P products
            GW_LF(F_LN | F_FEEDS);                      // LayerNorm feeding the next MLP of the same chain
            GW_LF(F_OUT);                               // per-node products P = x W^T
            GW_LF(F_RELU | F_OUT);
            default: layer_fast(std::integral_constant<int, -1>{}, l, acc, use, last_layer, bs, i0, nvalid, ln_slot); break;
          }
#undef GW_LF
          if (has_ln) ++ln_slot;
          ++li;
          tr.ev(900 + l);
          continue;
        }
        const int N = L.N, nval = L.n_valid;
        const int np = (N + 63) >> 6;
        const float wsi = L.wscale_inv;
        const uint32_t bias_o = OFF_PAR + 4 * cofs + l * 1024;
        const bool has_add0 = L.add[0].kind != SRC_NONE, has_add1 = L.add[1].kind != SRC_NONE;
        const bool has_res = L.residual.kind != SRC_NONE, has_out = L.out != nullptr;
        const bool relu = L.relu != 0, feeds = L.feeds_next != 0;
        const uint32_t g_o = OFF_LNP + 4 * cofs + (ln_slot * 2) * 1024, b_o = g_o + 1024;
        if (has_ln) ++ln_slot;
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256 + 16 * hq;

        // Epilogue operands that come from global memory, prefetched one chunk ahead into pf0 / pf1:
        //   pf0 = addend 0 (before the activation) or, on layers without addends, the residual (after LayerNorm); pf1 = addend 1
        const RowSrc& src0 = has_add0 ? L.add[0] : L.residual;
        const bool has0 = has_add0 || has_res;
        int r0[4] = {0, 0, 0, 0}, r1[4] = {0, 0, 0, 0};
        float pf0[16] = {}, pf1[16] = {};
        auto prefetch = [&](int s) {
          if (64 * s + 16 * hq >= N || ABL3(ABL_LOADS)) return;
          if (has0) load16(src0, bs, r0, 64 * s + cofs, lc, pf0);
          if (has_add1) load16(L.add[1], bs, r1, 64 * s + cofs, lc, pf1);
        };
        if (has0) rows_of(src0, i0, rl, r0);
        if (has_add1) rows_of(L.add[1], i0, rl, r1);
        prefetch(0);
        if (!last_layer) {
          tr.ev(600 + l);
          mbar_wait(bar_full_d + 8 * acc, use & 1, ch.status);
          tc_fence_after();
          tr.ev(610 + l);
        }

        // LayerNorm statistics (first pass over the accumulator).  Each thread reduces its 16 columns of each of its 4 rows
        // around a pivot (the row's first value it sees), the partial (mean, M2) pairs are merged with Chan's formula over
        // the 4 lanes and 4 warps that share a row: one pass, no cancellation, one barrier.
        float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f};
        if (has_ln && !ABL3(ABL_LN)) {
          float pv[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
          int cnt = 0;
          for (int s = 0; s < np; ++s) {
            if (64 * s + 16 * hq >= N) break;
            float v[16];
            tmem_ld_16x256b_x2(taddr + 64 * s, v);
            tmem_ld_16x256b_x2(taddr + 64 * s + (16u << 16), v + 8);
            tmem_wait_ld();
            const float4 b4 = *reinterpret_cast<const float4*>(smem + bias_o + 256 * s);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], wsi, ((i >> 2) & 1) ? ((i & 1) ? b4.w : b4.z) : ((i & 1) ? b4.y : b4.x));
            if (s == 0) {
#pragma unroll
              for (int k = 0; k < 4; ++k) pv[k] = v[8 * (k >> 1) + 2 * (k & 1)];
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int k = 2 * (i >> 3) + ((i >> 1) & 1);
              const float d = v[i] - pv[k];
              s1[k] += d;
              s2[k] = fmaf(d, d, s2[k]);
            }
            cnt += 4;
          }
          float m2[4];
          const float fc = (float)cnt, ic = cnt > 0 ? 1.0f / fc : 0.f;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float a = s1[k] * ic;
            mean[k] = pv[k] + a;
            m2[k] = fmaxf(s2[k] - s1[k] * a, 0.f);
          }
          float nn = fc;  // values per partial; the four lanes of a row hold equally many
#pragma unroll
          for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float mb = __shfl_xor_sync(0xffffffffu, mean[k], o), qb = __shfl_xor_sync(0xffffffffu, m2[k], o);
              const float d = mb - mean[k];
              mean[k] = 0.5f * (mean[k] + mb);
              m2[k] = (m2[k] + qb) + d * d * (0.5f * nn);
            }
            nn *= 2.f;
          }
          if (lc == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) ln_x[hq * 128 + rt[k]] = mean[k], ln_y[hq * 128 + rt[k]] = m2[k];
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // the four warps of this lane quadrant
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float n = 0.f, mu = 0.f, q2 = 0.f;
#pragma unroll
            for (int w = 0; w < WSPLIT; ++w) {  // same order in every thread of the row
              const int cw = (N - 16 * w + 63) >> 6;  // chunks in which warp w owns columns
              const float nw = cw > 0 ? 16.f * (float)cw : 0.f;
              if (nw > 0.f) {
                const float mw = ln_x[w * 128 + rt[k]], qw = ln_y[w * 128 + rt[k]];
                const float d = mw - mu, nt = n + nw;
                mu += d * (nw / nt);
                q2 += qw + d * d * (n * nw / nt);
                n = nt;
              }
            }
            mean[k] = mu;
            rstd[k] = 1.0f / sqrtf(q2 / (float)nval + 1e-5f);
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + q) : "memory");  // ln_x / ln_y may be rewritten by the next LayerNorm
        }

        for (int s = 0; s < np; ++s) {
          const bool have = 64 * s + 16 * hq < N;  // warp-uniform
          const int col = 64 * s + cofs;
          float v[16];
          if (have && !ABL3(ABL_TMEM)) {
            tmem_ld_16x256b_x2(taddr + 64 * s, v);
            tmem_ld_16x256b_x2(taddr + 64 * s + (16u << 16), v + 8);
            tmem_wait_ld();
            const float4 b4 = *reinterpret_cast<const float4*>(smem + bias_o + 256 * s);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = fmaf(v[i], wsi, ((i >> 2) & 1) ? ((i & 1) ? b4.w : b4.z) : ((i & 1) ? b4.y : b4.x));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
          }
          if (s + 1 == np) {  // my last read of this accumulator
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty_d + 8 * acc);
          }
          if (have) {
            const bool ld_ok = !ABL3(ABL_LOADS);
            if (has_add0 && ld_ok) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += pf0[i];
            }
            if (has_add1 && ld_ok) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += pf1[i];
            }
            if (relu) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (has_ln) {
              const float4 g4 = *reinterpret_cast<const float4*>(smem + g_o + 256 * s);
              const float4 e4 = *reinterpret_cast<const float4*>(smem + b_o + 256 * s);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                const int k = 2 * (i >> 3) + ((i >> 1) & 1);
                const float gx = ((i >> 2) & 1) ? ((i & 1) ? g4.w : g4.z) : ((i & 1) ? g4.y : g4.x);
                const float ex = ((i >> 2) & 1) ? ((i & 1) ? e4.w : e4.z) : ((i & 1) ? e4.y : e4.x);
                v[i] = fmaf((v[i] - mean[k]) * rstd[k], gx, ex);
              }
            }
            if (!has_add0 && has_res && ld_ok) {
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += pf0[i];
            }
          }
          if (s + 1 < np) prefetch(s + 1);  // in flight while this chunk is converted and stored
          if (have) {
            if (nval < N) {  // padded output columns (e.g. 78 of 80) must stay exactly zero
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (col + 2 * ((i >> 2) & 1) + (i & 1) >= nval) v[i] = 0.f;
            }
            if (has_out && !ABL3(ABL_STORES)) {
              float* ob = L.out + ((size_t)bs * rows + i0) * (size_t)L.ldo + col;
              const bool inside = 64 * s + 16 * hq + 16 <= L.out_cols;  // warp-uniform
              if (inside && vec4_ok(ob - col, L.ldo)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (rt[k] < nvalid) {
                    const int i = 8 * (k >> 1) + 2 * (k & 1);
                    *reinterpret_cast<float4*>(ob + (size_t)rt[k] * (size_t)L.ldo) = make_float4(v[i], v[i + 1], v[i + 4], v[i + 5]);
                  }
                }
              } else if (inside && vec2_ok(ob - col, L.ldo)) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (rt[k] < nvalid) {
                    float* orow = ob + (size_t)rt[k] * (size_t)L.ldo;
                    const int i = 8 * (k >> 1) + 2 * (k & 1);
                    *reinterpret_cast<float2*>(orow) = make_float2(v[i], v[i + 1]);
                    *reinterpret_cast<float2*>(orow + 2) = make_float2(v[i + 4], v[i + 5]);
                  }
                }
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  if (rt[k] < nvalid) {
                    float* orow = ob + (size_t)rt[k] * (size_t)L.ldo;
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                      const int i = 8 * (k >> 1) + 4 * j + 2 * (k & 1), c = col + 2 * j;
                      if (c < L.out_cols) orow[2 * j] = v[i];
                      if (c + 1 < L.out_cols) orow[2 * j + 1] = v[i + 1];
                    }
                  }
                }
              }
            }
          }
          if (feeds) {
            const uint32_t slot = (fi + s) % A_SLOTS;
            if (!ABL3(ABL_CONVERT)) store_operand16(smem + OFF_A + slot * A_SLOT_BYTES, rt, hq, lc, v, split, amax);
            publish(slot);
          }
          tr.ev(700 + 10 * l + s);
        }
        if (feeds) fi += np;
        ++li;
        tr.ev(900 + l);
      }
    }
    if (split && ch.status && amax > 60000.f) atomicOr(ch.status, 1);  // operand left the fp16 range: results invalid
  }

  tc_fence_before();
  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

}  // namespace t3

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static bool simple_kind(int k) { return k == SRC_STREAM || k == SRC_BCAST || k == SRC_GATHER || k == SRC_BGATHER; }
// a source every thread may read with 8-byte loads over `need` columns
// byte offsets inside a fast-path source fit 32 bits: gathered rows are addressed relative to the sample, contiguous rows
// relative to the tile
static bool fits32(const RowSrc& s) {
  const bool g = s.kind == SRC_GATHER || s.kind == SRC_BGATHER || s.kind == SRC_GATHER_BCAST_RELU;
  const long long span = g ? (long long)s.src_rows : 128;
  return span * (long long)s.ld * 4 + 4096 < (1ll << 32);
}
static bool src_fast(const RowSrc& s, int need) {
  return simple_kind(s.kind) && s.width >= need && aligned16(s.base + s.col0) && !(s.ld & 3) && fits32(s);
}

// Marks which parts of a chain take the lean full-width path (ch.fast) and the epilogue kind of every layer.
static void tc3_mark_lean(TcChain& ch) {
  using namespace t3;
  // which parts take the lean full-width path
  ch.fast = 0;
  {
    bool ok = true;
    int wsum = 0;
    for (int a = 0; a < 2; ++a) {
      const RowSrc& s = ch.a0[a];
      if (s.kind == SRC_NONE) continue;
      const bool gbr = s.kind == SRC_GATHER_BCAST_RELU;
      ok = ok && (simple_kind(s.kind) || gbr) && !(s.width & 63) && aligned16(s.base + s.col0) && !(s.ld & 3) && fits32(s);
      if (gbr) ok = ok && aligned16(s.base2) && !(s.ld2 & 3);
      wsum += s.width;
    }
    if (ok && wsum == ch.K0 && ch.a0[0].kind != SRC_NONE) ch.fast |= (int32_t)0x80000000u;
  }
  for (int l = 0; l < ch.n_layers; ++l) {
    const TcLayer& L = ch.layer[l];
    bool ok = (L.N == 256 || (L.N == 128 && !L.ln_g)) && L.n_valid == L.N;
    for (int a = 0; a < 2; ++a)
      if (L.add[a].kind != SRC_NONE) ok = ok && src_fast(L.add[a], L.N);
    if (L.residual.kind != SRC_NONE) ok = ok && src_fast(L.residual, L.N);
    if (L.out) ok = ok && aligned16(L.out) && !(L.ldo & 3) && L.out_cols >= L.N && L.add[1].kind == SRC_NONE && L.ldo < (1 << 20);
    if (ok) ch.fast |= 1 << l;
    const int f = (L.add[0].kind != SRC_NONE ? F_ADD0 : 0) | (L.add[1].kind != SRC_NONE ? F_ADD1 : 0) | (L.relu ? F_RELU : 0) |
                  (L.ln_g ? F_LN : 0) | (L.residual.kind != SRC_NONE ? F_RES : 0) | (L.out ? F_OUT : 0) | (L.feeds_next ? F_FEEDS : 0);
    static const int kinds[] = {F_ADD0 | F_ADD1 | F_RELU | F_FEEDS, F_ADD0 | F_RELU | F_FEEDS, F_RELU | F_FEEDS, F_LN | F_RES | F_OUT,
                                F_LN | F_FEEDS, F_OUT, F_RELU | F_OUT, F_LN | F_RES | F_OUT | F_FEEDS};
    ch.layer[l].kind = -1;
    for (int k : kinds)
      if (k == f) ch.layer[l].kind = f;
  }
}
cudaError_t launch_chain_tc3(const TcChain& ch_in, cudaStream_t stream) {
  TcChain ch = ch_in;
  using namespace t3;
  static int num_sms[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (num_sms[dev] == 0) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    const void* fns[4] = {(const void*)gw_chain_tc3_kernel<true, 0>, (const void*)gw_chain_tc3_kernel<true, 1>,
                          (const void*)gw_chain_tc3_kernel<false, 0>, (const void*)gw_chain_tc3_kernel<false, 1>};
    for (int i = 0; i < 4; ++i) {
      e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
      if (e != cudaSuccess) return e;
    }
    num_sms[dev] = n;
  }
  const long long R = (long long)ch.rows_per_sample * ch.batch;
  if (R <= 0 || ch.n_layers <= 0) return cudaSuccess;
  // structural requirements of the kernel
  if (ch.n_layers > PAR_LAYERS || ch.K0 <= 0 || (ch.K0 & 63)) return cudaErrorInvalidValue;
  int n_ln = 0;
  if (ch.a0[1].kind != SRC_NONE && (ch.a0[0].width & 63)) return cudaErrorInvalidValue;  // a chunk never straddles two sources
  if (ch.layer[ch.n_layers - 1].feeds_next) return cudaErrorInvalidValue;

  for (int l = 0; l < ch.n_layers; ++l) {
    const TcLayer& L = ch.layer[l];
    if (!L.Wp || (L.K & 63) || (L.N & 15) || L.N > 256 || L.N <= 0 || L.n_valid <= 0 || L.n_valid > L.N) return cudaErrorInvalidValue;
    if (L.feeds_next && (L.N & 63)) return cudaErrorInvalidValue;
    if (L.ln_g && L.add[0].kind != SRC_NONE) return cudaErrorInvalidValue;  // addends are applied before ReLU, not before LayerNorm
    if (L.ln_g && (L.n_valid != L.N || ++n_ln > 2)) return cudaErrorInvalidValue;
    if (L.add[0].kind == SRC_NONE && L.add[1].kind != SRC_NONE) return cudaErrorInvalidValue;
    for (int a = 0; a < 2; ++a)
      if (L.add[a].kind != SRC_NONE && L.add[a].kind != SRC_STREAM && L.add[a].kind != SRC_BCAST && L.add[a].kind != SRC_GATHER &&
          L.add[a].kind != SRC_BGATHER)
        return cudaErrorInvalidValue;
    if (L.residual.kind != SRC_NONE && L.residual.kind != SRC_STREAM && L.residual.kind != SRC_BCAST && L.residual.kind != SRC_GATHER &&
        L.residual.kind != SRC_BGATHER)
      return cudaErrorInvalidValue;
    if (L.add[0].kind != SRC_NONE && L.residual.kind != SRC_NONE) return cudaErrorInvalidValue;  // they share the prefetch registers
    if (l == 0 && L.K != ch.K0) return cudaErrorInvalidValue;
    if (l > 0 && !L.reuse_a && (!ch.layer[l - 1].feeds_next || ch.layer[l - 1].N != L.K)) return cudaErrorInvalidValue;
    if (L.reuse_a && (l == 0 || L.K != ch.layer[l - 1].K || ch.layer[l - 1].feeds_next || L.K > 64 * A_SLOTS)) return cudaErrorInvalidValue;  // the whole operand must still be resident
  }
  const int tiles = ((ch.rows_per_sample + TILE_M - 1) / TILE_M) * ch.batch;
  const int grid = tiles < num_sms[dev] ? tiles : num_sms[dev];
  tc3_mark_lean(ch);
  if (getenv("GW_TC3_NOFAST")) ch.fast = 0;
  const int32_t all = (int32_t)(0x80000000u | ((1u << ch.n_layers) - 1u));
  // (mode 2, per-part selection inside one kernel, measured slower than the general path: both paths' live state spills)
  const int mode = ch.fast == all ? 1 : 0;
  for (int a = 0; a < 2; ++a)
    if (ch.a0[a].kind == SRC_SEGSUM) return cudaErrorInvalidValue;  // reduce with gw_segsum_kernel first (a fused per-thread
                                                                    // reduction in stage 0 was measured slower than the kernel)
#define GW_LAUNCH3(SPLIT_, MODE_) gw_chain_tc3_kernel<SPLIT_, MODE_><<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ch)
  if (ch.split) {
    if (mode == 1) GW_LAUNCH3(true, 1); else GW_LAUNCH3(true, 0);
  } else {
    if (mode == 1) GW_LAUNCH3(false, 1); else GW_LAUNCH3(false, 0);
  }
#undef GW_LAUNCH3
  count_launch();
  return cudaGetLastError();
}

}  // namespace gw
