// gw_tc.cu -- the fused MLP-chain kernel on Blackwell tensor cores (tcgen05 + TMEM), precision GW_PREC_FP32_TC / BF16_TC.
//
// One persistent CTA per SM walks 128-row tiles of a gw::TcChain.  For every tile the whole chain
//     A0 = assemble(row sources)                                   (gather / CSR segment sum / relu(gather+const) ...)
//     for each layer:  D = A . W^T  (tcgen05.mma, fp32 accumulate in TMEM)
//                      v = D*s + bias + gathered addends ; ReLU | LayerNorm + residual
//                      v -> global (fp32)  and/or  v -> split fp16 hi/lo -> shared memory = A operand of the next layer
// runs without the activations ever leaving the SM.  This is the fusion BASELINE.json's north_star asks for: the
// reference's x[row]/x[col] gathers, cat, 3 Linear + LayerNorm, residual and scatter_sum (graph_net_block.py:131-135,
// 184-191) become one kernel per edge pass and one per node pass.
//
// fp32 fidelity on fp16 tensor cores: every fp32 operand a is split a = hi + lo (two fp16, 22 significand bits) and
// each product is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation in TMEM (the dropped lo*lo term is 2^-22
// relative).  Weights are pre-scaled by a power of two so that their lo parts stay in the fp16 normal range; the
// scale is undone exactly in the epilogue.  Cost: 3 kind::f16 MMAs per product = 1.5x a TF32 MMA, half of 3xTF32.
//
// CTA layout (448 threads, one persistent CTA per SM):
//   warp 14  lane 0: MMA issuer      -- tcgen05.mma.cta_group::1.kind::f16, M=128, N<=256, K=16 per instruction
//   warp 13  lane 0: weight producer -- streams pre-swizzled weight panels global->smem with cp.async.bulk (UBLKCP)
//   warps 8-12     : 160 movers      -- ALL global row traffic, in three groups of 64 that each own one staging buffer:
//                                       gather / stream / CSR segment-sum rows into padded fp32 staging pieces
//                                       (cp.async 16 B or index-hoisted vector loads, 8 lanes per 128 B row line) and
//                                       coalesced stores of finished rows out of staging
//   warps 0-7      : 256 workers     -- thread (q,lane,h) owns tile row 32q+lane (= TMEM lane) and the 32-column pieces
//                                       64s+32h; they convert staged rows to fp16 hi/lo operands and run every epilogue
//                                       from TMEM (bias, addend, ReLU, LayerNorm, residual) touching only TMEM and
//                                       shared memory (per-layer bias / gamma / beta are staged in shared memory once)
// Shared memory: 2 A slots x [128 x 64] (hi|lo) = 64 KB, 3 weight stages x [256 x 64] = 96 KB, 3 staging pieces
// [128 x 32] fp32 = 54 KB, mbarriers.  TMEM: 512 columns = two 128x256 fp32 accumulators, so the MMAs of layer l+1
// overlap the epilogue of layer l chunk by chunk, and the movers prefetch the next tile during the last epilogue.
// Tiles are batch-major (same rows of consecutive samples are neighbours) so broadcast constants are read from HBM
// once and served to the other samples from L2.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gw_internal.h"
#include "gw_ops.h"
#include "gw_tc_ptx.cuh"

namespace gw {

// Timing-attribution build (-DGW_ABLATE, tools/ablate.py): parts of the pipeline can be skipped at run time.  Results are
// WRONG under any non-zero mask; the product build compiles every ABL() to false.
#ifdef GW_ABLATE
#define ABL(bit) ((ch.ablate & (bit)) != 0)
#else
#define ABL(bit) false
#endif
enum { ABL_FENCE = 1, ABL_TAKE = 2, ABL_CONVERT = 4, ABL_MOVE = 8, ABL_LN = 16, ABL_TMEM = 32, ABL_MMA = 64, ABL_WEIGHTS = 128 };

constexpr int TILE_M = 128;
constexpr int A_SLOTS = 2, B_STAGES = 2, ST_BUFS = 5;
constexpr int A_HALF_BYTES = TILE_M * 128;      // [128 rows x 64 halfs]
constexpr int A_SLOT_BYTES = 2 * A_HALF_BYTES;  // hi | lo
constexpr int B_STAGE_BYTES = 256 * 128;        // [256 rows x 64 halfs], hi OR lo panel
constexpr int ST_STRIDE = 128;                  // staging row: 32 fp32; 16-byte chunk k of row r sits at chunk k ^ (r & 7)
constexpr int ST_BYTES = TILE_M * ST_STRIDE;    // one [128 rows x 32 cols] fp32 piece
constexpr int NUM_MOVERS = 160;   // five warps; warp g owns staging buffer g (pieces p with p % 5 == g)
constexpr int MOVER_GROUP = 32;
#ifndef GW_WSPLIT
#define GW_WSPLIT 2
#endif
constexpr int WSPLIT = GW_WSPLIT;          // worker threads per tile row (2 or 4): each owns 64/WSPLIT columns of every 64-column chunk
constexpr int WCOLS = 64 / WSPLIT;         // columns per worker thread per chunk (32 or 16)
constexpr int NUM_WORKERS = 128 * WSPLIT;  // thread (row, hq): tile row = TMEM lane; 4 worker warps per scheduler hide each other's latencies
constexpr int WORKER_WARPS = NUM_WORKERS / 32;
constexpr int PAR_LAYERS = 6;     // per-layer parameter rows staged in shared memory (bias; LayerNorm gamma/beta for <= 2 layers)
constexpr int NUM_THREADS = NUM_WORKERS + NUM_MOVERS + 64;  // workers first, then 5 mover warps, weight producer, MMA issuer
constexpr int WARP_MOVER0 = WORKER_WARPS, WARP_PRODUCER = WORKER_WARPS + 5, WARP_MMA = WORKER_WARPS + 6;
constexpr int OFF_A = 0;
constexpr int OFF_B = A_SLOTS * A_SLOT_BYTES;
constexpr int OFF_ST = OFF_B + B_STAGES * B_STAGE_BYTES;
constexpr int OFF_BAR = OFF_ST + ST_BUFS * ST_BYTES;
constexpr int NUM_BARS = 2 * A_SLOTS + 2 * B_STAGES + 4 + 2 * ST_BUFS;
constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
constexpr int OFF_PAR = OFF_TMEM + 16;                       // float bias[PAR_LAYERS][256]
constexpr int OFF_LNP = OFF_PAR + PAR_LAYERS * 1024;         // float gamma_beta[2][2][256]
constexpr int OFF_LN = OFF_LNP + 4 * 1024;                   // float ln_x[256], ln_y[256]: row statistics exchange
constexpr int SMEM_BYTES = OFF_LN + 2 * NUM_WORKERS * 4;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory limit");
static_assert(OFF_B % 1024 == 0 && A_SLOT_BYTES % 1024 == 0 && B_STAGE_BYTES % 1024 == 0, "SWIZZLE_128B needs 1 KB alignment");
static_assert(OFF_ST % 16 == 0 && OFF_BAR % 8 == 0, "alignment");

__device__ __forceinline__ void named_bar_workers() { asm volatile("bar.sync 1, %0;" ::"n"(NUM_WORKERS) : "memory"); }

// ------------------------------------------------------------------------------------------------------------------
// movers: global rows <-> padded fp32 staging pieces
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_arrive_noinc(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}

__device__ __forceinline__ bool vec4_ok(const float* base, int ld, int col) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ld & 3) == 0) && ((col & 3) == 0);
}
__device__ __forceinline__ bool is_simple(int kind) {
  return kind == SRC_STREAM || kind == SRC_BCAST || kind == SRC_GATHER || kind == SRC_BGATHER;
}
// row pointer of a simple source for sample b, local row i
__device__ __forceinline__ const float* simple_row(const RowSrc& s, int b, int i) {
  switch (s.kind) {
    case SRC_STREAM: return s.base + ((size_t)b * s.src_rows + i) * s.ld + s.col0;
    case SRC_BCAST: return s.base + (size_t)i * s.ld + s.col0;
    case SRC_GATHER: return s.base + ((size_t)b * s.src_rows + __ldg(s.idx + i)) * s.ld + s.col0;
    default: return s.base + (size_t)__ldg(s.idx + i) * s.ld + s.col0;  // SRC_BGATHER
  }
}
// 4 consecutive columns [c, c+4) of a row pointer; columns >= width read as 0
__device__ __forceinline__ float4 load4(const float* row, int c, int width, bool vec) {
  if (vec && c + 4 <= width) return __ldg(reinterpret_cast<const float4*>(row + c));
  float4 v;
  v.x = (c + 0 < width) ? __ldg(row + c + 0) : 0.f;
  v.y = (c + 1 < width) ? __ldg(row + c + 1) : 0.f;
  v.z = (c + 2 < width) ? __ldg(row + c + 2) : 0.f;
  v.w = (c + 3 < width) ? __ldg(row + c + 3) : 0.f;
  return v;
}
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

// value of source s at (sample b, local row i, columns [c, c+4)) -- any source kind
__device__ __forceinline__ float4 src_load4(const RowSrc& s, int b, int i, int c) {
  if (c >= s.width) return make_float4(0.f, 0.f, 0.f, 0.f);
  const bool vec = vec4_ok(s.base, s.ld, s.col0 + c);
  switch (s.kind) {
    case SRC_STREAM:
    case SRC_BCAST:
    case SRC_GATHER:
    case SRC_BGATHER:
      return load4(simple_row(s, b, i), c, s.width, vec);
    case SRC_SEGSUM: {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      const int j0 = __ldg(s.ptr + i), j1 = __ldg(s.ptr + i + 1);
      for (int j = j0; j < j1; ++j) {  // left to right: the order scatter_add visits the reference edge list
        const int e = s.perm ? __ldg(s.perm + j) : j;
        acc = add4(acc, load4(s.base + ((size_t)b * s.src_rows + e) * s.ld + s.col0, c, s.width, vec));
      }
      return acc;
    }
    case SRC_GATHER_BCAST_RELU: {
      float4 v = load4(s.base + ((size_t)b * s.src_rows + __ldg(s.idx + i)) * s.ld + s.col0, c, s.width, vec);
      v = add4(v, load4(s.base2 + (size_t)i * s.ld2, c, s.width, vec4_ok(s.base2, s.ld2, c)));
      return make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
    }
    default:
      return make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Split NV fp32 values (NV = 16 or 32) into fp16 hi/lo (or bf16) and store them as NV/8 16-byte chunks per part into the
// swizzled K-major operand tile: row r, logical 16B chunk j (j0 .. j0+NV/8-1 of the row's 8) lives at chunk position j ^ (r & 7).
template <int NV>
__device__ __forceinline__ void store_operand_piece(uint8_t* slot, int r, int j0, const float (&v)[NV], bool split, float& amax) {
  uint8_t* row_hi = slot + r * 128;
  uint8_t* row_lo = row_hi + A_HALF_BYTES;
#pragma unroll
  for (int c = 0; c < NV / 8; ++c) {
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a0 = v[8 * c + 2 * e], a1 = v[8 * c + 2 * e + 1];
      amax = fmaxf(amax, fmaxf(fabsf(a0), fabsf(a1)));
      if (split) {
        const __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
        const __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
        __half2 hh = __halves2half2(h0, h1), ll = __halves2half2(l0, l1);
        hi[e] = *reinterpret_cast<uint32_t*>(&hh);
        lo[e] = *reinterpret_cast<uint32_t*>(&ll);
      } else {
        __nv_bfloat162 bb = __floats2bfloat162_rn(a0, a1);
        hi[e] = *reinterpret_cast<uint32_t*>(&bb);
        lo[e] = 0;
      }
    }
    const int chunk = ((j0 + c) ^ (r & 7)) * 16;
    *reinterpret_cast<uint4*>(row_hi + chunk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    if (split) *reinterpret_cast<uint4*>(row_lo + chunk) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
  }
}

struct Retire {  // what a mover needs to know to finish a staged piece: store it out (if it has an output)
  float* out;    // first output row of the tile (null: nothing to store)
  int ldo, out_cols, c0, nvalid;
};

// ------------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NUM_THREADS, 1) gw_chain_tc_kernel(const __grid_constant__ TcChain ch) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rows = ch.rows_per_sample, batch = ch.batch;
  const int tiles_per_sample = (rows + TILE_M - 1) / TILE_M;
  const int num_tiles = tiles_per_sample * batch;  // batch-major: tile -> (sample = tile % batch, row block = tile / batch)
  const bool split = ch.split != 0;
  const int parts = split ? 2 : 1;

  const uint32_t bar_full_a = sbase + OFF_BAR;              // [A_SLOTS] workers -> MMA
  const uint32_t bar_empty_a = bar_full_a + 8 * A_SLOTS;    // [A_SLOTS] MMA -> workers (tcgen05.commit)
  const uint32_t bar_full_b = bar_empty_a + 8 * A_SLOTS;    // [B_STAGES] bulk copy -> MMA
  const uint32_t bar_empty_b = bar_full_b + 8 * B_STAGES;   // [B_STAGES] MMA -> producer
  const uint32_t bar_full_d = bar_empty_b + 8 * B_STAGES;   // [2] MMA -> workers: accumulator complete
  const uint32_t bar_empty_d = bar_full_d + 16;             // [2] workers -> MMA: accumulator drained
  const uint32_t bar_st_ready = bar_empty_d + 16;           // [ST_BUFS] movers -> workers: staging piece filled / free
  const uint32_t bar_st_done = bar_st_ready + 8 * ST_BUFS;  // [ST_BUFS] workers (one h group) -> movers: piece consumed / produced
  volatile uint32_t* tmem_ptr = reinterpret_cast<volatile uint32_t*>(smem + OFF_TMEM);

  if (threadIdx.x == 0) {
    if (sbase & 1023u) {  // SWIZZLE_128B operand tiles must be 1 KB aligned
      if (ch.status) atomicOr(ch.status, 4);
      __trap();
    }
    for (int i = 0; i < A_SLOTS; ++i) mbar_init(bar_full_a + 8 * i, NUM_WORKERS), mbar_init(bar_empty_a + 8 * i, 1);
    for (int i = 0; i < B_STAGES; ++i) mbar_init(bar_full_b + 8 * i, 1), mbar_init(bar_empty_b + 8 * i, 1);
    for (int i = 0; i < 2; ++i) mbar_init(bar_full_d + 8 * i, 1), mbar_init(bar_empty_d + 8 * i, NUM_WORKERS);
    for (int i = 0; i < ST_BUFS; ++i) mbar_init(bar_st_ready + 8 * i, MOVER_GROUP), mbar_init(bar_st_done + 8 * i, NUM_WORKERS);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) {  // TMEM: all 512 columns (two fp32 accumulators of 256 columns)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(sbase + OFF_TMEM), "r"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // per-layer column parameters -> shared memory (zero beyond n_valid), read by the epilogues as broadcast LDS.128
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

  // Warp ids are assigned by priority: the SM's warp arbiter favours the highest eligible warp id, so the roles that
  // others wait FOR (MMA issuer 14, weight producer 13, movers 8-12) sit above the workers (0-7), which spend much of
  // their time polling mbarriers.  (With the opposite order the pollers starve the very warps they are waiting on.)
  if (warp == WARP_PRODUCER) {
    // ===================================== weight producer =====================================================
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
              if (ABL(ABL_WEIGHTS)) {
                mbar_arrive(bar_full_b + 8 * stage);
                continue;
              }
              mbar_expect_tx(bar_full_b + 8 * stage, panel);
              bulk_g2s(sbase + OFF_B + stage * B_STAGE_BYTES, w + (size_t)(kc * parts + part) * panel, panel,
                       bar_full_b + 8 * stage);
            }
          }
        }
      }
    }
  } else if (warp == WARP_MMA) {
    // ===================================== MMA issuer ==========================================================
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
            tr.ev(200 + kc);
            const uint32_t a_hi = sbase + OFF_A + slot * A_SLOT_BYTES, a_lo = a_hi + A_HALF_BYTES;
            {  // hi weight panel: A_hi.B_hi (+ A_lo.B_hi)
              const uint32_t stage = bi % B_STAGES, nb = bi / B_STAGES;
              mbar_wait(bar_full_b + 8 * stage, nb & 1, ch.status);
              tc_fence_after();
              const uint32_t b = sbase + OFF_B + stage * B_STAGE_BYTES;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (!ABL(ABL_MMA)) tc_mma_f16(d_tmem, umma_desc(a_hi + 32 * ks), umma_desc(b + 32 * ks), idesc, (kc | ks) != 0);
              if (split && !ABL(ABL_MMA)) {
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) tc_mma_f16(d_tmem, umma_desc(a_lo + 32 * ks), umma_desc(b + 32 * ks), idesc, 1);
              }
              tc_commit(bar_empty_b + 8 * stage);
              ++bi;
            }
            if (split) {  // lo weight panel: A_hi.B_lo
              const uint32_t stage = bi % B_STAGES, nb = bi / B_STAGES;
              mbar_wait(bar_full_b + 8 * stage, nb & 1, ch.status);
              tc_fence_after();
              const uint32_t b = sbase + OFF_B + stage * B_STAGE_BYTES;
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                if (!ABL(ABL_MMA)) tc_mma_f16(d_tmem, umma_desc(a_hi + 32 * ks), umma_desc(b + 32 * ks), idesc, 1);
              tc_commit(bar_empty_b + 8 * stage);
              ++bi;
            }
            if (last_use) tc_commit(bar_empty_a + 8 * slot);  // the operand slot may be refilled
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
  } else if (warp >= WARP_MOVER0) {
    // ===================================== movers ==============================================================
    // Mover warp g owns staging buffer g and fills the pieces p with p % ST_BUFS == g, so ST_BUFS pieces are in flight.
    //  * aligned stream / broadcast / gather rows: 16-byte cp.async (LDGSTS) straight into staging, 8 lanes per 128-byte
    //    row line (a TMA bulk copy per 128-byte row was measured at ~40 cycles of issue each: too slow at this size);
    //  * everything else (CSR segment sums, unaligned rows such as the 102-wide features): loads + st.shared.
    // Staging rows are 128 B with the 16-byte chunks XOR-swizzled by (row & 7): conflict-free for the movers' 8-lanes-per-
    // row writes and for the workers' thread-per-row reads.  Output pieces are drained with coalesced 16-byte stores
    // when the buffer is recycled.
    const int mgroup = warp - WARP_MOVER0;
    const int rsub = lane >> 3, ck = lane & 7;
    const uint32_t st_buf = sbase + OFF_ST + mgroup * ST_BYTES;
    const uint32_t bar_ready = bar_st_ready + 8 * mgroup, bar_done = bar_st_done + 8 * mgroup;
    uint32_t pn = 0, my_use = 0;
    Tracer tr;
    tr.init(ch.trace, 2 + (mgroup < 3 ? mgroup : 5), lane == 0 && mgroup < 3);
    Retire prev;  // my previous piece (same buffer): must be retired before the buffer is refilled
    prev.out = nullptr, prev.ldo = 0, prev.out_cols = 0, prev.c0 = 0, prev.nvalid = 0;
    // swizzled staging address of (row r, my chunk)
    auto st_addr = [&](int r) -> uint32_t { return st_buf + r * ST_STRIDE + ((ck ^ (r & 7)) << 4); };

    auto retire = [&]() {  // wait until the workers are done with my previous piece, then store its output rows
      mbar_wait(bar_done, (my_use - 1) & 1, ch.status);
      if (prev.out && !ABL(ABL_MOVE)) {
        const int col = prev.c0 + 4 * ck;
        const bool vec = vec4_ok(prev.out, prev.ldo, col) && col + 4 <= prev.out_cols;
        if (col < prev.out_cols) {
#pragma unroll 8
          for (int j = 0; j < 32; ++j) {
            const int r = rsub + 4 * j;
            if (r < prev.nvalid) {
              const float4 v = lds128(st_addr(r));
              float* o = prev.out + (size_t)r * prev.ldo + col;
              if (vec) {
                *reinterpret_cast<float4*>(o) = v;
              } else {
                o[0] = v.x;
                if (col + 1 < prev.out_cols) o[1] = v.y;
                if (col + 2 < prev.out_cols) o[2] = v.z;
                if (col + 3 < prev.out_cols) o[3] = v.w;
              }
            }
          }
        }
        __syncwarp();  // all reads of the buffer precede its refill
      }
    };
    // One staging piece from one source; s0 == SRC_NONE is an output-only piece (the buffer is just handed to the
    // workers).  Only called for pieces this warp owns.
    auto fill = [&](const RowSrc& s0, int c, int bs, int i0, int nvalid, const Retire& rt) {
      tr.ev(10000 + (int)pn);
      if (my_use > 0) retire();
      tr.ev(20000 + (int)pn);
      prev = rt;
      ++my_use;
      const int kind = s0.kind;
      if (kind == SRC_NONE || ABL(ABL_MOVE)) {
        mbar_arrive(bar_ready);
        return;
      }
      const int col = c + 4 * ck;
      const int width = s0.width, ldi = s0.ld, col0 = s0.col0;
      const float* base = s0.base;
      const int rmax = nvalid - 1;
      if (is_simple(kind) && vec4_ok(base, ldi, col0 + c) && c + 32 <= width) {
        const bool gathered = kind == SRC_GATHER || kind == SRC_BGATHER;
        const float* cbase = base + col0 + col;
        const size_t ld = (size_t)ldi;
        const size_t boff = (kind == SRC_STREAM || kind == SRC_GATHER) ? (size_t)bs * (size_t)s0.src_rows : 0;
        const int32_t* idx = s0.idx;
#pragma unroll
        for (int part = 0; part < 4; ++part) {
          int rowi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int i = i0 + min(rsub + 4 * (8 * part + j), rmax);
            rowi[j] = gathered ? __ldg(idx + i) : i;
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) cp_async16(st_addr(rsub + 4 * (8 * part + j)), cbase + (boff + (size_t)rowi[j]) * ld);
        }
        cp_async_arrive_noinc(bar_ready);
        tr.ev(30000 + (int)pn);
        return;
      }
      if (kind == SRC_SEGSUM && !s0.perm && vec4_ok(base, ldi, col0 + c) && c + 32 <= width) {
        // CSR segment sums over contiguous edge rows (in-degree 6/7 on the mesh and decoder graphs)
        const float* tb = base + (size_t)bs * s0.src_rows * ldi + col0 + col;
        const int32_t* ptr = s0.ptr;
        const size_t ld = (size_t)ldi;
#pragma unroll 1
        for (int part = 0; part < 4; ++part) {
          int j0[8], j1[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int i = i0 + min(rsub + 4 * (8 * part + j), rmax);
            j0[j] = __ldg(ptr + i), j1[j] = __ldg(ptr + i + 1);
          }
#pragma unroll 2
          for (int j = 0; j < 8; ++j) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int e = j0[j];
            for (; e + 4 <= j1[j]; e += 4) {  // four rows in flight, summed left to right (reference order)
              const float4 a0 = __ldg(reinterpret_cast<const float4*>(tb + (size_t)e * ld));
              const float4 a1 = __ldg(reinterpret_cast<const float4*>(tb + (size_t)(e + 1) * ld));
              const float4 a2 = __ldg(reinterpret_cast<const float4*>(tb + (size_t)(e + 2) * ld));
              const float4 a3 = __ldg(reinterpret_cast<const float4*>(tb + (size_t)(e + 3) * ld));
              acc = add4(add4(add4(add4(acc, a0), a1), a2), a3);
            }
            for (; e < j1[j]; ++e) acc = add4(acc, __ldg(reinterpret_cast<const float4*>(tb + (size_t)e * ld)));
            sts128(st_addr(rsub + 4 * (8 * part + j)), acc);
          }
        }
      } else {  // generic (unaligned / partial-width / permuted) path
#pragma unroll 2
        for (int j = 0; j < 32; ++j) {
          const int r = rsub + 4 * j;
          sts128(st_addr(r), src_load4(s0, bs, i0 + min(r, rmax), col));
        }
      }
      mbar_arrive(bar_ready);  // release: my part of the piece is written (32 arrivals complete the phase)
    };
    // every mover and every worker enumerates the pieces identically; `mine()` advances the shared piece counter
    auto mine = [&]() -> bool { return (pn++ % ST_BUFS) == (uint32_t)mgroup; };

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int bs = tile % batch, i0 = (tile / batch) * TILE_M;
      const int nvalid = min(TILE_M, rows - i0);
      Retire no_out;
      no_out.out = nullptr, no_out.ldo = 0, no_out.out_cols = 0, no_out.c0 = 0, no_out.nvalid = nvalid;
      // stage-0 operand pieces: per 64-column chunk, per source part, per half
      const int nk0 = ch.K0 >> 6, w0 = ch.a0[0].width, w1 = ch.a0[1].kind != SRC_NONE ? ch.a0[1].width : 0;
      for (int c = 0; c < nk0; ++c) {
        const int colc = 64 * c;
        const bool first = colc < w0;
        const int rel = first ? colc : colc - w0;
        if (first || rel < w1) {
          const RowSrc& src = first ? ch.a0[0] : ch.a0[1];
          if (src.kind == SRC_GATHER_BCAST_RELU) {  // relu(gather + broadcast) is staged as its two addends
            for (int h = 0; h < 2; ++h)
              if (mine()) {
                RowSrc g = src;
                g.kind = SRC_GATHER;
                fill(g, rel + 32 * h, bs, i0, nvalid, no_out);
              }
            for (int h = 0; h < 2; ++h)
              if (mine()) {
                RowSrc t;
                t.kind = SRC_BCAST, t.base = src.base2, t.ld = src.ld2, t.width = src.width, t.col0 = 0;
                fill(t, rel + 32 * h, bs, i0, nvalid, no_out);
              }
          } else {
            for (int h = 0; h < 2; ++h)
              if (mine()) fill(src, rel + 32 * h, bs, i0, nvalid, no_out);  // columns past the width read as 0
          }
        } else {  // zero padding of K0
          for (int h = 0; h < 2; ++h)
            if (mine()) {
              RowSrc z = ch.a0[0];
              z.width = 0, z.kind = SRC_STREAM;
              fill(z, 32 * h, bs, i0, nvalid, no_out);
            }
        }
      }
      for (int l = 0; l < ch.n_layers; ++l) {
        const TcLayer& L = ch.layer[l];
        const bool has_add0 = L.add[0].kind != SRC_NONE, has_add1 = L.add[1].kind != SRC_NONE;
        const bool has_ro = L.residual.kind != SRC_NONE || L.out != nullptr;
        if (!has_add0 && !has_ro) continue;
        const int N = L.N;
        const int np = (N + 63) >> 6;
        for (int s = 0; s < np; ++s) {
          for (int a = 0; a < 2; ++a) {
            if (!(a == 0 ? has_add0 : has_add1)) continue;
            for (int h = 0; h < 2; ++h) {
              const int c0 = 64 * s + 32 * h;
              if (c0 < N && mine()) fill(L.add[a], c0, bs, i0, nvalid, no_out);
            }
          }
          if (has_ro)
            for (int h = 0; h < 2; ++h) {
              const int c0 = 64 * s + 32 * h;
              if (c0 < N && mine()) {
                Retire rt = no_out;
                if (L.out) {
                  rt.out = L.out + ((size_t)bs * rows + i0) * L.ldo;
                  rt.ldo = L.ldo, rt.out_cols = L.out_cols, rt.c0 = c0;
                }
                fill(L.residual, c0, bs, i0, nvalid, rt);
              }
            }
        }
      }
    }
    if (my_use > 0) retire();  // drain my last piece
  } else {
    // ===================================== workers: operand conversion + epilogues ============================
    // Thread (q, lane, hq): tile row r = 32q + lane (= TMEM lane; a warp can only read the lane quadrant warp % 4, which
    // is also its scheduler), columns [64 s + WCOLS*hq, +WCOLS) of every 64-column chunk s.  Staging pieces are the two
    // 32-column halves of a chunk; thread hq reads its WCOLS columns of the half h = hq / (WSPLIT/2).
    const int q = warp & 3;
    const int hq = warp >> 2;                     // 0 .. WSPLIT-1
    const int h = hq / (WSPLIT / 2);              // which 32-column staging half holds my columns
    const int sub = hq % (WSPLIT / 2);            // which WCOLS-column part of that half
    const int r = 32 * q + lane;                  // tile row == TMEM lane
    const int wtid = hq * 128 + r;
    const int rsw = r & 7;
    constexpr int NQ = WCOLS / 4;                 // 16-byte groups per thread per piece (8 or 4)
    const int j0 = hq * (WCOLS / 8);              // my first logical 16-byte chunk in an operand row (8 chunks of 8 halfs)
    const uint32_t st_row = sbase + OFF_ST + r * ST_STRIDE;  // + buffer * ST_BYTES + ((k ^ rsw) << 4) for 16-byte group k
    const uint32_t par_base = sbase + OFF_PAR, lnp_base = sbase + OFF_LNP;
    float* ln_x = reinterpret_cast<float*>(smem + OFF_LN);
    float* ln_y = ln_x + NUM_WORKERS;
    uint32_t fi = 0, li = 0, pn = 0;
    float amax = 0.f;
    Tracer tr;
    tr.init(ch.trace, 5 + (hq & 1), q == 0 && lane == 0 && hq < 2);

    auto piece_wait = [&](uint32_t p) { mbar_wait(bar_st_ready + 8 * (p % ST_BUFS), (p / ST_BUFS) & 1, ch.status); };
    auto piece_done = [&](uint32_t p) { mbar_arrive(bar_st_done + 8 * (p % ST_BUFS)); };
    // A staging buffer serves pieces of either column half in turn.  Threads whose columns are in the other half still
    // wait for the piece's ready phase and arrive on its done barrier ("observe" it), so that every worker sees every
    // phase of every barrier in order and a buffer is never refilled while some warp has yet to pass the previous phase
    // (a parity wait would otherwise be ambiguous by two phases).  take(): my half -> read my WCOLS floats and release.
    auto take = [&](uint32_t p, bool mine, float (&v)[WCOLS], bool accumulate) {
      tr.ev(10000 + (int)p);
      piece_wait(p);
      tr.ev(20000 + (int)p);
      if (mine && !ABL(ABL_TAKE)) {
        const uint32_t a = st_row + (p % ST_BUFS) * ST_BYTES;
#pragma unroll
        for (int k = 0; k < NQ; ++k) {
          const float4 t = lds128(a + (((sub * NQ + k) ^ rsw) << 4));
          if (accumulate) {
            v[4 * k] += t.x, v[4 * k + 1] += t.y, v[4 * k + 2] += t.z, v[4 * k + 3] += t.w;
          } else {
            v[4 * k] = t.x, v[4 * k + 1] = t.y, v[4 * k + 2] = t.z, v[4 * k + 3] = t.w;
          }
        }
      }
      piece_done(p);
    };
    auto piece_write = [&](uint32_t p, const float (&v)[WCOLS]) {
      const uint32_t a = st_row + (p % ST_BUFS) * ST_BYTES;
#pragma unroll
      for (int k = 0; k < NQ; ++k)
        sts128(a + (((sub * NQ + k) ^ rsw) << 4), make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]));
    };
    // v = acc * wscale_inv + bias for my WCOLS columns (bias row lives in shared memory, zero past n_valid)
    auto scale_bias = [&](float (&v)[WCOLS], float wsi, uint32_t bias_s) {
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float4 b4 = lds128(bias_s + 16 * k);
        v[4 * k] = fmaf(v[4 * k], wsi, b4.x), v[4 * k + 1] = fmaf(v[4 * k + 1], wsi, b4.y);
        v[4 * k + 2] = fmaf(v[4 * k + 2], wsi, b4.z), v[4 * k + 3] = fmaf(v[4 * k + 3], wsi, b4.w);
      }
    };

    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      // ---- stage 0: staged fp32 rows -> fp16 hi/lo operand chunks -----------------------------------------------
      {
        const int nk0 = ch.K0 >> 6, w0 = ch.a0[0].width;
        for (int c = 0; c < nk0; ++c, ++fi) {
          const uint32_t slot = fi % A_SLOTS, n = fi / A_SLOTS;
          const int colc = 64 * c;
          const RowSrc& src = (colc < w0) ? ch.a0[0] : ch.a0[1];
          const bool two = (colc < w0 || ch.a0[1].kind != SRC_NONE) && src.kind == SRC_GATHER_BCAST_RELU;
          float v[WCOLS];
          take(pn + 0, h == 0, v, false);
          take(pn + 1, h == 1, v, false);
          if (two) {
            take(pn + 2, h == 0, v, true);
            take(pn + 3, h == 1, v, true);
#pragma unroll
            for (int j = 0; j < WCOLS; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          pn += two ? 4 : 2;
          tr.ev(500 + c);
          mbar_wait(bar_empty_a + 8 * slot, (n & 1) ^ 1, ch.status);
          tr.ev(510 + c);
          if (!ABL(ABL_CONVERT)) store_operand_piece<WCOLS>(smem + OFF_A + slot * A_SLOT_BYTES, r, j0, v, split, amax);
          if (!ABL(ABL_FENCE)) fence_proxy_async();
          mbar_arrive(bar_full_a + 8 * slot);
          tr.ev(520 + c);
        }
      }
      // ---- layers -------------------------------------------------------------------------------------------------
      int ln_slot = 0;
      for (int l = 0; l < ch.n_layers; ++l, ++li) {
        const TcLayer& L = ch.layer[l];
        const uint32_t acc = li & 1, use = li >> 1;
        const int N = L.N, nval = L.n_valid;
        const int np = (N + 63) >> 6;
        const float wsi = L.wscale_inv;
        const uint32_t bias_s = par_base + l * 1024;
        const bool has_add0 = L.add[0].kind != SRC_NONE, has_add1 = L.add[1].kind != SRC_NONE;
        const bool has_res = L.residual.kind != SRC_NONE;
        const bool has_out = L.out != nullptr;
        const bool has_ro = has_res || has_out;
        const bool relu = L.relu != 0, has_ln = L.ln_g != nullptr, feeds = L.feeds_next != 0;
        const uint32_t g_s = lnp_base + (ln_slot * 2) * 1024, b_s = g_s + 1024;
        if (has_ln) ++ln_slot;
        tr.ev(600 + l);
        mbar_wait(bar_full_d + 8 * acc, use & 1, ch.status);
        tc_fence_after();
        tr.ev(610 + l);
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + acc * 256;
        float mean = 0.f, rstd = 1.f;
        if (has_ln && !ABL(ABL_LN)) {
          // LayerNorm statistics of this row (shared by the WSPLIT threads of the row): mean, then centred second moment,
          // like torch's CPU kernel, both straight from TMEM (cheaper than holding the row in registers).
          float s1 = 0.f;
          for (int s = 0; s < np; ++s) {
            const int c0 = 64 * s + WCOLS * hq;
            if (c0 >= N) break;
            float v[WCOLS];
            tmem_ldw(taddr + c0, v);
            scale_bias(v, wsi, bias_s + 4 * c0);
#pragma unroll
            for (int j = 0; j < WCOLS; ++j) s1 += v[j];
          }
          ln_x[wtid] = s1;
          named_bar_workers();
          float tot = 0.f;
#pragma unroll
          for (int t = 0; t < WSPLIT; ++t) tot += ln_x[t * 128 + r];  // same order for every thread of the row
          mean = tot / (float)nval;
          float s2 = 0.f;
          for (int s = 0; s < np; ++s) {
            const int c0 = 64 * s + WCOLS * hq;
            if (c0 >= N) break;
            float v[WCOLS];
            tmem_ldw(taddr + c0, v);
            scale_bias(v, wsi, bias_s + 4 * c0);
#pragma unroll
            for (int j = 0; j < WCOLS; ++j) {
              const float x = v[j] - mean;
              s2 = fmaf(x, x, s2);
            }
          }
          ln_y[wtid] = s2;
          named_bar_workers();
          float tot2 = 0.f;
#pragma unroll
          for (int t = 0; t < WSPLIT; ++t) tot2 += ln_y[t * 128 + r];
          rstd = 1.0f / sqrtf(tot2 / (float)nval + 1e-5f);
        }
        for (int s = 0; s < np; ++s) {
          const int c0 = 64 * s + WCOLS * hq;
          const bool have = c0 < N;                    // my columns of this chunk exist
          const bool have1 = 64 * s + 32 < N;          // the second 32-column half of this chunk exists
          float v[WCOLS];
          if (have && !ABL(ABL_TMEM)) {
            tmem_ldw(taddr + c0, v);
            scale_bias(v, wsi, bias_s + 4 * c0);
          } else {
#pragma unroll
            for (int j = 0; j < WCOLS; ++j) v[j] = 0.f;
          }
          // staged pieces of this chunk in the movers' order: add0(h0,h1), add1(h0,h1), residual/out(h0,h1)
          if (has_add0) {
            take(pn++, h == 0 && have, v, true);
            if (have1) take(pn++, h == 1 && have, v, true);
          }
          if (has_add1) {
            take(pn++, h == 0 && have, v, true);
            if (have1) take(pn++, h == 1 && have, v, true);
          }
          if (have) {
            if (relu) {
#pragma unroll
              for (int j = 0; j < WCOLS; ++j) v[j] = fmaxf(v[j], 0.f);
            }
            if (has_ln) {
              const float nm = -mean * rstd;
#pragma unroll
              for (int k = 0; k < NQ; ++k) {
                const float4 g4 = lds128(g_s + 4 * c0 + 16 * k), e4 = lds128(b_s + 4 * c0 + 16 * k);
                v[4 * k] = fmaf(fmaf(v[4 * k], rstd, nm), g4.x, e4.x);
                v[4 * k + 1] = fmaf(fmaf(v[4 * k + 1], rstd, nm), g4.y, e4.y);
                v[4 * k + 2] = fmaf(fmaf(v[4 * k + 2], rstd, nm), g4.z, e4.z);
                v[4 * k + 3] = fmaf(fmaf(v[4 * k + 3], rstd, nm), g4.w, e4.w);
              }
            }
            if (nval < N) {  // padded output columns (e.g. 78 of 80) must stay exactly zero
#pragma unroll
              for (int j = 0; j < WCOLS; ++j)
                if (c0 + j >= nval) v[j] = 0.f;
            }
          }
          if (has_ro) {
            for (int hh = 0; hh < (have1 ? 2 : 1); ++hh, ++pn) {
              piece_wait(pn);
              if (hh == h && have) {
                if (has_res && !ABL(ABL_TAKE)) {
                  const uint32_t a = st_row + (pn % ST_BUFS) * ST_BYTES;
#pragma unroll
                  for (int k = 0; k < NQ; ++k) {
                    const float4 t = lds128(a + (((sub * NQ + k) ^ rsw) << 4));
                    v[4 * k] += t.x, v[4 * k + 1] += t.y, v[4 * k + 2] += t.z, v[4 * k + 3] += t.w;
                  }
                }
                if (has_out && !ABL(ABL_TAKE)) piece_write(pn, v);  // in place: each thread overwrites exactly the bytes it read
              }
              piece_done(pn);
            }
          }
          tr.ev(700 + 10 * l + s);
          if (feeds) {  // publish this 64-column chunk of the next operand once all WSPLIT column parts are written
            const uint32_t f = fi + s, slot = f % A_SLOTS, n = f / A_SLOTS;
            mbar_wait(bar_empty_a + 8 * slot, (n & 1) ^ 1, ch.status);
            tr.ev(800 + 10 * l + s);
            if (!ABL(ABL_CONVERT)) store_operand_piece<WCOLS>(smem + OFF_A + slot * A_SLOT_BYTES, r, j0, v, split, amax);
            if (!ABL(ABL_FENCE)) fence_proxy_async();
            mbar_arrive(bar_full_a + 8 * slot);
          }
        }
        if (feeds) fi += np;
        tc_fence_before();
        mbar_arrive(bar_empty_d + 8 * acc);  // this thread no longer reads the accumulator
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

// ------------------------------------------------------------------------------------------------------------------
// weight packing (one-off per weight set)
// ------------------------------------------------------------------------------------------------------------------
static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
size_t tc_packed_bytes(int K_src, int N_src, int parts) {
  return (size_t)(round_up(K_src, 64) / 64) * parts * round_up(N_src, 16) * 128;
}

// dst image: for chunk kc, part p: panel of N rows x 128 B; element (n, k): 16B chunk ((k%64)/8) ^ (n&7), half k%8
// perm16 (gw_tc3.cu): inside every group of 16 output rows and of 16 K columns, packed position a holds logical index
// f(a) = 4*((a>>1)&3) + 2*(a>>3) + (a&1), the order in which a tcgen05.ld.16x256b fragment gives each thread 4 consecutive features.
__host__ __device__ inline int perm16_f(int a) { return (a & ~15) | (4 * ((a >> 1) & 3) + 2 * ((a >> 3) & 1) + (a & 1)); }
__global__ void gw_pack_weights_kernel(const float* __restrict__ W, int ldw, int K_src, int N_src, int Kp, int Np,
                                       float wscale, int parts, int perm16, uint8_t* __restrict__ dst) {
  const size_t total = (size_t)Np * Kp;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int n = (int)(e / Kp), k = (int)(e % Kp);
    const int ns = perm16 ? perm16_f(n) : n, ks = perm16 ? perm16_f(k) : k;
    const float w = (ns < N_src && ks < K_src) ? W[(size_t)ns * ldw + ks] * wscale : 0.f;
    const int kc = k >> 6, kk = k & 63;
    const size_t panel = (size_t)Np * 128;
    const size_t off = (size_t)n * 128 + (size_t)((((kk >> 3) ^ (n & 7)) << 4) + ((kk & 7) << 1));
    if (parts == 2) {
      const __half hi = __float2half_rn(w);
      const __half lo = __float2half_rn(w - __half2float(hi));
      *reinterpret_cast<__half*>(dst + (size_t)(kc * 2 + 0) * panel + off) = hi;
      *reinterpret_cast<__half*>(dst + (size_t)(kc * 2 + 1) * panel + off) = lo;
    } else {
      *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)kc * panel + off) = __float2bfloat16_rn(w);
    }
  }
}

cudaError_t launch_pack_weights(const float* W, int ldw, int K_src, int N_src, float wscale, int parts, int perm16, void* dst,
                                cudaStream_t stream) {
  const int Kp = round_up(K_src, 64), Np = round_up(N_src, 16);
  gw_pack_weights_kernel<<<256, 256, 0, stream>>>(W, ldw, K_src, N_src, Kp, Np, wscale, parts, perm16, static_cast<uint8_t*>(dst));
  count_launch();
  return cudaGetLastError();
}

__global__ void gw_absmax_kernel(const float* __restrict__ W, int ldw, int K_src, int N_src, float* out_max) {
  float m = 0.f;
  const size_t total = (size_t)N_src * K_src;
  for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(W[(e / K_src) * (size_t)ldw + (e % K_src)]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(out_max), __float_as_int(m));  // m >= 0: int order == float order
}

cudaError_t launch_absmax(const float* W, int ldw, int K_src, int N_src, float* out_max, cudaStream_t stream) {
  gw_absmax_kernel<<<64, 256, 0, stream>>>(W, ldw, K_src, N_src, out_max);
  count_launch();
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------------------------
cudaError_t launch_chain_tc(const TcChain& ch, cudaStream_t stream) {
  static int num_sms[64] = {0};
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (dev < 0 || dev >= 64) return cudaErrorInvalidDevice;
  if (num_sms[dev] == 0) {
    int n = 0;
    e = cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gw_chain_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
    num_sms[dev] = n;
  }
  const long long R = (long long)ch.rows_per_sample * ch.batch;
  if (R <= 0 || ch.n_layers <= 0) return cudaSuccess;
  // structural requirements of the kernel
  if (ch.n_layers > PAR_LAYERS || ch.K0 <= 0 || (ch.K0 & 63)) return cudaErrorInvalidValue;
  int n_ln = 0;
  if (ch.a0[1].kind != SRC_NONE && (ch.a0[0].width & 31)) return cudaErrorInvalidValue;
  for (int l = 0; l < ch.n_layers; ++l) {
    const TcLayer& L = ch.layer[l];
    if (!L.Wp || (L.K & 63) || (L.N & 15) || L.N > 256 || L.N <= 0 || L.n_valid <= 0 || L.n_valid > L.N) return cudaErrorInvalidValue;
    if (L.feeds_next && (L.N & 63)) return cudaErrorInvalidValue;
    if (L.ln_g && L.add[0].kind != SRC_NONE) return cudaErrorInvalidValue;   // addends are applied before ReLU, not before LayerNorm
    if (L.ln_g && (L.n_valid != L.N || ++n_ln > 2)) return cudaErrorInvalidValue;
    if (L.add[0].kind == SRC_NONE && L.add[1].kind != SRC_NONE) return cudaErrorInvalidValue;
    if (l == 0 && L.K != ch.K0) return cudaErrorInvalidValue;
    if (l > 0 && !L.reuse_a && (!ch.layer[l - 1].feeds_next || ch.layer[l - 1].N != L.K)) return cudaErrorInvalidValue;
    if (L.reuse_a && (l == 0 || L.K != ch.layer[l - 1].K || ch.layer[l - 1].feeds_next || L.K > 64 * A_SLOTS)) return cudaErrorInvalidValue;  // the whole operand must still be resident
  }
  const int tiles = ((ch.rows_per_sample + TILE_M - 1) / TILE_M) * ch.batch;
  const int grid = tiles < num_sms[dev] ? tiles : num_sms[dev];
  gw_chain_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(ch);
  count_launch();
  return cudaGetLastError();
}

}  // namespace gw
