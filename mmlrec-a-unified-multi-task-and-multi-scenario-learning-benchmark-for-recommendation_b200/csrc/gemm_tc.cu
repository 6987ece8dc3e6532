// K3, bf16 throughput mode: grouped GEMM on the 5th-generation tensor cores (tcgen05), operands
// staged by TMA into 128B-swizzled shared memory, fp32 accumulators in TMEM, warp-specialised
// persistent CTAs (1 per SM):
//   warp 0        TMA producer (one elected lane)
//   warp 1        TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..9    epilogue: tcgen05.ld -> bias (smem) / activation -> smem transpose -> ReLU-mask /
//                 accumulate -> coalesced fp32 and/or bf16 stores
// D[M,N] = A * B^T.  Each operand is either K-major (row-major [rows,K]) or MN-major ([K,rows]);
// the MN-major form is what wgrad needs (dW = dZ^T X reads both activations "transposed"), so no
// transposed copy of any activation is ever written.  The bias gradient of a wgrad problem
// (row sums of A) is produced by one extra N=16 MMA per k-step against a constant all-ones B tile.
#include <cuda.h>

#include "common.cuh"

namespace mmlrec {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
constexpr int TC_STAGES = 5;
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_ACC_COLS = 256;                 // TMEM columns per accumulator stage (128 main + 16 row-sum, padded)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_EPI_WARPS = 8;                 // two warps per TMEM lane quarter, 64 columns each
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_MAX_PROBLEMS = 96;              // tile table cached in shared memory
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;    // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;    // 16 KB
constexpr int TC_ONES_BYTES = 16 * 128;          // 16 rows x 128 B of bf16 1.0
constexpr int TC_STAGE_LD = 34;                  // transpose tile row stride: even (64-bit accesses), conflict-free per half-warp
constexpr int TC_STAGE_TILE_FLOATS = 32 * TC_STAGE_LD;
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + TC_ONES_BYTES + 256 +
                              TC_EPI_WARPS * TC_STAGE_TILE_FLOATS * 4 + 2 * TC_BN * 4 +
                              (2 * TC_MAX_PROBLEMS + 2) * 4;

struct alignas(128) TcRecord {
  CUtensorMap tmA;
  CUtensorMap tmB;
  float* C_f32; int64_t ldc_f32;
  uint16_t* C_bf16; int64_t ldc_bf16;
  const float* bias;
  const uint16_t* mask; int64_t ldmask;
  float* rowsum_a;
  int32_t M, N, K;
  int32_t act, accumulate;
  int32_t a_mn, b_mn;
  int32_t tiles_n;
};

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tmap), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_ld1(uint32_t taddr, uint32_t& r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): 128B swizzle.
//  K-major : rows of 128 B (64 bf16 of K), 8-row groups 1024 B apart (SBO); LBO unused.
//  MN-major: K-rows of 128 B (64 bf16 of M/N), 8-row groups 1024 B apart (SBO); the next 64 M/N
//            elements live in the next TMA box, 8192 B further (LBO).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, bool mn_major) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);                       // start address  [0,14)
  d |= (uint64_t)(mn_major ? (8192 >> 4) : 1) << 16;             // leading byte offset [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;                              // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                                        // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                                        // SWIZZLE_128B
  return d;
}
// instruction descriptor (InstrDescriptor): bf16 x bf16 -> fp32
__device__ __forceinline__ uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                    // c_format = F32
  d |= 1u << 7;                    // a_format = BF16
  d |= 1u << 10;                   // b_format = BF16
  d |= (a_mn ? 1u : 0u) << 15;     // a_major
  d |= (b_mn ? 1u : 0u) << 16;     // b_major
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;
}

// Branch-free store of one full 32x32 chunk from the transpose tile (lane = column pair, lanes 0-15 even
// rows, lanes 16-31 odd rows): bias (+ReLU), optional read-modify-write, 8-byte fp32 / 4-byte bf16x2 stores.
__device__ __forceinline__ void st_global_f2(float* p, float2 v) {
  asm volatile("st.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_global_u32(void* p, uint32_t v) {
  asm volatile("st.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float2 ld_global_f2(const float* p) {
  float2 v;
  asm volatile("ld.global.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}

template <bool RELU, bool F32, bool BF16, bool ACC>
__device__ __forceinline__ void store_chunk_fast(const float2* __restrict__ sread, float* __restrict__ pf, int64_t ldcf,
                                                 uint16_t* __restrict__ pb, int64_t ldcb, float bx, float by) {
  // all shared-memory reads first (explicit global-space stores below cannot alias them, but a generic
  // store would make the compiler serialise LDS -> math -> store per row)
  float2 x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = sread[i * TC_STAGE_LD];
  float2 old[16];
  if (F32 && ACC) {
#pragma unroll
    for (int i = 0; i < 16; ++i) old[i] = ld_global_f2(pf + (int64_t)(2 * i) * ldcf);
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float2 v = x[i];
    v.x += bx; v.y += by;
    if (RELU) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); }
    if (F32 && ACC) { v.x += old[i].x; v.y += old[i].y; }
    if (F32) st_global_f2(pf + (int64_t)(2 * i) * ldcf, v);
    if (BF16) st_global_u32(pb + (int64_t)(2 * i) * ldcb, pack_bf16x2(v.x, v.y));
  }
}

struct TileCoord { int pi, tm, tn; };
// prefix / tiles_n live in shared memory (copied once per CTA): the lookup costs no global latency
__device__ __forceinline__ TileCoord locate_tile(int t, const int32_t* s_prefix, const int32_t* s_tiles_n, int n_problems) {
  int lo = 0, hi = n_problems - 1;
  while (lo < hi) {  // last problem whose first tile is <= t
    int mid = (lo + hi + 1) >> 1;
    if (s_prefix[mid] <= t) lo = mid; else hi = mid - 1;
  }
  const int local = t - s_prefix[lo];
  const int tiles_n = s_tiles_n[lo];
  TileCoord c;
  c.pi = lo; c.tm = local / tiles_n; c.tn = local - c.tm * tiles_n;
  return c;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_grouped_tc_kernel(const TcRecord* __restrict__ recs, const int32_t* __restrict__ prefix, int n_problems, int total_tiles,
                       const int32_t* __restrict__ tile_order, const int32_t* __restrict__ cta_start,
                       long long* __restrict__ dbg) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment is required by the 128B swizzle atom
  // (offset arithmetic on the shared array itself, so the compiler keeps emitting LDS/STS, not generic LD/ST)
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + TC_STAGES * TC_A_BYTES;
  unsigned char* sOnes = smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + TC_ONES_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem base slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_ACC_STAGES);
  float* stage_s = reinterpret_cast<float*>(sOnes + TC_ONES_BYTES + 256);           // [EPI_WARPS][32][33]
  float* bias_s = stage_s + TC_EPI_WARPS * TC_STAGE_TILE_FLOATS;                     // [2][128]
  int32_t* s_prefix = reinterpret_cast<int32_t*>(bias_s + 2 * TC_BN);                // [MAX_PROBLEMS + 1]
  int32_t* s_tiles_n = s_prefix + TC_MAX_PROBLEMS + 1;                               // [MAX_PROBLEMS]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + TC_STAGES);
  const uint32_t tfull_bar = smem_u32(bars + 2 * TC_STAGES), tempty_bar = smem_u32(bars + 2 * TC_STAGES + TC_ACC_STAGES);

  // all-ones B tile for the row-sum MMA; tile table -> shared memory
  for (int i = threadIdx.x; i < TC_ONES_BYTES / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  for (int i = threadIdx.x; i <= n_problems; i += TC_THREADS) s_prefix[i] = prefix[i];
  for (int i = threadIdx.x; i < n_problems; i += TC_THREADS) s_tiles_n[i] = recs[i].tiles_n;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < TC_ACC_STAGES; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, TC_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // ones tile (generic writes) -> async proxy
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // this CTA's tiles: a host-computed balanced schedule (tile_order[cta_start[b] .. cta_start[b+1])) or,
  // without one, round-robin over the tile index space
  const int sched_begin = tile_order ? cta_start[blockIdx.x] : (int)blockIdx.x;
  const int sched_end = tile_order ? cta_start[blockIdx.x + 1] : total_tiles;
  const int sched_step = tile_order ? 1 : (int)gridDim.x;
  // optional per-tile clock stamps of CTA 0 (debug timeline): dbg[tile_iter * 16 + slot]
  const bool stamp = dbg != nullptr && blockIdx.x == 0;
#define TC_STAMP(iter, slot) do { if (stamp && (iter) < 64) dbg[(iter) * 16 + (slot)] = clock64(); } while (0)

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int pit = 0;
      for (int ti = sched_begin; ti < sched_end; ti += sched_step, ++pit) {
        const int t = tile_order ? __ldg(tile_order + ti) : ti;
        TC_STAMP(pit, 0);
        const TileCoord tc = locate_tile(t, s_prefix, s_tiles_n, n_problems);
        const TcRecord* R = recs + tc.pi;
        const int K = R->K, a_mn = R->a_mn, b_mn = R->b_mn;
        const int m0 = tc.tm * TC_BM, n0 = tc.tn * TC_BN;
        const int num_kb = (K + TC_BK - 1) / TC_BK;
        TC_STAMP(pit, 1);
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          if (kb == 0) TC_STAMP(pit, 2);
          const uint32_t fb = full_bar + 8 * stage;
          mbar_expect_tx(fb, TC_A_BYTES + TC_B_BYTES);
          const uint32_t a_dst = smem_u32(sA + stage * TC_A_BYTES), b_dst = smem_u32(sB + stage * TC_B_BYTES);
          const int k0 = kb * TC_BK;
          if (!a_mn) {
            tma_load_2d(a_dst, &R->tmA, fb, k0, m0);                 // box {64 k, 128 rows}
          } else {
            tma_load_2d(a_dst, &R->tmA, fb, m0, k0);                 // box {64 m, 64 k} x 2
            tma_load_2d(a_dst + 8192, &R->tmA, fb, m0 + 64, k0);
          }
          if (!b_mn) {
            tma_load_2d(b_dst, &R->tmB, fb, k0, n0);
          } else {
            tma_load_2d(b_dst, &R->tmB, fb, n0, k0);
            tma_load_2d(b_dst + 8192, &R->tmB, fb, n0 + 64, k0);
          }
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        TC_STAMP(pit, 3);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint64_t ones_desc = make_smem_desc(smem_u32(sOnes), false);
      int mit = 0;
      for (int ti = sched_begin; ti < sched_end; ti += sched_step, ++mit) {
        const int t = tile_order ? __ldg(tile_order + ti) : ti;
        const TileCoord tc = locate_tile(t, s_prefix, s_tiles_n, n_problems);
        const TcRecord* R = recs + tc.pi;
        const int K = R->K;
        const bool a_mn = R->a_mn != 0, b_mn = R->b_mn != 0;
        const bool rowsum = (R->rowsum_a != nullptr) && tc.tn == 0;
        const uint32_t idesc = make_idesc(TC_BM, TC_BN, a_mn, b_mn);
        const uint32_t idesc_ones = make_idesc(TC_BM, 16, a_mn, false);
        const int num_kb = (K + TC_BK - 1) / TC_BK;
        TC_STAMP(mit, 4);
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        TC_STAMP(mit, 5);
        const uint32_t d_tmem = tmem_base + acc * TC_ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          if (kb == 0) TC_STAMP(mit, 6);
          const uint32_t a_addr = smem_u32(sA + stage * TC_A_BYTES), b_addr = smem_u32(sB + stage * TC_B_BYTES);
          const uint64_t a_desc = make_smem_desc(a_addr, a_mn), b_desc = make_smem_desc(b_addr, b_mn);
          // advancing K by 16 elements: 32 B inside the swizzle row (K-major) or two 8-row groups (MN-major)
          const uint64_t a_step = a_mn ? (2048 >> 4) : (32 >> 4), b_step = b_mn ? (2048 >> 4) : (32 >> 4);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint32_t accumulate = (kb | k) != 0 ? 1u : 0u;
            tc_mma(d_tmem, a_desc + a_step * k, b_desc + b_step * k, idesc, accumulate);
            if (rowsum) tc_mma(d_tmem + TC_BN, a_desc + a_step * k, ones_desc + 2 * k, idesc_ones, accumulate);
          }
          tc_commit(empty_bar + 8 * stage);   // frees the smem slot when these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar + 8 * acc);       // accumulator ready for the epilogue
        TC_STAMP(mit, 7);
        if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // warp -> TMEM lane quarter q (rows 32q..32q+31 of the tile) and column half (64 columns).
    // Per 32-column chunk: tcgen05.ld (thread = row) -> + bias (smem) -> activation -> transpose through a
    // padded smem tile -> (thread = column) ReLU-mask / accumulate / fp32 + bf16 stores, all coalesced.
    const int q = warp & 3;
    const int ew = warp - 2;                   // 0..7
    const int half = ew >> 2;                  // 0: columns [0,64)  1: columns [64,128)
    float* my_stage = stage_s + ew * TC_STAGE_TILE_FLOATS;
    int acc = 0; uint32_t acc_phase = 0;
    int it = 0;
    for (int ti = sched_begin; ti < sched_end; ti += sched_step, ++it) {
      const int t = tile_order ? __ldg(tile_order + ti) : ti;
      const TileCoord tc = locate_tile(t, s_prefix, s_tiles_n, n_problems);
      const TcRecord* R = recs + tc.pi;
      const int M = R->M, N = R->N;
      const int m_base = tc.tm * TC_BM + q * 32;
      const int n0 = tc.tn * TC_BN;
      float* const cf = R->C_f32; const int64_t ldcf = R->ldc_f32;
      uint16_t* const cb = R->C_bf16; const int64_t ldcb = R->ldc_bf16;
      const float* const bias = R->bias;
      const uint16_t* const mask = R->mask; const int64_t ldmask = R->ldmask;
      const int act = R->act, accumulate = R->accumulate;
      float* const rowsum_out = (tc.tn == 0) ? R->rowsum_a : nullptr;
      const bool estamp = stamp && ew == 0 && lane == 0;
      if (estamp && it < 64) dbg[it * 16 + 8] = clock64();
      const int my_m = m_base + lane;
      const bool row_ok = my_m < M;
      const int rows_valid = min(32, M - m_base);              // warp-uniform (may be <= 0)
      // Operands the epilogue needs from global memory are fetched NOW (128-bit loads, all in flight
      // together) so that their latency hides behind this tile's MMAs:
      //  * the ReLU mask, in row layout (thread = row), 64 B per thread per chunk
      //  * the bias of "my" two columns per chunk (column layout is used for bias + activation)
      uint4 mk[2][4];
      float2 bcol[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int nc = n0 + half * 64 + c * 32;
        const int col = nc + 2 * (lane & 15);
        bcol[c].x = (bias != nullptr && col < N) ? __ldg(bias + col) : 0.f;
        bcol[c].y = (bias != nullptr && col + 1 < N) ? __ldg(bias + col + 1) : 0.f;
        if (mask != nullptr) {
#pragma unroll
          for (int v4 = 0; v4 < 4; ++v4) {
            mk[c][v4] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
            if (row_ok && nc + v4 * 8 + 8 <= N)
              mk[c][v4] = __ldg(reinterpret_cast<const uint4*>(mask + (int64_t)my_m * ldmask + nc) + v4);
            else if (row_ok && nc + v4 * 8 < N) {  // ragged tail: element-wise
              uint32_t w[4] = {0, 0, 0, 0};
              for (int e = 0; e < 8 && nc + v4 * 8 + e < N; ++e)
                w[e >> 1] |= (uint32_t)mask[(int64_t)my_m * ldmask + nc + v4 * 8 + e] << ((e & 1) * 16);
              mk[c][v4] = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      if (estamp && it < 64) dbg[it * 16 + 9] = clock64();
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      if (estamp && it < 64) dbg[it * 16 + 10] = clock64();
      const uint32_t t_row = tmem_base + acc * TC_ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int cc = half * 64 + c * 32;     // column offset inside the tile
        const int nc = n0 + cc;
        if (nc >= N || rows_valid <= 0) break; // warp-uniform
        uint32_t r[32];
        tc_ld32(t_row + cc, r);
        tc_wait_ld();
        if (estamp && it < 64 && c == 0) dbg[it * 16 + 12] = clock64();
        // phase 1 (thread = row): ReLU mask on the raw accumulator, then 16 x STS.64 into the tile
        if (mask != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {       // keep where the bf16 mask value is > 0
            const uint4 w4 = mk[c][j >> 3];
            const uint32_t word = ((j >> 1) & 3) == 0 ? w4.x : ((j >> 1) & 3) == 1 ? w4.y : ((j >> 1) & 3) == 2 ? w4.z : w4.w;
            const uint32_t mb = (j & 1) ? (word >> 16) : (word & 0xFFFFu);
            if (!((mb & 0x8000u) == 0 && (mb & 0x7FFFu) != 0)) r[j] = 0u;
          }
        }
        float2* srow = reinterpret_cast<float2*>(my_stage + lane * TC_STAGE_LD);
#pragma unroll
        for (int j2 = 0; j2 < 16; ++j2) srow[j2] = make_float2(__uint_as_float(r[2 * j2]), __uint_as_float(r[2 * j2 + 1]));
        __syncwarp();
        if (estamp && it < 64 && c == 0) dbg[it * 16 + 13] = clock64();
        // phase 2 (thread = column pair; lanes 0-15 row 2i, lanes 16-31 row 2i+1): bias + activation,
        // read-modify-write, coalesced 8-byte fp32 / 4-byte bf16x2 stores
        const int sub = lane >> 4;
        const int col = nc + 2 * (lane & 15);
        const float2* sread = reinterpret_cast<const float2*>(my_stage + sub * TC_STAGE_LD + 2 * (lane & 15));
        float* pf = cf ? cf + (int64_t)(m_base + sub) * ldcf + col : nullptr;
        uint16_t* pb = cb ? cb + (int64_t)(m_base + sub) * ldcb + col : nullptr;
        const float bx = bcol[c].x, by = bcol[c].y;
        const bool fast = rows_valid == 32 && nc + 32 <= N && (act == MMLREC_ACT_NONE || act == MMLREC_ACT_RELU) &&
                          (cf == nullptr || (ldcf & 1) == 0) && (cb == nullptr || (ldcb & 1) == 0);   // warp-uniform
        if (fast) {
          const int kind = (act == MMLREC_ACT_RELU ? 8 : 0) | (cf ? 4 : 0) | (cb ? 2 : 0) | ((cf && accumulate) ? 1 : 0);
          switch (kind) {
            case 4:  store_chunk_fast<false, true, false, false>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 5:  store_chunk_fast<false, true, false, true>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 2:  store_chunk_fast<false, false, true, false>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 6:  store_chunk_fast<false, true, true, false>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 7:  store_chunk_fast<false, true, true, true>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 12: store_chunk_fast<true, true, false, false>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 10: store_chunk_fast<true, false, true, false>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 14: store_chunk_fast<true, true, true, false>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 13: store_chunk_fast<true, true, false, true>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            case 15: store_chunk_fast<true, true, true, true>(sread, pf, ldcf, pb, ldcb, bx, by); break;
            default: break;  // unreachable: a problem always has at least one output
          }
        } else {
          const bool c0 = col < N, c1 = col + 1 < N;
          const bool vec_f = c1 && ((ldcf & 1) == 0), vec_b = c1 && ((ldcb & 1) == 0);
#pragma unroll 1
          for (int i = 0; i < 16; ++i) {
            const int rr = 2 * i + sub;
            float2 x = sread[i * TC_STAGE_LD];
            x.x += bx; x.y += by;
            if (act == MMLREC_ACT_RELU) { x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); }
            else if (act != MMLREC_ACT_NONE) { x.x = apply_act(x.x, act); x.y = apply_act(x.y, act); }
            if (rr < rows_valid && c0) {
              if (pf) {
                if (vec_f) {
                  float2* d = reinterpret_cast<float2*>(pf);
                  if (accumulate) { const float2 o = *d; x.x += o.x; x.y += o.y; }
                  *d = x;
                } else {
                  if (accumulate) { x.x += pf[0]; if (c1) x.y += pf[1]; }
                  pf[0] = x.x;
                  if (c1) pf[1] = x.y;
                }
              }
              if (pb) {
                if (vec_b) *reinterpret_cast<uint32_t*>(pb) = pack_bf16x2(x.x, x.y);
                else { pb[0] = float_to_bf16_bits(x.x); if (c1) pb[1] = float_to_bf16_bits(x.y); }
              }
            }
            if (pf) pf += 2 * ldcf;
            if (pb) pb += 2 * ldcb;
          }
        }
        __syncwarp();
        if (estamp && it < 64 && c == 0) dbg[it * 16 + 14] = clock64();
      }
      if (estamp && it < 64) dbg[it * 16 + 11] = clock64();
      if (rowsum_out != nullptr && half == 0) {
        uint32_t rs;
        tc_ld1(t_row + TC_BN, rs);
        tc_wait_ld();
        if (m_base + lane < M) rowsum_out[m_base + lane] = __uint_as_float(rs);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
  }
}

// ---------------------------------------------------------------------------------------------
// host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (EncodeTiledFn)p;
  }
  return fn;
}

// operand stored as row-major [outer, inner] with row stride ld (elements)
static int encode_operand(CUtensorMap* tm, const uint16_t* base, int64_t ld, int64_t inner, int64_t outer,
                          int box_inner, int box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable (no driver?)"); return -2; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -3; }
  return 0;
}

}  // namespace mmlrec

using namespace mmlrec;

extern "C" int64_t mmlrec_tc_record_bytes(void) { return (int64_t)sizeof(TcRecord); }
extern "C" int32_t mmlrec_tc_num_tiles(int32_t M, int32_t N) { return cdiv(M, TC_BM) * cdiv(N, TC_BN); }

extern "C" int mmlrec_tc_encode_problem(const MmlrecGemmTcDesc* d, void* record_host) {
  MMLREC_CHECK_ARG(d && record_host, "null argument");
  MMLREC_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "bad sizes");
  MMLREC_CHECK_ARG(((uintptr_t)d->A & 15) == 0 && ((uintptr_t)d->B & 15) == 0, "operands must be 16-byte aligned");
  MMLREC_CHECK_ARG((d->lda & 7) == 0 && (d->ldb & 7) == 0, "operand row strides must be multiples of 8 elements");
  MMLREC_CHECK_ARG(d->C_f32 == nullptr || ((d->ldc_f32 & 3) == 0 && ((uintptr_t)d->C_f32 & 15) == 0), "C_f32 alignment");
  MMLREC_CHECK_ARG(d->C_bf16 == nullptr || ((d->ldc_bf16 & 7) == 0 && ((uintptr_t)d->C_bf16 & 15) == 0), "C_bf16 alignment");
  MMLREC_CHECK_ARG(d->mask == nullptr || ((d->ldmask & 7) == 0 && ((uintptr_t)d->mask & 15) == 0), "mask alignment");
  MMLREC_CHECK_ARG(d->C_f32 || d->C_bf16, "no output");
  TcRecord rec;
  memset(&rec, 0, sizeof(rec));
  int rc;
  if (!d->a_mn_major) rc = encode_operand(&rec.tmA, d->A, d->lda, d->K, d->M, TC_BK, TC_BM);
  else                rc = encode_operand(&rec.tmA, d->A, d->lda, d->M, d->K, 64, TC_BK);
  if (rc) return rc;
  if (!d->b_mn_major) rc = encode_operand(&rec.tmB, d->B, d->ldb, d->K, d->N, TC_BK, TC_BN);
  else                rc = encode_operand(&rec.tmB, d->B, d->ldb, d->N, d->K, 64, TC_BK);
  if (rc) return rc;
  rec.C_f32 = d->C_f32; rec.ldc_f32 = d->ldc_f32; rec.C_bf16 = d->C_bf16; rec.ldc_bf16 = d->ldc_bf16;
  rec.bias = d->bias; rec.mask = d->mask; rec.ldmask = d->ldmask; rec.rowsum_a = d->colsum;
  rec.M = d->M; rec.N = d->N; rec.K = d->K; rec.act = d->act; rec.accumulate = d->accumulate;
  rec.a_mn = d->a_mn_major; rec.b_mn = d->b_mn_major; rec.tiles_n = cdiv(d->N, TC_BN);
  memcpy(record_host, &rec, sizeof(rec));
  return 0;
}

static int tc_sm_count() {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  return sm_count;
}

static int launch_tc(const void* records, const int32_t* tile_prefix, int32_t n_problems, int32_t total_tiles,
                     const int32_t* tile_order, const int32_t* cta_start, int32_t n_ctas, long long* dbg, void* stream) {
  MMLREC_CHECK_ARG(records && tile_prefix && n_problems > 0 && total_tiles >= 0, "bad args");
  MMLREC_CHECK_ARG(((uintptr_t)records & 127) == 0, "record table must be 128-byte aligned");
  MMLREC_CHECK_ARG(n_problems <= TC_MAX_PROBLEMS, "too many problems in one launch (split the table)");
  MMLREC_CHECK_ARG((tile_order == nullptr) == (cta_start == nullptr), "tile_order and cta_start come together");
  if (total_tiles == 0) return 0;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(gemm_grouped_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return (int)e; }
    opted = true;
  }
  int grid = tile_order ? n_ctas : (total_tiles < tc_sm_count() ? total_tiles : tc_sm_count());
  MMLREC_CHECK_ARG(grid > 0, "empty grid");
  gemm_grouped_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(
      reinterpret_cast<const TcRecord*>(records), tile_prefix, n_problems, total_tiles, tile_order, cta_start, dbg);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int32_t mmlrec_tc_sm_count(void) { return tc_sm_count(); }

extern "C" int mmlrec_gemm_grouped_tc(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                      int32_t total_tiles, void* stream) {
  return launch_tc(records, tile_prefix, n_problems, total_tiles, nullptr, nullptr, 0, nullptr, stream);
}

extern "C" int mmlrec_gemm_grouped_tc_scheduled(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                                int32_t total_tiles, const int32_t* tile_order,
                                                const int32_t* cta_start, int32_t n_ctas, void* stream) {
  return launch_tc(records, tile_prefix, n_problems, total_tiles, tile_order, cta_start, n_ctas, nullptr, stream);
}

extern "C" int mmlrec_gemm_grouped_tc_debug(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                            int32_t total_tiles, int64_t* stamps, void* stream) {
  return launch_tc(records, tile_prefix, n_problems, total_tiles, nullptr, nullptr, 0,
                   reinterpret_cast<long long*>(stamps), stream);
}
