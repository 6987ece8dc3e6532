// K3, bf16 throughput mode: grouped GEMM on the 5th-generation tensor cores (tcgen05), operands
// staged by TMA into 128B-swizzled shared memory, fp32 accumulators in TMEM, warp-specialised
// persistent CTAs (1 per SM):
//   warp 0        TMA producer (one elected lane), 4-stage mbarrier ring of {A 16 KB, B 16 KB} k-blocks
//   warp 1        TMEM allocator + tcgen05.mma issuer (one elected lane), 3 accumulator stages
//   warps 2..9    epilogue: tcgen05.ld (thread = row) -> ReLU-mask / bias / activation in registers ->
//                 128B-swizzled shared-memory box -> ONE bulk tensor store per box (cp.async.bulk.tensor,
//                 fp32 and/or bf16; `accumulate` = cp.reduce.async.bulk.tensor .add).  The TMA unit clips
//                 rows >= M / columns >= N, so ragged tiles take the same path as full ones.
// D[M,N] = A * B^T.  Each operand is either K-major (row-major [rows,K]) or MN-major ([K,rows]);
// the MN-major form is what wgrad needs (dW = dZ^T X reads both activations "transposed"), so no
// transposed copy of any activation is ever written.  The bias gradient of a wgrad problem
// (row sums of A) is produced by one extra N=16 MMA per k-step against a constant all-ones B tile.
#include "tc_common.cuh"

namespace mmlrec {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
constexpr int TC_STAGES = 4;
constexpr int TC_ACC_STAGES = 3;
constexpr int TC_ACC_COLS = 160;                 // TMEM columns per accumulator stage (128 main + 16 row-sum, padded to 32)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_EPI_WARPS = 8;                 // two warps per TMEM lane quarter, 64 columns each
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_MAX_PROBLEMS = 96;              // tile table cached in shared memory
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;    // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;    // 16 KB
constexpr int TC_ONES_BYTES = 16 * 128;          // 16 rows x 128 B of bf16 1.0
constexpr int TC_OUT_BUF_BYTES = 32 * 128;       // one TMA-store box: 32 rows x 128 B (32 fp32 or 64 bf16 columns), 128B-swizzled
constexpr int TC_OUT_BUFS = 2;                   // per epilogue warp (a store reads one while the next chunk fills the other)
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * (TC_A_BYTES + TC_B_BYTES) +
                              TC_EPI_WARPS * TC_OUT_BUFS * TC_OUT_BUF_BYTES + TC_ONES_BYTES + 256 +
                              TC_EPI_WARPS * 64 * 4 + (2 * TC_MAX_PROBLEMS + 2) * 4;

struct alignas(128) TcRecord {
  CUtensorMap tmA;
  CUtensorMap tmB;
  CUtensorMap tmC32;                            // fp32 output, box {32 cols, 32 rows}
  CUtensorMap tmC16;                            // bf16 output, box {64 cols, 32 rows}
  const float* bias;
  const uint16_t* mask; int64_t ldmask;
  float* rowsum_a;
  int32_t M, N, K;
  int32_t act, accumulate;
  int32_t a_mn, b_mn;
  int32_t tiles_n;
  int32_t has_f32, has_bf16;
};

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_grouped_tc_kernel(const TcRecord* __restrict__ recs, const int32_t* __restrict__ prefix, int n_problems, int total_tiles,
                       const int32_t* __restrict__ tile_order, const int32_t* __restrict__ cta_start,
                       long long* __restrict__ dbg) {
  extern __shared__ unsigned char smem_dyn[];
  if (dbg != nullptr && threadIdx.x == 0) {   // profiling: kernel entry of this CTA on the global timer
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); dbg[1024 + blockIdx.x * 8 + 5] = (long long)gt;
  }
  // 1024-byte alignment is required by the 128B swizzle atom
  // (offset arithmetic on the shared array itself, so the compiler keeps emitting LDS/STS, not generic LD/ST)
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + TC_STAGES * TC_A_BYTES;
  unsigned char* sOut = smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES);                    // [EPI_WARPS][OUT_BUFS][4096], 1024-aligned
  unsigned char* sOnes = sOut + TC_EPI_WARPS * TC_OUT_BUFS * TC_OUT_BUF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + TC_ONES_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+ACC) tmem_full, [2S+ACC,2S+2ACC) tmem_empty, then tmem base slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_ACC_STAGES);
  float* bias_s = reinterpret_cast<float*>(sOnes + TC_ONES_BYTES + 256);             // [EPI_WARPS][64]
  int32_t* s_prefix = reinterpret_cast<int32_t*>(bias_s + TC_EPI_WARPS * 64);        // [MAX_PROBLEMS + 1]
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
  fence_async_smem();  // ones tile (generic writes) -> async proxy
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
  // ... and of every CTA on the global timer (ns): dbg[1024 + cta * 8 + {0 setup done, 1 producer done, 2 MMA issuer
  // done, 3 epilogue warp 0 done, 4 number of tiles, 5 kernel entry, 6 TMEM released}]
#define TC_CTA_STAMP(slot) do { if (dbg != nullptr) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
                                  dbg[1024 + blockIdx.x * 8 + (slot)] = (long long)gt_; } } while (0)
  if (threadIdx.x == 0) { TC_CTA_STAMP(0); if (dbg != nullptr) dbg[1024 + blockIdx.x * 8 + 4] = (sched_end - sched_begin + sched_step - 1) / sched_step; }

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
        if (pit == 0) { tma_prefetch_desc(&R->tmA); tma_prefetch_desc(&R->tmB); }
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
        // descriptors of the next tile's problem (a descriptor fetch is otherwise exposed in front of its first load)
        const int tn_i = ti + sched_step;
        if (tn_i < sched_end) {
          const int t2 = tile_order ? __ldg(tile_order + tn_i) : tn_i;
          const TileCoord tc2 = locate_tile(t2, s_prefix, s_tiles_n, n_problems);
          if (tc2.pi != tc.pi) { tma_prefetch_desc(&recs[tc2.pi].tmA); tma_prefetch_desc(&recs[tc2.pi].tmB); }
        }
      }
      TC_CTA_STAMP(1);
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
          // 16-wide k-slices that lie entirely beyond K hold only TMA zero fill: skip their MMAs
          const int k_left = K - kb * TC_BK;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            if (k * 16 < k_left) {
              const uint32_t accumulate = (kb | k) != 0 ? 1u : 0u;
              tc_mma(d_tmem, a_desc + a_step * k, b_desc + b_step * k, idesc, accumulate);
              if (rowsum) tc_mma(d_tmem + TC_BN, a_desc + a_step * k, ones_desc + 2 * k, idesc_ones, accumulate);
            }
          }
          tc_commit(empty_bar + 8 * stage);   // frees the smem slot when these MMAs retire
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(tfull_bar + 8 * acc);       // accumulator ready for the epilogue
        TC_STAMP(mit, 7);
        if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      TC_CTA_STAMP(2);
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // warp -> TMEM lane quarter q (rows 32q..32q+31 of the tile) and column half (64 columns = two 32-column
    // chunks).  Both chunks are pulled out of TMEM at once (thread = row), after which the accumulator stage is
    // handed back to the MMA warp; mask / bias / activation run on registers; each finished box (32 rows x 128 B)
    // is written to a 128B-swizzled staging buffer with conflict-free 16-byte stores and leaves with one bulk
    // tensor store issued by lane 0.
    const int q = warp & 3;
    const int ew = warp - 2;                   // 0..7
    const int half = ew >> 2;                  // 0: columns [0,64)  1: columns [64,128)
    float* bias_w = bias_s + ew * 64;
    const uint32_t out_base = smem_u32(sOut + ew * TC_OUT_BUFS * TC_OUT_BUF_BYTES);
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    int acc = 0; uint32_t acc_phase = 0;
    int it = 0;
    uint32_t buf_i = 0;                        // staging buffers are used round-robin
    for (int ti = sched_begin; ti < sched_end; ti += sched_step, ++it) {
      const int t = tile_order ? __ldg(tile_order + ti) : ti;
      const TileCoord tc = locate_tile(t, s_prefix, s_tiles_n, n_problems);
      const TcRecord* R = recs + tc.pi;
      const int M = R->M, N = R->N;
      const int m_base = tc.tm * TC_BM + q * 32;
      const int n0 = tc.tn * TC_BN + half * 64;
      const bool has_f32 = R->has_f32 != 0, has_bf16 = R->has_bf16 != 0;
      const float* const bias = R->bias;
      const uint16_t* const mask = R->mask; const int64_t ldmask = R->ldmask;
      const int act = R->act, accumulate = R->accumulate;
      float* const rowsum_out = (tc.tn == 0 && half == 0) ? R->rowsum_a : nullptr;
      const bool estamp = stamp && ew == 0 && lane == 0;
      if (estamp && it < 64) dbg[it * 16 + 8] = clock64();
      const int my_m = m_base + lane;
      const bool row_ok = my_m < M;
      const bool rows_any = m_base < M;                        // warp-uniform
      const bool c_ok[2] = {rows_any && n0 < N, rows_any && n0 + 32 < N};   // warp-uniform
      if (lane == 0) {
        if (has_f32) tma_prefetch_desc(&R->tmC32);
        if (has_bf16) tma_prefetch_desc(&R->tmC16);
      }
      // Operands the epilogue needs from global memory are fetched NOW so that their latency hides behind this
      // tile's MMAs: the bias of the warp's 64 columns (-> shared-memory slot, read back as broadcasts) and the
      // ReLU mask in row layout (thread = row, 64 B per chunk)
      __syncwarp();                                            // the previous tile's reads of the slot are done
      {
        const int c_lo = n0 + lane, c_hi = n0 + 32 + lane;
        bias_w[lane] = (bias != nullptr && c_lo < N) ? __ldg(bias + c_lo) : 0.f;
        bias_w[32 + lane] = (bias != nullptr && c_hi < N) ? __ldg(bias + c_hi) : 0.f;
      }
      uint4 mk[2][4];
      if (mask != nullptr) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const int nc = n0 + c * 32;
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
      __syncwarp();                                            // bias slot visible to the whole warp
      if (estamp && it < 64) dbg[it * 16 + 9] = clock64();
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      if (estamp && it < 64) dbg[it * 16 + 10] = clock64();
      const uint32_t t_row = tmem_base + acc * TC_ACC_COLS + ((uint32_t)(q * 32) << 16);
      uint32_t r0[32], r1[32];
      uint32_t rs = 0;
      if (c_ok[0]) tc_ld32(t_row + half * 64, r0);
      if (c_ok[1]) tc_ld32(t_row + half * 64 + 32, r1);
      if (rowsum_out != nullptr) tc_ld1(t_row + TC_BN, rs);
      tc_wait_ld();
      // the accumulator stage is free as soon as this warp's values sit in registers
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar + 8 * acc);
      if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      if (estamp && it < 64) dbg[it * 16 + 12] = clock64();
      if (rowsum_out != nullptr && row_ok) rowsum_out[my_m] = __uint_as_float(rs);
      uint32_t packed[2][16];
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (!c_ok[c]) continue;                                // warp-uniform
        uint32_t (&r)[32] = c == 0 ? r0 : r1;
        epilogue_math(r, mk[c], mask != nullptr, smem_u32(bias_w) + c * 128, act);
        if (estamp && it < 64 && c == 0) dbg[it * 16 + 13] = clock64();
        if (has_bf16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) packed[c][j] = pack_bf16x2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
        }
        if (has_f32) {
          // staging buffer: wait until the store issued two boxes ago has finished reading it
          if (lane == 0) bulk_wait_read<TC_OUT_BUFS - 1>();
          __syncwarp();
          if (estamp && it < 64 && c == 0) dbg[it * 16 + 14] = clock64();
          const uint32_t buf = out_base + (buf_i & (TC_OUT_BUFS - 1)) * TC_OUT_BUF_BYTES;
          ++buf_i;
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16)                    // 16-byte chunk c16 of the row lands at chunk (c16 ^ row%8)
            st_shared_v4(buf + row_off + (((uint32_t)c16 ^ sw) << 4), r[4 * c16], r[4 * c16 + 1], r[4 * c16 + 2], r[4 * c16 + 3]);
          fence_async_smem();
          __syncwarp();
          if (estamp && it < 64 && c == 0) dbg[it * 16 + 15] = clock64();
          if (lane == 0) {
            if (accumulate) tma_reduce_add_2d(&R->tmC32, buf, n0 + c * 32, m_base);
            else tma_store_2d(&R->tmC32, buf, n0 + c * 32, m_base);
            bulk_commit();
          }
        }
      }
      if (has_bf16 && c_ok[0]) {
        if (lane == 0) bulk_wait_read<TC_OUT_BUFS - 1>();
        __syncwarp();
        const uint32_t buf = out_base + (buf_i & (TC_OUT_BUFS - 1)) * TC_OUT_BUF_BYTES;
        ++buf_i;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          if (!c_ok[c]) continue;                              // columns >= N are clipped by the store
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_shared_v4(buf + row_off + (((uint32_t)(4 * c + i) ^ sw) << 4), packed[c][4 * i], packed[c][4 * i + 1],
                         packed[c][4 * i + 2], packed[c][4 * i + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&R->tmC16, buf, n0, m_base);
          bulk_commit();
        }
      }
      if (estamp && it < 64) dbg[it * 16 + 11] = clock64();
    }
    if (lane == 0) bulk_wait_all();            // the staging buffers must outlive the last stores
    if (ew == 0 && lane == 0) TC_CTA_STAMP(3);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TC_TMEM_COLS));
    if (lane == 0) TC_CTA_STAMP(6);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int encode_operand(CUtensorMap* tm, const uint16_t* base, int64_t ld, int64_t inner, int64_t outer,
                          int box_inner, int box_outer) {
  return tc_encode_map(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, ld, inner, outer, box_inner, box_outer);
}

}  // namespace mmlrec

using namespace mmlrec;

extern "C" int64_t mmlrec_tc_record_bytes(void) { return (int64_t)sizeof(TcRecord); }
extern "C" int32_t mmlrec_tc_num_tiles(int32_t M, int32_t N) { return cdiv(M, TC_BM) * cdiv(N, TC_BN); }

extern "C" int mmlrec_tc_encode_problem(const MmlrecGemmTcDesc* d, void* record_host) {
  MMLREC_CHECK_ARG(d && record_host, "null argument");
  MMLREC_CHECK_ARG(!d->c_transposed, "transposed stores are a CTA-pair kernel feature");
  MMLREC_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "bad sizes");
  MMLREC_CHECK_ARG(((uintptr_t)d->A & 15) == 0 && ((uintptr_t)d->B & 15) == 0, "operands must be 16-byte aligned");
  MMLREC_CHECK_ARG((d->lda & 7) == 0 && (d->ldb & 7) == 0, "operand row strides must be multiples of 8 elements");
  MMLREC_CHECK_ARG(d->C_f32 == nullptr || ((d->ldc_f32 & 3) == 0 && ((uintptr_t)d->C_f32 & 15) == 0), "C_f32 alignment");
  MMLREC_CHECK_ARG(d->C_bf16 == nullptr || ((d->ldc_bf16 & 7) == 0 && ((uintptr_t)d->C_bf16 & 15) == 0), "C_bf16 alignment");
  MMLREC_CHECK_ARG(d->mask == nullptr || ((d->ldmask & 7) == 0 && ((uintptr_t)d->mask & 15) == 0), "mask alignment");
  MMLREC_CHECK_ARG(d->C_f32 || d->C_bf16, "no output");
  MMLREC_CHECK_ARG(!(d->accumulate && d->C_bf16), "accumulate applies to an fp32-only output (the sum is formed in memory)");
  TcRecord rec;
  memset(&rec, 0, sizeof(rec));
  int rc;
  if (!d->a_mn_major) rc = encode_operand(&rec.tmA, d->A, d->lda, d->K, d->M, TC_BK, TC_BM);
  else                rc = encode_operand(&rec.tmA, d->A, d->lda, d->M, d->K, 64, TC_BK);
  if (rc) return rc;
  if (!d->b_mn_major) rc = encode_operand(&rec.tmB, d->B, d->ldb, d->K, d->N, TC_BK, TC_BN);
  else                rc = encode_operand(&rec.tmB, d->B, d->ldb, d->N, d->K, 64, TC_BK);
  if (rc) return rc;
  if (d->C_f32) {
    rc = tc_encode_map(&rec.tmC32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->C_f32, d->ldc_f32, d->N, d->M, 32, 32);
    if (rc) return rc;
  }
  if (d->C_bf16) {
    rc = tc_encode_map(&rec.tmC16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->C_bf16, d->ldc_bf16, d->N, d->M, 64, 32);
    if (rc) return rc;
  }
  rec.has_f32 = d->C_f32 != nullptr; rec.has_bf16 = d->C_bf16 != nullptr;
  rec.bias = d->bias; rec.mask = d->mask; rec.ldmask = d->ldmask; rec.rowsum_a = d->colsum;
  rec.M = d->M; rec.N = d->N; rec.K = d->K; rec.act = d->act; rec.accumulate = d->accumulate;
  rec.a_mn = d->a_mn_major; rec.b_mn = d->b_mn_major; rec.tiles_n = cdiv(d->N, TC_BN);
  memcpy(record_host, &rec, sizeof(rec));
  return 0;
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
                                            int32_t total_tiles, const int32_t* tile_order, const int32_t* cta_start,
                                            int32_t n_ctas, int64_t* stamps, void* stream) {
  return launch_tc(records, tile_prefix, n_problems, total_tiles, tile_order, cta_start, n_ctas,
                   reinterpret_cast<long long*>(stamps), stream);
}
