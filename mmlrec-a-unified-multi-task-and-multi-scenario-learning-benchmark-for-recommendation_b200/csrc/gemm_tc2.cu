// K3, bf16 throughput mode, CTA-pair version: the same grouped GEMM as gemm_tc.cu, executed by clusters of two
// CTAs (two SMs of one TPC) with `tcgen05.mma.cta_group::2`.  A pair computes one 256 x BN tile (BN = 256 or 128):
// each CTA stages ITS 128 rows of A and ITS half of the BN rows of B (TMA, 128B swizzle), the leader CTA issues
// one MMA that reads both CTAs' shared memory, and each CTA's TMEM receives the accumulator of its own 128 rows.
// Per 128 x 128 of output a CTA therefore pulls half the operand bytes of the one-CTA kernel (16 KB of A per
// 256 columns instead of per 128, B split between the two) -- these problems are bound by L2 -> SM operand
// traffic (K <= 256 forward, batch-contraction backward), not by the tensor pipe, so that is what counts.
//   warp 0        TMA producer (both CTAs; completion bytes of both land on the LEADER's full barrier)
//   warp 1        TMEM allocator (both CTAs, cta_group::2) + MMA issuer (leader only; commits are multicast to
//                 the barriers of both CTAs)
//   warps 2..9    epilogue of the CTA's own 128 x BN accumulator: tcgen05.ld -> mask / bias / activation ->
//                 128B-swizzled staging box -> bulk tensor store (or .add reduction); the accumulator stage is
//                 released by an mbarrier arrive on the leader (remote for the peer CTA)
// D[M,N] = A * B^T, operands K-major or MN-major, bias gradient by an extra N=16 MMA against an all-ones tile,
// exactly as in gemm_tc.cu; split-K is expressed by the caller as several problems over K slices.
#include <cstdlib>
#include <cstring>

#include "tc_common.cuh"

namespace mmlrec {

constexpr int T2_BM = 128;                       // rows per CTA; a pair covers 256
constexpr int T2_BK = 64;
constexpr int T2_STAGES = 4;
constexpr int T2_ACC_STAGES = 2;
constexpr int T2_ACC_COLS = 256;                 // TMEM columns per accumulator stage
constexpr int T2_TMEM_COLS = 512;
constexpr int T2_EPI_WARPS = 16;                 // four per TMEM lane quarter: each owns 32 rows x BN/4 columns of a tile
constexpr int T2_THREADS = 64 + 32 * T2_EPI_WARPS;
constexpr int T2_MAX_PROBLEMS = 96;
constexpr int T2_A_BYTES = T2_BM * T2_BK * 2;    // 16 KB: this CTA's 128 rows of A
constexpr int T2_B_BYTES = 128 * T2_BK * 2;      // 16 KB: this CTA's half of B (BN/2 <= 128 rows)
constexpr int T2_ONES_BYTES = 16 * 128;
constexpr int T2_OUT_BUF_BYTES = 32 * 128;       // one staging box per epilogue warp

struct alignas(128) Tc2Record {
  CUtensorMap tmA;                              // K-major: box {64 k, 128 rows}; MN-major: box {64 m, 64 k}
  CUtensorMap tmB;                              // K-major: box {64 k, BN/2 rows}; MN-major: box {64 n, 64 k}
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
  int32_t bn;                                   // 256 or 128
  int32_t rowsum_col;                           // TMEM column (inside the stage) that receives the row sums
  uint32_t* bits_out; const uint32_t* mask_bits; // ReLU bit masks (see MmlrecGemmTcDesc)
  int32_t bits_out_chunks, bits_out_chunk0, mask_bits_chunks, mask_bits_chunk0;
  int32_t epi;                                  // epilogue variant (EPI_*), chosen at encode time
  float* c_t; int64_t ldc_t;                    // EPI_F32_TRANSPOSED: C[n * ldc_t + m] = D[m][n]
  float* colsum_b;                              // with it: column sums of B over K (ones-tile MMA into TMEM columns 128..)
};

// the scalar part of a record, cached in shared memory once per CTA (every role reads it at every tile)
struct Tc2Meta {
  const float* bias;
  const uint16_t* mask; int64_t ldmask;
  float* rowsum_a;
  int32_t M, N, K;
  int32_t act, accumulate;
  int32_t a_mn, b_mn;
  int32_t has_f32, has_bf16;
  int32_t bn, rowsum_col;
  int32_t bits_out_chunks, bits_out_chunk0, mask_bits_chunks, mask_bits_chunk0;
  int32_t epi;
  uint32_t* bits_out; const uint32_t* mask_bits;
  float* c_t; int64_t ldc_t;
  float* colsum_b;
};

// Epilogue variants.  The generic one handles every combination of mask / bias / activation / output precision at
// ~10 instructions per accumulator element; the step's hot problems (K <= 256 forward layers, K = 128 dgrad) are bound
// by exactly that instruction stream, so the common combinations get straight-line variants of 1.5-3 instructions per
// element built on packed arithmetic (FADD2, F2FP.RELU.PACK, PRMT sign replication).
enum : int {
  EPI_GENERIC = 0,
  EPI_BF16_FWD_BITS = 1,     // bf16 out = relu(acc + bias), + 1 bit per element ("> 0") for the dgrad through this ReLU
  EPI_BF16_FWD = 2,          // bf16 out = relu(acc + bias)
  EPI_BF16_MASKBITS = 3,     // bf16 out = acc where the producer's ReLU bit is set (dgrad into a hidden layer)
  EPI_F32_FWD = 4,           // fp32 out = relu(acc + bias)
  EPI_F32_PLAIN = 5,         // fp32 out (+)= acc (wgrad incl. bias-gradient row sums, dgrad into fp32 gradient buffers)
  EPI_F32_TRANSPOSED = 6,    // fp32 C[n][m] = acc[m][n], straight from registers (thread = row m: a warp's 32 stores of
                             // one column are 128 contiguous bytes) -- weight gradients computed as X^T dZ
};

constexpr int T2_SMEM_BYTES = 1024 + T2_STAGES * (T2_A_BYTES + T2_B_BYTES) + T2_EPI_WARPS * T2_OUT_BUF_BYTES +
                              T2_ONES_BYTES + 256 + T2_EPI_WARPS * 64 * 4 + (2 * T2_MAX_PROBLEMS + 2) * 4 +
                              T2_MAX_PROBLEMS * (int)sizeof(Tc2Meta);

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of this CTA's box whose completion bytes are credited to the LEADER CTA's mbarrier
// (cute SM100_TMA_2SM_LOAD_2D: the barrier address with the peer bit cleared names CTA 0 of the pair)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"((uint64_t)tmap), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc2_mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  const uint32_t z = 0;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(z) : "memory");
}
// completion of all MMAs issued so far -> one arrive on the barrier at this address in BOTH CTAs of the pair
__device__ __forceinline__ void tc2_commit(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask) : "memory");
}
// arrive on the barrier at local address `bar` of CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 remAddr32;\n"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n"
      "}\n" ::"r"(bar), "r"(cta) : "memory");
}

// ReLU mask of one 32-column chunk (thread = row): zero the accumulator where the bf16 mask value is not > 0.
// __vcmpgts2 compares the two signed halfwords of a word with 0: a bf16 is > 0 iff its bits, read as int16, are > 0.
__device__ __forceinline__ void apply_mask32(uint32_t (&r)[32], const uint4 (&mk)[4]) {
#pragma unroll
  for (int v4 = 0; v4 < 4; ++v4) {
    const uint32_t w[4] = {mk[v4].x, mk[v4].y, mk[v4].z, mk[v4].w};
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const uint32_t m = __vcmpgts2(w[p], 0u);                       // 0xFFFF per halfword that is > 0
      r[v4 * 8 + 2 * p] &= (uint32_t)((int32_t)(m << 16) >> 31);
      r[v4 * 8 + 2 * p + 1] &= (uint32_t)((int32_t)m >> 31);
    }
  }
}
// 64 B of row `my_m` of the mask starting at column nc (columns >= N read as "keep")
__device__ __forceinline__ void load_mask32(uint4 (&mk)[4], const uint16_t* mask, int64_t ldmask, int my_m, bool row_ok, int nc, int N) {
#pragma unroll
  for (int v4 = 0; v4 < 4; ++v4) {
    mk[v4] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    if (row_ok && nc + v4 * 8 + 8 <= N)
      mk[v4] = __ldg(reinterpret_cast<const uint4*>(mask + (int64_t)my_m * ldmask + nc) + v4);
    else if (row_ok && nc + v4 * 8 < N) {   // ragged tail: element-wise
      uint32_t w[4] = {0, 0, 0, 0};
      for (int e = 0; e < 8 && nc + v4 * 8 + e < N; ++e)
        w[e >> 1] |= (uint32_t)__ldg(mask + (int64_t)my_m * ldmask + nc + v4 * 8 + e) << ((e & 1) * 16);
      mk[v4] = make_uint4(w[0], w[1], w[2], w[3]);
    }
  }
}

// ---- packed arithmetic of the straight-line epilogue variants ------------------------------------------------------
// (a0, a1) += (b0, b1): one FADD2
__device__ __forceinline__ void add2(uint32_t& a0, uint32_t& a1, float b0, float b1) {
  uint64_t x, y;
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "r"(a0), "r"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(b0), "f"(b1));
  asm("add.rn.f32x2 %0, %0, %1;" : "+l"(x) : "l"(y));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(a0), "=r"(a1) : "l"(x));
}
// {bf16(max(hi, 0)), bf16(max(lo, 0))}: one F2FP.RELU
__device__ __forceinline__ uint32_t pack_relu_bf16x2(uint32_t lo, uint32_t hi) {
  uint32_t r;
  asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(__uint_as_float(hi)), "f"(__uint_as_float(lo)));
  return r;
}
// One 32-column chunk of a ReLU forward layer (thread = row): p[i] = bf16x2 of relu(acc + bias) for columns (2i, 2i+1);
// returns the chunk's ReLU bit word: bit i <-> column 2i, bit 16 + i <-> column 2i + 1 (see MmlrecGemmTcDesc).
// A post-ReLU bf16 half h is in [0, 0x7FFF]: h + 0x7FFF carries into bit 15 exactly when h != 0 and never further.
template <bool BITS>
__device__ __forceinline__ uint32_t epi_bias_relu_pack(const uint32_t (&r)[32], uint32_t bias_addr, uint32_t (&p)[16]) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const float4 b = ld_shared_f4(bias_addr + 16 * g);
    uint32_t a0 = r[4 * g], a1 = r[4 * g + 1], a2 = r[4 * g + 2], a3 = r[4 * g + 3];
    add2(a0, a1, b.x, b.y);
    add2(a2, a3, b.z, b.w);
    p[2 * g] = pack_relu_bf16x2(a0, a1);
    p[2 * g + 1] = pack_relu_bf16x2(a2, a3);
  }
  uint32_t w = 0;
  if (BITS) {
#pragma unroll
    for (int i = 0; i < 16; ++i) w = (w >> 1) | ((p[i] + 0x7FFF7FFFu) & 0x80008000u);
  }
  return w;
}
// dgrad through a ReLU: p[i] = bf16x2(acc) of columns (2i, 2i+1), zeroed where the column's bit of w is clear.  The two
// bits of a pair are moved to the sign positions of bytes 0 and 2; PRMT in sign-replication mode widens them to halfwords.
__device__ __forceinline__ void epi_maskbits_pack(const uint32_t (&r)[32], uint32_t w, uint32_t (&p)[16]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t s = i <= 7 ? (w << (7 - i)) : (w >> (i - 7));
    uint32_t m;   // (__byte_perm drops the sign-replication bit of the selector digits: PTX directly)
    asm("prmt.b32 %0, %1, 0, 0xAA88;" : "=r"(m) : "r"(s));
    p[i] = pack_bf16x2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1])) & m;
  }
}
// position of column j's bit inside a chunk's ReLU bit word
__device__ __forceinline__ constexpr int relu_bit_pos(int j) { return (j >> 1) + ((j & 1) << 4); }
// chunk c (0 / 1) of the warp's bf16 staging box: four conflict-free 16-byte stores per row
__device__ __forceinline__ void stage_bf16_chunk(uint32_t buf, uint32_t row_off, uint32_t sw, int c, const uint32_t (&p)[16]) {
#pragma unroll
  for (int i = 0; i < 4; ++i)
    st_shared_v4(buf + row_off + (((uint32_t)(4 * c + i) ^ sw) << 4), p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]);
}
__device__ __forceinline__ void stage_f32_chunk(uint32_t buf, uint32_t row_off, uint32_t sw, const uint32_t (&r)[32]) {
#pragma unroll
  for (int c16 = 0; c16 < 8; ++c16)
    st_shared_v4(buf + row_off + (((uint32_t)c16 ^ sw) << 4), r[4 * c16], r[4 * c16 + 1], r[4 * c16 + 2], r[4 * c16 + 3]);
}

// per-tile values of one epilogue warp that the straight-line variants need
struct EpiTile {
  const Tc2Record* R;
  uint32_t buf, row_off, sw, bias_addr;      // staging box, this row's offset / swizzle phase in it, the warp's bias slot
  uint32_t mb0, mb1;                         // ReLU bit words of this row (mask variants)
  uint32_t* bits_word;                       // where this row's bit word of chunk 0 goes (chunk 1: + 32 words)
  int m_base, n_first, lane, accumulate;
  bool row_ok, c1_ok;
  float* c_t; int64_t ldc_t; int N;          // EPI_F32_TRANSPOSED
};

// The warp's 32 x 64 block, chunk by chunk: tcgen05.ld -> (stage handed back after the last load) -> packed math ->
// staging box.  KIND is warp-uniform.  bf16 variants fill the warp's 64-column box (the caller issues its bulk store);
// fp32 variants send one 32-column box per chunk themselves.
template <int KIND>
__device__ __forceinline__ void epi_tile_lean(const EpiTile& e, uint32_t taddr, uint32_t release_bar) {
  constexpr bool BF16 = KIND == EPI_BF16_FWD_BITS || KIND == EPI_BF16_FWD || KIND == EPI_BF16_MASKBITS;
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    if (c == 1 && !e.c1_ok) break;
    uint32_t r[32];
    tc_ld32(taddr + 32 * c, r);
    tc_wait_ld();
    if (c == 1 || !e.c1_ok) {                  // this warp's last values sit in registers: the stage is free
      tc_fence_before();
      __syncwarp();
      if (e.lane == 0) mbar_arrive_cluster(release_bar, 0);
    }
    if (BF16) {
      uint32_t p[16];
      if (KIND == EPI_BF16_MASKBITS) {
        epi_maskbits_pack(r, c == 0 ? e.mb0 : e.mb1, p);
      } else {
        const uint32_t w = epi_bias_relu_pack<KIND == EPI_BF16_FWD_BITS>(r, e.bias_addr + 128u * c, p);
        if (KIND == EPI_BF16_FWD_BITS && e.row_ok) e.bits_word[32 * c] = w;
      }
      if (c == 0) {
        if (e.lane == 0) bulk_wait_read<0>();  // the previous tile's store has read the staging box
        __syncwarp();
      }
      stage_bf16_chunk(e.buf, e.row_off, e.sw, c, p);
    } else if (KIND == EPI_F32_TRANSPOSED) {
      if (e.row_ok) {
        float* col = e.c_t + (int64_t)(e.n_first + 32 * c) * e.ldc_t + (e.m_base + e.lane);
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (e.n_first + 32 * c + j < e.N) col[(int64_t)j * e.ldc_t] = __uint_as_float(r[j]);
      }
    } else {
      if (KIND == EPI_F32_FWD) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          const float4 b = ld_shared_f4(e.bias_addr + 128u * c + 16 * g);
          add2(r[4 * g], r[4 * g + 1], b.x, b.y);
          add2(r[4 * g + 2], r[4 * g + 3], b.z, b.w);
#pragma unroll
          for (int k = 0; k < 4; ++k) r[4 * g + k] = __float_as_uint(fmaxf(__uint_as_float(r[4 * g + k]), 0.f));
        }
      }
      if (e.lane == 0) bulk_wait_read<0>();
      __syncwarp();
      stage_f32_chunk(e.buf, e.row_off, e.sw, r);
      fence_async_smem();
      __syncwarp();
      if (e.lane == 0) {
        if (e.accumulate) tma_reduce_add_2d(&e.R->tmC32, e.buf, e.n_first + 32 * c, e.m_base);
        else tma_store_2d(&e.R->tmC32, e.buf, e.n_first + 32 * c, e.m_base);
        bulk_commit();
      }
    }
  }
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(T2_THREADS, 1)
gemm_grouped_tc2_kernel(const Tc2Record* __restrict__ recs, const int32_t* __restrict__ prefix, int n_problems, int total_tiles,
                        const int32_t* __restrict__ tile_order, const int32_t* __restrict__ pair_start,
                        long long* __restrict__ dbg) {
  // programmatic dependent launch: the next grid of the stream may be scheduled from now on; everything up to pdl_wait()
  // below (barriers, TMEM, the problem tables -- written once at plan build, never by a predecessor) overlaps the tail
  // of the previous kernel
  pdl_launch_dependents();
  extern __shared__ unsigned char smem_dyn[];
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  unsigned char* sA = smem;
  unsigned char* sB = smem + T2_STAGES * T2_A_BYTES;
  unsigned char* sOut = smem + T2_STAGES * (T2_A_BYTES + T2_B_BYTES);
  unsigned char* sOnes = sOut + T2_EPI_WARPS * T2_OUT_BUF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + T2_ONES_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * T2_STAGES + 2 * T2_ACC_STAGES);
  float* bias_s = reinterpret_cast<float*>(sOnes + T2_ONES_BYTES + 256);             // [EPI_WARPS][64]
  int32_t* s_prefix = reinterpret_cast<int32_t*>(bias_s + T2_EPI_WARPS * 64);
  int32_t* s_tiles_n = s_prefix + T2_MAX_PROBLEMS + 1;
  Tc2Meta* s_meta = reinterpret_cast<Tc2Meta*>(s_tiles_n + T2_MAX_PROBLEMS + 1);     // 8-byte aligned (offsets above are)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();               // 0 = leader (issues the MMAs)
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + T2_STAGES);
  const uint32_t tfull_bar = smem_u32(bars + 2 * T2_STAGES), tempty_bar = smem_u32(bars + 2 * T2_STAGES + T2_ACC_STAGES);

  for (int i = threadIdx.x; i < T2_ONES_BYTES / 4; i += T2_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  for (int i = threadIdx.x; i <= n_problems; i += T2_THREADS) s_prefix[i] = prefix[i];
  for (int i = threadIdx.x; i < n_problems; i += T2_THREADS) {
    const Tc2Record* R = recs + i;
    s_tiles_n[i] = R->tiles_n;
    Tc2Meta m;
    m.bias = R->bias; m.mask = R->mask; m.ldmask = R->ldmask; m.rowsum_a = R->rowsum_a;
    m.M = R->M; m.N = R->N; m.K = R->K; m.act = R->act; m.accumulate = R->accumulate;
    m.a_mn = R->a_mn; m.b_mn = R->b_mn; m.has_f32 = R->has_f32; m.has_bf16 = R->has_bf16;
    m.bn = R->bn; m.rowsum_col = R->rowsum_col; m.epi = R->epi; m.c_t = R->c_t; m.ldc_t = R->ldc_t; m.colsum_b = R->colsum_b;
    m.bits_out = R->bits_out; m.mask_bits = R->mask_bits;
    m.bits_out_chunks = R->bits_out_chunks; m.bits_out_chunk0 = R->bits_out_chunk0;
    m.mask_bits_chunks = R->mask_bits_chunks; m.mask_bits_chunk0 = R->mask_bits_chunk0;
    s_meta[i] = m;
  }
  if (threadIdx.x == 0) {
    // full: one arrive (the leader's expect_tx) + the bytes of both CTAs; empty / tmem_full: one multicast commit;
    // tmem_empty (used on the leader only): every epilogue warp of both CTAs
    for (int s = 0; s < T2_STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < T2_ACC_STAGES; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 2 * T2_EPI_WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_async_smem();
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(T2_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();                                     // the peer's barriers / ones tile / TMEM exist before anyone uses them
  tc_fence_after();
  pdl_wait();                                             // the previous kernel's outputs (our operands, masks, biases) are final
  const uint32_t tmem_base = *tmem_slot;
  const int sched_begin = tile_order ? pair_start[pair] : pair;
  const int sched_end = tile_order ? pair_start[pair + 1] : total_tiles;
  const int sched_step = tile_order ? 1 : n_pairs;
  const bool stamp = dbg != nullptr && blockIdx.x == 0;   // per-tile clock64 stamps of CTA 0: dbg[tile_iter * 16 + slot]
#define T2_STAMP(iter, slot) do { if (stamp && (iter) < 64) dbg[(iter) * 16 + (slot)] = clock64(); } while (0)
#define T2_CTA_STAMP(slot) do { if (dbg != nullptr) { unsigned long long gt_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_)); \
                                  dbg[1024 + blockIdx.x * 8 + (slot)] = (long long)gt_; } } while (0)
  if (threadIdx.x == 0) { T2_CTA_STAMP(0); if (dbg != nullptr) dbg[1024 + blockIdx.x * 8 + 4] = (sched_end - sched_begin + sched_step - 1) / sched_step; }

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int pit = 0;
      for (int ti = sched_begin; ti < sched_end; ti += sched_step, ++pit) {
        const int t = tile_order ? __ldg(tile_order + ti) : ti;
        T2_STAMP(pit, 0);
        const TileCoord tc = locate_tile(t, s_prefix, s_tiles_n, n_problems);
        const Tc2Record* R = recs + tc.pi;
        const Tc2Meta& mt = s_meta[tc.pi];
        const int K = mt.K, a_mn = mt.a_mn, b_mn = mt.b_mn, bn = mt.bn;
        if (stamp && pit < 64) dbg[pit * 16 + 1] = ((long long)tc.pi << 32) | (long long)((K + 63) / 64);
        if (pit == 0) { tma_prefetch_desc(&R->tmA); tma_prefetch_desc(&R->tmB); }
        const int half_n = bn >> 1;
        const int m0 = tc.tm * 256 + (int)rank * T2_BM;             // this CTA's rows of A
        const int n0 = tc.tn * bn + (int)rank * half_n;             // this CTA's rows of B
        const uint32_t stage_bytes = 2u * (uint32_t)(T2_A_BYTES + half_n * 128);   // both CTAs' boxes
        const int num_kb = (K + T2_BK - 1) / T2_BK;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t fb = full_bar + 8 * stage;
          if (rank == 0) mbar_expect_tx(fb, stage_bytes);
          const uint32_t a_dst = smem_u32(sA + stage * T2_A_BYTES), b_dst = smem_u32(sB + stage * T2_B_BYTES);
          const int k0 = kb * T2_BK;
          if (!a_mn) {
            tma_load_2d_pair(a_dst, &R->tmA, fb, k0, m0);
          } else {
            tma_load_2d_pair(a_dst, &R->tmA, fb, m0, k0);
            tma_load_2d_pair(a_dst + 8192, &R->tmA, fb, m0 + 64, k0);
          }
          if (!b_mn) {
            tma_load_2d_pair(b_dst, &R->tmB, fb, k0, n0);           // box {64 k, BN/2 rows}
          } else {
            tma_load_2d_pair(b_dst, &R->tmB, fb, n0, k0);
            if (half_n == 128) tma_load_2d_pair(b_dst + 8192, &R->tmB, fb, n0 + 64, k0);
          }
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
        }
        T2_STAMP(pit, 3);
        const int tn_i = ti + sched_step;
        if (tn_i < sched_end) {
          const int t2 = tile_order ? __ldg(tile_order + tn_i) : tn_i;
          const TileCoord tc2 = locate_tile(t2, s_prefix, s_tiles_n, n_problems);
          if (tc2.pi != tc.pi) { tma_prefetch_desc(&recs[tc2.pi].tmA); tma_prefetch_desc(&recs[tc2.pi].tmB); }
        }
      }
      T2_CTA_STAMP(1);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    if (lane == 0 && rank == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint64_t ones_desc = make_smem_desc(smem_u32(sOnes), false);
      // the same all-ones bytes as an A operand of 128 rows per CTA: a zero stride between the 8-row groups makes all of
      // them alias the first 1 KB (every element is 1.0, so the swizzle pattern is irrelevant)
      const uint64_t ones_a_desc = ones_desc & ~((uint64_t)0x3FFF << 32);
      int mit = -1;
      for (int ti = sched_begin; ti < sched_end; ti += sched_step) {
        ++mit;
        const int t = tile_order ? __ldg(tile_order + ti) : ti;
        const TileCoord tc = locate_tile(t, s_prefix, s_tiles_n, n_problems);
        const Tc2Meta& mt = s_meta[tc.pi];
        const int K = mt.K, bn = mt.bn;
        const bool a_mn = mt.a_mn != 0, b_mn = mt.b_mn != 0;
        const bool rowsum = (mt.rowsum_a != nullptr) && tc.tn == 0;
        const int rowsum_col = mt.rowsum_col;
        const bool rowsum_shared = rowsum_col < bn;      // the row sums live in unused columns of the main accumulator
        const uint32_t idesc = make_idesc(256, bn, a_mn, b_mn);
        const uint32_t idesc_ones = make_idesc(256, 16, a_mn, false);
        // column sums of B (bias gradient of a transposed-product wgrad): D2[256 x bn] += ones[256 x 16] B^T into the
        // spare columns 128.. of the stage; every row of D2 is the column-sum vector
        const bool colsum_b = mt.colsum_b != nullptr && tc.tm == 0;
        const uint32_t idesc_csb = make_idesc(256, bn, false, b_mn);
        const int num_kb = (K + T2_BK - 1) / T2_BK;
        T2_STAMP(mit, 4);
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        T2_STAMP(mit, 5);
        const uint32_t d_tmem = tmem_base + acc * T2_ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
          if (kb == 0) T2_STAMP(mit, 6);
          const uint32_t a_addr = smem_u32(sA + stage * T2_A_BYTES), b_addr = smem_u32(sB + stage * T2_B_BYTES);
          const uint64_t a_desc = make_smem_desc(a_addr, a_mn), b_desc = make_smem_desc(b_addr, b_mn);
          const uint64_t a_step = a_mn ? (2048 >> 4) : (32 >> 4), b_step = b_mn ? (2048 >> 4) : (32 >> 4);
          const int k_left = K - kb * T2_BK;
#pragma unroll
          for (int k = 0; k < T2_BK / 16; ++k) {
            if (k * 16 < k_left) {
              const uint32_t accumulate = (kb | k) != 0 ? 1u : 0u;
              tc2_mma(d_tmem, a_desc + a_step * k, b_desc + b_step * k, idesc, accumulate);
              // (shared columns were just zeroed / kept by the main MMA: B rows >= N are TMA zero fill)
              if (rowsum) tc2_mma(d_tmem + rowsum_col, a_desc + a_step * k, ones_desc + 2 * k, idesc_ones,
                                  rowsum_shared ? 1u : accumulate);
              if (colsum_b) tc2_mma(d_tmem + 128, ones_a_desc + 2 * k, b_desc + b_step * k, idesc_csb, accumulate);
            }
          }
          tc2_commit(empty_bar + 8 * stage);             // frees the stage in both CTAs
          if (++stage == T2_STAGES) { stage = 0; phase ^= 1; }
        }
        tc2_commit(tfull_bar + 8 * acc);                 // accumulator ready, both CTAs
        T2_STAMP(mit, 7);
        if (++acc == T2_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
      T2_CTA_STAMP(2);
    }
  } else {
    // ===================== epilogue (warps 2..17, both CTAs) =====================
    // warp -> TMEM lane quarter q (rows 32q.. of this CTA's 128) and column quarter (BN/4 columns = one or two
    // 32-column chunks).  Per chunk: tcgen05.ld (thread = row) -> ReLU mask / bias / activation in registers ->
    // 128B-swizzled staging box (conflict-free 16-byte stores) -> one bulk tensor store (or .add reduction) by lane 0.
    // The mask of the next chunk and the bias / first mask chunk of the NEXT tile are fetched while the current
    // chunk is being processed, so no global latency sits between the accumulator becoming ready and the store.
    const int q = warp & 3;
    const int ew = warp - 2;
    const int cq = ew >> 2;                                     // column quarter
    float* bias_w = bias_s + ew * 64;
    const uint32_t bias_addr = smem_u32(bias_w);
    const uint32_t buf = smem_u32(sOut + ew * T2_OUT_BUF_BYTES);
    const uint32_t row_off = (uint32_t)lane * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    int acc = 0; uint32_t acc_phase = 0;
    int it = -1;
    const bool estamp = stamp && ew == 0 && lane == 0;
    // per-tile state that is prefetched one tile ahead
    TileCoord tc{0, 0, 0};
    float bv0 = 0.f, bv1 = 0.f;
    uint32_t mbits0 = 0xFFFFFFFFu, mbits1 = 0xFFFFFFFFu;       // ReLU bit words of this row for the warp's two chunks
    auto prefetch_tile = [&](int ti_) {
      const int t_ = tile_order ? __ldg(tile_order + ti_) : ti_;
      tc = locate_tile(t_, s_prefix, s_tiles_n, n_problems);
      const Tc2Meta& m_ = s_meta[tc.pi];
      // fp32 output of a 128-wide tile: every epilogue warp takes ONE 32-column chunk (its own store box), so the 16
      // warps work side by side; otherwise a warp owns 64 columns (one bf16 store box) and, with BN = 128, the warps
      // of column quarters 2 and 3 have none
      const bool narrow_ = m_.bn == 128 && m_.has_f32 != 0;
      const int nf = tc.tn * m_.bn + (narrow_ ? cq * 32 : cq * 64);
      const bool active_ = narrow_ || cq * 64 < m_.bn;
      bv0 = (active_ && m_.bias != nullptr && nf + lane < m_.N) ? __ldg(m_.bias + nf + lane) : 0.f;
      bv1 = (active_ && !narrow_ && m_.bias != nullptr && nf + 32 + lane < m_.N) ? __ldg(m_.bias + nf + 32 + lane) : 0.f;
      const int mm = tc.tm * 256 + (int)rank * T2_BM + q * 32 + lane;
      if (active_ && m_.mask_bits != nullptr) {
        // one coalesced 128-byte read per chunk: the words of the warp's 32 rows are adjacent
        const int64_t w0 = ((int64_t)(mm >> 5) * m_.mask_bits_chunks + m_.mask_bits_chunk0 + (nf >> 5)) * 32 + (mm & 31);
        mbits0 = (mm < m_.M && nf < m_.N) ? __ldg(m_.mask_bits + w0) : 0u;
        mbits1 = (!narrow_ && mm < m_.M && nf + 32 < m_.N) ? __ldg(m_.mask_bits + w0 + 32) : 0u;
      }
    };
    if (sched_begin < sched_end) prefetch_tile(sched_begin);
    for (int ti = sched_begin; ti < sched_end; ti += sched_step) {
      ++it;
      if (estamp && it < 64) dbg[it * 16 + 8] = clock64();
      const Tc2Record* R = recs + tc.pi;
      const Tc2Meta& mt = s_meta[tc.pi];
      const int M = mt.M, N = mt.N, bn = mt.bn;
      const bool narrow = bn == 128 && mt.has_f32 != 0;         // see prefetch_tile: one chunk per warp
      const int nchunks = narrow ? 1 : (cq * 64 < bn) ? 2 : 0;
      const int m_base = tc.tm * 256 + (int)rank * T2_BM + q * 32;
      const int col0 = narrow ? cq * 32 : cq * 64;              // first accumulator column of this warp
      const int n_first = tc.tn * bn + col0;
      const bool has_f32 = mt.has_f32 != 0, has_bf16 = mt.has_bf16 != 0;
      const uint32_t* const mask_bits = mt.mask_bits;
      const uint16_t* const mask = mask_bits != nullptr ? nullptr : mt.mask; const int64_t ldmask = mt.ldmask;
      uint32_t* const bits_out = mt.bits_out;
      const uint32_t mb[2] = {mbits0, mbits1};
      const int act = mt.act, accumulate = mt.accumulate;
      float* const rowsum_out = (tc.tn == 0 && cq == 0) ? mt.rowsum_a : nullptr;
      const int rowsum_col = mt.rowsum_col;
      const int my_m = m_base + lane;
      const bool row_ok = my_m < M;
      const bool rows_any = m_base < M && nchunks > 0;
      __syncwarp();                                            // the previous tile's reads of the bias slot are done
      bias_w[lane] = bv0;
      bias_w[32 + lane] = bv1;
      __syncwarp();
      if (estamp && it < 64) dbg[it * 16 + 9] = clock64();
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      if (estamp && it < 64) dbg[it * 16 + 10] = clock64();
      const uint32_t t_row = tmem_base + acc * T2_ACC_COLS + ((uint32_t)(q * 32) << 16);
      uint32_t rs = 0;
      if (rowsum_out != nullptr) tc_ld1(t_row + rowsum_col, rs);
      // column sums of B: every row of the second accumulator holds them; the leader's row-0 warps read one chunk each
      float* const csb = (mt.colsum_b != nullptr && tc.tm == 0 && rank == 0 && q == 0 && narrow) ? mt.colsum_b : nullptr;
      uint32_t rc[32];
      if (csb != nullptr) tc_ld32(t_row + 128 + cq * 32, rc);
      const bool first_ok = rows_any && n_first < N;
      const int kind = mt.epi;
      if (kind != EPI_GENERIC && nchunks > 0) {
        // straight-line variants (see EPI_*)
        const bool c1_ok = nchunks == 2 && rows_any && n_first + 32 < N;
        if (rowsum_out != nullptr) {
          tc_wait_ld();
          if (row_ok) rowsum_out[my_m] = __uint_as_float(rs);
        }
        if (csb != nullptr) {
          tc_wait_ld();
          if (lane == 0) {
            const int n0 = tc.tn * bn + cq * 32;
#pragma unroll
            for (int j = 0; j < 32; ++j) if (n0 + j < N) csb[n0 + j] = __uint_as_float(rc[j]);
          }
        }
        if (!first_ok) {                                         // nothing of this warp's block is inside the problem
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_bar + 8 * acc, 0);
        } else {
          EpiTile e;
          e.R = R; e.buf = buf; e.row_off = row_off; e.sw = sw; e.bias_addr = bias_addr;
          e.mb0 = mbits0; e.mb1 = mbits1;
          e.bits_word = bits_out == nullptr ? nullptr
              : bits_out + ((int64_t)(my_m >> 5) * mt.bits_out_chunks + mt.bits_out_chunk0 + (n_first >> 5)) * 32 + (my_m & 31);
          e.m_base = m_base; e.n_first = n_first; e.lane = lane; e.accumulate = accumulate;
          e.row_ok = row_ok; e.c1_ok = c1_ok;
          e.c_t = mt.c_t; e.ldc_t = mt.ldc_t; e.N = N;
          const uint32_t taddr = t_row + col0, rel = tempty_bar + 8 * acc;
          switch (kind) {
            case EPI_BF16_FWD_BITS: epi_tile_lean<EPI_BF16_FWD_BITS>(e, taddr, rel); break;
            case EPI_BF16_FWD:      epi_tile_lean<EPI_BF16_FWD>(e, taddr, rel); break;
            case EPI_BF16_MASKBITS: epi_tile_lean<EPI_BF16_MASKBITS>(e, taddr, rel); break;
            case EPI_F32_FWD:       epi_tile_lean<EPI_F32_FWD>(e, taddr, rel); break;
            case EPI_F32_TRANSPOSED: epi_tile_lean<EPI_F32_TRANSPOSED>(e, taddr, rel); break;
            default:                epi_tile_lean<EPI_F32_PLAIN>(e, taddr, rel); break;
          }
        }
        if (estamp && it < 64) dbg[it * 16 + 12] = clock64();
      } else
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        if (c >= nchunks) break;
        const int nc = n_first + c * 32;
        const bool c_ok = rows_any && nc < N;                  // warp-uniform
        uint32_t r[32];
        if (c_ok) tc_ld32(t_row + col0 + c * 32, r);
        tc_wait_ld();
        if (c == nchunks - 1) {
          // the accumulator stage is free once this warp's last values sit in registers: tell the leader
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(tempty_bar + 8 * acc, 0);
          if (estamp && it < 64) dbg[it * 16 + 12] = clock64();
          if (rowsum_out != nullptr && row_ok) rowsum_out[my_m] = __uint_as_float(rs);
        }
        if (!c_ok) continue;
        if (mask_bits != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = (mb[c] >> relu_bit_pos(j)) & 1u ? r[j] : 0u;
        } else if (mask != nullptr) {   // bf16 mask (no bit array for this activation): loaded at use
          uint4 mk[4];
          load_mask32(mk, mask, ldmask, my_m, row_ok, nc, N);
          apply_mask32(r, mk);
        }
        {
          uint4 unused[4];
          epilogue_math(r, unused, false, bias_addr + (uint32_t)c * 128u, act);
        }
        if (bits_out != nullptr && row_ok) {
          // "output > 0" of this row's 32 columns as one word (a float is > 0 iff its bits, read as int32, are > 0)
          uint32_t w = 0;
#pragma unroll
          for (int j = 0; j < 32; ++j) w |= ((int32_t)r[j] > 0 ? 1u : 0u) << relu_bit_pos(j);
          bits_out[((int64_t)(my_m >> 5) * mt.bits_out_chunks + mt.bits_out_chunk0 + (nc >> 5)) * 32 + (my_m & 31)] = w;
        }
        if (has_f32) {
          if (lane == 0) bulk_wait_read<0>();                  // the staging box of the previous store has been read
          __syncwarp();
#pragma unroll
          for (int c16 = 0; c16 < 8; ++c16)
            st_shared_v4(buf + row_off + (((uint32_t)c16 ^ sw) << 4), r[4 * c16], r[4 * c16 + 1], r[4 * c16 + 2], r[4 * c16 + 3]);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            if (accumulate) tma_reduce_add_2d(&R->tmC32, buf, nc, m_base);
            else tma_store_2d(&R->tmC32, buf, nc, m_base);
            bulk_commit();
          }
        } else {   // bf16 only: the chunk goes straight into its half of the 64-column box
          if (c == 0) {
            if (lane == 0) bulk_wait_read<0>();
            __syncwarp();
          }
#pragma unroll
          for (int i = 0; i < 4; ++i)
            st_shared_v4(buf + row_off + (((uint32_t)(4 * c + i) ^ sw) << 4),
                         pack_bf16x2(__uint_as_float(r[8 * i]), __uint_as_float(r[8 * i + 1])),
                         pack_bf16x2(__uint_as_float(r[8 * i + 2]), __uint_as_float(r[8 * i + 3])),
                         pack_bf16x2(__uint_as_float(r[8 * i + 4]), __uint_as_float(r[8 * i + 5])),
                         pack_bf16x2(__uint_as_float(r[8 * i + 6]), __uint_as_float(r[8 * i + 7])));
        }
      }
      if (nchunks == 0) {                                      // a warp without columns only hands the stage back
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_bar + 8 * acc, 0);
      }
      // next tile's bias / first mask chunk: in flight while the bf16 box leaves and while waiting for its accumulator
      const int m_base_cur = m_base, n_first_cur = n_first;
      if (ti + sched_step < sched_end) prefetch_tile(ti + sched_step);
      if (has_bf16 && first_ok) {
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&R->tmC16, buf, n_first_cur, m_base_cur);
          bulk_commit();
        }
      }
      if (estamp && it < 64) dbg[it * 16 + 11] = clock64();
      if (++acc == T2_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_all();
    if (ew == 0 && lane == 0) T2_CTA_STAMP(3);
  }
  tc_fence_before();
  __syncthreads();
  __syncwarp();
  cluster_sync_all();                                     // nobody leaves while the peer may still read its smem / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(T2_TMEM_COLS));
    if (lane == 0) T2_CTA_STAMP(6);
  }
}

}  // namespace mmlrec

using namespace mmlrec;

extern "C" int64_t mmlrec_tc2_record_bytes(void) { return (int64_t)sizeof(Tc2Record); }

// tile width of a problem: 256 columns where that halves the passes over A, unless the bias-gradient row sums then
// have no spare accumulator columns (a stage is 256 TMEM columns wide)
static int tc2_pick_bn(const MmlrecGemmTcDesc* d) {
  if (d->N <= 128) return 128;
  if (d->colsum != nullptr && d->N > 240) return 128;
  return 256;
}
extern "C" int32_t mmlrec_tc2_num_tiles(const MmlrecGemmTcDesc* d) {
  return cdiv(d->M, 256) * cdiv(d->N, tc2_pick_bn(d));
}

extern "C" int mmlrec_tc2_encode_problem(const MmlrecGemmTcDesc* d, void* record_host) {
  MMLREC_CHECK_ARG(d && record_host, "null argument");
  MMLREC_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "bad sizes");
  MMLREC_CHECK_ARG(((uintptr_t)d->A & 15) == 0 && ((uintptr_t)d->B & 15) == 0, "operands must be 16-byte aligned");
  MMLREC_CHECK_ARG((d->lda & 7) == 0 && (d->ldb & 7) == 0, "operand row strides must be multiples of 8 elements");
  MMLREC_CHECK_ARG(d->C_f32 == nullptr || d->c_transposed || ((d->ldc_f32 & 3) == 0 && ((uintptr_t)d->C_f32 & 15) == 0), "C_f32 alignment");
  MMLREC_CHECK_ARG(d->C_bf16 == nullptr || ((d->ldc_bf16 & 7) == 0 && ((uintptr_t)d->C_bf16 & 15) == 0), "C_bf16 alignment");
  MMLREC_CHECK_ARG(d->mask == nullptr || ((d->ldmask & 7) == 0 && ((uintptr_t)d->mask & 15) == 0), "mask alignment");
  MMLREC_CHECK_ARG(!d->c_transposed || (d->C_f32 != nullptr && d->C_bf16 == nullptr && d->bias == nullptr && d->mask == nullptr &&
                                        d->mask_bits == nullptr && d->relu_bits_out == nullptr && d->colsum == nullptr &&
                                        (d->colsum_b == nullptr || d->N <= 128) &&
                                        d->act == MMLREC_ACT_NONE && !d->accumulate),
                   "a transposed store takes the plain fp32 product only");
  MMLREC_CHECK_ARG(d->colsum_b == nullptr || d->c_transposed, "colsum_b comes with c_transposed");
  MMLREC_CHECK_ARG((d->C_f32 != nullptr) != (d->C_bf16 != nullptr),
                   "the CTA-pair kernel writes one output precision per problem (list the problem twice for both)");
  Tc2Record rec;
  memset(&rec, 0, sizeof(rec));
  const int bn = tc2_pick_bn(d);
  const CUtensorMapDataType BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc;
  if (!d->a_mn_major) rc = tc_encode_map(&rec.tmA, BF, 2, d->A, d->lda, d->K, d->M, T2_BK, T2_BM);
  else                rc = tc_encode_map(&rec.tmA, BF, 2, d->A, d->lda, d->M, d->K, 64, T2_BK);
  if (rc) return rc;
  if (!d->b_mn_major) rc = tc_encode_map(&rec.tmB, BF, 2, d->B, d->ldb, d->K, d->N, T2_BK, bn / 2);
  else                rc = tc_encode_map(&rec.tmB, BF, 2, d->B, d->ldb, d->N, d->K, 64, T2_BK);
  if (rc) return rc;
  if (d->C_f32 && !d->c_transposed) {
    rc = tc_encode_map(&rec.tmC32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d->C_f32, d->ldc_f32, d->N, d->M, 32, 32);
    if (rc) return rc;
  }
  if (d->C_bf16) {
    rc = tc_encode_map(&rec.tmC16, BF, 2, d->C_bf16, d->ldc_bf16, d->N, d->M, 64, 32);
    if (rc) return rc;
  }
  rec.has_f32 = d->C_f32 != nullptr; rec.has_bf16 = d->C_bf16 != nullptr;
  rec.bias = d->bias; rec.mask = d->mask; rec.ldmask = d->ldmask; rec.rowsum_a = d->colsum;
  rec.M = d->M; rec.N = d->N; rec.K = d->K; rec.act = d->act; rec.accumulate = d->accumulate;
  rec.a_mn = d->a_mn_major; rec.b_mn = d->b_mn_major; rec.bn = bn; rec.tiles_n = cdiv(d->N, bn);
  rec.rowsum_col = bn == 128 ? 128 : 240;
  {
    const bool any_mask = d->mask != nullptr || d->mask_bits != nullptr;
    const char* force = getenv("MMLREC_TC2_EPILOGUE");                 // "generic": A/B switch for the profiles
    int epi = EPI_GENERIC;
    if (force != nullptr && strcmp(force, "generic") == 0) {
    } else if (d->C_bf16 != nullptr) {
      if (!any_mask && d->act == MMLREC_ACT_RELU && d->colsum == nullptr)
        epi = d->relu_bits_out != nullptr ? EPI_BF16_FWD_BITS : EPI_BF16_FWD;
      else if (d->mask_bits != nullptr && d->bias == nullptr && d->act == MMLREC_ACT_NONE && d->relu_bits_out == nullptr &&
               d->colsum == nullptr)
        epi = EPI_BF16_MASKBITS;
    } else if (!any_mask && d->relu_bits_out == nullptr) {
      if (d->act == MMLREC_ACT_RELU && d->colsum == nullptr && !d->accumulate) epi = EPI_F32_FWD;
      else if (d->act == MMLREC_ACT_NONE && d->bias == nullptr) epi = EPI_F32_PLAIN;
    }
    if (d->c_transposed) { epi = EPI_F32_TRANSPOSED; rec.c_t = d->C_f32; rec.ldc_t = d->ldc_f32; rec.colsum_b = d->colsum_b; }
    rec.epi = epi;
  }
  rec.bits_out = d->relu_bits_out; rec.bits_out_chunks = d->bits_out_chunks; rec.bits_out_chunk0 = d->bits_out_chunk0;
  rec.mask_bits = d->mask_bits; rec.mask_bits_chunks = d->mask_bits_chunks; rec.mask_bits_chunk0 = d->mask_bits_chunk0;
  MMLREC_CHECK_ARG(d->relu_bits_out == nullptr || d->bits_out_chunks > 0, "bits_out_chunks");
  MMLREC_CHECK_ARG(d->mask_bits == nullptr || d->mask_bits_chunks > 0, "mask_bits_chunks");
  memcpy(record_host, &rec, sizeof(rec));
  return 0;
}

extern "C" int mmlrec_gemm_grouped_tc2(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                       int32_t total_tiles, const int32_t* tile_order, const int32_t* pair_start,
                                       int32_t n_pairs, int64_t* stamps, void* stream) {
  MMLREC_CHECK_ARG(records && tile_prefix && n_problems > 0 && total_tiles >= 0, "bad args");
  MMLREC_CHECK_ARG(((uintptr_t)records & 127) == 0, "record table must be 128-byte aligned");
  MMLREC_CHECK_ARG(n_problems <= T2_MAX_PROBLEMS, "too many problems in one launch (split the table)");
  MMLREC_CHECK_ARG((tile_order == nullptr) == (pair_start == nullptr), "tile_order and pair_start come together");
  if (total_tiles == 0) return 0;
  static bool opted = false;
  if (!opted) {
    cudaError_t e = cudaFuncSetAttribute(gemm_grouped_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T2_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_tc2: smem opt-in failed: %s", cudaGetErrorString(e)); return (int)e; }
    opted = true;
  }
  const int max_pairs = tc_sm_count() / 2;
  int pairs = tile_order ? n_pairs : (total_tiles < max_pairs ? total_tiles : max_pairs);
  MMLREC_CHECK_ARG(pairs > 0 && pairs <= max_pairs, "bad pair count");
  launch_pdl_if(pdl_enabled_gemm(), gemm_grouped_tc2_kernel, dim3(2 * pairs), dim3(T2_THREADS), (size_t)T2_SMEM_BYTES, stream,
             reinterpret_cast<const Tc2Record*>(records), tile_prefix, n_problems, total_tiles, tile_order, pair_start,
             reinterpret_cast<long long*>(stamps));
  MMLREC_RETURN_LAUNCH(1);
}
