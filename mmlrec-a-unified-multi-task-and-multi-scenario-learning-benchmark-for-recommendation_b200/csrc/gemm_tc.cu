// K3, bf16 throughput mode: grouped GEMM on the 5th-generation tensor cores (tcgen05), operands
// staged by TMA into 128B-swizzled shared memory, fp32 accumulators in TMEM, warp-specialised
// persistent CTAs (1 per SM):
//   warp 0        TMA producer (one elected lane)
//   warp 1        TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2..5    epilogue: tcgen05.ld -> bias / activation / ReLU-mask -> fp32 and/or bf16 stores
// D[M,N] = A * B^T.  Each operand is either K-major (row-major [rows,K]) or MN-major ([K,rows]);
// the MN-major form is what wgrad needs (dW = dZ^T X reads both activations "transposed"), so no
// transposed copy of any activation is ever written.  The bias gradient of a wgrad problem
// (row sums of A) is produced by one extra N=16 MMA per k-step against a constant all-ones B tile.
#include <cuda.h>

#include "common.cuh"

namespace mmlrec {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64;
constexpr int TC_STAGES = 4;
constexpr int TC_ACC_STAGES = 2;
constexpr int TC_ACC_COLS = 256;                 // TMEM columns per accumulator stage (128 main + 16 row-sum, padded)
constexpr int TC_TMEM_COLS = 512;
constexpr int TC_THREADS = 192;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 2;    // 16 KB
constexpr int TC_B_BYTES = TC_BN * TC_BK * 2;    // 16 KB
constexpr int TC_ONES_BYTES = 16 * 128;          // 16 rows x 128 B of bf16 1.0
constexpr int TC_SMEM_BYTES = 1024 /*align slack*/ + TC_STAGES * (TC_A_BYTES + TC_B_BYTES) + TC_ONES_BYTES + 256;

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
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
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

struct TileCoord { int pi, tm, tn; };
__device__ __forceinline__ TileCoord locate_tile(int t, const int32_t* __restrict__ prefix, int n_problems,
                                                 const TcRecord* __restrict__ recs) {
  int pi = 0;
  while (pi + 1 < n_problems && prefix[pi + 1] <= t) ++pi;
  int local = t - prefix[pi];
  int tiles_n = recs[pi].tiles_n;
  TileCoord c;
  c.pi = pi; c.tm = local / tiles_n; c.tn = local - c.tm * tiles_n;
  return c;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_grouped_tc_kernel(const TcRecord* __restrict__ recs, const int32_t* __restrict__ prefix, int n_problems, int total_tiles) {
  extern __shared__ unsigned char smem_dyn[];
  // 1024-byte alignment is required by the 128B swizzle atom
  unsigned char* smem = reinterpret_cast<unsigned char*>(((uintptr_t)smem_dyn + 1023) & ~(uintptr_t)1023);
  unsigned char* sA = smem;
  unsigned char* sB = smem + TC_STAGES * TC_A_BYTES;
  unsigned char* sOnes = smem + TC_STAGES * (TC_A_BYTES + TC_B_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOnes + TC_ONES_BYTES);
  // bars: [0,S) full, [S,2S) empty, [2S,2S+2) tmem_full, [2S+2,2S+4) tmem_empty, then tmem base slot
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * TC_STAGES + 2 * TC_ACC_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + TC_STAGES);
  const uint32_t tfull_bar = smem_u32(bars + 2 * TC_STAGES), tempty_bar = smem_u32(bars + 2 * TC_STAGES + TC_ACC_STAGES);

  // all-ones B tile for the row-sum MMA
  for (int i = threadIdx.x; i < TC_ONES_BYTES / 4; i += TC_THREADS) reinterpret_cast<uint32_t*>(sOnes)[i] = 0x3F803F80u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(full_bar + 8 * s, 1); mbar_init(empty_bar + 8 * s, 1); }
    for (int s = 0; s < TC_ACC_STAGES; ++s) { mbar_init(tfull_bar + 8 * s, 1); mbar_init(tempty_bar + 8 * s, 4); }
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = locate_tile(t, prefix, n_problems, recs);
        const TcRecord* R = recs + tc.pi;
        const int K = R->K, a_mn = R->a_mn, b_mn = R->b_mn;
        const int m0 = tc.tm * TC_BM, n0 = tc.tn * TC_BN;
        const int num_kb = (K + TC_BK - 1) / TC_BK;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
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
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const uint64_t ones_desc = make_smem_desc(smem_u32(sOnes), false);
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TileCoord tc = locate_tile(t, prefix, n_problems, recs);
        const TcRecord* R = recs + tc.pi;
        const int K = R->K;
        const bool a_mn = R->a_mn != 0, b_mn = R->b_mn != 0;
        const bool rowsum = (R->rowsum_a != nullptr) && tc.tn == 0;
        const uint32_t idesc = make_idesc(TC_BM, TC_BN, a_mn, b_mn);
        const uint32_t idesc_ones = make_idesc(TC_BM, 16, a_mn, false);
        const int num_kb = (K + TC_BK - 1) / TC_BK;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * TC_ACC_COLS;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          tc_fence_after();
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
        if (++acc == TC_ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const TileCoord tc = locate_tile(t, prefix, n_problems, recs);
      const TcRecord* R = recs + tc.pi;
      const int M = R->M, N = R->N;
      const int m = tc.tm * TC_BM + q * 32 + lane;
      const int n0 = tc.tn * TC_BN;
      float* const cf = R->C_f32; const int64_t ldcf = R->ldc_f32;
      uint16_t* const cb = R->C_bf16; const int64_t ldcb = R->ldc_bf16;
      const float* const bias = R->bias;
      const uint16_t* const mask = R->mask; const int64_t ldmask = R->ldmask;
      const int act = R->act, accumulate = R->accumulate;
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + acc * TC_ACC_COLS + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int c = 0; c < TC_BN / 32; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= N) break;                  // warp-uniform
        uint32_t r[32];
        tc_ld32(t_row + c * 32, r);
        tc_wait_ld();
        if (m < M) {
          const bool full = nc + 32 <= N;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float x = __uint_as_float(r[j]);
            if (bias && (full || nc + j < N)) x += __ldg(bias + nc + j);
            v[j] = apply_act(x, act);
          }
          if (mask) {
            const uint16_t* mrow = mask + (int64_t)m * ldmask + nc;
            if (full) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint4 mv = __ldg(reinterpret_cast<const uint4*>(mrow) + j8);
                uint32_t w4[4] = {mv.x, mv.y, mv.z, mv.w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  // bf16 > 0  <=>  sign bit clear and magnitude non-zero
                  uint32_t lo = w4[u] & 0xFFFFu, hi = w4[u] >> 16;
                  if (!((lo & 0x8000u) == 0 && (lo & 0x7FFFu) != 0)) v[j8 * 8 + u * 2] = 0.f;
                  if (!((hi & 0x8000u) == 0 && (hi & 0x7FFFu) != 0)) v[j8 * 8 + u * 2 + 1] = 0.f;
                }
              }
            } else {
              for (int j = 0; j < 32 && nc + j < N; ++j) {
                uint32_t b = mrow[j];
                if (!((b & 0x8000u) == 0 && (b & 0x7FFFu) != 0)) v[j] = 0.f;
              }
            }
          }
          if (cf) {
            float* crow = cf + (int64_t)m * ldcf + nc;
            if (full) {
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float4 o = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
                if (accumulate) {
                  float4 old = reinterpret_cast<float4*>(crow)[j4];
                  o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
                }
                reinterpret_cast<float4*>(crow)[j4] = o;
              }
            } else {
              for (int j = 0; j < 32 && nc + j < N; ++j) crow[j] = accumulate ? crow[j] + v[j] : v[j];
            }
          }
          if (cb) {
            uint16_t* brow = cb + (int64_t)m * ldcb + nc;
            if (full) {
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                uint4 o;
                o.x = pack_bf16x2(v[8 * j8], v[8 * j8 + 1]);
                o.y = pack_bf16x2(v[8 * j8 + 2], v[8 * j8 + 3]);
                o.z = pack_bf16x2(v[8 * j8 + 4], v[8 * j8 + 5]);
                o.w = pack_bf16x2(v[8 * j8 + 6], v[8 * j8 + 7]);
                reinterpret_cast<uint4*>(brow)[j8] = o;
              }
            } else {
              for (int j = 0; j < 32 && nc + j < N; ++j) brow[j] = float_to_bf16_bits(v[j]);
            }
          }
        }
      }
      if (R->rowsum_a != nullptr && tc.tn == 0) {
        uint32_t rs;
        tc_ld1(t_row + TC_BN, rs);
        tc_wait_ld();
        if (m < M) R->rowsum_a[m] = __uint_as_float(rs);
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

extern "C" int mmlrec_gemm_grouped_tc(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                      int32_t total_tiles, void* stream) {
  MMLREC_CHECK_ARG(records && tile_prefix && n_problems > 0 && total_tiles >= 0, "bad args");
  MMLREC_CHECK_ARG(((uintptr_t)records & 127) == 0, "record table must be 128-byte aligned");
  if (total_tiles == 0) return 0;
  static int sm_count = 0;
  static bool opted = false;
  if (!opted) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(gemm_grouped_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) { set_error("gemm_tc: smem opt-in failed: %s", cudaGetErrorString(e)); return (int)e; }
    opted = true;
  }
  int grid = total_tiles < sm_count ? total_tiles : sm_count;
  gemm_grouped_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(
      reinterpret_cast<const TcRecord*>(records), tile_prefix, n_problems, total_tiles);
  MMLREC_RETURN_LAUNCH(1);
}
