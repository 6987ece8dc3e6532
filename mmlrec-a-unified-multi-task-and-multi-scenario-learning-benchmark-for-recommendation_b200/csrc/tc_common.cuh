// PTX wrappers shared by the tcgen05 grouped-GEMM kernels (gemm_tc.cu: one CTA per tile; gemm_tc2.cu: CTA pairs).
#pragma once
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"

namespace mmlrec {

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
// shared -> global bulk tensor store of one box (coordinates {column, row}); out-of-range rows / columns are clipped
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
// same box, added element-wise into global memory (fp32)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"((uint64_t)tmap), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
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
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

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

// One 32-column chunk of one accumulator row (thread = row): ReLU-mask of the raw accumulator (keep where the bf16
// mask value is > 0), + bias (broadcast reads of the warp's shared-memory bias slot), activation.
__device__ __forceinline__ float4 ld_shared_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void epilogue_math(uint32_t (&r)[32], const uint4 (&mk)[4], bool has_mask, uint32_t bias32, int act) {
  if (has_mask) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint4 w4 = mk[j >> 3];
      const uint32_t word = ((j >> 1) & 3) == 0 ? w4.x : ((j >> 1) & 3) == 1 ? w4.y : ((j >> 1) & 3) == 2 ? w4.z : w4.w;
      const uint32_t mb = (j & 1) ? (word >> 16) : (word & 0xFFFFu);
      if (!((mb & 0x8000u) == 0 && (mb & 0x7FFFu) != 0)) r[j] = 0u;
    }
  }
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 b = ld_shared_f4(bias32 + 16 * j4);
    r[4 * j4 + 0] = __float_as_uint(__uint_as_float(r[4 * j4 + 0]) + b.x);
    r[4 * j4 + 1] = __float_as_uint(__uint_as_float(r[4 * j4 + 1]) + b.y);
    r[4 * j4 + 2] = __float_as_uint(__uint_as_float(r[4 * j4 + 2]) + b.z);
    r[4 * j4 + 3] = __float_as_uint(__uint_as_float(r[4 * j4 + 3]) + b.w);
  }
  if (act == MMLREC_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(fmaxf(__uint_as_float(r[j]), 0.f));
  } else if (act != MMLREC_ACT_NONE) {
#pragma unroll   // (a rolled loop would index r[] dynamically and push the whole accumulator row into local memory)
    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(apply_act(__uint_as_float(r[j]), act));
  }
}


// host side: tensor-map encoding through the driver entry point (no link-time libcuda dependency)
typedef CUresult (*TcEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline TcEncodeTiledFn tc_get_encode_fn() {
  static TcEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = (TcEncodeTiledFn)p;
  }
  return fn;
}

// 2-D array stored row-major [outer, inner] with row stride ld (elements of `elem_bytes`), 128B-swizzled boxes
static inline int tc_encode_map(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base, int64_t ld,
                                int64_t inner, int64_t outer, int box_inner, int box_outer) {
  TcEncodeTiledFn fn = tc_get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable (no driver?)"); return -2; }
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * (cuuint64_t)elem_bytes};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  static int promo = -1;   // MMLREC_TMA_L2_PROMOTION = 0 (none) / 64 / 128 / 256 (default): A/B switch for the profiles
  if (promo < 0) {
    const char* e = getenv("MMLREC_TMA_L2_PROMOTION");
    promo = e ? atoi(e) : 256;
  }
  const CUtensorMapL2promotion l2p = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                     : promo == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                     : promo == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  CUresult r = fn(tm, dt, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, l2p, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed with CUresult %d", (int)r); return -3; }
  return 0;
}

static inline int tc_sm_count() {
  static int sm_count = 0;
  if (!sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
  }
  return sm_count;
}

}  // namespace mmlrec
