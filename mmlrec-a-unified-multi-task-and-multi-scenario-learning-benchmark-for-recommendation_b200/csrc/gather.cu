// K1: multi-field embedding gather + concat (HBM-bound; see DESIGN.md "K1").
//
// One CTA owns a tile of R consecutive samples.  (1) the X tile is staged in shared memory with
// coalesced loads; (2) every (sample, field, 16-byte chunk) becomes one cp.async (LDGSTS) from
// the table row straight into the shared output tile -- no register staging, so a thread keeps
// many independent 128-bit loads in flight; dense features are copied from the staged X tile;
// (3) the finished tile is streamed out with coalesced 128-bit stores (fp32 and/or bf16).
#include "common.cuh"

namespace mmlrec {

struct GatherArgs {
  const float* X; int64_t ldx; int B;
  const float* emb; const int64_t* field_meta; int F_s; int D;
  const int32_t* dense_xcol; int F_d; int dense_out_col;
  float* out_f32; int64_t ld_f32;
  uint16_t* out_bf16; int64_t ld_bf16;
  int32_t* oob_flag;
  int R;        // rows per tile
  int x_cols;   // columns of X staged per row (max referenced column + 1)
  int in_dim;   // F_s*D + F_d
  int W;        // shared tile row width in floats (multiple of 4, >= in_dim, >= ld_bf16 if bf16 out)
  int n_tiles;
  const float* const* shards;  // non-null: row-sharded tables, shards[o] = base of rank o's shard (peer memory)
  int n_shards;
  const float* staged;   // non-null: rows were delivered by their owners into [B][F_s*D] staging rows (peer.cu)
  int flags;             // 1: table rows with an L2 evict-last hint; 2: X with an evict-first hint; 4: small tables through L1
  int small_vocab;       // tables with at most this many rows count as small
};

constexpr int kGatherThreads = 256;

// L2 policies: the tables are the only data of this kernel with reuse (small vocabularies are hit thousands of times
// per batch), X and the output stream through once.  Table rows are fetched with an evict-last hint, X with
// evict-first loads and the output with streaming stores, so the 126 MB L2 holds table rows instead of the stream.
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void cp_async_16_hint(void* smem_dst, const void* gmem_src, uint64_t policy) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;\n" ::"r"(d), "l"(gmem_src), "l"(policy));
}
// 16-byte cp.async that allocates in L1 (.ca): rows of SMALL tables are read by every SM thousands of times per
// batch; fetched with .cg they all hit the same few L2 lines (one slice serves the whole chip), from L1 they do not
__device__ __forceinline__ void cp_async_16_ca(void* smem_dst, const void* gmem_src) {
  uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ float4 ld_stream_f4(const float4* p, uint64_t policy) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
  return v;
}

__global__ void __launch_bounds__(kGatherThreads) gather_concat_kernel(const GatherArgs a) {
  pdl_prologue();
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* out_s = reinterpret_cast<float*>(smem_raw);                 // [R][W]
  float* x_s = out_s + (size_t)a.R * a.W;                            // [R][x_cols]
  int64_t* meta_s = reinterpret_cast<int64_t*>(x_s + (((size_t)a.R * a.x_cols + 1) & ~(size_t)1));  // [F_s][4]
  int32_t* dcol_s = reinterpret_cast<int32_t*>(meta_s + (size_t)a.F_s * 4);                        // [F_d]

  const int tid = threadIdx.x;
  for (int i = tid; i < a.F_s * 4; i += kGatherThreads) meta_s[i] = a.field_meta[i];
  for (int i = tid; i < a.F_d; i += kGatherThreads) dcol_s[i] = a.dense_xcol[i];

  const int chunks_per_field = a.D >> 2;
  const int sparse_chunks = a.F_s * chunks_per_field;
  const int W4 = a.W >> 2;
  const uint64_t pol_keep = policy_evict_last(), pol_stream = policy_evict_first();
  const bool x_vec = (a.ldx == a.x_cols) && ((a.ldx & 3) == 0) && ((reinterpret_cast<uintptr_t>(a.X) & 15) == 0);

  // Flat loops over the tile (every lane busy) with the row / column split done by a float reciprocal instead of an
  // integer division: q = trunc((i + 0.5) / d) is exact for i < 2^21 (the margin 0.5 / d exceeds the rounding error
  // i * 2^-22 / d), and a tile has < 2^15 elements.  The first version of this kernel spent ~70 % of its issue slots
  // on 32-bit divisions (profiles/ncu_gather_r02.txt).
  const float inv_chunks = 1.0f / (float)max(sparse_chunks, 1), inv_fd = 1.0f / (float)max(a.F_d, 1);
  const float inv_pad = 1.0f / (float)max(a.W - a.in_dim, 1), inv_w4 = 1.0f / (float)W4, inv_in = 1.0f / (float)a.in_dim;
  const float inv_xc = 1.0f / (float)a.x_cols, inv_nb4 = 1.0f / (float)max((int)(a.ld_bf16 >> 2), 1);
  auto fdiv = [](int i, float inv) { return __float2int_rz(((float)i + 0.5f) * inv); };
  const int cpf_shift = (chunks_per_field & (chunks_per_field - 1)) == 0 ? __ffs(chunks_per_field) - 1 : -1;
  const bool out_flat = a.out_f32 != nullptr && a.ld_f32 == a.W && ((reinterpret_cast<uintptr_t>(a.out_f32) & 15) == 0);
  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int r0 = tile * a.R;
    const int rows = min(a.R, a.B - r0);
    __syncthreads();  // previous tile fully written out / meta visible
    // (1) stage X tile: whole rows are contiguous, so the tile is one flat block of 128-bit streaming loads
    if (x_vec) {
      const float4* src = reinterpret_cast<const float4*>(a.X + (int64_t)r0 * a.ldx);
      float4* dst = reinterpret_cast<float4*>(x_s);
      const int n4 = (rows * a.x_cols) >> 2;
      if (a.flags & 2) { for (int i = tid; i < n4; i += kGatherThreads) dst[i] = ld_stream_f4(src + i, pol_stream); }
      else { for (int i = tid; i < n4; i += kGatherThreads) dst[i] = __ldg(src + i); }
    } else {
      for (int i = tid; i < rows * a.x_cols; i += kGatherThreads) {
        const int r = fdiv(i, inv_xc), c = i - r * a.x_cols;
        x_s[i] = __ldg(a.X + (int64_t)(r0 + r) * a.ldx + c);
      }
    }
    // zero the padding columns [in_dim, W) once per tile (they are never written by the gather)
    const int pad = a.W - a.in_dim;
    for (int i = tid; i < rows * pad; i += kGatherThreads) {
      const int r = fdiv(i, inv_pad), c = a.in_dim + (i - r * pad);
      out_s[r * a.W + c] = 0.f;
    }
    __syncthreads();
    // (2) gather: one cp.async per (row, field, 16-byte chunk); the chunks of a row sit in adjacent lanes, so a warp
    // instruction asks for each 32-byte sector once
    bool oob = false;
    for (int i = tid; i < rows * sparse_chunks; i += kGatherThreads) {
      const int r = fdiv(i, inv_chunks), ch = i - r * sparse_chunks;
      const int f = cpf_shift >= 0 ? (ch >> cpf_shift) : (ch / chunks_per_field);
      const int part = ch - f * chunks_per_field;
      const int64_t* m = meta_s + f * 4;
      int64_t id = (int64_t)x_s[r * a.x_cols + (int)m[2]];  // fp32 -> int64 truncation == .long()
      if (id < 0 || id >= m[1]) { oob = true; id = id < 0 ? 0 : m[1] - 1; }
      float* dst = out_s + r * a.W + (int)m[3] + (part << 2);
      if (a.staged) {
        cp_async_16(dst, a.staged + (int64_t)(r0 + r) * ((int64_t)a.F_s * a.D) + (int64_t)f * a.D + (part << 2));
        continue;
      }
      const float* src;
      if (a.shards) {   // owner = id mod R holds the row at local row id / R (same table offsets on every rank)
        const int64_t q = id / a.n_shards;
        src = a.shards[(int)(id - q * a.n_shards)] + m[0] + q * a.D + (part << 2);
      } else {
        src = a.emb + m[0] + id * a.D + (part << 2);
      }
      if ((a.flags & 4) && m[1] <= a.small_vocab) cp_async_16_ca(dst, src);
      else if (a.flags & 1) cp_async_16_hint(dst, src, pol_keep);
      else cp_async_16(dst, src);
    }
    for (int i = tid; i < rows * a.F_d; i += kGatherThreads) {
      const int r = fdiv(i, inv_fd), d = i - r * a.F_d;
      out_s[r * a.W + a.dense_out_col + d] = x_s[r * a.x_cols + dcol_s[d]];
    }
    if (oob && a.oob_flag) *a.oob_flag = 1;
    cp_async_commit_wait_all();
    __syncthreads();
    // (3) stream the tile out
    if (out_flat) {   // the output rows are as wide as the shared tile: the tile is one contiguous block in memory too
      const float4* src = reinterpret_cast<const float4*>(out_s);
      float4* dst = reinterpret_cast<float4*>(a.out_f32 + (int64_t)r0 * a.ld_f32);
      for (int i = tid; i < rows * W4; i += kGatherThreads) __stcs(dst + i, src[i]);
    } else if (a.out_f32) {
      if ((a.ld_f32 & 3) == 0 && a.ld_f32 >= a.W) {
        for (int i = tid; i < rows * W4; i += kGatherThreads) {
          const int r = fdiv(i, inv_w4), c4 = i - r * W4;
          __stcs(reinterpret_cast<float4*>(a.out_f32 + (int64_t)(r0 + r) * a.ld_f32) + c4, reinterpret_cast<const float4*>(out_s)[i]);
        }
      } else {
        for (int i = tid; i < rows * a.in_dim; i += kGatherThreads) {
          const int r = fdiv(i, inv_in), c = i - r * a.in_dim;
          a.out_f32[(int64_t)(r0 + r) * a.ld_f32 + c] = out_s[r * a.W + c];
        }
      }
    }
    if (a.out_bf16) {
      const int nb4 = (int)(a.ld_bf16 >> 2);  // ld_bf16 is a multiple of 8 (16 B)
      for (int i = tid; i < rows * nb4; i += kGatherThreads) {
        const int r = fdiv(i, inv_nb4), c4 = i - r * nb4;
        const float4 v = *reinterpret_cast<const float4*>(out_s + r * a.W + (c4 << 2));
        uint2 o;
        o.x = pack_bf16x2(v.x, v.y);
        o.y = pack_bf16x2(v.z, v.w);
        *reinterpret_cast<uint2*>(a.out_bf16 + (int64_t)(r0 + r) * a.ld_bf16 + (c4 << 2)) = o;
      }
    }
  }
}

}  // namespace mmlrec

static int gather_launch(const float* X, int64_t ldx, int32_t B, const float* emb, const float* const* shards,
                         int32_t n_shards, const float* staged, const int64_t* field_meta, int32_t F_s, int32_t D,
                         const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                         float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                         int32_t* oob_flag, void* stream);

extern "C" int mmlrec_gather_concat(const float* X, int64_t ldx, int32_t B, const float* emb,
                                    const int64_t* field_meta, int32_t F_s, int32_t D,
                                    const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                                    float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                                    int32_t* oob_flag, void* stream) {
  return gather_launch(X, ldx, B, emb, nullptr, 0, nullptr, field_meta, F_s, D, dense_xcol, F_d, dense_out_col, out_f32, ld_f32,
                       out_bf16, ld_bf16, oob_flag, stream);
}

extern "C" int mmlrec_gather_concat_sharded(const float* X, int64_t ldx, int32_t B, const float* const* shards,
                                            int32_t n_shards, const int64_t* field_meta, int32_t F_s, int32_t D,
                                            const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                                            float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                                            int32_t* oob_flag, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(shards != nullptr && n_shards > 0, "no shards");
  return gather_launch(X, ldx, B, nullptr, shards, n_shards, nullptr, field_meta, F_s, D, dense_xcol, F_d, dense_out_col, out_f32,
                       ld_f32, out_bf16, ld_bf16, oob_flag, stream);
}

extern "C" int mmlrec_gather_concat_staged(const float* X, int64_t ldx, int32_t B, const float* staged_rows,
                                           const int64_t* field_meta, int32_t F_s, int32_t D,
                                           const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                                           float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                                           void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(staged_rows != nullptr, "no staging rows");
  return gather_launch(X, ldx, B, nullptr, nullptr, 0, staged_rows, field_meta, F_s, D, dense_xcol, F_d, dense_out_col,
                       out_f32, ld_f32, out_bf16, ld_bf16, nullptr, stream);
}

static int gather_launch(const float* X, int64_t ldx, int32_t B, const float* emb, const float* const* shards,
                         int32_t n_shards, const float* staged, const int64_t* field_meta, int32_t F_s, int32_t D,
                         const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                         float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                         int32_t* oob_flag, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(B >= 0 && F_s >= 0 && F_d >= 0 && F_s + F_d > 0, "bad sizes");
  MMLREC_CHECK_ARG(F_s == 0 || (D > 0 && (D & 3) == 0), "embedding dim must be a positive multiple of 4");
  MMLREC_CHECK_ARG(out_f32 || out_bf16, "no output");
  if (B == 0) return 0;
  GatherArgs a;
  a.shards = shards; a.n_shards = n_shards;
  a.staged = staged;
  { static int fl = -1; if (fl < 0) { const char* e = getenv("MMLREC_GATHER_FLAGS"); fl = e ? atoi(e) : 5; } a.flags = fl; }
  { static int sv = -1; if (sv < 0) { const char* e = getenv("MMLREC_GATHER_SMALL_VOCAB"); sv = e ? atoi(e) : 8192; } a.small_vocab = sv; }
  a.X = X; a.ldx = ldx; a.B = B; a.emb = emb; a.field_meta = field_meta; a.F_s = F_s; a.D = D;
  a.dense_xcol = dense_xcol; a.F_d = F_d; a.dense_out_col = dense_out_col;
  a.out_f32 = out_f32; a.ld_f32 = ld_f32; a.out_bf16 = out_bf16; a.ld_bf16 = ld_bf16; a.oob_flag = oob_flag;
  a.in_dim = F_s * D + F_d;
  MMLREC_CHECK_ARG(dense_out_col + F_d <= a.in_dim || F_d == 0, "dense block outside the row");
  a.W = (a.in_dim + 3) & ~3;
  if (out_bf16) {
    MMLREC_CHECK_ARG((ld_bf16 & 7) == 0 && ld_bf16 >= a.in_dim, "ld_bf16 must be a multiple of 8 and >= in_dim");
    if (ld_bf16 > a.W) a.W = (int)ld_bf16;
  }
  MMLREC_CHECK_ARG(out_f32 == nullptr || ld_f32 >= a.in_dim, "ld_f32 < in_dim");
  a.x_cols = (int)ldx;  // stage whole rows: every column of X is a feature in the reference layout
  MMLREC_CHECK_ARG(a.x_cols >= F_s + F_d, "ldx smaller than the feature count");
  // rows per tile: ~32 KB of shared memory per CTA (6 CTAs resident per SM, the rest of the 228 KB stays L1 for the
  // rows of small tables); measured sweep in profiles/gather_sweep_r02.txt
  size_t row_bytes = (size_t)(a.W + a.x_cols) * 4;
  int R = (int)(24576 / row_bytes);
  { static int tb = 0; if (!tb) { const char* e = getenv("MMLREC_GATHER_TILE_BYTES"); tb = e ? atoi(e) : 32768; } R = (int)(tb / row_bytes); }
  R = R < 1 ? 1 : (R > 64 ? 64 : R);
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
  while (R > 1 && (int64_t)mmlrec::cdiv(B, R) < 2 * n_sm) R >>= 1;  // enough tiles to fill the chip
  a.R = R;
  a.n_tiles = cdiv(B, R);
  size_t smem = (size_t)R * a.W * 4 + ((((size_t)R * a.x_cols + 1) & ~(size_t)1) * 4) + (size_t)F_s * 32 + (size_t)F_d * 4 + 16;
  static int smem_opt_in = 0;
  if (smem > 48 * 1024 && smem > (size_t)smem_opt_in) {
    cudaError_t e = cudaFuncSetAttribute(gather_concat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gather: smem opt-in failed (%zu bytes)", smem); return (int)e; }
    smem_opt_in = (int)smem;
  }
  int grid = a.n_tiles < n_sm * 16 ? a.n_tiles : n_sm * 16;
  launch_pdl(gather_concat_kernel, dim3(grid), dim3(kGatherThreads), smem, stream, a);
  MMLREC_RETURN_LAUNCH(1);
}
