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
};

constexpr int kGatherThreads = 256;

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

  for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
    const int r0 = tile * a.R;
    const int rows = min(a.R, a.B - r0);
    __syncthreads();  // previous tile fully written out / meta visible
    // (1) stage X tile
    for (int i = tid; i < rows * a.x_cols; i += kGatherThreads) {
      int r = i / a.x_cols, c = i - r * a.x_cols;
      x_s[r * a.x_cols + c] = __ldg(a.X + (int64_t)(r0 + r) * a.ldx + c);
    }
    // zero the padding columns [in_dim, W) once per tile (they are never written by the gather)
    const int pad = a.W - a.in_dim;
    for (int i = tid; i < rows * pad; i += kGatherThreads) {
      int r = i / pad, c = a.in_dim + (i - r * pad);
      out_s[r * a.W + c] = 0.f;
    }
    __syncthreads();
    // (2) gather: one cp.async per (row, field, chunk)
    bool oob = false;
    for (int i = tid; i < rows * sparse_chunks; i += kGatherThreads) {
      int r = i / sparse_chunks, ch = i - r * sparse_chunks;
      int f = ch / chunks_per_field, part = ch - f * chunks_per_field;
      const int64_t* m = meta_s + f * 4;
      int64_t id = (int64_t)x_s[r * a.x_cols + (int)m[2]];  // fp32 -> int64 truncation == .long()
      if (id < 0 || id >= m[1]) { oob = true; id = id < 0 ? 0 : m[1] - 1; }
      const float* src;
      if (a.staged) {
        cp_async_16(out_s + r * a.W + (int)m[3] + (part << 2),
                    a.staged + (int64_t)(r0 + r) * ((int64_t)a.F_s * a.D) + (int64_t)f * a.D + (part << 2));
        continue;
      }
      if (a.shards) {   // owner = id mod R holds the row at local row id / R (same table offsets on every rank)
        const int64_t q = id / a.n_shards;
        src = a.shards[(int)(id - q * a.n_shards)] + m[0] + q * a.D + (part << 2);
      } else {
        src = a.emb + m[0] + id * a.D + (part << 2);
      }
      cp_async_16(out_s + r * a.W + (int)m[3] + (part << 2), src);
    }
    for (int i = tid; i < rows * a.F_d; i += kGatherThreads) {
      int r = i / a.F_d, d = i - r * a.F_d;
      out_s[r * a.W + a.dense_out_col + d] = x_s[r * a.x_cols + dcol_s[d]];
    }
    if (oob && a.oob_flag) *a.oob_flag = 1;
    cp_async_commit_wait_all();
    __syncthreads();
    // (3) stream the tile out
    if (a.out_f32) {
      if ((a.ld_f32 & 3) == 0 && a.ld_f32 >= a.W) {
        for (int i = tid; i < rows * W4; i += kGatherThreads) {
          int r = i / W4, c4 = i - r * W4;
          float4 v = *reinterpret_cast<const float4*>(out_s + r * a.W + (c4 << 2));
          __stcs(reinterpret_cast<float4*>(a.out_f32 + (int64_t)(r0 + r) * a.ld_f32) + c4, v);
        }
      } else {
        for (int i = tid; i < rows * a.in_dim; i += kGatherThreads) {
          int r = i / a.in_dim, c = i - r * a.in_dim;
          a.out_f32[(int64_t)(r0 + r) * a.ld_f32 + c] = out_s[r * a.W + c];
        }
      }
    }
    if (a.out_bf16) {
      const int nb4 = (int)(a.ld_bf16 >> 2);  // ld_bf16 is a multiple of 8 (16 B)
      for (int i = tid; i < rows * nb4; i += kGatherThreads) {
        int r = i / nb4, c4 = i - r * nb4;
        float4 v = *reinterpret_cast<const float4*>(out_s + r * a.W + (c4 << 2));
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
  // rows per tile: ~24 KB of shared memory per CTA so that ~8 CTAs are resident per SM
  size_t row_bytes = (size_t)(a.W + a.x_cols) * 4;
  int R = (int)(24576 / row_bytes);
  R = R < 1 ? 1 : (R > 64 ? 64 : R);
  while (R > 1 && (int64_t)mmlrec::cdiv(B, R) < 2 * 148) R >>= 1;  // enough tiles to fill the chip
  a.R = R;
  a.n_tiles = cdiv(B, R);
  size_t smem = (size_t)R * a.W * 4 + ((((size_t)R * a.x_cols + 1) & ~(size_t)1) * 4) + (size_t)F_s * 32 + (size_t)F_d * 4 + 16;
  static int smem_opt_in = 0;
  if (smem > 48 * 1024 && smem > (size_t)smem_opt_in) {
    cudaError_t e = cudaFuncSetAttribute(gather_concat_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gather: smem opt-in failed (%zu bytes)", smem); return (int)e; }
    smem_opt_in = (int)smem;
  }
  int grid = a.n_tiles < 148 * 16 ? a.n_tiles : 148 * 16;
  launch_pdl(gather_concat_kernel, dim3(grid), dim3(kGatherThreads), smem, stream, a);
  MMLREC_RETURN_LAUNCH(1);
}
