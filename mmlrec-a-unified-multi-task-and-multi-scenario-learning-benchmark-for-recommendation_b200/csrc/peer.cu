// Row-sharded embedding tables over NVLink peer memory (BASELINE config 5).
//
// owner(id) = id mod R; the owner stores row id at local row id / R of its shard.  Every shard and
// every exchange buffer is a cudaMalloc allocation exported with CUDA IPC, so a kernel on rank r can
// address rank o's memory directly through NVSwitch.  All exchanges are ONE-SIDED stores into small,
// dense buffers at slots fixed by (rank, sample, field) -- only owned entries cross the wire (all-to-all
// volume), nothing is appended atomically, so every reduction order is deterministic:
//   forward   push_ids: key (id / R) << 32 | pos into the OWNER's request buffer        [ids all-to-all]
//             barrier; serve_rows: the owner reads its LOCAL rows (2 MB pages, full HBM speed -- random
//             16-byte reads of a multi-GB PEER mapping run ~10x slower, profiles/sharded_timeline_r01.txt)
//             and stores them into the REQUESTER's staging rows                          [rows all-to-all]
//             barrier; K1 assembles dnn_input from the staging rows (gather.cu, `staged` path).
//             The owner sorts the request keys on a side stream while forward/backward run.
//   backward  push_grads: the D gradient floats of every (sample, field) into the owner's receive rows
//             [row-grad all-to-all]; the dense-gradient all-reduce is the barrier; K2 on the owner.
//   (`shards` path of gather.cu: K1 reading rows straight from the owners' shards -- no barriers, the
//   better choice while the tables fit the peer TLB reach.)
// The barrier is a flag exchange through peer memory (one CTA, ~3 us) instead of a collective launch.
#include "common.cuh"

namespace mmlrec {

// rq_keys of one owner: [2 parities][F_s][B_all] uint64, sentinel ~0 (slot not owned by this rank)
// rows_in of one requester: [b][F_s*D] float;  rx_grad of one owner: [B_all][F_s*D] float
struct PushArgs {
  const float* X; int64_t ldx; int b;
  const float* d_input; int64_t ld;
  const int64_t* field_meta; int F_s; int D;
  int rank; int R; int B_all;
  uint64_t* const* rq_keys; float* const* rx_grad;
  const MmlrecHyper* hyper; int32_t* oob_flag;
};

__device__ __forceinline__ int64_t clamp_id(const float* X, int64_t ldx, int i, const int64_t* m, bool& oob) {
  int64_t id = (int64_t)__ldg(X + (int64_t)i * ldx + (int)m[2]);   // fp32 carrier, truncation == .long()
  if (id < 0 || id >= m[1]) { oob = true; id = id < 0 ? 0 : m[1] - 1; }
  return id;
}

__global__ void __launch_bounds__(256) emb_push_ids_kernel(const PushArgs a, int step_offset) {
  const int par = (a.hyper->step + step_offset) & 1;
  const int64_t n = (int64_t)a.b * a.F_s;
  bool oob = false;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(t % a.F_s);
    const int i = (int)(t / a.F_s);
    const int64_t id = clamp_id(a.X, a.ldx, i, a.field_meta + j * 4, oob);
    const int64_t q = id / a.R;
    const int64_t pos = (int64_t)a.rank * a.b + i;
    a.rq_keys[(int)(id - q * a.R)][((int64_t)par * a.F_s + j) * a.B_all + pos] = ((uint64_t)(uint32_t)q << 32) | (uint32_t)pos;
  }
  if (oob && a.oob_flag) *a.oob_flag = 1;
}

__global__ void __launch_bounds__(256) emb_push_grads_kernel(const PushArgs a) {
  const int dv = a.D >> 2;
  const int64_t n = (int64_t)a.b * a.F_s * dv;
  const int64_t row_w = (int64_t)a.F_s * a.D;
  bool oob = false;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int part = (int)(t % dv);
    const int64_t ij = t / dv;
    const int j = (int)(ij % a.F_s);
    const int i = (int)(ij / a.F_s);
    const int64_t* m = a.field_meta + j * 4;
    const int64_t id = clamp_id(a.X, a.ldx, i, m, oob);
    const int o = (int)(id % a.R);
    const int64_t pos = (int64_t)a.rank * a.b + i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.d_input + (int64_t)i * a.ld + (int)m[3]) + part);
    reinterpret_cast<float4*>(a.rx_grad[o] + pos * row_w + (int64_t)j * a.D)[part] = g;
  }
}

// owner: answer every request that landed in rq_keys with the local row, stored into the requester's staging rows
struct ServeArgs {
  const uint64_t* rq_keys; const float* emb; const int64_t* field_meta; int F_s; int D; int b; int B_all;
  float* const* rows_in; const MmlrecHyper* hyper;
};

__global__ void __launch_bounds__(256) emb_serve_rows_kernel(const ServeArgs a, int step_offset) {
  const int par = (a.hyper->step + step_offset) & 1;
  const int dv = a.D >> 2;
  const int64_t n = (int64_t)a.F_s * a.B_all * dv;
  const int64_t row_w = (int64_t)a.F_s * a.D;
  const uint64_t* keys = a.rq_keys + (int64_t)par * a.F_s * a.B_all;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int part = (int)(t % dv);
    const int64_t jp = t / dv;           // = j * B_all + pos: consecutive threads walk one field's slots
    const uint64_t key = keys[jp];
    if (key == ~0ull) continue;
    const int j = (int)(jp / a.B_all);
    const int pos = (int)(jp - (int64_t)j * a.B_all);
    const int r = pos / a.b, i = pos - r * a.b;
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.emb + a.field_meta[j * 4] + (int64_t)(key >> 32) * a.D) + part);
    reinterpret_cast<float4*>(a.rows_in[r] + (int64_t)i * row_w + (int64_t)j * a.D)[part] = v;
  }
}

// Cross-GPU barrier through peer memory: every rank stores its new epoch into slot [rank] of every peer's
// flag vector and waits until all R slots of its own vector reached that epoch.  Work queued on the stream
// before the barrier (peer stores of earlier kernels) is complete and system-visible when it starts.
__global__ void peer_barrier_kernel(int32_t* const* peer_flags, int32_t* local_epoch, int rank, int R, int32_t* err_flag) {
  __shared__ int s_e;
  if (threadIdx.x == 0) { s_e = *local_epoch + 1; *local_epoch = s_e; }
  __syncthreads();
  const int e = s_e;
  __threadfence_system();
  if ((int)threadIdx.x < R) {
    volatile int32_t* theirs = peer_flags[threadIdx.x] + rank;
    *theirs = e;
    volatile int32_t* mine = peer_flags[rank] + threadIdx.x;
    const long long t0 = clock64();
    while (*mine < e) {
      if (clock64() - t0 > 120000000000ll) {   // ~60 s: a peer died (a slow one -- plan build, graph capture -- gets this long);
                                              // fail loudly (err flag, read by BaseModel.check_ids) instead of hanging the GPU
        if (err_flag) *err_flag = 2;
        break;
      }
    }
  }
  __threadfence_system();
}

// Dense-gradient all-reduce (SUM) over peer memory, the data-parallel towers' one collective: rank r adds up slice r
// of every rank's gradient buffer (peer loads, fixed rank order 0..R-1 -> every rank ends up with bit-identical sums)
// and stores the result into slice r of EVERY rank's output buffer (peer stores): a reduce-scatter and an all-gather in
// one pass, 2 x (R-1)/R x n x 4 bytes over NVLink per GPU in each direction.  The caller brackets it with two flag
// barriers (inputs complete everywhere / outputs landed everywhere).  Replaces ncclAllReduce on this path: 56 us for
// 4.3 MB on 8 GPUs there, latency-dominated (profiles/sharded_timeline_8gpu_r02.txt).
__global__ void __launch_bounds__(256) peer_allreduce_kernel(const float* const* __restrict__ in, float* const* __restrict__ out,
                                                             int64_t n4, int rank, int R) {
  __shared__ const float4* s_src[32];
  __shared__ float4* s_dst[32];
  if ((int)threadIdx.x < R) {
    s_src[threadIdx.x] = reinterpret_cast<const float4*>(in[threadIdx.x]);
    s_dst[threadIdx.x] = reinterpret_cast<float4*>(out[threadIdx.x]);
  }
  __syncthreads();
  const int64_t per = (n4 + R - 1) / R;
  const int64_t lo = (int64_t)rank * per, hi = lo + per < n4 ? lo + per : n4;
  for (int64_t i = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p0 = 0; p0 < R; p0 += 8) {
      float4 v[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) if (p0 + q < R) v[q] = __ldcg(s_src[p0 + q] + i);   // up to 8 peer loads in flight, L1 bypassed
#pragma unroll
      for (int q = 0; q < 8; ++q) if (p0 + q < R) { s.x += v[q].x; s.y += v[q].y; s.z += v[q].z; s.w += v[q].w; }
    }
    for (int p = 0; p < R; ++p) s_dst[p][i] = s;
  }
}

}  // namespace mmlrec

using namespace mmlrec;

extern "C" int mmlrec_peer_allreduce_f32(const float* const* peer_in, float* const* peer_out, int64_t n, int32_t rank,
                                         int32_t R, void* stream) {
  MMLREC_CHECK_ARG(peer_in && peer_out && n > 0 && (n & 3) == 0, "buffers of a multiple of 4 floats");
  MMLREC_CHECK_ARG(R > 0 && R <= 32 && rank >= 0 && rank < R, "bad rank / world");
  const int64_t per = ((n >> 2) + R - 1) / R;
  int grid = (int)((per + 255) / 256);
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 1; }
  if (grid > 4 * n_sm) grid = 4 * n_sm;
  if (grid < 1) grid = 1;
  peer_allreduce_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(peer_in, peer_out, n >> 2, rank, R);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_peer_alloc(void** ptr, int64_t bytes) {
  MMLREC_CHECK_ARG(ptr && bytes > 0, "bad args");
  cudaError_t e = cudaMalloc(ptr, (size_t)bytes);
  if (e != cudaSuccess) { set_error("mmlrec_peer_alloc: %s", cudaGetErrorString(e)); return (int)e; }
  e = cudaMemset(*ptr, 0, (size_t)bytes);
  if (e != cudaSuccess) { set_error("mmlrec_peer_alloc: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int mmlrec_peer_free(void* ptr) {
  cudaError_t e = cudaFree(ptr);
  if (e != cudaSuccess) { set_error("mmlrec_peer_free: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int mmlrec_peer_export(void* ptr, unsigned char* handle64) {
  MMLREC_CHECK_ARG(ptr && handle64, "bad args");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, ptr);
  if (e != cudaSuccess) { set_error("mmlrec_peer_export: %s", cudaGetErrorString(e)); return (int)e; }
  memcpy(handle64, &h, 64);
  return 0;
}

extern "C" int mmlrec_peer_import(const unsigned char* handle64, void** ptr) {
  MMLREC_CHECK_ARG(ptr && handle64, "bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  cudaError_t e = cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) { set_error("mmlrec_peer_import: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int mmlrec_peer_close(void* ptr) {
  cudaError_t e = cudaIpcCloseMemHandle(ptr);
  if (e != cudaSuccess) { set_error("mmlrec_peer_close: %s", cudaGetErrorString(e)); return (int)e; }
  return 0;
}

extern "C" int mmlrec_peer_fill_u64(uint64_t* p, int64_t n, uint64_t v, void* stream) {
  MMLREC_CHECK_ARG(p && n >= 0, "bad args");
  if (v == ~0ull || v == 0) {
    cudaError_t e = cudaMemsetAsync(p, v ? 0xff : 0, (size_t)n * 8, (cudaStream_t)stream);
    if (e != cudaSuccess) { set_error("mmlrec_peer_fill_u64: %s", cudaGetErrorString(e)); return (int)e; }
    return 0;
  }
  set_error("mmlrec_peer_fill_u64: only 0 and ~0 are supported");
  return 1;
}

static int push_grid(int64_t n) {
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
  return (int)((n + 255) / 256 < n_sm * 8 ? (n + 255) / 256 : n_sm * 8);
}

extern "C" int mmlrec_emb_push_ids(const float* X, int64_t ldx, int32_t b, const int64_t* field_meta, int32_t F_s,
                                   int32_t rank, int32_t R, int32_t B_all, uint64_t* const* rq_keys,
                                   const MmlrecHyper* hyper, int32_t step_offset, int32_t* oob_flag, void* stream) {
  MMLREC_CHECK_ARG(X && field_meta && rq_keys && hyper, "null argument");
  MMLREC_CHECK_ARG(b > 0 && F_s > 0 && R > 0 && rank >= 0 && rank < R && B_all >= (int64_t)R * b, "bad sizes");
  PushArgs a{X, ldx, b, nullptr, 0, field_meta, F_s, 0, rank, R, B_all, rq_keys, nullptr, hyper, oob_flag};
  emb_push_ids_kernel<<<push_grid((int64_t)b * F_s), 256, 0, (cudaStream_t)stream>>>(a, step_offset);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_emb_serve_rows(const uint64_t* rq_keys, const float* emb, const int64_t* field_meta, int32_t F_s,
                                     int32_t D, int32_t b, int32_t B_all, float* const* rows_in,
                                     const MmlrecHyper* hyper, int32_t step_offset, void* stream) {
  MMLREC_CHECK_ARG(rq_keys && emb && field_meta && rows_in && hyper, "null argument");
  MMLREC_CHECK_ARG(b > 0 && F_s > 0 && D > 0 && (D & 3) == 0 && B_all >= b, "bad sizes");
  ServeArgs a{rq_keys, emb, field_meta, F_s, D, b, B_all, rows_in, hyper};
  emb_serve_rows_kernel<<<push_grid((int64_t)F_s * B_all * (D >> 2)), 256, 0, (cudaStream_t)stream>>>(a, step_offset);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_emb_push_grads(const float* X, int64_t ldx, int32_t b, const float* d_input, int64_t ld,
                                     const int64_t* field_meta, int32_t F_s, int32_t D, int32_t rank, int32_t R,
                                     int32_t B_all, float* const* rx_grad, void* stream) {
  MMLREC_CHECK_ARG(X && d_input && field_meta && rx_grad, "null argument");
  MMLREC_CHECK_ARG(b > 0 && F_s > 0 && D > 0 && (D & 3) == 0 && (ld & 3) == 0, "bad sizes");
  MMLREC_CHECK_ARG(R > 0 && rank >= 0 && rank < R && B_all >= (int64_t)R * b, "bad rank / world / B_all");
  PushArgs a{X, ldx, b, d_input, ld, field_meta, F_s, D, rank, R, B_all, nullptr, rx_grad, nullptr, nullptr};
  emb_push_grads_kernel<<<push_grid((int64_t)b * F_s * (D >> 2)), 256, 0, (cudaStream_t)stream>>>(a);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_peer_barrier(int32_t* const* peer_flags, int32_t* local_epoch, int32_t rank, int32_t R,
                                   int32_t* err_flag, void* stream) {
  MMLREC_CHECK_ARG(peer_flags && local_epoch && R > 0 && R <= 32 && rank >= 0 && rank < R, "bad args");
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(peer_flags, local_epoch, rank, R, err_flag);
  MMLREC_RETURN_LAUNCH(1);
}
