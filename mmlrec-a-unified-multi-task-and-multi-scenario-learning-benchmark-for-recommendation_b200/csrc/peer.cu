// Row-sharded embedding tables over NVLink peer memory (BASELINE config 5).
//
// owner(id) = id mod R; the owner stores row id at local row id / R of its shard.  Every shard and
// every receive buffer is a cudaMalloc allocation exported with CUDA IPC, so a kernel on rank r can
// address rank o's memory directly through NVSwitch:
//   forward   K1 reads each row straight from its owner's shard (gather.cu, `shards` path): the
//             id / row all-to-all of a two-sided design collapses into 16-byte peer loads;
//   backward  the push kernel below writes every (key, gradient row) of the local batch into the
//             OWNER's receive buffer at a slot fixed by (rank, sample) -- only owned entries cross
//             the wire (all-to-all volume), nothing is atomically appended, so the owner's sort +
//             segmented reduce (K2) sees a deterministic order;
//   barrier   the dense-gradient all-reduce that follows the push orders "all pushes done" before
//             "owner sorts"; receive buffers are double-buffered by step parity so a fast rank's next
//             push never lands in a buffer a slow owner is still reading.
#include "common.cuh"

namespace mmlrec {

// rx_keys of one owner: [2 parities][F_s][B_all] uint64, sentinel ~0 (nothing received for that slot)
// rx_grad of one owner: [2 parities][B_all][F_s*D] float
struct PushArgs {
  const float* X; int64_t ldx; int b;
  const float* d_input; int64_t ld;
  const int64_t* field_meta; int F_s; int D;
  int rank; int R; int B_all;
  uint64_t* const* rx_keys; float* const* rx_grad;
  const MmlrecHyper* hyper;
};

__global__ void __launch_bounds__(256) emb_push_rows_kernel(const PushArgs a) {
  const int par = a.hyper->step & 1;
  const int dv = a.D >> 2;
  const int64_t n = (int64_t)a.b * a.F_s * dv;
  const int64_t row_w = (int64_t)a.F_s * a.D;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
    const int part = (int)(t % dv);
    const int64_t ij = t / dv;
    const int j = (int)(ij % a.F_s);
    const int i = (int)(ij / a.F_s);
    const int64_t* m = a.field_meta + j * 4;
    int64_t id = (int64_t)__ldg(a.X + (int64_t)i * a.ldx + (int)m[2]);
    id = id < 0 ? 0 : (id >= m[1] ? m[1] - 1 : id);   // same clamp as the gather
    const int o = (int)(id % a.R);
    const int64_t pos = (int64_t)a.rank * a.b + i;
    const float4 g = __ldg(reinterpret_cast<const float4*>(a.d_input + (int64_t)i * a.ld + (int)m[3]) + part);
    float* dst = a.rx_grad[o] + ((int64_t)par * a.B_all + pos) * row_w + (int64_t)j * a.D;
    reinterpret_cast<float4*>(dst)[part] = g;
    if (part == 0)
      a.rx_keys[o][((int64_t)par * a.F_s + j) * a.B_all + pos] = ((uint64_t)(uint32_t)(id / a.R) << 32) | (uint32_t)pos;
  }
}

}  // namespace mmlrec

using namespace mmlrec;

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

extern "C" int mmlrec_emb_push_rows(const float* X, int64_t ldx, int32_t b, const float* d_input, int64_t ld,
                                    const int64_t* field_meta, int32_t F_s, int32_t D, int32_t rank, int32_t R,
                                    int32_t B_all, uint64_t* const* rx_keys, float* const* rx_grad,
                                    const MmlrecHyper* hyper, void* stream) {
  MMLREC_CHECK_ARG(X && d_input && field_meta && rx_keys && rx_grad && hyper, "null argument");
  MMLREC_CHECK_ARG(b > 0 && F_s > 0 && D > 0 && (D & 3) == 0 && (ld & 3) == 0, "bad sizes");
  MMLREC_CHECK_ARG(R > 0 && rank >= 0 && rank < R && B_all >= (int64_t)R * b, "bad rank / world / B_all");
  PushArgs a{X, ldx, b, d_input, ld, field_meta, F_s, D, rank, R, B_all, rx_keys, rx_grad, hyper};
  const int64_t n = (int64_t)b * F_s * (D >> 2);
  const int grid = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  emb_push_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a);
  MMLREC_RETURN_LAUNCH(1);
}
