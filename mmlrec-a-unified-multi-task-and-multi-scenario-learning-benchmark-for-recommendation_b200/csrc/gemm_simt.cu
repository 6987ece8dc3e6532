// K3, fp32 parity mode: grouped GEMM on the CUDA cores (FFMA).  One launch executes a table of
// independent problems C = act(A*B^T + bias) [mask] [+= C] with arbitrary element strides, so the
// same kernel serves forward, dgrad (B read through transposed strides) and wgrad (A, B read
// through transposed strides; the bias gradient falls out as the row sums of A).
// This is the mode whose results must match the reference to 1e-5; the throughput mode is
// gemm_tc.cu (bf16 tcgen05).
#include "common.cuh"

namespace mmlrec {

constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
constexpr int kGemmThreads = 256;

__global__ void __launch_bounds__(kGemmThreads)
gemm_grouped_f32_kernel(const MmlrecGemmF32* __restrict__ problems, const int32_t* __restrict__ tile_prefix,
                        int n_problems) {
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  __shared__ MmlrecGemmF32 P;
  __shared__ int s_tile;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int t = blockIdx.x, pi = 0;
    while (pi + 1 < n_problems && tile_prefix[pi + 1] <= t) ++pi;
    P = problems[pi];
    s_tile = t - tile_prefix[pi];
  }
  __syncthreads();
  const int tiles_n = (P.N + BN - 1) / BN;
  const int tm = s_tile / tiles_n, tn = s_tile - tm * tiles_n;
  const int m0 = tm * BM, n0 = tn * BN;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float rowsum = 0.f;
  const bool do_rowsum = (P.rowsum_a != nullptr) && tn == 0;
  const bool a_k_fast = (P.a_cs == 1), b_k_fast = (P.b_cs == 1);

  for (int k0 = 0; k0 < P.K; k0 += BK) {
#pragma unroll
    for (int l = 0; l < (BM * BK) / kGemmThreads; ++l) {
      int idx = tid + l * kGemmThreads;
      int m, k;
      if (a_k_fast) { k = idx & (BK - 1); m = idx >> 4; } else { m = idx & (BM - 1); k = idx >> 6; }
      int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < P.M && gk < P.K) ? __ldg(P.A + gm * P.a_rs + gk * P.a_cs) : 0.f;
    }
#pragma unroll
    for (int l = 0; l < (BN * BK) / kGemmThreads; ++l) {
      int idx = tid + l * kGemmThreads;
      int n, k;
      if (b_k_fast) { k = idx & (BK - 1); n = idx >> 4; } else { n = idx & (BN - 1); k = idx >> 6; }
      int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < P.N && gk < P.K) ? __ldg(P.B + gn * P.b_rs + gk * P.b_cs) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (do_rowsum && tid < BM) {
#pragma unroll
      for (int k = 0; k < BK; ++k) rowsum += As[k][tid];
    }
    __syncthreads();
  }
  if (do_rowsum && tid < BM && m0 + tid < P.M) P.rowsum_a[m0 + tid] = rowsum;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int gm = m0 + ty * 4 + i;
    if (gm >= P.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int gn = n0 + tx * 4 + j;
      if (gn >= P.N) continue;
      float v = acc[i][j];
      if (P.bias) v += __ldg(P.bias + gn);
      v = apply_act(v, P.act);
      if (P.mask && !(__ldg(P.mask + gm * P.ldmask + gn) > 0.f)) v = 0.f;
      float* c = P.C + gm * P.ldc + gn;
      if (P.accumulate) v += *c;
      *c = v;
    }
  }
}

}  // namespace mmlrec

extern "C" int mmlrec_gemm_grouped_f32(const MmlrecGemmF32* problems, const int32_t* tile_prefix, int32_t n_problems,
                                       int32_t total_tiles, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(n_problems > 0 && total_tiles >= 0, "bad sizes");
  if (total_tiles == 0) return 0;
  gemm_grouped_f32_kernel<<<total_tiles, kGemmThreads, 0, (cudaStream_t)stream>>>(problems, tile_prefix, n_problems);
  MMLREC_RETURN_LAUNCH(1);
}
