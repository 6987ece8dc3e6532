// Fused non-GEMM stages of the step: BatchNorm(+activation), gate softmax + expert mixture,
// prediction heads + BCE (forward and backward in one pass), the flat dense optimizer and a few
// element-wise utilities.  All of these are HBM/L2-bound streaming kernels; cross-CTA reductions
// use per-CTA partials + a "last CTA reduces in fixed order" epilogue (deterministic, no float
// atomics).
#include <cstdlib>

#include "common.cuh"

namespace mmlrec {

// last-CTA-done ticket: returns true in exactly one CTA (the last to arrive), after which all
// other CTAs' global writes are visible to it.  The counter is reset for the next launch.
__device__ __forceinline__ bool last_block_ticket(int32_t* counter, int n_blocks) {
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = atomicAdd(counter, 1);
    s_last = (t == n_blocks - 1) ? 1 : 0;
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d (training: batch statistics, two-pass variance; eval: running statistics)
// grid = ceil(N/32); block = 256 = 8 row-groups x 32 columns
// ------------------------------------------------------------------------------------------------
constexpr int kBnThreads = 256;

__device__ __forceinline__ float bn_block_colsum(float v, float (*red)[33]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();
  red[w][lane] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += red[i][lane];
  return s;
}

__global__ void __launch_bounds__(kBnThreads)
bn_forward_kernel(const float* Z, int64_t ldz, int M, int N, const float* gamma, const float* beta,
                  float* running_mean, float* running_var, int64_t* nbt, int n_tracked,
                  float* save_mean, float* save_invstd, float* Y, int64_t ldy, uint16_t* Yb, int64_t ldyb,
                  int act, int training) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const bool cv = col < N;
  const float momentum = 0.1f, eps = 1e-5f;
  float mean, invstd;
  if (training == 2) {   // synchronised statistics (bn_combine_kernel wrote them): only normalise
    mean = cv ? save_mean[col] : 0.f;
    invstd = cv ? save_invstd[col] : 0.f;
  } else if (training) {
    float s = 0.f;
    if (cv) for (int r = w; r < M; r += 8) s += Z[r * ldz + col];
    s = bn_block_colsum(s, red);
    mean = s / (float)M;
    float q = 0.f;
    if (cv) for (int r = w; r < M; r += 8) { float d = Z[r * ldz + col] - mean; q += d * d; }
    q = bn_block_colsum(q, red);
    invstd = 1.f / sqrtf(q / (float)M + eps);
    if (cv && w == 0) {
      save_mean[col] = mean;
      save_invstd[col] = invstd;
      float unbiased = M > 1 ? q / (float)(M - 1) : q;
      running_mean[col] = (1.f - momentum) * running_mean[col] + momentum * mean;
      running_var[col] = (1.f - momentum) * running_var[col] + momentum * unbiased;
    }
    if (blockIdx.x == 0 && nbt && threadIdx.x < n_tracked) nbt[threadIdx.x] += 1;
  } else {
    mean = cv ? running_mean[col] : 0.f;
    invstd = cv ? 1.f / sqrtf(running_var[col] + eps) : 0.f;
  }
  if (!cv) return;
  const float g = gamma[col], b = beta[col];
  for (int r = w; r < M; r += 8) {
    float y = apply_act((Z[r * ldz + col] - mean) * invstd * g + b, act);
    if (Y) Y[r * ldy + col] = y;
    if (Yb) Yb[r * ldyb + col] = float_to_bf16_bits(y);
  }
}

__global__ void __launch_bounds__(kBnThreads)
bn_backward_kernel(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int M, int N, const float* gamma,
                   const float* save_mean, const float* save_invstd, float* dZ, int64_t lddz,
                   uint16_t* dZb, int64_t lddzb, float* dgamma, float* dbeta, const float* ext_sums, int M_total) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const bool cv = col < N;
  const float mean = cv ? save_mean[col] : 0.f, invstd = cv ? save_invstd[col] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (ext_sums != nullptr) {   // synchronised: sums over the GLOBAL batch (all-reduced bn_backward_sums_kernel output)
    s1 = cv ? ext_sums[col] : 0.f;
    s2 = cv ? ext_sums[N + col] : 0.f;
  } else {
    if (cv) for (int r = w; r < M; r += 8) {
      float dy = dY[r * lddy + col];
      s1 += dy;
      s2 += dy * (Z[r * ldz + col] - mean) * invstd;
    }
    s1 = bn_block_colsum(s1, red);
    s2 = bn_block_colsum(s2, red);
  }
  if (!cv) return;
  if (w == 0 && ext_sums == nullptr) { dgamma[col] = s2; dbeta[col] = s1; }
  const float g = gamma[col], inv_m = 1.f / (float)M_total;
  for (int r = w; r < M; r += 8) {
    float xh = (Z[r * ldz + col] - mean) * invstd;
    float dz = g * invstd * (dY[r * lddy + col] - s1 * inv_m - xh * s2 * inv_m);
    if (dZ) dZ[r * lddz + col] = dz;
    if (dZb) dZb[r * lddzb + col] = float_to_bf16_bits(dz);
  }
}

// Synchronised BatchNorm under data parallelism (the global batch is what the single-process reference normalises over).
// forward : bn_stats_kernel (per rank: mean_r, sum (x - mean_r)^2) -> all-gather -> bn_combine_kernel (Chan's pairwise
//           formula over equal-sized ranks, running statistics with the global count) -> bn_forward_kernel(training = 2)
// backward: bn_backward_sums_kernel (per rank: sum dy, sum dy * xhat; also THIS rank's dgamma / dbeta partials, which the
//           dense-gradient all-reduce adds up) -> all-reduce -> bn_backward_kernel(ext_sums, M_total)
__global__ void __launch_bounds__(kBnThreads) bn_stats_kernel(const float* Z, int64_t ldz, int M, int N, float* stats) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const bool cv = col < N;
  float s = 0.f;
  if (cv) for (int r = w; r < M; r += 8) s += Z[r * ldz + col];
  s = bn_block_colsum(s, red);
  const float mean = s / (float)M;
  float q = 0.f;
  if (cv) for (int r = w; r < M; r += 8) { float d = Z[r * ldz + col] - mean; q += d * d; }
  q = bn_block_colsum(q, red);
  if (cv && w == 0) { stats[col] = mean; stats[N + col] = q; }
}

__global__ void bn_combine_kernel(const float* all_stats, int R, int M, int N, float* running_mean, float* running_var,
                                  int64_t* nbt, int n_tracked, float* save_mean, float* save_invstd) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const float momentum = 0.1f, eps = 1e-5f;
  if (blockIdx.x == 0 && nbt && (int)threadIdx.x < n_tracked) nbt[threadIdx.x] += 1;
  if (col >= N) return;
  float mean = 0.f;
  for (int r = 0; r < R; ++r) mean += all_stats[(int64_t)r * 2 * N + col];
  mean /= (float)R;
  float m2 = 0.f;
  for (int r = 0; r < R; ++r) {
    const float d = all_stats[(int64_t)r * 2 * N + col] - mean;
    m2 += all_stats[(int64_t)r * 2 * N + N + col] + (float)M * d * d;
  }
  const float total = (float)M * (float)R;
  save_mean[col] = mean;
  save_invstd[col] = 1.f / sqrtf(m2 / total + eps);
  running_mean[col] = (1.f - momentum) * running_mean[col] + momentum * mean;
  running_var[col] = (1.f - momentum) * running_var[col] + momentum * (total > 1.f ? m2 / (total - 1.f) : m2);
}

__global__ void __launch_bounds__(kBnThreads)
bn_backward_sums_kernel(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int M, int N, const float* save_mean,
                        const float* save_invstd, float* sums, float* dgamma, float* dbeta) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + lane;
  const bool cv = col < N;
  const float mean = cv ? save_mean[col] : 0.f, invstd = cv ? save_invstd[col] : 0.f;
  float s1 = 0.f, s2 = 0.f;
  if (cv) for (int r = w; r < M; r += 8) {
    float dy = dY[r * lddy + col];
    s1 += dy;
    s2 += dy * (Z[r * ldz + col] - mean) * invstd;
  }
  s1 = bn_block_colsum(s1, red);
  s2 = bn_block_colsum(s2, red);
  if (cv && w == 0) { sums[col] = s1; sums[N + col] = s2; dgamma[col] = s2; dbeta[col] = s1; }
}

// ------------------------------------------------------------------------------------------------
// gate head + softmax + mixture, forward.  grid = (ceil(B/8), n_gates), one warp per sample.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gate_mix_forward_kernel(const MmlrecGate* gates, int B) {
  __shared__ MmlrecGate G;
  if (threadIdx.x == 0) G = gates[blockIdx.y];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + w;
  if (b >= B) return;
  const float* gin = G.gate_in + (int64_t)b * G.ld_gate_in;
  float logit = -INFINITY;  // lane e holds logit e
  for (int e = 0; e < G.n_e; ++e) {
    const float* wr = G.Wg + (int64_t)e * G.ld_Wg;
    float s = 0.f;
    for (int h = lane; h < G.Hg; h += 32) s = fmaf(gin[h], __ldg(wr + h), s);
    s = warp_sum(s);
    if (lane == e) logit = s;
  }
  float mx = logit;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float ex = lane < G.n_e ? expf(logit - mx) : 0.f;
  float den = warp_sum(ex);
  float p = ex / den;
  if (lane < G.n_e) G.probs[(int64_t)b * G.n_e + lane] = p;
  // uniform trip count: every lane reaches the shuffles, only loads / stores are guarded
  for (int h0 = 0; h0 < G.H; h0 += 32) {
    const int h = h0 + lane;
    const bool hv = h < G.H;
    float acc = 0.f;
    for (int e = 0; e < G.n_e; ++e) {
      float pe = __shfl_sync(0xffffffffu, p, e);
      if (hv) acc = fmaf(pe, G.expert[e][(int64_t)b * G.ld_expert + h], acc);
    }
    if (hv) {
      G.mix[(int64_t)b * G.ld_mix + h] = acc;
      if (G.mix_bf16) G.mix_bf16[(int64_t)b * G.ld_mix_bf16 + h] = float_to_bf16_bits(acc);
    }
  }
}

// backward A: per gate.  grid = (ceil(B/64), n_gates); 8 warps x 8 samples; smem holds the gate
// inputs and dlogits of the 64 samples for the deterministic dWg partial.
constexpr int kGateRows = 64;
__global__ void __launch_bounds__(256)
gate_mix_backward_gate_kernel(const MmlrecGate* gates, int n_gates, int B, float* scratch, int64_t per_gate_scratch,
                              int32_t* counters) {
  extern __shared__ __align__(16) float smem_f[];
  __shared__ MmlrecGate G;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int r0 = blockIdx.x * kGateRows;
  const int rows = min(kGateRows, B - r0);
  for (int gi = blockIdx.y; gi < n_gates; gi += gridDim.y) {
    __syncthreads();
    if (threadIdx.x == 0) G = gates[gi];
    __syncthreads();
    if (G.d_mix == nullptr) continue;
    float* gin_s = smem_f;                         // [64][Hg]
    float* dl_s = smem_f + kGateRows * G.Hg;       // [64][n_e]
    for (int i = threadIdx.x; i < kGateRows * G.Hg; i += 256) {
      int r = i / G.Hg, h = i - r * G.Hg;
      gin_s[i] = r < rows ? G.gate_in[(int64_t)(r0 + r) * G.ld_gate_in + h] : 0.f;
    }
    __syncthreads();
    for (int rr = 0; rr < 8; ++rr) {
      const int r = w * 8 + rr;
      const int b = r0 + r;
      float dl = 0.f;
      if (r < rows) {
        const float* dm = G.d_mix + (int64_t)b * G.ld_d_mix;
        float dp = 0.f;
        for (int e = 0; e < G.n_e; ++e) {
          const float* eo = G.expert[e] + (int64_t)b * G.ld_expert;
          float s = 0.f;
          for (int h = lane; h < G.H; h += 32) s = fmaf(dm[h], eo[h], s);
          s = warp_sum(s);
          if (lane == e) dp = s;
        }
        float p = lane < G.n_e ? G.probs[(int64_t)b * G.n_e + lane] : 0.f;
        float dot = warp_sum(p * dp);
        dl = p * (dp - dot);
        // d(gate_in)
        for (int h0 = 0; h0 < G.Hg; h0 += 32) {  // uniform trip count (shuffles inside)
          const int h = h0 + lane;
          const bool hv = h < G.Hg;
          float acc = 0.f;
          for (int e = 0; e < G.n_e; ++e) {
            const float de = __shfl_sync(0xffffffffu, dl, e);
            if (hv) acc = fmaf(de, __ldg(G.Wg + (int64_t)e * G.ld_Wg + h), acc);
          }
          if (hv) {
            if (G.relu_mask_gate_in && !(gin_s[r * G.Hg + h] > 0.f)) acc = 0.f;
            if (G.d_gate_in) {
              float* dst = G.d_gate_in + (int64_t)b * G.ld_d_gate_in + h;
              if (G.accumulate_d_gate_in) acc += *dst;
              *dst = acc;
            }
            if (G.d_gate_in_bf16) G.d_gate_in_bf16[(int64_t)b * G.ld_d_gate_in_bf16 + h] = float_to_bf16_bits(acc);
          }
        }
      }
      if (lane < G.n_e) dl_s[r * G.n_e + lane] = dl;
    }
    __syncthreads();
    // per-CTA partial of dWg[e][h] = sum_r dl[r][e] * gin[r][h]
    float* part = scratch + (int64_t)gi * per_gate_scratch + (int64_t)blockIdx.x * G.n_e * G.Hg;
    for (int i = threadIdx.x; i < G.n_e * G.Hg; i += 256) {
      int e = i / G.Hg, h = i - e * G.Hg;
      float acc = 0.f;
      for (int r = 0; r < kGateRows; ++r) acc = fmaf(dl_s[r * G.n_e + e], gin_s[r * G.Hg + h], acc);
      part[i] = acc;
    }
    if (last_block_ticket(counters + gi, gridDim.x)) {
      const float* base = scratch + (int64_t)gi * per_gate_scratch;
      for (int i = threadIdx.x; i < G.n_e * G.Hg; i += 256) {
        float acc = 0.f;
        for (int c = 0; c < (int)gridDim.x; ++c) acc += base[(int64_t)c * G.n_e * G.Hg + i];
        G.dWg[(int64_t)(i / G.Hg) * G.ld_Wg + (i % G.Hg)] = acc;
      }
    }
  }
}

// backward B: per expert.  grid = (ceil(B/8), n_experts), one warp per sample.
__global__ void __launch_bounds__(256) gate_mix_backward_expert_kernel(const MmlrecExpertGrad* experts, int B) {
  __shared__ MmlrecExpertGrad E;
  if (threadIdx.x == 0) E = experts[blockIdx.y];
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * 8 + w;
  if (b >= B) return;
  float pu = 0.f;  // lane u holds the probability user u assigns to this expert
  if (lane < E.n_users) pu = E.user_probs[lane][(int64_t)b * E.user_prob_ld[lane] + E.user_prob_col[lane]];
  for (int h0 = 0; h0 < E.H; h0 += 32) {  // uniform trip count (shuffles inside)
    const int h = h0 + lane;
    const bool hv = h < E.H;
    float acc = 0.f;
    for (int u = 0; u < E.n_users; ++u) {
      const float pw = __shfl_sync(0xffffffffu, pu, u);
      if (hv) acc = fmaf(pw, E.user_d_mix[u][(int64_t)b * E.user_d_mix_ld[u] + h], acc);
    }
    if (hv) {
      if (E.relu_mask && !(E.expert[(int64_t)b * E.ld_expert + h] > 0.f)) acc = 0.f;
      if (E.d_expert) E.d_expert[(int64_t)b * E.ld_d_expert + h] = acc;
      if (E.d_expert_bf16) E.d_expert_bf16[(int64_t)b * E.ld_d_expert_bf16 + h] = float_to_bf16_bits(acc);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// heads + loss (forward + backward).  grid = ceil(B/16); 8 warps x 2 samples; a warp owns a sample:
// per task a coalesced row load + warp reduction gives the logit (lane t keeps task t), lanes 0..T-1
// evaluate sigmoid / BCE / dz, then the row is revisited for d_h and the per-warp dw slab.  CTA
// partials (loss, dbias, dw) go to scratch; the last CTA reduces them in a fixed order.
// ------------------------------------------------------------------------------------------------
constexpr int kHeadRows = 16;
constexpr int kHeadMaxK = 8;  // H <= 256

__global__ void __launch_bounds__(256)
heads_kernel(const MmlrecHead* heads, int T, int B, const float* y, int64_t ldy, float* pred, int64_t ld_pred,
             float* loss, int esmm, int training, float* scratch, int stride_cta, int32_t* counter, int grad_mode,
             const float* mask, int64_t ldm) {
  pdl_prologue();
  extern __shared__ __align__(16) float dw_s[];            // [8 warps][T][hmax]
  __shared__ MmlrecHead Hd[MMLREC_MAX_TASKS];
  __shared__ float part_s[8][MMLREC_MAX_TASKS][2];         // per warp: loss, dbias partials
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < T) Hd[tid] = heads[tid];
  __syncthreads();
  const bool bwd = training && y != nullptr;
  const bool cum_bias = (esmm & 2) != 0;   // MMLREC_HEADS_CUMULATIVE_BIAS
  esmm &= 1;
  const int hmax = bwd ? (stride_cta / T) - 2 : 0;
  float* my_dw = dw_s + (size_t)w * T * hmax;
  if (bwd) for (int i = lane; i < T * hmax; i += 32) my_dw[i] = 0.f;
  float loss_acc = 0.f, dz_acc = 0.f;                       // lane t accumulates task t over this warp's rows
  for (int rr = 0; rr < 2; ++rr) {
    const int b = blockIdx.x * kHeadRows + w * 2 + rr;
    if (b >= B) break;  // warp-uniform
    // logits: lane t keeps z_t
    float z = 0.f;
    for (int t = 0; t < T; ++t) {
      const float* hrow = Hd[t].h + (int64_t)b * Hd[t].ld_h;
      float s = 0.f;
      for (int h = lane; h < Hd[t].H; h += 32) s = fmaf(hrow[h], __ldg(Hd[t].w + h), s);
      s = warp_sum(s);
      if (lane == t) {
        z = s + (Hd[t].bias ? *Hd[t].bias : 0.f) + (Hd[t].bias2 ? *Hd[t].bias2 : 0.f);
        if (cum_bias) for (int q = 0; q < t; ++q) z += Hd[q].bias ? *Hd[q].bias : 0.f;
      }
    }
    // probabilities, loss terms, dz (lanes < T)
    const float z0 = __shfl_sync(0xffffffffu, z, 0);
    float dz = 0.f, l = 0.f, cross = 0.f;
    if (lane < T) {
      const int t = lane;
      if (Hd[t].kind == MMLREC_HEAD_SIGMOID_BCE) {
        const float p = 1.f / (1.f + expf(-z));
        float out = p, scale = 1.f;  // scale = d(out)/dp
        if (esmm && t == 1) {        // esmm.py:59  ctcvr = ctr * cvr
          const float p0 = 1.f / (1.f + expf(-z0));
          out = p0 * p;
          scale = p0;
        }
        // scenario mask (mmoe.py:101-106): the prediction is multiplied by the sample's domain-mask entry of this head,
        // the loss is BCE(masked prediction, y, weight = mask) (basemodel.py:273-282)
        const float mk = mask ? mask[(int64_t)b * ldm + Hd[t].mask_col] : 1.f;
        out *= mk;
        scale *= mk;
        pred[(int64_t)b * ld_pred + t] = out;
        if (y) {
          const float yy = y[(int64_t)b * ldy + t];
          float gout;
          if (grad_mode) {   // y carries dL/d(pred) from autograd (differentiable forward()): no loss is formed here
            gout = yy;
          } else {
            // F.binary_cross_entropy: (y-1)*max(log1p(-x),-100) - y*max(log(x),-100)
            l = mk * ((yy - 1.f) * fmaxf(log1pf(-out), -100.f) - yy * fmaxf(logf(out), -100.f));
            // binary_cross_entropy_backward: (x-y)/max((1-x)*x, 1e-12); sigmoid_backward: g*(1-p)*p
            gout = mk * (out - yy) / fmaxf((1.f - out) * out, 1e-12f);
          }
          dz = gout * scale * (1.f - p) * p;
          if (esmm && t == 1) cross = gout * p;  // d loss_1 / d p0
        }
      } else {  // identity + MSE (F.mse_loss, reduction='sum')
        pred[(int64_t)b * ld_pred + t] = z;
        if (y) {
          if (grad_mode) { dz = y[(int64_t)b * ldy + t]; }
          else {
            const float d = z - y[(int64_t)b * ldy + t];
            l = d * d;
            dz = 2.f * d;
          }
        }
      }
    }
    if (esmm) {  // head 0 also receives gradient through out1 = p0*p1
      const float c1 = __shfl_sync(0xffffffffu, cross, 1);
      if (lane == 0 && y) {
        const float p0 = 1.f / (1.f + expf(-z));
        dz += c1 * (1.f - p0) * p0;
      }
    }
    if (!bwd) continue;
    loss_acc += l;
    dz_acc += dz;
    // d_h and the dw slab
    for (int t = 0; t < T; ++t) {
      const MmlrecHead& hd = Hd[t];
      const float dzt = __shfl_sync(0xffffffffu, dz, t);
      const float* hrow = hd.h + (int64_t)b * hd.ld_h;
      for (int h = lane; h < hd.H; h += 32) {
        const float hv = hrow[h];
        my_dw[t * hmax + h] = fmaf(dzt, hv, my_dw[t * hmax + h]);
        float g = dzt * __ldg(hd.w + h);
        if (hd.relu_mask && !(hv > 0.f)) g = 0.f;
        if (hd.d_h) hd.d_h[(int64_t)b * hd.ld_d_h + h] = g;
        if (hd.d_h_bf16) hd.d_h_bf16[(int64_t)b * hd.ld_d_h_bf16 + h] = float_to_bf16_bits(g);
      }
    }
  }
  if (!bwd) return;
  if (lane < T) { part_s[w][lane][0] = loss_acc; part_s[w][lane][1] = dz_acc; }
  __syncthreads();
  // CTA partial -> scratch [cta][T][2 + hmax]  (warps summed in fixed order)
  float* cta_out = scratch + (int64_t)blockIdx.x * stride_cta;
  const int per_t = 2 + hmax;
  for (int i = tid; i < T * per_t; i += 256) {
    const int t = i / per_t, k = i - t * per_t;
    float s = 0.f;
    if (k < 2) {
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) s += part_s[ww][t][k];
    } else {
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) s += dw_s[((size_t)ww * T + t) * hmax + (k - 2)];
    }
    cta_out[i] = s;
  }
}

// Deterministic reduction of the per-CTA partials [n_cta][T][2 + hmax]: 32 outputs x 8 partial-groups per CTA
// (each thread sums every 8th partial, the 8 group sums are added in a fixed order).  Outputs are re-indexed so
// that the 2T scalars (loss_t, dbias_t) are outputs 0..2T-1 and land in CTA 0, which also forms the total loss.
__global__ void __launch_bounds__(256)
heads_reduce_kernel(const MmlrecHead* heads, int T, float* loss, int esmm, const float* scratch, int stride_cta, int n_cta) {
  pdl_prologue();
  __shared__ float red[8][33];
  __shared__ float tot_s[MMLREC_MAX_TASKS][2];
  const int ix = threadIdx.x & 31, iy = threadIdx.x >> 5;
  const int per_t = stride_cta / T, hmax = per_t - 2;
  const int n_out = T * per_t;
  const int j = blockIdx.x * 32 + ix;
  int t = 0, k = 0;
  if (j < 2 * T) { t = j >> 1; k = j & 1; }
  else { const int jj = j - 2 * T; t = jj / hmax; k = 2 + (jj - t * hmax); }
  const int i = t * per_t + k;
  float s = 0.f;
  if (j < n_out) {   // 8 independent loads in flight per thread (the loop is L2-latency bound), fixed order
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
    int c = iy;
    for (; c + 56 < n_cta; c += 64) {
      const float* q = scratch + (int64_t)c * stride_cta + i;
      a0 += q[0]; a1 += q[(int64_t)8 * stride_cta]; a2 += q[(int64_t)16 * stride_cta]; a3 += q[(int64_t)24 * stride_cta];
      a4 += q[(int64_t)32 * stride_cta]; a5 += q[(int64_t)40 * stride_cta]; a6 += q[(int64_t)48 * stride_cta]; a7 += q[(int64_t)56 * stride_cta];
    }
    for (; c < n_cta; c += 8) a0 += scratch[(int64_t)c * stride_cta + i];
    s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
  }
  red[iy][ix] = s;
  __syncthreads();
  if (iy == 0 && j < n_out) {
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) v += red[q][ix];
    if (k < 2) tot_s[t][k] = v;
    else if (k - 2 < heads[t].H) heads[t].dw[k - 2] = v;
  }
  if (blockIdx.x != 0) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    float total = 0.f;
    for (int q = 0; q < T; ++q) { loss[q] = tot_s[q][0]; total += tot_s[q][0]; }
    loss[T] = total;
    if (esmm) {  // one shared bias receives both heads' gradient
      if (heads[0].dbias) *heads[0].dbias = tot_s[0][1] + tot_s[1][1];
    } else {
      for (int q = 0; q < T; ++q) if (heads[q].dbias) *heads[q].dbias = tot_s[q][1];
    }
    for (int q = 0; q < T; ++q) if (heads[q].dbias2) *heads[q].dbias2 = tot_s[q][1];
  }
}

// ------------------------------------------------------------------------------------------------
// heads + loss, one launch (the training path for T <= TM tasks of width <= 32 * KM).  A warp owns R samples and
// fetches ALL their tower rows first (R * T * KM independent loads per lane in flight -- the two-kernel version above
// walks samples and tasks one dependent L2 round trip at a time), keeps them in registers for the logits, d_h and the
// dw partial, and lane r * TM + t evaluates sigmoid / BCE / dz of (sample r, task t).  CTA partials go to scratch; the
// LAST CTA to finish (atomic ticket) adds them up in index order, so the result does not depend on which CTA that is.
// ------------------------------------------------------------------------------------------------
template <int TM, int KM, int R>
__global__ void __launch_bounds__(256)
heads_fast_kernel(const MmlrecHead* heads, int T, int B, const float* y, int64_t ldy, float* pred, int64_t ld_pred,
                  float* loss, int esmm, float* scratch, int stride_cta, int32_t* counter, int grad_mode,
                  const float* mask, int64_t ldm) {
  static_assert(R * TM <= 32, "one lane per (sample, task)");
  pdl_prologue();
  constexpr int PW = 2 + KM * 32;                          // per warp and task: loss, dz, dw[KM * 32]
  __shared__ MmlrecHead Hd[TM];
  __shared__ float red_s[8][TM][PW];
  __shared__ float tot_s[TM][2];
  __shared__ int last_s;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  if (tid < T) Hd[tid] = heads[tid];
  __syncthreads();
  const int row0 = (blockIdx.x * 8 + w) * R;
  float hv[R][TM][KM], wv[TM][KM];
#pragma unroll
  for (int t = 0; t < TM; ++t)
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const int h = lane + 32 * k;
      const bool ok = t < T && h < Hd[t].H;
      wv[t][k] = ok ? __ldg(Hd[t].w + h) : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r)
        hv[r][t][k] = (ok && row0 + r < B) ? Hd[t].h[(int64_t)(row0 + r) * Hd[t].ld_h + h] : 0.f;
    }
  // logits: after the butterfly every lane holds every sum; lane r * TM + t keeps z of (sample r, task t)
  const int my_r = lane / TM, my_t = lane - my_r * TM;
  const int my_b = row0 + my_r;
  const bool mine = my_r < R && my_t < T && my_b < B;
  float z = 0.f;
#pragma unroll
  for (int r = 0; r < R; ++r)
#pragma unroll
    for (int t = 0; t < TM; ++t) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < KM; ++k) s = fmaf(hv[r][t][k], wv[t][k], s);
      s = warp_sum(s);
      if (lane == r * TM + t) z = s;
    }
  const bool cum_bias = (esmm & 2) != 0;   // MMLREC_HEADS_CUMULATIVE_BIAS
  const bool shared_bias = (esmm & 4) != 0;   // one bias parameter behind every head: its gradient is the sum
  esmm &= 1;
  if (mine) {
    z += (Hd[my_t].bias ? *Hd[my_t].bias : 0.f) + (Hd[my_t].bias2 ? *Hd[my_t].bias2 : 0.f);
    if (cum_bias) for (int q = 0; q < my_t; ++q) z += Hd[q].bias ? *Hd[q].bias : 0.f;
  }
  const float z0 = __shfl_sync(0xffffffffu, z, my_r * TM < 32 ? my_r * TM : 0);   // task 0 of the same sample (esmm)
  float dz = 0.f, l = 0.f, cross = 0.f;
  if (mine) {
    if (Hd[my_t].kind == MMLREC_HEAD_SIGMOID_BCE) {
      const float p = 1.f / (1.f + expf(-z));
      float out = p, scale = 1.f;
      if (esmm && my_t == 1) {                               // esmm.py:59  ctcvr = ctr * cvr
        const float p0 = 1.f / (1.f + expf(-z0));
        out = p0 * p;
        scale = p0;
      }
      const float mk = mask ? mask[(int64_t)my_b * ldm + Hd[my_t].mask_col] : 1.f;   // scenario mask, see heads_kernel
      out *= mk;
      scale *= mk;
      pred[(int64_t)my_b * ld_pred + my_t] = out;
      const float yy = y[(int64_t)my_b * ldy + my_t];
      float gout;
      if (grad_mode) {
        gout = yy;
      } else {   // same formulas as heads_kernel (F.binary_cross_entropy forward / backward, sigmoid_backward)
        l = mk * ((yy - 1.f) * fmaxf(log1pf(-out), -100.f) - yy * fmaxf(logf(out), -100.f));
        gout = mk * (out - yy) / fmaxf((1.f - out) * out, 1e-12f);
      }
      dz = gout * scale * (1.f - p) * p;
      if (esmm && my_t == 1) cross = gout * p;
    } else {
      pred[(int64_t)my_b * ld_pred + my_t] = z;
      if (grad_mode) { dz = y[(int64_t)my_b * ldy + my_t]; }
      else { const float d = z - y[(int64_t)my_b * ldy + my_t]; l = d * d; dz = 2.f * d; }
    }
  }
  if (esmm) {   // head 0 also receives gradient through out1 = p0 * p1
    const float c1 = __shfl_sync(0xffffffffu, cross, (lane + 1) & 31);
    if (mine && my_t == 0) {
      const float p0 = 1.f / (1.f + expf(-z));
      dz += c1 * (1.f - p0) * p0;
    }
  }
  // this warp's partials (loss, dz, dw) and the d_h rows.  Heads that read the SAME tower output (MLP: one shared
  // stack, sharedbottom without towers) share its gradient buffer: the first head of such a group writes the sum.
  float dza[R][TM];
#pragma unroll
  for (int t = 0; t < TM; ++t)
#pragma unroll
    for (int r = 0; r < R; ++r) dza[r][t] = __shfl_sync(0xffffffffu, dz, r * TM + t);
#pragma unroll
  for (int t = 0; t < TM; ++t) {
    float ls = 0.f, ds = 0.f;
#pragma unroll
    for (int r = 0; r < R; ++r) { ls += __shfl_sync(0xffffffffu, l, r * TM + t); ds += dza[r][t]; }
    if (t >= T) continue;   // warp-uniform
    const MmlrecHead& hd = Hd[t];
    if (lane == 0) { red_s[w][t][0] = ls; red_s[w][t][1] = ds; }
    bool first = true;
#pragma unroll
    for (int q = 0; q < TM; ++q) if (q < t && Hd[q].h == hd.h) first = false;
#pragma unroll
    for (int k = 0; k < KM; ++k) {
      const int h = lane + 32 * k;
      float dw = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) dw = fmaf(dza[r][t], hv[r][t][k], dw);
      red_s[w][t][2 + h] = dw;
      if (h >= hd.H || !first) continue;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (row0 + r >= B) break;
        float g = dza[r][t] * wv[t][k];
#pragma unroll
        for (int q = 0; q < TM; ++q) if (q > t && q < T && Hd[q].h == hd.h) g = fmaf(dza[r][q], wv[q][k], g);
        if (hd.relu_mask && !(hv[r][t][k] > 0.f)) g = 0.f;
        if (hd.d_h) hd.d_h[(int64_t)(row0 + r) * hd.ld_d_h + h] = g;
        if (hd.d_h_bf16) hd.d_h_bf16[(int64_t)(row0 + r) * hd.ld_d_h_bf16 + h] = float_to_bf16_bits(g);
      }
    }
  }
  __syncthreads();
  // CTA partial -> scratch [cta][T][2 + hmax] (warps summed in fixed order)
  const int per_t = stride_cta / T, n_out = T * per_t;
  float* cta_out = scratch + (int64_t)blockIdx.x * stride_cta;
  for (int i = tid; i < n_out; i += 256) {
    const int t = i / per_t, k = i - t * per_t;
    float s = 0.f;
    if (k < PW) {
#pragma unroll
      for (int ww = 0; ww < 8; ++ww) s += red_s[ww][t][k];
    }
    cta_out[i] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last_s = atomicAdd(counter, 1) == (int)gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!last_s) return;
  if (tid == 0) *counter = 0;                                // ready for the next launch / graph replay
  __threadfence();
  const int n_cta = (int)gridDim.x;
  for (int i = tid; i < n_out; i += 256) {   // (a 4-way split of the CTA range with 32 loads in flight measured slower)
    float a[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) a[q] = 0.f;
    int c = 0;
    for (; c + 16 <= n_cta; c += 16) {
#pragma unroll
      for (int q = 0; q < 16; ++q) a[q] += __ldcg(scratch + (int64_t)(c + q) * stride_cta + i);
    }
    for (; c < n_cta; ++c) a[0] += __ldcg(scratch + (int64_t)c * stride_cta + i);
    float v = 0.f;
#pragma unroll
    for (int q = 0; q < 16; ++q) v += a[q];
    const int t = i / per_t, k = i - t * per_t;
    if (k < 2) tot_s[t][k] = v;
    else if (k < PW) red_s[0][t][k] = v;                     // (the per-warp partials are consumed: reuse as dw[t][k])
  }
  __syncthreads();
  // heads that share one final layer (mlp.py:28: every task reads the same logit) share its gradient: the first head of
  // such a group writes the sum over the group, in task order
  for (int i = tid; i < T * KM * 32; i += 256) {
    const int t = i / (KM * 32), h = i - t * (KM * 32);
    if (h >= Hd[t].H) continue;
    bool first = true;
    for (int q = 0; q < t; ++q) first = first && Hd[q].dw != Hd[t].dw;
    if (!first) continue;
    float v = red_s[0][t][2 + h];
    for (int q = t + 1; q < T; ++q) if (Hd[q].dw == Hd[t].dw) v += red_s[0][q][2 + h];
    Hd[t].dw[h] = v;
  }
  if (tid == 0) {
    float total = 0.f;
    for (int q = 0; q < T; ++q) { loss[q] = tot_s[q][0]; total += tot_s[q][0]; }
    loss[T] = total;
    if (esmm || shared_bias) {
      float v = 0.f;
      for (int q = 0; q < T; ++q) v += tot_s[q][1];
      if (Hd[0].dbias) *Hd[0].dbias = v;
    } else if (cum_bias) {   // bias q enters the logits of tasks q, q+1, ...
      for (int q = 0; q < T; ++q) {
        float v = 0.f;
        for (int t = q; t < T; ++t) v += tot_s[t][1];
        if (Hd[q].dbias) *Hd[q].dbias = v;
      }
    } else {
      for (int q = 0; q < T; ++q) if (Hd[q].dbias) *Hd[q].dbias = tot_s[q][1];
    }
    for (int q = 0; q < T; ++q) if (Hd[q].dbias2) *Hd[q].dbias2 = tot_s[q][1];
  }
}

// ------------------------------------------------------------------------------------------------
// flat dense optimizer + element-wise utilities
// ------------------------------------------------------------------------------------------------
__global__ void dense_optimizer_kernel(float* p, const float* g, float* s1, float* s2, int64_t n,
                                       const MmlrecHyper* hyper, uint16_t* shadow) {
  const MmlrecHyper hp = *hyper;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float pv = p[i], a = s1 ? s1[i] : 0.f, b = s2 ? s2[i] : 0.f;
    optimizer_update(pv, g[i], a, b, hp);
    p[i] = pv;
    if (s1) s1[i] = a;
    if (s2) s2[i] = b;
    if (shadow) shadow[i] = float_to_bf16_bits(pv);
  }
}

// Same update over 4 elements per thread (128-bit accesses), with the gradient given as `n_slices` partial buffers
// `slice_stride` floats apart that are added in slice order: the split-K wgrad problems write their partial tiles
// straight into gradient slices 1..S-1 (slice 0 is the ordinary gradient buffer every other kernel writes), so the
// split needs no separate reduction pass and stays deterministic.
__global__ void __launch_bounds__(256)
dense_optimizer_sliced_kernel(float4* p, const float4* g, float4* s1, float4* s2, int64_t n4, const MmlrecHyper* hyper,
                              uint2* shadow, int n_slices, int64_t slice_stride4) {
  pdl_prologue();
  const MmlrecHyper hp = *hyper;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 gv = g[i];
    for (int k = 1; k < n_slices; ++k) {
      const float4 t = g[(int64_t)k * slice_stride4 + i];
      gv.x += t.x; gv.y += t.y; gv.z += t.z; gv.w += t.w;
    }
    float4 pv = p[i];
    float4 a = s1 ? s1[i] : make_float4(0.f, 0.f, 0.f, 0.f), b = s2 ? s2[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    optimizer_update(pv.x, gv.x, a.x, b.x, hp);
    optimizer_update(pv.y, gv.y, a.y, b.y, hp);
    optimizer_update(pv.z, gv.z, a.z, b.z, hp);
    optimizer_update(pv.w, gv.w, a.w, b.w, hp);
    p[i] = pv;
    if (s1) s1[i] = a;
    if (s2) s2[i] = b;
    if (shadow) shadow[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
  }
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  pdl_prologue();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

__global__ void cast_f32_bf16_kernel(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int rows, int cols,
                                     int cols_pad) {
  const int64_t n = (int64_t)rows * cols_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / cols_pad), c = (int)(i - (int64_t)r * cols_pad);
    dst[r * ld_dst + c] = c < cols ? float_to_bf16_bits(src[r * ld_src + c]) : (uint16_t)0;
  }
}

__global__ void cast_bf16_f32_kernel(const uint16_t* src, int64_t ld_src, float* dst, int64_t ld_dst, int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    dst[r * ld_dst + c] = bf16_bits_to_float(src[r * ld_src + c]);
  }
}

__global__ void mul_kernel(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo, int rows,
                           int cols, int accumulate) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    float v = a[r * lda + c] * b[r * ldb + c];
    if (accumulate) v += out[r * ldo + c];
    out[r * ldo + c] = v;
  }
}

__device__ __forceinline__ float apply_dkind(float g, float v, int dkind) {
  if (dkind == 1) return v > 0.f ? g : 0.f;
  if (dkind == 2) return g * v * (1.f - 0.5f * v);   // d/dz [2*sigmoid(z)] = y * (1 - y/2)
  return g;
}

__global__ void copy_cols_kernel(const float* src, int64_t ld_src, float* d32, int64_t ld32, uint16_t* d16, int64_t ld16,
                                 int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float v = src[r * ld_src + c];
    if (d32) d32[r * ld32 + c] = v;
    if (d16) d16[r * ld16 + c] = float_to_bf16_bits(v);
  }
}

__global__ void mul_forward_kernel(const float* a, int64_t lda, const float* b, int64_t ldb, float* o32, int64_t ld32,
                                   uint16_t* o16, int64_t ld16, int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float v = a[r * lda + c] * b[r * ldb + c];
    if (o32) o32[r * ld32 + c] = v;
    if (o16) o16[r * ld16 + c] = float_to_bf16_bits(v);
  }
}

__global__ void mul_backward_kernel(const float* d_out, int64_t ld_dout, const float* a, int64_t lda, const float* b,
                                    int64_t ldb, float* da32, uint16_t* da16, int64_t ld_da, int dkind_a, int acc_a,
                                    float* db32, uint16_t* db16, int64_t ld_db, int dkind_b, int acc_b, int rows, int cols) {
  const int64_t n = (int64_t)rows * cols;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i - (int64_t)r * cols);
    const float g = d_out[r * ld_dout + c], av = a[r * lda + c], bv = b[r * ldb + c];
    if (da32 || da16) {
      float v = apply_dkind(g * bv, av, dkind_a);
      if (da32) { if (acc_a) v += da32[r * ld_da + c]; da32[r * ld_da + c] = v; }
      if (da16) da16[r * ld_da + c] = float_to_bf16_bits(v);
    }
    if (db32 || db16) {
      float v = apply_dkind(g * av, bv, dkind_b);
      if (db32) { if (acc_b) v += db32[r * ld_db + c]; db32[r * ld_db + c] = v; }
      if (db16) db16[r * ld_db + c] = float_to_bf16_bits(v);
    }
  }
}

// AITM information transfer (aitm.py:82-91): two tokens (p = the previous task's transferred feature, q = this task's
// own feature), each with value / key / query projections, one row of [V_p K_p Q_p V_q K_q Q_q] (6H floats):
//   s_j = <K_j, Q_j> / sqrt(H),  a = softmax_j(s),  out = a_p V_p + a_q V_q
// One warp per row; the row's two attention weights are kept for the backward pass.
__global__ void aitm_attention_forward_kernel(const float* vkq, int64_t ld, int rows, int H, float denom, float* o32,
                                              int64_t ld32, uint16_t* o16, int64_t ld16, float* attn) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* x = vkq + (int64_t)r * ld;
    float sp = 0.f, sq = 0.f;
    for (int h = lane; h < H; h += 32) {
      sp = fmaf(x[H + h], x[2 * H + h], sp);
      sq = fmaf(x[4 * H + h], x[5 * H + h], sq);
    }
    sp = warp_sum(sp) / denom;
    sq = warp_sum(sq) / denom;
    const float m = fmaxf(sp, sq);
    const float ep = expf(sp - m), eq = expf(sq - m);
    const float ap = ep / (ep + eq), aq = eq / (ep + eq);
    if (attn && lane == 0) { attn[2 * (int64_t)r] = ap; attn[2 * (int64_t)r + 1] = aq; }
    for (int h = lane; h < H; h += 32) {
      const float v = ap * x[h] + aq * x[3 * H + h];
      if (o32) o32[(int64_t)r * ld32 + h] = v;
      if (o16) o16[(int64_t)r * ld16 + h] = float_to_bf16_bits(v);
    }
  }
}

// d_vkq (fp32 or bf16, same 6H layout) from d_out: dV_j = a_j d_out; with g_j = <d_out, V_j>:
// ds_j = a_j (g_j - sum_i a_i g_i); dK_j = ds_j Q_j / sqrt(H); dQ_j = ds_j K_j / sqrt(H)
__global__ void aitm_attention_backward_kernel(const float* d_out, int64_t ld_dout, const float* vkq, int64_t ld,
                                               const float* attn, int rows, int H, float denom, float* d32, uint16_t* d16,
                                               int64_t ld_d) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < rows; r += gridDim.x * wpb) {
    const float* x = vkq + (int64_t)r * ld;
    const float* g = d_out + (int64_t)r * ld_dout;
    const float ap = attn[2 * (int64_t)r], aq = attn[2 * (int64_t)r + 1];
    float gp = 0.f, gq = 0.f;
    for (int h = lane; h < H; h += 32) {
      const float go = g[h];
      gp = fmaf(go, x[h], gp);
      gq = fmaf(go, x[3 * H + h], gq);
    }
    gp = warp_sum(gp);
    gq = warp_sum(gq);
    const float mean = ap * gp + aq * gq;
    const float dsp = ap * (gp - mean) / denom, dsq = aq * (gq - mean) / denom;
    for (int h = lane; h < H; h += 32) {
      const float go = g[h];
      const float v[6] = {ap * go, dsp * x[2 * H + h], dsp * x[H + h], aq * go, dsq * x[5 * H + h], dsq * x[4 * H + h]};
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        if (d32) d32[(int64_t)r * ld_d + j * H + h] = v[j];
        if (d16) d16[(int64_t)r * ld_d + j * H + h] = float_to_bf16_bits(v[j]);
      }
    }
  }
}

// APG (apg.py:87-99, the shipped use_uv_shared / no-P variant): the per-sample k x k matrix of the layer,
//   kk[b, j] = sum_i nk[b, i] * wkk[b, i * k + j] + bkk[b, j],
// where wkk [B, k*k] and bkk [B, k] are generated from the (detached) scene embedding by two Linear layers.  Byte work:
// every sample streams its own k*k matrix once (4 k^2 B per sample each way); one warp per sample, lanes over j so the
// matrix rows are read coalesced.
__global__ void apg_mix_forward_kernel(const float* nk, int64_t ld_nk, const float* wkk, int64_t ld_w, const float* bkk,
                                       int64_t ld_b, int B, int k, float* o32, int64_t ld32, uint16_t* o16, int64_t ld16) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += gridDim.x * wpb) {
    const float* x = nk + (int64_t)r * ld_nk;
    const float* w = wkk + (int64_t)r * ld_w;
    for (int j = lane; j < k; j += 32) {
      float acc = 0.f;
      for (int i = 0; i < k; ++i) acc = fmaf(x[i], w[i * k + j], acc);
      acc += bkk[(int64_t)r * ld_b + j];
      if (o32) o32[(int64_t)r * ld32 + j] = acc;
      if (o16) o16[(int64_t)r * ld16 + j] = float_to_bf16_bits(acc);
    }
  }
}

// d(nk)[b, i] = sum_j d(kk)[b, j] wkk[b, i*k + j];  d(wkk)[b, i*k + j] = nk[b, i] d(kk)[b, j];  d(bkk) = d(kk).
// All three are assigned (each input has this stage as its only consumer), in fp32 or bf16 as the producers' GEMMs read them.
__global__ void apg_mix_backward_kernel(const float* d_kk, int64_t ld_dkk, const float* nk, int64_t ld_nk, const float* wkk,
                                        int64_t ld_w, int B, int k, float* dnk32, uint16_t* dnk16, int64_t ld_dnk,
                                        float* dw32, uint16_t* dw16, int64_t ld_dw, float* db32, uint16_t* db16,
                                        int64_t ld_db) {
  const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
  for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < B; r += gridDim.x * wpb) {
    const float* g = d_kk + (int64_t)r * ld_dkk;
    const float* x = nk + (int64_t)r * ld_nk;
    const float* w = wkk + (int64_t)r * ld_w;
    for (int i = 0; i < k; ++i) {
      float acc = 0.f;
      for (int j = lane; j < k; j += 32) acc = fmaf(g[j], w[i * k + j], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        if (dnk32) dnk32[(int64_t)r * ld_dnk + i] = acc;
        if (dnk16) dnk16[(int64_t)r * ld_dnk + i] = float_to_bf16_bits(acc);
      }
    }
    for (int e = lane; e < k * k; e += 32) {
      const int i = e / k, j = e - i * k;
      const float v = x[i] * g[j];
      if (dw32) dw32[(int64_t)r * ld_dw + e] = v;
      if (dw16) dw16[(int64_t)r * ld_dw + e] = float_to_bf16_bits(v);
    }
    for (int j = lane; j < k; j += 32) {
      if (db32) db32[(int64_t)r * ld_db + j] = g[j];
      if (db16) db16[(int64_t)r * ld_db + j] = float_to_bf16_bits(g[j]);
    }
  }
}

// out[n] = sum_b Z[b, n] (fp32 or bf16 input): the bias gradient of a Linear whose weight is stored [K, N] (its wgrad
// problem has dZ as the B operand, so the GEMM's row-sum path does not apply).  One CTA per 32 columns, 32 row groups,
// partials added in row-group order: deterministic.
__global__ void colsum_kernel(const float* z32, const uint16_t* z16, int64_t ld, int B, int N, float* out) {
  __shared__ float part[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float acc = 0.f;
  if (n < N)
    for (int r = ty; r < B; r += 32)
      acc += z32 ? z32[(int64_t)r * ld + n] : bf16_bits_to_float(z16[(int64_t)r * ld + n]);
  part[ty][tx] = acc;
  __syncthreads();
  if (ty == 0 && n < N) {
    float t = 0.f;
    for (int g = 0; g < 32; ++g) t += part[g][tx];
    out[n] = t;
  }
}

// SNR-trans / MSSM gate (snr_trans.py:9-50, mssm.py:9-60): out_i = sum_j z_ij * (x_j @ M_ij) with a hard-concrete gate
// per (output i, input j): s = sigmoid(log u - log(1 - u) + log(alpha) / beta), z = clamp(s * (eps - gamma) + gamma, 0, 1).
// SNR-trans: u, z are scalars per connection (Z = 1, u trained); MSSM: vectors over the output unit v (Z = U, u a
// constant: the reference keeps it in a plain list).  The trans matrices M_ij [U, U] are constants in both (never
// registered, never updated).  All outputs of the gate are ONE Linear over the concatenated inputs with the derived weight
//   W_eff[i * U + v][j * U + u] = z_ij[v] * M_ij[u][v]          (nn.Linear layout [n_out * U, n_in * U])
__device__ __forceinline__ float snr_gate_s(float u, float alpha) {
  const float logit = logf(u) - logf(1.f - u) + logf(alpha) / 0.9f;
  return 1.f / (1.f + expf(-logit));
}

__global__ void snr_gate_weights_kernel(const float* u, const float* alpha, const float* M, int n_out, int n_in, int U,
                                        int Z, float* w_eff, int64_t ld_w, uint16_t* w16) {
  const int64_t n = (int64_t)n_out * U * n_in * U;
  const float a = alpha[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(i % (n_in * U));
    const int row = (int)(i / (n_in * U));
    const int oi = row / U, v = row - oi * U, ij = col / U, uu = col - ij * U;
    const float ug = u[(int64_t)(oi * n_in + ij) * Z + (Z > 1 ? v : 0)];
    const float s_ = snr_gate_s(ug, a) * 1.2f - 0.1f;   // eps - gamma = 1.1 + 0.1, gamma = -0.1
    const float z = fminf(fmaxf(s_, 0.f), 1.f);
    const float w = z * M[(((int64_t)oi * n_in + ij) * U + uu) * U + v];
    w_eff[row * ld_w + col] = w;
    if (w16) w16[row * ld_w + col] = float_to_bf16_bits(w);
  }
}

// Z = 1: dz_ij = sum_{u,v} dW_eff[i*U+v][j*U+u] * M_ij[u][v]; Z = U: dz_ij[v] = sum_u (same term).  One CTA per (i, j).
__global__ void snr_gate_fold_kernel(const float* d_w_eff, int64_t ld_w, const float* M, int n_in, int U, int Z,
                                     float* dz) {
  const int oi = blockIdx.x / n_in, ij = blockIdx.x - oi * n_in;
  const float* m = M + (int64_t)blockIdx.x * U * U;
  if (Z > 1) {   // a warp per output unit v, lanes over u, fixed order
    const int lane = threadIdx.x & 31;
    for (int v = threadIdx.x >> 5; v < U; v += blockDim.x >> 5) {
      float acc = 0.f;
      for (int uu = lane; uu < U; uu += 32)
        acc = fmaf(d_w_eff[(int64_t)(oi * U + v) * ld_w + ij * U + uu], m[(int64_t)uu * U + v], acc);
      acc = warp_sum(acc);
      if (lane == 0) dz[(int64_t)blockIdx.x * U + v] = acc;
    }
    return;
  }
  float acc = 0.f;
  for (int e = threadIdx.x; e < U * U; e += blockDim.x) {
    const int v = e / U, uu = e - v * U;            // consecutive threads walk a row of dW_eff
    acc = fmaf(d_w_eff[(int64_t)(oi * U + v) * ld_w + ij * U + uu], m[(int64_t)uu * U + v], acc);
  }
  __shared__ float part[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) dz[blockIdx.x] = t;
  }
}

// chain rule through the hard-concrete gate: the clamp passes a gradient where 0 < s' <= 1 (snr_trans.py:41-44)
__global__ void snr_gate_chain_kernel(const float* dz, const float* u, const float* alpha, int n, float* d_u,
                                      float* d_alpha) {
  const float a = alpha[0];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float uu = u[i];
    const float s = snr_gate_s(uu, a);
    const float s_ = s * 1.2f - 0.1f;
    const float dlogit = (s_ > 0.f && s_ <= 1.f) ? dz[i] * 1.2f * s * (1.f - s) : 0.f;
    if (d_u) d_u[i] = dlogit * (1.f / uu + 1.f / (1.f - uu));   // MSSM's u is a constant: no d_u
    acc += dlogit;
  }
  __shared__ float part[32];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) d_alpha[0] = t / (a * 0.9f);
  }
}

// i indexes (t, n, k) of w_eff [T*N, ld_w]; spec[t] and shared are [K, N] row-major
__global__ void star_weights_kernel(const int64_t* spec_ptrs, const int64_t* spec_b_ptrs, const float* shared,
                                    const float* shared_b, int T, int K, int N, float* w_eff, int64_t ld_w,
                                    uint16_t* w16, float* b_eff) {
  const int64_t n = (int64_t)T * N * K;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const int64_t tn = i / K;
    const int nn = (int)(tn % N), t = (int)(tn / N);
    const float* sp = reinterpret_cast<const float*>(spec_ptrs[t]);
    const float v = sp[(int64_t)k * N + nn] * shared[(int64_t)k * N + nn];
    w_eff[tn * ld_w + k] = v;
    if (w16) w16[tn * ld_w + k] = float_to_bf16_bits(v);
    if (k == 0) b_eff[tn] = reinterpret_cast<const float*>(spec_b_ptrs[t])[nn] + shared_b[nn];
  }
}

__global__ void star_fold_kernel(const float* d_w_eff, int64_t ld_w, const float* d_b_eff, const int64_t* spec_ptrs,
                                 const float* shared, const int32_t* live, int T, int K, int N, float* d_shared,
                                 float* d_shared_b, float* d_spec_last, float* d_spec_b_last) {
  const int64_t n = (int64_t)K * N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i / N), nn = (int)(i - (int64_t)k * N);
    float acc = 0.f;
    for (int t = 0; t < T; ++t) {
      if (!live[t]) continue;
      const float g = d_w_eff[((int64_t)t * N + nn) * ld_w + k];
      acc = fmaf(g, reinterpret_cast<const float*>(spec_ptrs[t])[i], acc);
      if (t == T - 1 && d_spec_last) d_spec_last[i] = g * shared[i];
    }
    d_shared[i] = acc;
    if (k == 0) {
      float bacc = 0.f;
      for (int t = 0; t < T; ++t) if (live[t]) bacc += d_b_eff[(int64_t)t * N + nn];
      d_shared_b[nn] = bacc;
      if (d_spec_b_last && live[T - 1]) d_spec_b_last[nn] = d_b_eff[(int64_t)(T - 1) * N + nn];
    }
  }
}

static inline int grid_for(int64_t n) {
  int64_t g = (n + 255) / 256;
  static int cap = 0;   // 8 resident CTAs of 256 threads per SM
  if (!cap) {
    int dev = 0, n_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cap = 8 * (n_sm > 0 ? n_sm : 1);
  }
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace mmlrec

using namespace mmlrec;

extern "C" int mmlrec_bn_forward(const float* Z, int64_t ldz, int32_t M, int32_t N, const float* gamma,
                                 const float* beta, float* running_mean, float* running_var,
                                 int64_t* num_batches_tracked, int32_t n_tracked, float* save_mean,
                                 float* save_invstd, float* Y, int64_t ldy, uint16_t* Y_bf16, int64_t ldy_bf16,
                                 int32_t act, int32_t training, void* stream) {
  MMLREC_CHECK_ARG(M > 0 && N > 0, "bad sizes");
  MMLREC_CHECK_ARG(n_tracked <= 256, "too many tracked counters");
  MMLREC_CHECK_ARG(Y || Y_bf16, "no output");
  bn_forward_kernel<<<cdiv(N, 32), kBnThreads, 0, (cudaStream_t)stream>>>(
      Z, ldz, M, N, gamma, beta, running_mean, running_var, num_batches_tracked, n_tracked, save_mean, save_invstd, Y,
      ldy, Y_bf16, ldy_bf16, act, training);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_bn_backward(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int32_t M, int32_t N,
                                  const float* gamma, const float* save_mean, const float* save_invstd, float* dZ,
                                  int64_t lddz, uint16_t* dZ_bf16, int64_t lddz_bf16, float* dgamma, float* dbeta,
                                  void* stream) {
  MMLREC_CHECK_ARG(M > 0 && N > 0, "bad sizes");
  bn_backward_kernel<<<cdiv(N, 32), kBnThreads, 0, (cudaStream_t)stream>>>(
      dY, lddy, Z, ldz, M, N, gamma, save_mean, save_invstd, dZ, lddz, dZ_bf16, lddz_bf16, dgamma, dbeta, nullptr, M);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_bn_stats(const float* Z, int64_t ldz, int32_t M, int32_t N, float* stats, void* stream) {
  MMLREC_CHECK_ARG(Z && stats && M > 0 && N > 0, "bad args");
  bn_stats_kernel<<<cdiv(N, 32), kBnThreads, 0, (cudaStream_t)stream>>>(Z, ldz, M, N, stats);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_bn_combine(const float* all_stats, int32_t R, int32_t M, int32_t N, float* running_mean,
                                 float* running_var, int64_t* num_batches_tracked, int32_t n_tracked, float* save_mean,
                                 float* save_invstd, void* stream) {
  MMLREC_CHECK_ARG(all_stats && R > 0 && M > 0 && N > 0 && running_mean && running_var && save_mean && save_invstd, "bad args");
  MMLREC_CHECK_ARG(n_tracked <= 256, "too many tracked counters");
  bn_combine_kernel<<<cdiv(N, 256), 256, 0, (cudaStream_t)stream>>>(all_stats, R, M, N, running_mean, running_var,
                                                                   num_batches_tracked, n_tracked, save_mean, save_invstd);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_bn_backward_sums(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int32_t M, int32_t N,
                                       const float* save_mean, const float* save_invstd, float* sums, float* dgamma,
                                       float* dbeta, void* stream) {
  MMLREC_CHECK_ARG(dY && Z && sums && dgamma && dbeta && M > 0 && N > 0, "bad args");
  bn_backward_sums_kernel<<<cdiv(N, 32), kBnThreads, 0, (cudaStream_t)stream>>>(dY, lddy, Z, ldz, M, N, save_mean, save_invstd,
                                                                              sums, dgamma, dbeta);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_bn_backward_synced(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int32_t M, int32_t N,
                                         const float* gamma, const float* save_mean, const float* save_invstd, float* dZ,
                                         int64_t lddz, uint16_t* dZ_bf16, int64_t lddz_bf16, const float* global_sums,
                                         int32_t M_total, void* stream) {
  MMLREC_CHECK_ARG(M > 0 && N > 0 && global_sums && M_total >= M, "bad args");
  bn_backward_kernel<<<cdiv(N, 32), kBnThreads, 0, (cudaStream_t)stream>>>(
      dY, lddy, Z, ldz, M, N, gamma, save_mean, save_invstd, dZ, lddz, dZ_bf16, lddz_bf16, nullptr, nullptr, global_sums, M_total);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_gate_mix_forward(const MmlrecGate* gates, int32_t n_gates, int32_t B, void* stream) {
  MMLREC_CHECK_ARG(n_gates > 0 && B > 0, "bad sizes");
  dim3 grid(cdiv(B, 8), n_gates);
  gate_mix_forward_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gates, B);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int64_t mmlrec_gate_mix_backward_scratch(int32_t n_gates, int32_t max_ne, int32_t max_hg, int32_t B) {
  return (int64_t)n_gates * cdiv(B, kGateRows) * max_ne * max_hg;
}

extern "C" int mmlrec_gate_mix_backward(const MmlrecGate* gates, int32_t n_gates, const MmlrecExpertGrad* experts,
                                        int32_t n_experts, int32_t B, int32_t max_ne, int32_t max_hg,
                                        int32_t serialize_gates, float* scratch, int32_t* counters, void* stream) {
  MMLREC_CHECK_ARG(n_gates > 0 && B > 0 && max_ne > 0 && max_ne <= MMLREC_MAX_GATE_EXPERTS && max_hg > 0, "bad sizes");
  const int n_cta = cdiv(B, kGateRows);
  const int64_t per_gate = (int64_t)n_cta * max_ne * max_hg;
  size_t smem = (size_t)kGateRows * (size_t)(max_hg + max_ne) * sizeof(float);
  MMLREC_CHECK_ARG(smem <= 200 * 1024, "gate input too wide for the shared-memory tile");
  static size_t opted = 0;
  if (smem > 48 * 1024 && smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(gate_mix_backward_gate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gate_mix_backward: smem opt-in failed"); return (int)e; }
    opted = smem;
  }
  dim3 grid(n_cta, serialize_gates ? 1 : n_gates);
  gate_mix_backward_gate_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(gates, n_gates, B, scratch, per_gate, counters);
  MMLREC_CHECK_LAUNCH(1);
  if (n_experts > 0) {
    dim3 g2(cdiv(B, 8), n_experts);
    gate_mix_backward_expert_kernel<<<g2, 256, 0, (cudaStream_t)stream>>>(experts, B);
    MMLREC_CHECK_LAUNCH(1);
  }
  return 0;
}

extern "C" int64_t mmlrec_heads_scratch(int32_t T, int32_t max_h, int32_t B) {
  return (int64_t)cdiv(B, kHeadRows) * T * (2 + max_h);
}

static int heads_launch(const MmlrecHead* heads, int32_t T, int32_t B, const float* y, int64_t ldy,
                        float* pred, int64_t ld_pred, float* loss, int32_t esmm, int32_t training,
                        float* scratch, int64_t scratch_floats, int32_t* counters, int grad_mode, void* stream,
                        const float* mask = nullptr, int64_t ldm = 0);

extern "C" int mmlrec_heads_forward_backward(const MmlrecHead* heads, int32_t T, int32_t B, const float* y, int64_t ldy,
                                             float* pred, int64_t ld_pred, float* loss, int32_t esmm, int32_t training,
                                             float* scratch, int64_t scratch_floats, int32_t* counters, void* stream) {
  return heads_launch(heads, T, B, y, ldy, pred, ld_pred, loss, esmm, training, scratch, scratch_floats, counters, 0, stream);
}

extern "C" int mmlrec_heads_backward_external(const MmlrecHead* heads, int32_t T, int32_t B, const float* d_pred, int64_t ld_d_pred,
                                              float* pred, int64_t ld_pred, float* loss, int32_t esmm,
                                              float* scratch, int64_t scratch_floats, int32_t* counters, void* stream) {
  MMLREC_CHECK_ARG(d_pred != nullptr, "no upstream gradient");
  return heads_launch(heads, T, B, d_pred, ld_d_pred, pred, ld_pred, loss, esmm, 1, scratch, scratch_floats, counters, 1, stream);
}

extern "C" int mmlrec_heads_forward_backward_masked(const MmlrecHead* heads, int32_t T, int32_t B, const float* y,
                                                    int64_t ldy, const float* mask, int64_t ld_mask, float* pred,
                                                    int64_t ld_pred, float* loss, int32_t flags, int32_t training,
                                                    float* scratch, int64_t scratch_floats, int32_t* counters,
                                                    void* stream) {
  MMLREC_CHECK_ARG(mask != nullptr && (flags & 1) == 0, "a scenario mask (not with the ESMM product head)");
  return heads_launch(heads, T, B, y, ldy, pred, ld_pred, loss, flags, training, scratch, scratch_floats, counters, 0, stream,
                      mask, ld_mask);
}

static int heads_launch(const MmlrecHead* heads, int32_t T, int32_t B, const float* y, int64_t ldy,
                        float* pred, int64_t ld_pred, float* loss, int32_t esmm, int32_t training,
                        float* scratch, int64_t scratch_floats, int32_t* counters, int grad_mode, void* stream,
                        const float* mask, int64_t ldm) {
  MMLREC_CHECK_ARG(T > 0 && T < MMLREC_MAX_TASKS && B > 0, "bad sizes");
  MMLREC_CHECK_ARG(!(esmm & 1) || T == 2, "esmm needs exactly two heads");
  MMLREC_CHECK_ARG((esmm & ~7) == 0 && (esmm & 3) != 3, "bad head flags");
  MMLREC_CHECK_ARG(!((esmm & 4) && training && y != nullptr && !(T <= 8)), "shared bias: one-launch kernel only");
  const int n_cta = cdiv(B, kHeadRows);
  int stride_cta = 0;
  size_t smem = 0;
  if (training && y != nullptr) {
    MMLREC_CHECK_ARG(scratch && counters && scratch_floats > 0 && scratch_floats % n_cta == 0, "bad scratch");
    stride_cta = (int)(scratch_floats / n_cta);
    MMLREC_CHECK_ARG(stride_cta % T == 0 && stride_cta / T - 2 <= 32 * kHeadMaxK, "head width > 256 unsupported");
    smem = (size_t)8 * T * (stride_cta / T - 2) * sizeof(float);
    MMLREC_CHECK_ARG(smem <= 160 * 1024, "too many / too wide heads for the per-warp dw slabs");
    static size_t opted = 0;
    if (smem > 40 * 1024 && smem > opted) {
      cudaError_t e = cudaFuncSetAttribute(heads_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("heads: smem opt-in failed"); return (int)e; }
      opted = smem;
    }
  }
  if (training && y != nullptr && getenv("MMLREC_HEADS_TWO_KERNELS") == nullptr) {
    // one-launch path: rows fetched up front, last CTA reduces (heads_fast_kernel)
    const int hmax = stride_cta / T - 2;
    if (T <= 4 && hmax <= 64) {
      launch_pdl(heads_fast_kernel<4, 2, 4>, dim3(cdiv(B, 32)), dim3(256), 0, stream, heads, T, B, y, ldy, pred, ld_pred, loss,
                 esmm, scratch, stride_cta, counters, grad_mode, mask, ldm);
      MMLREC_RETURN_LAUNCH(1);
    }
    if (T <= 8 && hmax <= 128) {
      launch_pdl(heads_fast_kernel<8, 4, 2>, dim3(cdiv(B, 16)), dim3(256), 0, stream, heads, T, B, y, ldy, pred, ld_pred, loss,
                 esmm, scratch, stride_cta, counters, grad_mode, mask, ldm);
      MMLREC_RETURN_LAUNCH(1);
    }
    // wide heads (129..256 columns) whose biases are cumulative or shared -- MLP / ESCM on the KuaiRec shape, [512, 256] --
    // exist only in the one-launch kernel; everything else that wide keeps the two-kernel path below
    if ((esmm & 6) && T <= 4 && hmax <= 256) {
      launch_pdl(heads_fast_kernel<4, 8, 2>, dim3(cdiv(B, 16)), dim3(256), 0, stream, heads, T, B, y, ldy, pred, ld_pred, loss,
                 esmm, scratch, stride_cta, counters, grad_mode, mask, ldm);
      MMLREC_RETURN_LAUNCH(1);
    }
  }
  MMLREC_CHECK_ARG(!((esmm & 6) && training && y != nullptr), "cumulative / shared biases are handled by the one-launch kernel only");
  launch_pdl(heads_kernel, dim3(n_cta), dim3(256), smem, stream, heads, T, B, y, ldy, pred, ld_pred, loss, esmm, training, scratch,
             stride_cta, counters, grad_mode, mask, ldm);
  if (!(training && y != nullptr)) { MMLREC_RETURN_LAUNCH(1); }
  MMLREC_CHECK_LAUNCH(1);
  launch_pdl(heads_reduce_kernel, dim3(cdiv(stride_cta, 32)), dim3(256), 0, stream, heads, T, loss, esmm, scratch, stride_cta, n_cta);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_dense_optimizer_step(float* param, const float* grad, float* state1, float* state2, int64_t n,
                                           const MmlrecHyper* hyper, uint16_t* bf16_shadow, void* stream) {
  MMLREC_CHECK_ARG(n >= 0 && hyper, "bad args");
  if (n == 0) return 0;
  dense_optimizer_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, state1, state2, n, hyper, bf16_shadow);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_dense_optimizer_step_sliced(float* param, const float* grad, float* state1, float* state2, int64_t n,
                                                  const MmlrecHyper* hyper, uint16_t* bf16_shadow, int32_t n_slices,
                                                  int64_t slice_stride, void* stream) {
  MMLREC_CHECK_ARG(n >= 0 && hyper && n_slices >= 1, "bad args");
  MMLREC_CHECK_ARG((n & 3) == 0 && (slice_stride & 3) == 0, "n and slice_stride must be multiples of 4");
  MMLREC_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad | (uintptr_t)state1 | (uintptr_t)state2) & 15) == 0 &&
                   ((uintptr_t)bf16_shadow & 7) == 0, "buffers must be 16-byte aligned");
  if (n == 0) return 0;
  launch_pdl(dense_optimizer_sliced_kernel, dim3(grid_for(n / 4)), dim3(256), 0, stream,
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(state1),
      reinterpret_cast<float4*>(state2), n / 4, hyper, reinterpret_cast<uint2*>(bf16_shadow), n_slices, slice_stride / 4);
  MMLREC_RETURN_LAUNCH(1);
}

namespace mmlrec {
// dst[seg.dst + i] = sum_{s < S} src[s * slice_stride + seg.src + i]   (fixed order: deterministic split-K wgrad)
__global__ void __launch_bounds__(256) sum_slices_kernel(const int64_t* seg, float* dst, const float* src, int S,
                                                         int64_t slice_stride) {
  pdl_prologue();
  const int64_t d0 = seg[blockIdx.y * 3 + 0], s0 = seg[blockIdx.y * 3 + 1], n = seg[blockIdx.y * 3 + 2];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = src[s0 + i];
    for (int k = 1; k < S; ++k) acc += src[(int64_t)k * slice_stride + s0 + i];
    dst[d0 + i] = acc;
  }
}
}  // namespace mmlrec

extern "C" int mmlrec_sum_slices(const int64_t* segments, int32_t n_segments, int64_t max_n, float* dst, const float* src,
                                 int32_t S, int64_t slice_stride, void* stream) {
  MMLREC_CHECK_ARG(segments && dst && src && n_segments > 0 && S > 0 && max_n > 0, "bad args");
  int gx = (int)((max_n + 255) / 256);
  if (gx > 592) gx = 592;
  dim3 grid(gx, n_segments);
  mmlrec::launch_pdl(mmlrec::sum_slices_kernel, grid, dim3(256), 0, stream, segments, dst, src, S, slice_stride);
  MMLREC_RETURN_LAUNCH(1);
}

namespace mmlrec {
// L2 regularisation of the dense parameters (model/basemodel.py:524-540: reg = sum l2 * w^2 over the registered weights,
// added to the loss before backward): grad += 2 * coef * p, and the value of the term.  coef holds the per-element l2
// (0 for biases / BatchNorm / anything not registered; NEGATIVE where the gradient entry is written by nobody else and
// must be assigned instead of accumulated); CTA partials of the value are added by the finish kernel in
// block order (deterministic).
__global__ void __launch_bounds__(256) l2_grad_kernel(const float* __restrict__ p, float* g, const float* __restrict__ coef,
                                                      int64_t n, float* partial) {
  __shared__ float red[8];
  float acc = 0.f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float c = coef[i];
    if (c != 0.f) {   // c < 0: no backward kernel ever writes this entry (an unused expert): ASSIGN, the buffer holds stale data
      const float w = p[i], a = fabsf(c);
      g[i] = (c > 0.f ? g[i] : 0.f) + 2.f * a * w;
      acc = fmaf(a * w, w, acc);
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k];
    partial[blockIdx.x] = t;
  }
}
__global__ void l2_finish_kernel(const float* partial, int n_part, float* reg_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float t = 0.f;
  for (int k = 0; k < n_part; ++k) t += partial[k];
  *reg_out = t;
}
}  // namespace mmlrec

extern "C" int64_t mmlrec_l2_scratch(void) { return 1024; }

extern "C" int mmlrec_l2_regularize(const float* param, float* grad, const float* l2_coef, int64_t n, float* reg_out,
                                    float* scratch, void* stream) {
  MMLREC_CHECK_ARG(param && grad && l2_coef && reg_out && scratch && n > 0, "bad args");
  int grid = grid_for(n);
  if (grid > 1024) grid = 1024;
  mmlrec::l2_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(param, grad, l2_coef, n, scratch);
  MMLREC_CHECK_LAUNCH(1);
  mmlrec::l2_finish_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(scratch, grid, reg_out);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_fill_f32(float* p, int64_t n, float v, void* stream) {
  if (n <= 0) return 0;
  launch_pdl(fill_kernel, dim3(grid_for(n)), dim3(256), 0, stream, p, n, v);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_cast_f32_to_bf16(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst, int32_t rows,
                                       int32_t cols, int32_t cols_pad, void* stream) {
  MMLREC_CHECK_ARG(rows >= 0 && cols >= 0 && cols_pad >= cols, "bad sizes");
  if (rows == 0 || cols_pad == 0) return 0;
  cast_f32_bf16_kernel<<<grid_for((int64_t)rows * cols_pad), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst, rows,
                                                                                           cols, cols_pad);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_cast_bf16_to_f32(const uint16_t* src, int64_t ld_src, float* dst, int64_t ld_dst, int32_t rows,
                                       int32_t cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  cast_bf16_f32_kernel<<<grid_for((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst, ld_dst, rows, cols);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_mul_f32(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                              int32_t rows, int32_t cols, int32_t accumulate, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  mul_kernel<<<grid_for((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, out, ldo, rows, cols, accumulate);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_copy_cols(const float* src, int64_t ld_src, float* dst_f32, int64_t ld_f32, uint16_t* dst_bf16,
                                int64_t ld_bf16, int32_t rows, int32_t cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  copy_cols_kernel<<<grid_for((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst_f32, ld_f32, dst_bf16,
                                                                                   ld_bf16, rows, cols);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_mul_forward(const float* a, int64_t lda, const float* b, int64_t ldb, float* out_f32, int64_t ld_f32,
                                  uint16_t* out_bf16, int64_t ld_bf16, int32_t rows, int32_t cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  mul_forward_kernel<<<grid_for((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, out_f32, ld_f32,
                                                                                     out_bf16, ld_bf16, rows, cols);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_mul_backward(const float* d_out, int64_t ld_dout, const float* a, int64_t lda, const float* b,
                                   int64_t ldb, float* da_f32, uint16_t* da_bf16, int64_t ld_da, int32_t dkind_a,
                                   int32_t acc_a, float* db_f32, uint16_t* db_bf16, int64_t ld_db, int32_t dkind_b,
                                   int32_t acc_b, int32_t rows, int32_t cols, void* stream) {
  if (rows <= 0 || cols <= 0) return 0;
  mul_backward_kernel<<<grid_for((int64_t)rows * cols), 256, 0, (cudaStream_t)stream>>>(
      d_out, ld_dout, a, lda, b, ldb, da_f32, da_bf16, ld_da, dkind_a, acc_a, db_f32, db_bf16, ld_db, dkind_b, acc_b, rows,
      cols);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_aitm_attention_forward(const float* vkq, int64_t ld, int32_t rows, int32_t H, float* out_f32,
                                             int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16, float* attn,
                                             void* stream) {
  if (rows <= 0 || H <= 0) return 0;
  MMLREC_CHECK_ARG(vkq && (out_f32 || out_bf16), "null buffer");
  aitm_attention_forward_kernel<<<grid_for((int64_t)rows * 32), 256, 0, (cudaStream_t)stream>>>(
      vkq, ld, rows, H, sqrtf((float)H), out_f32, ld_f32, out_bf16, ld_bf16, attn);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_aitm_attention_backward(const float* d_out, int64_t ld_dout, const float* vkq, int64_t ld,
                                              const float* attn, int32_t rows, int32_t H, float* d_vkq_f32,
                                              uint16_t* d_vkq_bf16, int64_t ld_d, void* stream) {
  if (rows <= 0 || H <= 0) return 0;
  MMLREC_CHECK_ARG(d_out && vkq && attn && (d_vkq_f32 || d_vkq_bf16), "null buffer");
  aitm_attention_backward_kernel<<<grid_for((int64_t)rows * 32), 256, 0, (cudaStream_t)stream>>>(
      d_out, ld_dout, vkq, ld, attn, rows, H, sqrtf((float)H), d_vkq_f32, d_vkq_bf16, ld_d);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_apg_mix_forward(const float* nk, int64_t ld_nk, const float* wkk, int64_t ld_w, const float* bkk,
                                      int64_t ld_b, int32_t B, int32_t k, float* out_f32, int64_t ld_f32,
                                      uint16_t* out_bf16, int64_t ld_bf16, void* stream) {
  if (B <= 0 || k <= 0) return 0;
  MMLREC_CHECK_ARG(nk && wkk && bkk && (out_f32 || out_bf16), "null buffer");
  apg_mix_forward_kernel<<<grid_for((int64_t)B * 32), 256, 0, (cudaStream_t)stream>>>(nk, ld_nk, wkk, ld_w, bkk, ld_b, B, k,
                                                                                     out_f32, ld_f32, out_bf16, ld_bf16);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_apg_mix_backward(const float* d_kk, int64_t ld_dkk, const float* nk, int64_t ld_nk, const float* wkk,
                                       int64_t ld_w, int32_t B, int32_t k, float* d_nk_f32, uint16_t* d_nk_bf16,
                                       int64_t ld_dnk, float* d_wkk_f32, uint16_t* d_wkk_bf16, int64_t ld_dw,
                                       float* d_bkk_f32, uint16_t* d_bkk_bf16, int64_t ld_db, void* stream) {
  if (B <= 0 || k <= 0) return 0;
  MMLREC_CHECK_ARG(d_kk && nk && wkk && (d_nk_f32 || d_nk_bf16) && (d_wkk_f32 || d_wkk_bf16) && (d_bkk_f32 || d_bkk_bf16),
                   "null buffer");
  apg_mix_backward_kernel<<<grid_for((int64_t)B * 32), 256, 0, (cudaStream_t)stream>>>(
      d_kk, ld_dkk, nk, ld_nk, wkk, ld_w, B, k, d_nk_f32, d_nk_bf16, ld_dnk, d_wkk_f32, d_wkk_bf16, ld_dw, d_bkk_f32,
      d_bkk_bf16, ld_db);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_colsum(const float* z_f32, const uint16_t* z_bf16, int64_t ld, int32_t B, int32_t N, float* out,
                             void* stream) {
  if (N <= 0) return 0;
  MMLREC_CHECK_ARG((z_f32 || z_bf16) && out && B >= 0, "null buffer");
  colsum_kernel<<<(N + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(z_f32, z_bf16, ld, B, N, out);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_snr_gate_weights(const float* u, const float* alpha, const float* trans, int32_t n_out, int32_t n_in,
                                       int32_t U, int32_t zdim, float* w_eff, int64_t ld_w, uint16_t* w_eff_bf16,
                                       void* stream) {
  MMLREC_CHECK_ARG(u && alpha && trans && w_eff && n_out > 0 && n_in > 0 && U > 0, "bad argument");
  MMLREC_CHECK_ARG(zdim == 1 || zdim == U, "zdim must be 1 (SNR-trans) or U (MSSM)");
  snr_gate_weights_kernel<<<grid_for((int64_t)n_out * U * n_in * U), 256, 0, (cudaStream_t)stream>>>(
      u, alpha, trans, n_out, n_in, U, zdim, w_eff, ld_w, w_eff_bf16);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_snr_gate_fold(const float* d_w_eff, int64_t ld_w, const float* trans, const float* u,
                                    const float* alpha, int32_t n_out, int32_t n_in, int32_t U, int32_t zdim,
                                    float* dz_scratch, float* d_u, float* d_alpha, void* stream) {
  MMLREC_CHECK_ARG(d_w_eff && trans && u && alpha && dz_scratch && d_alpha && n_out > 0 && n_in > 0 && U > 0,
                   "bad argument");
  MMLREC_CHECK_ARG(zdim == 1 || zdim == U, "zdim must be 1 (SNR-trans) or U (MSSM)");
  snr_gate_fold_kernel<<<n_out * n_in, 256, 0, (cudaStream_t)stream>>>(d_w_eff, ld_w, trans, n_in, U, zdim, dz_scratch);
  MMLREC_CHECK_LAUNCH(1);
  snr_gate_chain_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(dz_scratch, u, alpha, n_out * n_in * zdim, d_u, d_alpha);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_star_weights(const int64_t* spec_ptrs, const int64_t* spec_b_ptrs, const float* shared,
                                   const float* shared_b, int32_t T, int32_t K, int32_t N, float* w_eff, int64_t ld_w,
                                   uint16_t* w_eff_bf16, float* b_eff, void* stream) {
  MMLREC_CHECK_ARG(T > 0 && K > 0 && N > 0 && ld_w >= K, "bad sizes");
  star_weights_kernel<<<grid_for((int64_t)T * N * K), 256, 0, (cudaStream_t)stream>>>(spec_ptrs, spec_b_ptrs, shared, shared_b,
                                                                                    T, K, N, w_eff, ld_w, w_eff_bf16, b_eff);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_star_fold(const float* d_w_eff, int64_t ld_w, const float* d_b_eff, const int64_t* spec_ptrs,
                                const float* shared, const int32_t* live, int32_t T, int32_t K, int32_t N, float* d_shared,
                                float* d_shared_b, float* d_spec_last, float* d_spec_b_last, void* stream) {
  MMLREC_CHECK_ARG(T > 0 && K > 0 && N > 0 && ld_w >= K, "bad sizes");
  star_fold_kernel<<<grid_for((int64_t)K * N), 256, 0, (cudaStream_t)stream>>>(d_w_eff, ld_w, d_b_eff, spec_ptrs, shared, live,
                                                                             T, K, N, d_shared, d_shared_b, d_spec_last,
                                                                             d_spec_b_last);
  MMLREC_RETURN_LAUNCH(1);
}
