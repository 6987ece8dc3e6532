// Level-fused gate stage (model/mmoe.py:80-88 / model/ple.py:127-152 and their backward): one record describes
// ALL gates of a level and the level's distinct expert activations.
//   * tiled kernels (default, second half of this file): a CTA stages the rows of 8 samples in shared memory;
//   * warp-per-sample kernels (first half; fallback for unaligned rows or levels whose tile exceeds 110 KB): a warp
//     owns one sample, softmax in registers (lane e <-> expert slot e), ONE pass over the expert rows with
//     batched 128-bit loads.
// Either way every activation row of the level is read once and every output row written once, and the CTA
// partials of dWg are reduced in a fixed order by gate_level_dwg_reduce_kernel (deterministic, no float atomics).
#include "common.cuh"

namespace mmlrec {

constexpr int GL_MAXG = MMLREC_LEVEL_MAX_GATES;   // register arrays are sized by this
constexpr int GL_FWD_WARPS = 8;            // forward: 8 warps x 1 sample
constexpr int GL_BWD_WARPS = 8;            // backward: 8 warps x 1 sample = 8 samples per CTA
constexpr int GL_BWD_ROWS = 8;
constexpr int GL_BATCH = 4;                // expert rows fetched together (4 x 128-bit loads in flight per lane)

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float dot4(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// logits + softmax of gate g for sample b: returns this lane's probability (lane e <-> slot e)
__device__ __forceinline__ float gate_probs(const MmlrecGateLevel& L, int g, int b, const float* wg_s, int lane) {
  const int Hg = L.Hg[g], ne = L.n_e[g];
  const float* gin = L.gate_in[g] + (int64_t)b * L.ld_gate_in[g];
  float logit = -INFINITY;
  for (int e0 = 0; e0 < ne; e0 += 4) {      // four independent dot products in flight
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const int e1 = min(e0 + 1, ne - 1), e2 = min(e0 + 2, ne - 1), e3 = min(e0 + 3, ne - 1);
    for (int h = lane; h < Hg; h += 32) {
      const float x = gin[h];
      s0 = fmaf(x, wg_s[e0 * Hg + h], s0);
      s1 = fmaf(x, wg_s[e1 * Hg + h], s1);
      s2 = fmaf(x, wg_s[e2 * Hg + h], s2);
      s3 = fmaf(x, wg_s[e3 * Hg + h], s3);
    }
    s0 = warp_sum(s0); s1 = warp_sum(s1); s2 = warp_sum(s2); s3 = warp_sum(s3);
    if (lane == e0) logit = s0;
    if (lane == e0 + 1 && e0 + 1 < ne) logit = s1;
    if (lane == e0 + 2 && e0 + 2 < ne) logit = s2;
    if (lane == e0 + 3 && e0 + 3 < ne) logit = s3;
  }
  float mx = logit;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const float ex = lane < ne ? expf(logit - mx) : 0.f;
  return ex / warp_sum(ex);
}

__device__ __forceinline__ void stage_level(const MmlrecGateLevel* lv, MmlrecGateLevel& L, float* wg_s, int* wg_off) {
  const int n_words = sizeof(MmlrecGateLevel) / 4;
  for (int i = threadIdx.x; i < n_words; i += blockDim.x)
    reinterpret_cast<uint32_t*>(&L)[i] = reinterpret_cast<const uint32_t*>(lv)[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    int at = 0;
    for (int g = 0; g < L.n_gates; ++g) { wg_off[g] = at; at += L.n_e[g] * L.Hg[g]; }
    wg_off[L.n_gates] = at;
  }
  __syncthreads();
  for (int g = 0; g < L.n_gates; ++g) {
    const int Hg = L.Hg[g];
    for (int i = threadIdx.x; i < L.n_e[g] * Hg; i += blockDim.x)
      wg_s[wg_off[g] + i] = L.Wg[g][(int64_t)(i / Hg) * L.ld_Wg[g] + (i % Hg)];
  }
  __syncthreads();
}

__global__ void __launch_bounds__(GL_FWD_WARPS * 32, 4) gate_level_forward_kernel(const MmlrecGateLevel* lv, int B) {
  __shared__ MmlrecGateLevel L;
  __shared__ float wg_s[MMLREC_LEVEL_MAX_WG];
  __shared__ int wg_off[GL_MAXG + 1];
  stage_level(lv, L, wg_s, wg_off);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int G = L.n_gates, H = L.H;
  const int b = blockIdx.x * GL_FWD_WARPS + w;
  if (b >= B) return;  // warp-uniform; no block-level sync follows
  float pg[GL_MAXG];
#pragma unroll
  for (int g = 0; g < GL_MAXG; ++g) {
    pg[g] = 0.f;
    if (g < G) {
      pg[g] = gate_probs(L, g, b, wg_s + wg_off[g], lane);
      if (lane < L.n_e[g]) L.probs[g][(int64_t)b * L.n_e[g] + lane] = pg[g];
    }
  }
  for (int h0 = 0; h0 < H; h0 += 128) {
    const int h = h0 + lane * 4;
    const bool hv = h < H;
    float4 acc[GL_MAXG];
#pragma unroll
    for (int g = 0; g < GL_MAXG; ++g) acc[g] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int u0 = 0; u0 < L.n_experts; u0 += GL_BATCH) {
      float4 eo[GL_BATCH];
#pragma unroll
      for (int k = 0; k < GL_BATCH; ++k)   // all loads of the batch are issued before any use
        eo[k] = (hv && u0 + k < L.n_experts) ? ld4(L.expert[u0 + k] + (int64_t)b * L.ld_expert + h)
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < GL_BATCH; ++k) {
        if (u0 + k >= L.n_experts) break;  // warp-uniform
#pragma unroll
        for (int g = 0; g < GL_MAXG; ++g) {
          const int slot = g < G ? L.slot[u0 + k][g] : -1;  // warp-uniform
          if (slot >= 0) {
            const float p = __shfl_sync(0xffffffffu, pg[g], slot);
            acc[g].x = fmaf(p, eo[k].x, acc[g].x); acc[g].y = fmaf(p, eo[k].y, acc[g].y);
            acc[g].z = fmaf(p, eo[k].z, acc[g].z); acc[g].w = fmaf(p, eo[k].w, acc[g].w);
          }
        }
      }
    }
    if (hv) {
#pragma unroll
      for (int g = 0; g < GL_MAXG; ++g) {
        if (g < G) {
          *reinterpret_cast<float4*>(L.mix[g] + (int64_t)b * L.ld_mix[g] + h) = acc[g];
          if (L.mix_bf16[g]) {
            uint2 o;
            o.x = pack_bf16x2(acc[g].x, acc[g].y);
            o.y = pack_bf16x2(acc[g].z, acc[g].w);
            *reinterpret_cast<uint2*>(L.mix_bf16[g] + (int64_t)b * L.ld_mix_bf16[g] + h) = o;
          }
        }
      }
    }
  }
}

// dynamic smem: dl_s [GL_BWD_ROWS][total_ne] + gin_s [GL_BWD_ROWS][total_hg]   (total_ne = sum n_e, total_hg = sum Hg)
__global__ void __launch_bounds__(GL_BWD_WARPS * 32, 3)
gate_level_backward_kernel(const MmlrecGateLevel* lv, int B, float* scratch) {
  extern __shared__ __align__(16) float dyn_s[];
  __shared__ MmlrecGateLevel L;
  __shared__ float wg_s[MMLREC_LEVEL_MAX_WG];
  __shared__ int wg_off[GL_MAXG + 1], ne_off[GL_MAXG + 1], hg_off[GL_MAXG + 1];
  stage_level(lv, L, wg_s, wg_off);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int G = L.n_gates, H = L.H;
  if (threadIdx.x == 0) {
    int a = 0, c = 0;
    for (int g = 0; g < G; ++g) { ne_off[g] = a; hg_off[g] = c; a += L.n_e[g]; c += L.Hg[g]; }
    ne_off[G] = a; hg_off[G] = c;
  }
  __syncthreads();
  const int total_wg = wg_off[G], total_ne = ne_off[G], total_hg = hg_off[G];
  float* dl_s = dyn_s;                            // [GL_BWD_ROWS][total_ne]
  float* gin_s = dyn_s + GL_BWD_ROWS * total_ne;  // [GL_BWD_ROWS][total_hg]
  const int r0 = blockIdx.x * GL_BWD_ROWS;
  for (int rr = 0; rr < 1; ++rr) {
    const int r = w;
    const int b = r0 + r;
    if (b >= B) {  // rows past the batch contribute zeros to the CTA partial
      for (int i = lane; i < total_ne; i += 32) dl_s[r * total_ne + i] = 0.f;
      for (int i = lane; i < total_hg; i += 32) gin_s[r * total_hg + i] = 0.f;
      continue;
    }
    float pg[GL_MAXG], dp[GL_MAXG];
    bool live[GL_MAXG];
#pragma unroll
    for (int g = 0; g < GL_MAXG; ++g) {
      live[g] = g < G && L.d_mix[g] != nullptr;
      pg[g] = (live[g] && lane < L.n_e[g]) ? L.probs[g][(int64_t)b * L.n_e[g] + lane] : 0.f;
      dp[g] = 0.f;
    }
    // one pass over the expert rows: d(expert) and the softmax-input dot products
    for (int h0 = 0; h0 < H; h0 += 128) {
      const int h = h0 + lane * 4;
      const bool hv = h < H;
      float4 dm[GL_MAXG];
#pragma unroll
      for (int g = 0; g < GL_MAXG; ++g)
        dm[g] = (live[g] && hv) ? ld4(L.d_mix[g] + (int64_t)b * L.ld_d_mix[g] + h) : make_float4(0.f, 0.f, 0.f, 0.f);
      for (int u0 = 0; u0 < L.n_experts; u0 += GL_BATCH) {
        float4 eo[GL_BATCH];
#pragma unroll
        for (int k = 0; k < GL_BATCH; ++k)
          eo[k] = (hv && u0 + k < L.n_experts) ? ld4(L.expert[u0 + k] + (int64_t)b * L.ld_expert + h)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int k = 0; k < GL_BATCH; ++k) {
          const int u = u0 + k;
          if (u >= L.n_experts) break;  // warp-uniform
          float4 de = make_float4(0.f, 0.f, 0.f, 0.f);
          bool used = false;
#pragma unroll
          for (int g = 0; g < GL_MAXG; ++g) {
            const int slot = live[g] ? L.slot[u][g] : -1;  // warp-uniform
            if (slot >= 0) {
              used = true;
              const float p = __shfl_sync(0xffffffffu, pg[g], slot);
              de.x = fmaf(p, dm[g].x, de.x); de.y = fmaf(p, dm[g].y, de.y);
              de.z = fmaf(p, dm[g].z, de.z); de.w = fmaf(p, dm[g].w, de.w);
              const float s = warp_sum(dot4(dm[g], eo[k]));
              if (lane == slot) dp[g] += s;
            }
          }
          if (used && hv) {
            if (L.expert_relu) {
              if (!(eo[k].x > 0.f)) de.x = 0.f;
              if (!(eo[k].y > 0.f)) de.y = 0.f;
              if (!(eo[k].z > 0.f)) de.z = 0.f;
              if (!(eo[k].w > 0.f)) de.w = 0.f;
            }
            if (L.d_expert[u]) *reinterpret_cast<float4*>(L.d_expert[u] + (int64_t)b * L.ld_d_expert + h) = de;
            if (L.d_expert_bf16[u]) {
              uint2 o;
              o.x = pack_bf16x2(de.x, de.y);
              o.y = pack_bf16x2(de.z, de.w);
              *reinterpret_cast<uint2*>(L.d_expert_bf16[u] + (int64_t)b * L.ld_d_expert_bf16 + h) = o;
            }
          }
        }
      }
    }
    // softmax backward, d(gate_in); dlogits and gate inputs are staged for the CTA-wide dWg partial
#pragma unroll
    for (int g = 0; g < GL_MAXG; ++g) {
      if (g >= G) continue;
      const int Hg = L.Hg[g], ne = L.n_e[g];
      float dl = 0.f;
      if (live[g]) {
        const float dot = warp_sum(pg[g] * dp[g]);
        dl = pg[g] * (dp[g] - dot);   // lane e <-> d logit e (0 for lanes >= n_e)
      }
      if (lane < ne) dl_s[r * total_ne + ne_off[g] + lane] = dl;
      const float* gin = L.gate_in[g] + (int64_t)b * L.ld_gate_in[g];
      const float* wg = wg_s + wg_off[g];
      for (int h0 = 0; h0 < Hg; h0 += 32) {   // uniform trip count (shuffles inside)
        const int h = h0 + lane;
        const bool hv = h < Hg;
        const float gv = hv ? gin[h] : 0.f;
        if (hv) gin_s[r * total_hg + hg_off[g] + h] = gv;
        if (!live[g]) continue;  // warp-uniform
        float acc = 0.f;
        for (int e = 0; e < ne; ++e) {
          const float de = __shfl_sync(0xffffffffu, dl, e);
          if (hv) acc = fmaf(de, wg[e * Hg + h], acc);
        }
        if (hv) {
          if (L.relu_mask_gate_in[g] && !(gv > 0.f)) acc = 0.f;
          if (L.d_gate_in[g]) {
            float* dst = L.d_gate_in[g] + (int64_t)b * L.ld_d_gate_in[g] + h;
            if (L.accumulate_d_gate_in[g]) acc += *dst;
            *dst = acc;
          }
          if (L.d_gate_in_bf16[g]) L.d_gate_in_bf16[g][(int64_t)b * L.ld_d_gate_in_bf16[g] + h] = float_to_bf16_bits(acc);
        }
      }
    }
  }
  __syncthreads();
  // CTA partial of dWg[g][e][h] = sum over the 32 staged samples of dl[r][e] * gin[r][h]  (fixed order)
  float* part = scratch + (int64_t)blockIdx.x * total_wg;
  for (int i = threadIdx.x; i < total_wg; i += blockDim.x) {
    int g = 0;
    while (g + 1 < G && wg_off[g + 1] <= i) ++g;
    const int local = i - wg_off[g], Hg = L.Hg[g];
    const int e = local / Hg, h = local - e * Hg;
    const float* dlp = dl_s + ne_off[g] + e;
    const float* gp = gin_s + hg_off[g] + h;
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < GL_BWD_ROWS; ++r) s = fmaf(dlp[r * total_ne], gp[r * total_hg], s);
    part[i] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// Tiled kernels (the default).  A CTA owns GT_ROWS samples.  Every row it needs -- expert activations,
// gate inputs, gate-head weights, and in backward d(mix) of the live gates -- is brought into shared
// memory ONCE with bulk async copies (cp.async.bulk, one per row, issued by one thread each and tracked
// by an mbarrier); the phases then run from shared memory with work mapped so that almost every issued
// instruction is a 128-bit LDS or an FMA.
//   forward   F1 logits: 8 lanes per (gate, expert) dot product   F2 softmax: thread <-> (sample, gate)
//             F3 mixtures: warp <-> sample, lane <-> 4 columns
//   backward  P1 dp[r][g][e] = <d_mix[r][g], expert[r][u(g,e)]>    8 lanes per dot product
//             P3 d_expert[r][u] = relu'(.) sum_{g uses u} p * d_mix[r][g]   warp <-> expert, lane <-> 4 columns
//             P2 dlogit = p * (dp - <p, dp>)                        thread <-> (sample, gate)
//             P4 d_gate_in[r][g] = dlogit Wg     warp <-> sample, gates in order (shared inputs accumulate)
//             P5 CTA partial of dWg = sum_r dlogit[r] (x) gate_in[r]   thread <-> 4 columns of a (gate, expert) row
// dynamic smem (floats): Wg [total_wg] | d_mix [R][G][H] (backward) | expert [R][E][H] | gate_in [R][total_hg]
//                        | p [R][total_ne] | dp/dlogit [R][total_ne] (backward)
// ------------------------------------------------------------------------------------------------
constexpr int GT_ROWS = 8;
constexpr int GT_THREADS = 256;
constexpr int GT_WARPS = GT_THREADS / 32;
constexpr int GT_MAXP = GL_MAXG * MMLREC_LEVEL_MAX_EXPERTS;
constexpr int GT_MAXSRC = 2 * GL_MAXG + MMLREC_LEVEL_MAX_EXPERTS;

struct __align__(8) GtPair { uint16_t a, c, n4, col; };   // operand offsets (floats), length in float4, column in [total_ne]

struct GtTables {
  int wg_off[GL_MAXG + 1], ne_off[GL_MAXG + 1], hg_off[GL_MAXG + 1], pair_off[GL_MAXG + 1];
  uint8_t live[GL_MAXG];
  GtPair pair[GT_MAXP];                 // live (gate, expert) pairs, grouped by gate
  uint16_t pair_eo[GT_MAXP];            // expert row offset (float4) of the pair inside one sample's expert block
  uint8_t pair_g[GT_MAXP], pair_e[GT_MAXP];
  uint8_t ucnt[MMLREC_LEVEL_MAX_EXPERTS], ug[MMLREC_LEVEL_MAX_EXPERTS][GL_MAXG];   // gates whose gradient reaches expert u
  uint8_t used[MMLREC_LEVEL_MAX_EXPERTS];                                          // expert u is read by some live gate
  uint16_t ucol[MMLREC_LEVEL_MAX_EXPERTS][GL_MAXG];
  // rows to stage per sample
  const float* src[GT_MAXSRC];
  int64_t src_ld[GT_MAXSRC];
  int dst_off[GT_MAXSRC], dst_stride[GT_MAXSRC], n4[GT_MAXSRC];
};

__device__ __forceinline__ void fma4(float4& acc, float s, const float4& v) {
  acc.x = fmaf(s, v.x, acc.x); acc.y = fmaf(s, v.y, acc.y); acc.z = fmaf(s, v.z, acc.z); acc.w = fmaf(s, v.w, acc.w);
}
__device__ __forceinline__ uint32_t gt_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void gt_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(gt_smem_u32(dst)), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void gt_bar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GT_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"
      "@p bra GT_WAIT_DONE;\n"
      "bra GT_WAIT_LOOP;\n"
      "GT_WAIT_DONE:\n"
      "}\n" ::"r"(bar), "r"(parity) : "memory");
}

// record -> shared memory, offsets, pair table, per-expert user lists.  `backward`: only gates that receive a
// gradient are live; forward: all gates.
__device__ __forceinline__ void gt_setup(const MmlrecGateLevel* lv, MmlrecGateLevel& L, GtTables& T, bool backward,
                                         uint64_t* bar) {
  const int tid = threadIdx.x;
  for (int i = tid; i < (int)(sizeof(MmlrecGateLevel) / 4); i += (int)blockDim.x)
    reinterpret_cast<uint32_t*>(&L)[i] = reinterpret_cast<const uint32_t*>(lv)[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(gt_smem_u32(bar)), "r"((int)blockDim.x));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int G = L.n_gates, E = L.n_experts, H = L.H;
  if (tid == 0) {
    int a = 0, c = 0, d = 0, p = 0;
    for (int g = 0; g < G; ++g) {
      T.wg_off[g] = a; T.ne_off[g] = c; T.hg_off[g] = d; T.pair_off[g] = p;
      const bool lv_g = !backward || L.d_mix[g] != nullptr;
      T.live[g] = lv_g ? 1 : 0;
      a += L.n_e[g] * L.Hg[g]; c += L.n_e[g]; d += L.Hg[g];
      if (lv_g) p += L.n_e[g];
    }
    T.wg_off[G] = a; T.ne_off[G] = c; T.hg_off[G] = d; T.pair_off[G] = p;
  }
  __syncthreads();
  for (int i = tid; i < E * G; i += (int)blockDim.x) {   // thread <-> (expert, gate)
    const int u = i / G, g = i - u * G;
    const int e = L.slot[u][g];
    if (e < 0 || !T.live[g]) continue;
    const int p = T.pair_off[g] + e;
    GtPair d;
    if (backward) { d.a = (uint16_t)(g * H); d.c = (uint16_t)(u * H); d.n4 = (uint16_t)(H >> 2); }
    else { d.a = (uint16_t)T.hg_off[g]; d.c = (uint16_t)(T.wg_off[g] + e * L.Hg[g]); d.n4 = (uint16_t)(L.Hg[g] >> 2); }
    d.col = (uint16_t)(T.ne_off[g] + e);
    T.pair[p] = d;
    T.pair_eo[p] = (uint16_t)(u * (H >> 2));
    T.pair_g[p] = (uint8_t)g; T.pair_e[p] = (uint8_t)e;
  }
  if (tid < E) {
    int n = 0;
    bool used = false;
    for (int g = 0; g < G; ++g) {
      const int sl = L.slot[tid][g];
      if (!T.live[g] || sl < 0) continue;
      used = true;
      if (backward && ((L.detach_mask[tid] >> g) & 1u)) continue;   // gate g treats this expert as a constant
      T.ug[tid][n] = (uint8_t)g; T.ucol[tid][n] = (uint16_t)(T.ne_off[g] + sl); ++n;
    }
    T.ucnt[tid] = (uint8_t)n;
    T.used[tid] = used ? 1 : 0;
  }
  __syncthreads();
}

// Issue one bulk copy per (source, sample) row and per live gate-head weight row; every thread arrives on the
// mbarrier once (with the bytes it asked for), then waits for the whole tile.
__device__ __forceinline__ void gt_stage(const MmlrecGateLevel& L, const GtTables& T, int n_src, int n_rows, int r0,
                                         float* dyn_s, float* wg_s, int n_pairs, uint64_t* bar) {
  const int tid = threadIdx.x;
  const uint32_t b32 = gt_smem_u32(bar);
  const int n_items = n_src * GT_ROWS;
  uint32_t bytes = 0;
  for (int it = tid; it < n_items + n_pairs; it += (int)blockDim.x) {
    if (it < n_items) {
      const int v = it >> 3, r = it & (GT_ROWS - 1);
      if (r < n_rows) bytes += (uint32_t)T.n4[v] << 4;
    } else {
      bytes += (uint32_t)L.Hg[T.pair_g[it - n_items]] << 2;
    }
  }
  if (bytes) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b32), "r"(bytes) : "memory");
  else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(b32) : "memory");
  for (int it = tid; it < n_items + n_pairs; it += (int)blockDim.x) {
    if (it < n_items) {
      const int v = it >> 3, r = it & (GT_ROWS - 1);
      const int n4 = T.n4[v];
      if (r < n_rows && n4)
        gt_bulk_g2s(dyn_s + T.dst_off[v] + r * T.dst_stride[v], T.src[v] + (int64_t)(r0 + r) * T.src_ld[v], (uint32_t)n4 << 4, b32);
    } else {
      const int p = it - n_items, g = T.pair_g[p], e = T.pair_e[p], Hg = L.Hg[g];
      gt_bulk_g2s(wg_s + T.wg_off[g] + e * Hg, L.Wg[g] + (int64_t)e * L.ld_Wg[g], (uint32_t)Hg << 2, b32);
    }
  }
  gt_bar_wait(b32, 0);
}

// NT = 256: one warp per sample.  NT = 512: two warps per sample (pairs / gates split between them) -- the phases are
// chains of dependent LDS -> FMA -> SHFL, so twice the resident warps hide twice the latency.
template <int NT>
__global__ void __launch_bounds__(NT, 2)
gate_level_forward_tiled_kernel(const MmlrecGateLevel* lv, int B) {
  pdl_prologue();
  constexpr int NW = NT / 32, PARTS = NW / GT_ROWS;
  static_assert(NW % GT_ROWS == 0, "warps must be a multiple of the samples per CTA");
  extern __shared__ __align__(128) float dyn_s[];
  __shared__ __align__(16) MmlrecGateLevel L;
  __shared__ GtTables T;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  gt_setup(lv, L, T, false, &bar);
  const int G = L.n_gates, E = L.n_experts, H = L.H, H4 = H >> 2;
  const int total_wg = T.wg_off[G], total_ne = T.ne_off[G], total_hg = T.hg_off[G];
  float* wg_s = dyn_s;
  float* eo_s = wg_s + total_wg;
  float* gin_s = eo_s + GT_ROWS * E * H;
  float* p_s = gin_s + GT_ROWS * total_hg;
  const int r0 = blockIdx.x * GT_ROWS;
  const int n_rows = min(GT_ROWS, B - r0);
  if (tid < E + G) {
    const int v = tid;
    const float* src = nullptr; int64_t ld = 0; int off = 0, stride = 0, n4 = 0;
    if (v < E) {
      if (T.used[v]) { src = L.expert[v]; ld = L.ld_expert; off = (int)(eo_s - dyn_s) + v * H; stride = E * H; n4 = H4; }
    } else {
      const int g = v - E;
      src = L.gate_in[g]; ld = L.ld_gate_in[g]; off = (int)(gin_s - dyn_s) + T.hg_off[g]; stride = total_hg; n4 = L.Hg[g] >> 2;
    }
    T.src[v] = src; T.src_ld[v] = ld; T.dst_off[v] = off; T.dst_stride[v] = stride; T.n4[v] = n4;
  }
  __syncthreads();
  gt_stage(L, T, E + G, n_rows, r0, dyn_s, wg_s, total_ne, &bar);
  // ---- F1: logits; a warp owns a sample, 8 lanes share one (gate, expert) dot product
  {
    const int sub = lane & 7, pl = lane >> 3;
    const int r = w % GT_ROWS, part = w / GT_ROWS;
    const int q_per = ((total_ne + 3) / 4 + PARTS - 1) / PARTS;           // pair-quads per part
    const int p_lo = 4 * q_per * part, p_hi = min(total_ne, p_lo + 4 * q_per);
    if (r < n_rows) {
      const float* gin_r = gin_s + r * total_hg;
#pragma unroll 3
      for (int p0 = p_lo; p0 < p_hi; p0 += 4) {   // uniform trip count: shuffles inside
        const int p = p0 + pl;
        const bool pv = p < total_ne;
        float s = 0.f;
        if (pv) {
          const GtPair d = T.pair[p];
          const float4* a = reinterpret_cast<const float4*>(gin_r + d.a);
          const float4* c = reinterpret_cast<const float4*>(wg_s + d.c);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
          for (int q = sub; q < d.n4; q += 8) {
            const float4 x = a[q], y = c[q];
            acc.x = fmaf(x.x, y.x, acc.x); acc.y = fmaf(x.y, y.y, acc.y);
            acc.z = fmaf(x.z, y.z, acc.z); acc.w = fmaf(x.w, y.w, acc.w);
          }
          s = (acc.x + acc.y) + (acc.z + acc.w);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (pv && sub == 0) p_s[r * total_ne + p] = s;
      }
    }
  }
  __syncthreads();
  // ---- F2: softmax per (sample, gate); probabilities are saved for the backward pass
  for (int it = tid; it < n_rows * G; it += NT) {
    const int r = it / G, g = it - r * G;
    const int ne = L.n_e[g];
    float* pr = p_s + r * total_ne + T.ne_off[g];
    float mx = -INFINITY;
    for (int e = 0; e < ne; ++e) mx = fmaxf(mx, pr[e]);
    float sum = 0.f;
    for (int e = 0; e < ne; ++e) { const float x = expf(pr[e] - mx); pr[e] = x; sum += x; }
    float* out = L.probs[g] + (int64_t)(r0 + r) * ne;
    for (int e = 0; e < ne; ++e) { const float x = pr[e] / sum; pr[e] = x; out[e] = x; }
  }
  __syncthreads();
  // ---- F3: mixtures; a warp owns a sample, lane <-> 4 columns
  if (w % GT_ROWS < n_rows) {
    const int r = w % GT_ROWS;
    const int b = r0 + r;
    const float4* eo = reinterpret_cast<const float4*>(eo_s + r * E * H);
    for (int g = w / GT_ROWS; g < G; g += PARTS) {
      const int ne = L.n_e[g], off = T.ne_off[g];
      const float* pr = p_s + r * total_ne + off;
      const uint16_t* eoff = T.pair_eo + off;
      float* out32 = L.mix[g] + (int64_t)b * L.ld_mix[g];
      uint16_t* out16 = L.mix_bf16[g] ? L.mix_bf16[g] + (int64_t)b * L.ld_mix_bf16[g] : nullptr;
      for (int q = lane; q < H4; q += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int e = 0; e < ne; ++e) fma4(acc, pr[e], eo[eoff[e] + q]);
        reinterpret_cast<float4*>(out32)[q] = acc;
        if (out16) {
          uint2 o;
          o.x = pack_bf16x2(acc.x, acc.y);
          o.y = pack_bf16x2(acc.z, acc.w);
          reinterpret_cast<uint2*>(out16)[q] = o;
        }
      }
    }
  }
}

__global__ void __launch_bounds__(GT_THREADS, 2)
gate_level_backward_tiled_kernel(const MmlrecGateLevel* lv, int B, float* scratch) {
  pdl_prologue();
  extern __shared__ __align__(128) float dyn_s[];
  __shared__ __align__(16) MmlrecGateLevel L;
  __shared__ GtTables T;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  gt_setup(lv, L, T, true, &bar);
  const int G = L.n_gates, E = L.n_experts, H = L.H, H4 = H >> 2;
  const int total_wg = T.wg_off[G], total_ne = T.ne_off[G], total_hg = T.hg_off[G], n_pairs = T.pair_off[G];
  float* wg_s = dyn_s;
  float* dm_s = wg_s + total_wg;
  float* eo_s = dm_s + GT_ROWS * G * H;
  float* gin_s = eo_s + GT_ROWS * E * H;
  float* p_s = gin_s + GT_ROWS * total_hg;
  float* dl_s = p_s + GT_ROWS * total_ne;
  const int r0 = blockIdx.x * GT_ROWS;
  const int n_rows = min(GT_ROWS, B - r0);
  const int per_row = 2 * G + E;
  if (tid < per_row) {
    const int v = tid;
    const float* src = nullptr; int64_t ld = 0; int off = 0, stride = 0, n4 = 0;
    if (v < G) {
      if (T.live[v]) { src = L.d_mix[v]; ld = L.ld_d_mix[v]; off = (int)(dm_s - dyn_s) + v * H; stride = G * H; n4 = H4; }
    } else if (v < G + E) {
      const int u = v - G;
      if (T.used[u]) { src = L.expert[u]; ld = L.ld_expert; off = (int)(eo_s - dyn_s) + u * H; stride = E * H; n4 = H4; }
    } else {
      const int g = v - G - E;
      if (T.live[g]) { src = L.gate_in[g]; ld = L.ld_gate_in[g]; off = (int)(gin_s - dyn_s) + T.hg_off[g]; stride = total_hg; n4 = L.Hg[g] >> 2; }
    }
    T.src[v] = src; T.src_ld[v] = ld; T.dst_off[v] = off; T.dst_stride[v] = stride; T.n4[v] = n4;
  }
  // probabilities, and zeros for the gate inputs of rows past the batch (they enter the CTA's dWg partial)
  for (int r = w; r < GT_ROWS; r += GT_WARPS) {
    const int b = r0 + r;
    for (int c = lane; c < total_ne; c += 32) {
      int g = 0;
      while (g + 1 < G && T.ne_off[g + 1] <= c) ++g;
      p_s[r * total_ne + c] = (T.live[g] && b < B) ? L.probs[g][(int64_t)b * L.n_e[g] + (c - T.ne_off[g])] : 0.f;
    }
    if (b >= B) for (int c = lane; c < total_hg; c += 32) gin_s[r * total_hg + c] = 0.f;
  }
  __syncthreads();
  gt_stage(L, T, per_row, n_rows, r0, dyn_s, wg_s, n_pairs, &bar);
  __syncthreads();   // p_s / zero rows written with ordinary stores by other warps

  // ---- P1: softmax-input dot products; a warp owns a sample, 8 lanes share one (gate, expert) pair
  {
    const int sub = lane & 7, pl = lane >> 3;
    for (int r = w; r < n_rows; r += GT_WARPS) {
      const float* dm_r = dm_s + r * G * H;
      const float* eo_r = eo_s + r * E * H;
#pragma unroll 3
      for (int p0 = 0; p0 < n_pairs; p0 += 4) {   // uniform trip count: shuffles inside
        const int p = p0 + pl;
        const bool pv = p < n_pairs;
        float s = 0.f;
        int col = 0;
        if (pv) {
          const GtPair d = T.pair[p];
          col = d.col;
          const float4* a = reinterpret_cast<const float4*>(dm_r + d.a);
          const float4* c = reinterpret_cast<const float4*>(eo_r + d.c);
          float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
          for (int q = sub; q < H4; q += 8) {
            const float4 x = a[q], y = c[q];
            acc.x = fmaf(x.x, y.x, acc.x); acc.y = fmaf(x.y, y.y, acc.y);
            acc.z = fmaf(x.z, y.z, acc.z); acc.w = fmaf(x.w, y.w, acc.w);
          }
          s = (acc.x + acc.y) + (acc.z + acc.w);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        if (pv && sub == 0) dl_s[r * total_ne + col] = s;
      }
    }
  }
  // ---- P3: d(expert) (needs only p and d_mix: runs before the softmax backward to share its barrier);
  // a warp owns an expert and walks the CTA's samples, lane <-> 4 columns
  for (int u = w; u < E; u += GT_WARPS) {
    const int cnt = T.ucnt[u];
    if (cnt == 0) continue;  // warp-uniform
    float* out32 = L.d_expert[u] ? L.d_expert[u] + (int64_t)r0 * L.ld_d_expert : nullptr;
    uint16_t* out16 = L.d_expert_bf16[u] ? L.d_expert_bf16[u] + (int64_t)r0 * L.ld_d_expert_bf16 : nullptr;
    const bool relu = L.expert_relu != 0;
    for (int q = lane; q < H4; q += 32) {
      float4 de[GT_ROWS];
#pragma unroll
      for (int r = 0; r < GT_ROWS; ++r) de[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int k = 0; k < cnt; ++k) {   // user gates outer, the CTA's samples inner: table lookups amortised
        const float* pk = p_s + T.ucol[u][k];
        const float4* dk = reinterpret_cast<const float4*>(dm_s + T.ug[u][k] * H) + q;
#pragma unroll
        for (int r = 0; r < GT_ROWS; ++r) fma4(de[r], pk[r * total_ne], dk[r * G * H4]);
      }
#pragma unroll
      for (int r = 0; r < GT_ROWS; ++r) {
        if (r >= n_rows) break;  // warp-uniform (rows past the batch hold stale smem: never stored)
        float4 d = de[r];
        if (relu) {
          const float4 x = reinterpret_cast<const float4*>(eo_s + (r * E + u) * H)[q];
          if (!(x.x > 0.f)) d.x = 0.f;
          if (!(x.y > 0.f)) d.y = 0.f;
          if (!(x.z > 0.f)) d.z = 0.f;
          if (!(x.w > 0.f)) d.w = 0.f;
        }
        if (out32) *reinterpret_cast<float4*>(out32 + (int64_t)r * L.ld_d_expert + 4 * q) = d;
        if (out16) {
          uint2 o;
          o.x = pack_bf16x2(d.x, d.y);
          o.y = pack_bf16x2(d.z, d.w);
          *reinterpret_cast<uint2*>(out16 + (int64_t)r * L.ld_d_expert_bf16 + 4 * q) = o;
        }
      }
    }
  }
  __syncthreads();
  // ---- P2: softmax backward in place (dp -> dlogit); zero for dead gates and rows past the batch
  for (int it = tid; it < GT_ROWS * G; it += GT_THREADS) {
    const int r = it / G, g = it - r * G;
    const int ne = L.n_e[g], base = r * total_ne + T.ne_off[g];
    if (!T.live[g] || r >= n_rows) {
      for (int e = 0; e < ne; ++e) dl_s[base + e] = 0.f;
      continue;
    }
    float dot = 0.f;
    for (int e = 0; e < ne; ++e) dot = fmaf(p_s[base + e], dl_s[base + e], dot);
    for (int e = 0; e < ne; ++e) dl_s[base + e] = p_s[base + e] * (dl_s[base + e] - dot);
  }
  __syncthreads();
  // ---- P4: d(gate_in); a warp owns a sample and walks the gates in order, so gates that share an
  // input (MMoE: every head reads the level input) accumulate without a race
  for (int r = w; r < n_rows; r += GT_WARPS) {
    const int b = r0 + r;
    for (int g = 0; g < G; ++g) {
      if (!T.live[g] || (!L.d_gate_in[g] && !L.d_gate_in_bf16[g])) continue;
      const int Hg = L.Hg[g], ne = L.n_e[g];
      const float4* wg = reinterpret_cast<const float4*>(wg_s + T.wg_off[g]);
      const float* dl = dl_s + r * total_ne + T.ne_off[g];
      for (int q = lane; q < (Hg >> 2); q += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int e = 0; e < ne; ++e) fma4(acc, dl[e], wg[e * (Hg >> 2) + q]);
        if (L.relu_mask_gate_in[g]) {
          const float4 x = reinterpret_cast<const float4*>(gin_s + r * total_hg + T.hg_off[g])[q];
          if (!(x.x > 0.f)) acc.x = 0.f;
          if (!(x.y > 0.f)) acc.y = 0.f;
          if (!(x.z > 0.f)) acc.z = 0.f;
          if (!(x.w > 0.f)) acc.w = 0.f;
        }
        if (L.d_gate_in[g]) {
          float4* dst = reinterpret_cast<float4*>(L.d_gate_in[g] + (int64_t)b * L.ld_d_gate_in[g]) + q;
          if (L.accumulate_d_gate_in[g]) { const float4 o = *dst; acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w; }
          *dst = acc;
        }
        if (L.d_gate_in_bf16[g]) {
          uint2 o;
          o.x = pack_bf16x2(acc.x, acc.y);
          o.y = pack_bf16x2(acc.z, acc.w);
          reinterpret_cast<uint2*>(L.d_gate_in_bf16[g] + (int64_t)b * L.ld_d_gate_in_bf16[g])[q] = o;
        }
      }
    }
  }
  // ---- P5: CTA partial of dWg (fixed order over the CTA's samples)
  float* part = scratch + (int64_t)blockIdx.x * total_wg;
  for (int i4 = tid; i4 < (total_wg >> 2); i4 += GT_THREADS) {
    const int i = i4 << 2;
    int g = 0;
    while (g + 1 < G && T.wg_off[g + 1] <= i) ++g;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    if (T.live[g]) {
      const int local = i - T.wg_off[g], Hg = L.Hg[g];
      const int e = local / Hg, h = local - e * Hg;
      const float* dlp = dl_s + T.ne_off[g] + e;
      const float* gp = gin_s + T.hg_off[g] + h;
#pragma unroll
      for (int r = 0; r < GT_ROWS; ++r) fma4(s, dlp[r * total_ne], *reinterpret_cast<const float4*>(gp + r * total_hg));
    }
    *reinterpret_cast<float4*>(part + i) = s;
  }
}

// deterministic reduction of the CTA partials: 32 outputs x 32 partial-groups per CTA (1024 threads).  A thread sums
// every 32nd partial of its output with all its loads in flight at once (the loop is pure L2 latency: 8 groups x 64
// dependent rounds took 12 us for 512 partials), the 32 group sums are added in a fixed order.
constexpr int GR_GROUPS = 32;
__global__ void __launch_bounds__(32 * GR_GROUPS)
gate_level_dwg_reduce_kernel(const MmlrecGateLevel* lv, const float* scratch, int n_cta, int total_wg) {
  pdl_prologue();
  __shared__ float red[GR_GROUPS][33];
  const int ix = threadIdx.x & 31, iy = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + ix;
  float s = 0.f;
  if (i < total_wg) {
    float a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = 0.f;
    int c = iy;
    for (; c + 7 * GR_GROUPS < n_cta; c += 8 * GR_GROUPS) {
#pragma unroll
      for (int q = 0; q < 8; ++q) a[q] += scratch[(int64_t)(c + q * GR_GROUPS) * total_wg + i];
    }
    for (; c < n_cta; c += GR_GROUPS) a[0] += scratch[(int64_t)c * total_wg + i];
    s = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
  red[iy][ix] = s;
  __syncthreads();
  if (iy != 0 || i >= total_wg) return;
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < GR_GROUPS; ++k) t += red[k][ix];
  int g = 0, off = 0;
  while (g + 1 < lv->n_gates && off + lv->n_e[g] * lv->Hg[g] <= i) { off += lv->n_e[g] * lv->Hg[g]; ++g; }
  if (lv->d_mix[g] != nullptr) {
    const int local = i - off, Hg = lv->Hg[g];
    lv->dWg[g][(int64_t)(local / Hg) * lv->ld_Wg[g] + (local % Hg)] = t;
  }
}

}  // namespace mmlrec

using namespace mmlrec;

extern "C" int mmlrec_gate_level_forward(const MmlrecGateLevel* level, int32_t B, void* stream) {
  MMLREC_CHECK_ARG(level && B > 0, "bad args");
  gate_level_forward_kernel<<<cdiv(B, GL_FWD_WARPS), GL_FWD_WARPS * 32, 0, (cudaStream_t)stream>>>(level, B);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int64_t mmlrec_gate_level_backward_scratch(int32_t total_wg, int32_t B) {
  return (int64_t)cdiv(B, GL_BWD_ROWS) * total_wg;
}

extern "C" int mmlrec_gate_level_backward(const MmlrecGateLevel* level, int32_t B, int32_t total_wg, int32_t total_ne,
                                          int32_t total_hg, float* scratch, int32_t* counter, void* stream) {
  MMLREC_CHECK_ARG(level && B > 0 && scratch && counter, "bad args");
  MMLREC_CHECK_ARG(total_wg > 0 && total_wg <= MMLREC_LEVEL_MAX_WG && total_ne > 0 && total_hg > 0, "sizes out of range");
  const size_t smem = (size_t)GL_BWD_ROWS * (total_ne + total_hg) * sizeof(float);   // staged dlogits + gate inputs
  MMLREC_CHECK_ARG(smem <= 160 * 1024, "gate inputs too wide for the fused kernel");
  static size_t opted = 0;   // static (record + staged weights) + dynamic shared memory can exceed 48 KB
  if (smem + 16 * 1024 > 48 * 1024 && smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(gate_level_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gate_level_backward: smem opt-in failed"); return (int)e; }
    opted = smem;
  }
  const int n_cta = cdiv(B, GL_BWD_ROWS);
  gate_level_backward_kernel<<<n_cta, GL_BWD_WARPS * 32, smem, (cudaStream_t)stream>>>(level, B, scratch);
  MMLREC_CHECK_LAUNCH(1);
  launch_pdl(gate_level_dwg_reduce_kernel, dim3(cdiv(total_wg, 32)), dim3(32 * GR_GROUPS), 0, stream, level, (const float*)scratch, n_cta, total_wg);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int64_t mmlrec_gate_level_backward_tiled_scratch(int32_t total_wg, int32_t B) {
  return (int64_t)cdiv(B, GT_ROWS) * total_wg;
}

/* shared memory (bytes) the tiled backward needs; callers use the warp-per-sample kernel when it does not fit */
extern "C" int64_t mmlrec_gate_level_backward_tiled_smem(int32_t n_gates, int32_t n_experts, int32_t H, int32_t total_wg,
                                                         int32_t total_ne, int32_t total_hg) {
  return ((int64_t)total_wg + (int64_t)GT_ROWS * ((int64_t)(n_gates + n_experts) * H + total_hg + 2 * total_ne)) * 4;
}

extern "C" int mmlrec_gate_level_backward_tiled(const MmlrecGateLevel* level, int32_t B, int32_t n_gates, int32_t n_experts,
                                                int32_t H, int32_t total_wg, int32_t total_ne, int32_t total_hg,
                                                float* scratch, void* stream) {
  MMLREC_CHECK_ARG(level && B > 0 && scratch, "bad args");
  MMLREC_CHECK_ARG(n_gates > 0 && n_gates <= GL_MAXG && n_experts > 0 && n_experts <= MMLREC_LEVEL_MAX_EXPERTS &&
                   H > 0 && H % 4 == 0 && total_wg > 0 && total_wg % 4 == 0 && total_hg % 4 == 0 && total_ne > 0,
                   "sizes out of range");
  const int64_t smem = mmlrec_gate_level_backward_tiled_smem(n_gates, n_experts, H, total_wg, total_ne, total_hg);
  MMLREC_CHECK_ARG(smem <= 110 * 1024, "level too large for the tiled kernel");
  static int64_t opted = 0;
  if (smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(gate_level_backward_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gate_level_backward_tiled: smem opt-in failed"); return (int)e; }
    opted = smem;
  }
  const int n_cta = cdiv(B, GT_ROWS);
  launch_pdl(gate_level_backward_tiled_kernel, dim3(n_cta), dim3(GT_THREADS), (size_t)smem, stream, level, B, scratch);
  MMLREC_CHECK_LAUNCH(1);
  launch_pdl(gate_level_dwg_reduce_kernel, dim3(cdiv(total_wg, 32)), dim3(32 * GR_GROUPS), 0, stream, level, (const float*)scratch, n_cta, total_wg);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int64_t mmlrec_gate_level_forward_tiled_smem(int32_t n_experts, int32_t H, int32_t total_wg, int32_t total_ne,
                                                        int32_t total_hg) {
  return ((int64_t)total_wg + (int64_t)GT_ROWS * ((int64_t)n_experts * H + total_hg + total_ne)) * 4;
}

extern "C" int mmlrec_gate_level_forward_tiled(const MmlrecGateLevel* level, int32_t B, int32_t n_experts, int32_t H,
                                               int32_t total_wg, int32_t total_ne, int32_t total_hg, void* stream) {
  MMLREC_CHECK_ARG(level && B > 0, "bad args");
  MMLREC_CHECK_ARG(n_experts > 0 && n_experts <= MMLREC_LEVEL_MAX_EXPERTS && H > 0 && H % 4 == 0 && total_wg > 0 &&
                   total_wg % 4 == 0 && total_hg % 4 == 0 && total_ne > 0 && total_ne <= GT_MAXP, "sizes out of range");
  const int64_t smem = mmlrec_gate_level_forward_tiled_smem(n_experts, H, total_wg, total_ne, total_hg);
  MMLREC_CHECK_ARG(smem <= 110 * 1024, "level too large for the tiled kernel");
  static int64_t opted = 0;
  if (smem > opted) {
    cudaError_t e = cudaFuncSetAttribute(gate_level_forward_tiled_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(gate_level_forward_tiled_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gate_level_forward_tiled: smem opt-in failed"); return (int)e; }
    opted = smem;
  }
  static int threads = 0;
  if (!threads) { const char* e = getenv("MMLREC_GATE_FWD_THREADS"); threads = (e && atoi(e) == 256) ? 256 : 512; }
  if (threads == 256)
    launch_pdl(gate_level_forward_tiled_kernel<256>, dim3(cdiv(B, GT_ROWS)), dim3(256), (size_t)smem, stream, level, B);
  else
    launch_pdl(gate_level_forward_tiled_kernel<512>, dim3(cdiv(B, GT_ROWS)), dim3(512), (size_t)smem, stream, level, B);
  MMLREC_RETURN_LAUNCH(1);
}
