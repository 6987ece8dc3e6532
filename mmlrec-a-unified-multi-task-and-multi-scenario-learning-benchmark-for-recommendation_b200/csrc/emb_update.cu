// K2: embedding backward = per-field sort of the batch ids + segmented warp-shuffle reduce of the
// gradient slices + the optimizer's row update, fused.  No atomics, deterministic summation order.
// Replaces F_s x embedding_dense_backward (dense [V,D] gradients, autograd of
// model/basemodel.py:475-477) + the dense torch.optim update of every table (basemodel.py:313).
#include "common.cuh"

namespace mmlrec {

// ------------------------------------------------------------------------------------------------
// optimizer clock (torch/optim/adam.py: bias_correction{1,2} = 1 - beta^step, formed in double)
// ------------------------------------------------------------------------------------------------
__global__ void hyper_advance_kernel(MmlrecHyper* h, float2* hist, int cap) {
  pdl_prologue();
  int t = h->step + 1;
  h->step = t;
  if (h->optimizer == MMLREC_OPT_ADAM) {
    double bc1 = 1.0 - pow(h->beta1_d, (double)t);
    double bc2 = 1.0 - pow(h->beta2_d, (double)t);
    h->step_size = (float)(h->lr_d / bc1);
    h->bc2_sqrt = (float)sqrt(bc2);
    // history of the per-step factors: the lazy catch-up of a row replays the steps it missed with exactly these
    if (hist != nullptr) hist[t & (cap - 1)] = make_float2(h->step_size, h->bc2_sqrt);
  }
}

// One Adam step with a zero gradient (what torch's dense Adam does to every row the batch did not touch), with the
// step's own bias-correction factors.  Same operations as optimizer_update(p, 0, ...): the sweep and the lazy
// catch-up both come through here, so they agree bit for bit.
__device__ __forceinline__ void adam_zero_grad_step(float& p, float& s1, float& s2, float step_size, float bc2_sqrt,
                                                    const MmlrecHyper& h) {
  s1 = s1 + h.one_minus_beta1 * (0.f - s1);
  s2 = s2 * h.beta2;
  const float denom = sqrtf(s2) / bc2_sqrt + h.eps;
  p = p - step_size * (s1 / denom);
}

// Bring one row (D floats at element offset `off`) from step `last` to step `upto` by replaying the zero-gradient
// steps last+1 .. upto in registers.
__device__ __forceinline__ void adam_replay_row(float* emb, float* m, float* v, int64_t off, int D, int last, int upto,
                                                const float2* __restrict__ hist, int cap, const MmlrecHyper& hp) {
  for (int d0 = 0; d0 < D; d0 += 4) {
    float4 p4 = *reinterpret_cast<float4*>(emb + off + d0);
    float4 m4 = *reinterpret_cast<float4*>(m + off + d0);
    float4 v4 = *reinterpret_cast<float4*>(v + off + d0);
    for (int j = last + 1; j <= upto; ++j) {
      const float2 hj = __ldg(hist + (j & (cap - 1)));
      adam_zero_grad_step(p4.x, m4.x, v4.x, hj.x, hj.y, hp);
      adam_zero_grad_step(p4.y, m4.y, v4.y, hj.x, hj.y, hp);
      adam_zero_grad_step(p4.z, m4.z, v4.z, hj.x, hj.y, hp);
      adam_zero_grad_step(p4.w, m4.w, v4.w, hj.x, hj.y, hp);
    }
    *reinterpret_cast<float4*>(emb + off + d0) = p4;
    *reinterpret_cast<float4*>(m + off + d0) = m4;
    *reinterpret_cast<float4*>(v + off + d0) = v4;
  }
}

// Exact lazy dense-Adam, part 1 (start of a training step, hyper already advanced to step t): every row the batch is
// about to read is brought up to step t-1.  One thread per (sample, field); among the threads that name the same row
// the one whose compare-and-swap moves row_touch from its old value to t-1 does the replay, the others do nothing
// (nobody reads the row before this kernel has finished).  Rows never touched (row_touch < 0) have zero moments: a
// zero-gradient step does not move them.
__global__ void emb_adam_catch_up_kernel(const float* __restrict__ X, int64_t ldx, int B, const int64_t* __restrict__ field_meta,
                                         int F_s, int D, float* emb, float* m, float* v, int32_t* row_touch,
                                         const MmlrecHyper* hyper, const float2* hist, int cap) {
  pdl_prologue();
  const MmlrecHyper hp = *hyper;
  const int target = hp.step - 1;
  const int64_t n = (int64_t)B * F_s;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(i / F_s), f = (int)(i - (int64_t)s * F_s);
    const int64_t table_off = field_meta[f * 4 + 0], vocab = field_meta[f * 4 + 1];
    int64_t id = (int64_t)X[(int64_t)s * ldx + field_meta[f * 4 + 2]];
    if (id < 0) id = 0;                                     // clamped like the gather (which flags it)
    if (id >= vocab) id = vocab - 1;
    const int64_t row = table_off / D + id;
    const int last = row_touch[row];
    if (last < 0 || last >= target) continue;
    if (atomicCAS(row_touch + row, last, target) != last) continue;
    adam_replay_row(emb, m, v, table_off + id * D, D, last, target, hist, cap, hp);
  }
}

// The same catch-up on the OWNER of row-sharded tables (csrc/peer.cu): the rows to bring up to date are named by the
// request keys the forward exchange delivered -- key = local_row << 32 | pos in rq_keys[parity][F_s][B_all], ~0 where
// the slot belongs to another owner -- and the kernel runs between the ids barrier and emb_serve_rows.
__global__ void emb_adam_catch_up_keys_kernel(const uint64_t* __restrict__ rq_keys, int B_all,
                                              const int64_t* __restrict__ field_meta, int F_s, int D, float* emb, float* m,
                                              float* v, int32_t* row_touch, const MmlrecHyper* hyper, int step_offset,
                                              const float2* hist, int cap) {
  pdl_prologue();
  const MmlrecHyper hp = *hyper;
  const int target = hp.step - 1;
  const int64_t n = (int64_t)F_s * B_all;
  const uint64_t* keys = rq_keys + (int64_t)((hp.step + step_offset) & 1) * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    if (key == ~0ull) continue;
    const int f = (int)(i / B_all);
    const int64_t off = field_meta[f * 4 + 0] + (int64_t)(key >> 32) * D;
    const int64_t row = off / D;
    const int last = row_touch[row];
    if (last < 0 || last >= target) continue;
    if (atomicCAS(row_touch + row, last, target) != last) continue;
    adam_replay_row(emb, m, v, off, D, last, target, hist, cap, hp);
  }
}

// part 2 (before anything reads the tables outside a training step -- predict, state_dict, a checkpoint -- and before
// the history ring wraps): every row is brought up to the current step.
__global__ void emb_adam_flush_kernel(float* emb, float* m, float* v, int32_t* row_touch, int64_t total_rows, int D,
                                      const MmlrecHyper* hyper, const float2* hist, int cap) {
  const MmlrecHyper hp = *hyper;
  const int target = hp.step;
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < total_rows; row += (int64_t)gridDim.x * blockDim.x) {
    const int last = row_touch[row];
    if (last < 0 || last >= target) continue;
    adam_replay_row(emb, m, v, row * D, D, last, target, hist, cap, hp);
    row_touch[row] = target;
  }
}

// ------------------------------------------------------------------------------------------------
// per-field sort: hybrid bitonic network on 64-bit keys (id << 32 | position).  Chunks of 4096 keys
// are sorted in shared memory; strides >= 4096 (only when B > 4096) run as global passes.
// Position in the low word makes the order stable and the result unique => deterministic.
// ------------------------------------------------------------------------------------------------
constexpr int kSortChunk = 4096;
constexpr int kSortThreads = 1024;

__device__ __forceinline__ void cmp_swap(uint64_t& a, uint64_t& b, bool asc) {
  if ((a > b) == asc) { uint64_t t = a; a = b; b = t; }
}

// mode 0: build keys from X and run all stages with k <= chunk.  mode 1: load keys, finish stage
// `k_fixed` (strides chunk/2 .. 1).  When `emit` the sorted ids / positions are written out.
__global__ void __launch_bounds__(kSortThreads)
sort_local_kernel(const float* X, int64_t ldx, int B, const int64_t* field_meta, uint64_t* keys, int n_pad,
                  int chunk, int mode, int k_fixed, int emit, int32_t* sorted_ids, int32_t* sorted_pos,
                  uint64_t* rx_keys = nullptr, const MmlrecHyper* hyper = nullptr) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* s = reinterpret_cast<uint64_t*>(smem_raw);
  const int f = blockIdx.y;
  const int base = blockIdx.x * chunk;
  uint64_t* kf = keys + (int64_t)f * n_pad;
  const int tid = threadIdx.x;
  if (mode == 0) {
    const int xcol = (int)field_meta[f * 4 + 2];
    const int64_t vocab = field_meta[f * 4 + 1];
    for (int i = tid; i < chunk; i += kSortThreads) {
      int gi = base + i;
      uint64_t key = ~0ull;
      if (gi < B) {
        int64_t id = (int64_t)__ldg(X + (int64_t)gi * ldx + xcol);
        // out-of-range ids are clamped exactly like the gather clamps them (which also raises the oob flag the
        // host turns into an IndexError): the update then goes to the row the forward pass read, never past a table
        if (id < 0) id = 0;
        if (id >= vocab) id = vocab - 1;
        key = ((uint64_t)(uint32_t)id << 32) | (uint32_t)gi;
      }
      s[i] = key;
    }
  } else if (mode == 2) {
    // keys were pushed by the peers into this rank's receive buffer [parity][F_s][B]; consume and reset
    uint64_t* rx = rx_keys + ((int64_t)(hyper->step & 1) * gridDim.y + f) * B;
    for (int i = tid; i < chunk; i += kSortThreads) {
      int gi = base + i;
      uint64_t key = ~0ull;
      if (gi < B) { key = rx[gi]; rx[gi] = ~0ull; }
      s[i] = key;
    }
  } else {
    for (int i = tid; i < chunk; i += kSortThreads) s[i] = kf[base + i];
  }
  __syncthreads();
  const int k_lo = mode != 1 ? 2 : k_fixed;
  const int k_hi = mode != 1 ? chunk : k_fixed;
  for (int k = k_lo; k <= k_hi; k <<= 1) {
    int j0 = k >> 1;
    if (j0 > (chunk >> 1)) j0 = chunk >> 1;
    for (int j = j0; j > 0; j >>= 1) {
      for (int t = tid; t < (chunk >> 1); t += kSortThreads) {
        int i = ((t / j) * (j << 1)) + (t % j);
        bool asc = (((base + i) & k) == 0);
        cmp_swap(s[i], s[i + j], asc);
      }
      __syncthreads();
    }
  }
  if (emit) {
    for (int i = tid; i < chunk; i += kSortThreads) {
      int gi = base + i;
      if (gi < B) {
        sorted_ids[(int64_t)f * B + gi] = (int32_t)(s[i] >> 32);
        sorted_pos[(int64_t)f * B + gi] = (int32_t)(s[i] & 0xffffffffu);
      }
    }
  } else {
    for (int i = tid; i < chunk; i += kSortThreads) kf[base + i] = s[i];
  }
}

__global__ void sort_global_step_kernel(uint64_t* keys, int n_pad, int j, int k) {
  const int f = blockIdx.y;
  uint64_t* kf = keys + (int64_t)f * n_pad;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < (n_pad >> 1); t += gridDim.x * blockDim.x) {
    int i = ((t / j) * (j << 1)) + (t % j);
    uint64_t a = kf[i], b = kf[i + j];
    bool asc = ((i & k) == 0);
    if ((a > b) == asc) { kf[i] = b; kf[i + j] = a; }
  }
}

// ------------------------------------------------------------------------------------------------
// segmented reduce + fused row update
// grid = (ceil(B/256), F_s); thread <-> one sorted position.  A run of equal ids is owned by the
// CTA that contains its first element; if the run leaves the CTA's chunk the whole CTA scans the
// remainder forward.  Within the chunk: segmented inclusive scan with warp shuffles, warps stitched
// through shared memory in order.
// ------------------------------------------------------------------------------------------------
constexpr int kSegThreads = 256;
constexpr int kSegWarps = kSegThreads / 32;

struct SegArgs {
  const float* d_input; int64_t ld; int B;
  const int32_t* sorted_ids; const int32_t* sorted_pos; const int64_t* field_meta; int D;
  float* emb; float* state1; float* state2; int32_t* row_touch; const MmlrecHyper* hyper;
  float* grad_rows_out;
  int64_t parity_stride;   // sharded tables: d_input += (hyper->step & 1) * parity_stride (double-buffered receive rows)
};

template <int W>
__device__ __forceinline__ void load_slice(float (&v)[W], const float* p) {
#pragma unroll
  for (int q = 0; q < W / 4; ++q) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p) + q);
    v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
  }
}

template <int W>
__device__ void seg_pass(const SegArgs& a, int d0, int f, int64_t table_off, int out_col,
                         int p, bool valid, int my_id, int my_pos, bool true_tail, bool owned,
                         int start_lane, bool chunk_open_owned, int chunk_tail_id, int last_tid,
                         float (*tail_s)[8], float (*carry_s)[8], float (*red_s)[8], const MmlrecHyper& hp) {
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int32_t* ids_f = a.sorted_ids + (int64_t)f * a.B;
  const int32_t* pos_f = a.sorted_pos + (int64_t)f * a.B;
  float v[W];
  if (valid) {
    load_slice<W>(v, a.d_input + (int64_t)my_pos * a.ld + out_col + d0);
  } else {
#pragma unroll
    for (int q = 0; q < W; ++q) v[q] = 0.f;
  }
  // segmented inclusive scan inside the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
#pragma unroll
    for (int q = 0; q < W; ++q) {
      float t = __shfl_up_sync(0xffffffffu, v[q], o);
      if (lane - o >= start_lane) v[q] += t;
    }
  }
  // stitch warps: lane 31 publishes the (possibly partial) sum of the warp's last run
  if (lane == 31) {
#pragma unroll
    for (int q = 0; q < W; ++q) tail_s[w][q] = v[q];
  }
  __syncthreads();
  if (tid < W) {
    float c = 0.f;
    // carry into warp 0 is 0 (a run entering the chunk from the left is not owned here)
    carry_s[0][tid] = 0.f;
    for (int ww = 1; ww < kSegWarps; ++ww) {
      // flags for warp ww-1 were stored in red_s[ww-1][0..1] by the caller (open, whole)
      bool open_prev = red_s[ww - 1][0] != 0.f, whole_prev = red_s[ww - 1][1] != 0.f;
      c = open_prev ? tail_s[ww - 1][tid] + (whole_prev ? c : 0.f) : 0.f;
      carry_s[ww][tid] = c;
    }
  }
  __syncthreads();
  // the first run-part of each warp (lanes whose start_lane == 0) receives the carry
  if (start_lane == 0) {
#pragma unroll
    for (int q = 0; q < W; ++q) v[q] += carry_s[w][q];
  }
  // run leaves the chunk: the whole CTA scans the remainder forward (block-uniform branch)
  int n_beyond = 0;  // elements of the chunk's last run that lie beyond the chunk
  if (chunk_open_owned) {
    const int tail_id2 = chunk_tail_id;
    float acc[W];
#pragma unroll
    for (int q = 0; q < W; ++q) acc[q] = 0.f;
    int base = (blockIdx.x + 1) * kSegThreads;
    int all_match = 1;
    while (all_match && base < a.B) {
      int q2 = base + tid;
      bool m = q2 < a.B && ids_f[q2] == tail_id2;
      if (m) {
        float t[W];
        load_slice<W>(t, a.d_input + (int64_t)pos_f[q2] * a.ld + out_col + d0);
#pragma unroll
        for (int q = 0; q < W; ++q) acc[q] += t[q];
      }
      const int n_match = __syncthreads_count(m ? 1 : 0);
      n_beyond += n_match;
      all_match = n_match == kSegThreads;
      base += kSegThreads;
    }
#pragma unroll
    for (int q = 0; q < W; ++q) acc[q] = warp_sum(acc[q]);
    __syncthreads();  // carry_s reads above are done in every warp before tail_s is reused
    if (lane == 0) {
#pragma unroll
      for (int q = 0; q < W; ++q) tail_s[w][q] = acc[q];
    }
    __syncthreads();
    if (tid == last_tid) {
#pragma unroll
      for (int q = 0; q < W; ++q) {
        float t = 0.f;
        for (int ww = 0; ww < kSegWarps; ++ww) t += tail_s[ww][q];
        v[q] += t;
      }
    }
  }
  // apply
  const bool apply = valid && owned && (true_tail || (chunk_open_owned && tid == last_tid));
  if (apply) {
    const int64_t off = table_off + (int64_t)my_id * a.D + d0;
    if (a.grad_rows_out) {  // reported at the run's last sorted position
      const int p_tail = true_tail ? p : p + n_beyond;
#pragma unroll
      for (int q = 0; q < W; ++q) a.grad_rows_out[((int64_t)f * a.B + p_tail) * a.D + d0 + q] = v[q];
    }
    if (a.emb) {
      float pr[W], s1[W], s2[W];
      load_slice<W>(pr, a.emb + off);
      if (a.state1) load_slice<W>(s1, a.state1 + off);
      if (a.state2) load_slice<W>(s2, a.state2 + off);
#pragma unroll
      for (int q = 0; q < W; ++q) {
        float x1 = a.state1 ? s1[q] : 0.f, x2 = a.state2 ? s2[q] : 0.f;
        optimizer_update(pr[q], v[q], x1, x2, hp);
        s1[q] = x1; s2[q] = x2;
      }
#pragma unroll
      for (int q = 0; q < W / 4; ++q) {
        reinterpret_cast<float4*>(a.emb + off)[q] = make_float4(pr[4 * q], pr[4 * q + 1], pr[4 * q + 2], pr[4 * q + 3]);
        if (a.state1) reinterpret_cast<float4*>(a.state1 + off)[q] = make_float4(s1[4 * q], s1[4 * q + 1], s1[4 * q + 2], s1[4 * q + 3]);
        if (a.state2) reinterpret_cast<float4*>(a.state2 + off)[q] = make_float4(s2[4 * q], s2[4 * q + 1], s2[4 * q + 2], s2[4 * q + 3]);
      }
      if (d0 == 0 && a.row_touch) a.row_touch[off / a.D] = hp.step;
    }
  }
  __syncthreads();  // shared scratch is reused by the next pass
}

__global__ void __launch_bounds__(kSegThreads) emb_seg_update_kernel(SegArgs a) {
  if (a.parity_stride) a.d_input += (int64_t)(a.hyper->step & 1) * a.parity_stride;
  __shared__ float tail_s[kSegWarps][8];
  __shared__ float carry_s[kSegWarps][8];
  __shared__ float flag_s[kSegWarps][8];   // [w][0] = open, [w][1] = whole, [w][2] = has true head
  __shared__ MmlrecHyper hp_s;
  const int f = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int p = blockIdx.x * kSegThreads + tid;
  const int32_t* ids_f = a.sorted_ids + (int64_t)f * a.B;
  if (tid == 0) hp_s = *a.hyper;
  // negative ids are the sort's sentinels (slots of the receive buffer this rank does not own)
  const bool valid = p < a.B && ids_f[p] >= 0;
  const int my_id = valid ? ids_f[p] : -1;
  const int my_pos = valid ? a.sorted_pos[(int64_t)f * a.B + p] : 0;
  const int prev_id = (valid && p > 0) ? ids_f[p - 1] : -2;
  const int next_id = (p + 1 < a.B) ? ids_f[p + 1] : -3;
  const bool is_head = valid && (p == 0 || prev_id != my_id);
  const bool true_tail = valid && (next_id != my_id);
  const unsigned head_mask = __ballot_sync(0xffffffffu, is_head);
  // in-warp segment start (lane 0 always starts a part)
  const unsigned le_mask = 0xffffffffu >> (31 - lane);
  const unsigned starts = (head_mask | 1u) & le_mask;
  const int start_lane = 31 - __clz(starts);
  // warp summary flags
  const bool last_lane_open = __shfl_sync(0xffffffffu, valid && !true_tail ? 1 : 0, 31) != 0;
  if (lane == 0) {
    flag_s[w][0] = last_lane_open ? 1.f : 0.f;                 // run continues past lane 31
    flag_s[w][1] = (head_mask == 0u) ? 1.f : 0.f;              // no true head inside this warp
    flag_s[w][2] = (head_mask != 0u) ? 1.f : 0.f;
  }
  __syncthreads();
  // owned: a true head exists at or before me inside the chunk
  bool owned = (head_mask & le_mask) != 0u;
  for (int ww = 0; ww < w; ++ww) owned = owned || (flag_s[ww][2] != 0.f);
  bool any_head = false;
  for (int ww = 0; ww < kSegWarps; ++ww) any_head = any_head || (flag_s[ww][2] != 0.f);
  if (!any_head) return;  // the whole chunk is the inside of a run owned by an earlier CTA
  const int n_valid = min(kSegThreads, a.B - blockIdx.x * kSegThreads);
  const int last_tid = n_valid - 1;
  // block-uniform: does the chunk's last run continue beyond the chunk (and is it owned here)?
  __shared__ int s_open, s_tail_id;
  if (tid == last_tid) { s_open = (valid && !true_tail && owned) ? 1 : 0; s_tail_id = my_id; }
  __syncthreads();
  const bool chunk_open_owned = s_open != 0;
  const int chunk_tail_id = s_tail_id;
  const int64_t table_off = a.field_meta[f * 4 + 0];
  const int out_col = (int)a.field_meta[f * 4 + 3];
  for (int d0 = 0; d0 < a.D; d0 += 8) {
    if (a.D - d0 >= 8)
      seg_pass<8>(a, d0, f, table_off, out_col, p, valid, my_id, my_pos, true_tail, owned, start_lane,
                  chunk_open_owned, chunk_tail_id, last_tid, tail_s, carry_s, flag_s, hp_s);
    else
      seg_pass<4>(a, d0, f, table_off, out_col, p, valid, my_id, my_pos, true_tail, owned, start_lane,
                  chunk_open_owned, chunk_tail_id, last_tid, tail_s, carry_s, flag_s, hp_s);
  }
}

// Stamp every row the batch touches with the current step, straight from the sorted ids.  Lets the
// dense-Adam sweep of the UNtouched rows start before the backward pass (it needs no gradient).
__global__ void emb_stamp_rows_kernel(const int32_t* sorted_ids, const int64_t* field_meta, int F_s, int B, int D,
                                      int32_t* row_touch, const MmlrecHyper* hyper) {
  const int step = hyper->step;
  const int64_t n = (int64_t)F_s * B;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int f = (int)(i / B);
    if (sorted_ids[i] >= 0) row_touch[field_meta[f * 4 + 0] / D + sorted_ids[i]] = step;   // < 0: sort sentinel
  }
}

// Zero-gradient update for every row not touched in this step (dense optimizer semantics of nn.Embedding(sparse=False):
// Adam rows keep moving by their momentum, RMSprop's square_avg keeps decaying; Adagrad / SGD rows are unaffected).
__global__ void emb_adam_sweep_kernel(float* emb, float* m, float* v, const int32_t* row_touch,
                                      int64_t total_rows, int D, const MmlrecHyper* hyper) {
  const MmlrecHyper hp = *hyper;
  const int dv = D >> 2;
  const int64_t n4 = total_rows * dv;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t row = i / dv;
    if (row_touch[row] == hp.step) continue;
    if (hp.optimizer == MMLREC_OPT_RMSPROP) {   // zero gradient: square_avg *= alpha, the row itself does not move
      float4 q = reinterpret_cast<float4*>(m)[i];
      q.x *= hp.alpha; q.y *= hp.alpha; q.z *= hp.alpha; q.w *= hp.alpha;
      reinterpret_cast<float4*>(m)[i] = q;
      continue;
    }
    float4 p4 = reinterpret_cast<float4*>(emb)[i];
    float4 m4 = reinterpret_cast<float4*>(m)[i];
    float4 v4 = reinterpret_cast<float4*>(v)[i];
    adam_zero_grad_step(p4.x, m4.x, v4.x, hp.step_size, hp.bc2_sqrt, hp);
    adam_zero_grad_step(p4.y, m4.y, v4.y, hp.step_size, hp.bc2_sqrt, hp);
    adam_zero_grad_step(p4.z, m4.z, v4.z, hp.step_size, hp.bc2_sqrt, hp);
    adam_zero_grad_step(p4.w, m4.w, v4.w, hp.step_size, hp.bc2_sqrt, hp);
    reinterpret_cast<float4*>(emb)[i] = p4;
    reinterpret_cast<float4*>(m)[i] = m4;
    reinterpret_cast<float4*>(v)[i] = v4;
  }
}

}  // namespace mmlrec

extern "C" int mmlrec_hyper_advance(MmlrecHyper* hyper, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(hyper, "null hyper");
  launch_pdl(hyper_advance_kernel, dim3(1), dim3(1), 0, stream, hyper, (float2*)nullptr, 1);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_hyper_advance_hist(MmlrecHyper* hyper, float* hist, int32_t cap, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(hyper && hist && cap > 1 && (cap & (cap - 1)) == 0, "hist ring must have a power-of-two capacity");
  launch_pdl(hyper_advance_kernel, dim3(1), dim3(1), 0, stream, hyper, reinterpret_cast<float2*>(hist), cap);
  MMLREC_RETURN_LAUNCH(1);
}

static int emb_sm_count() {
  static int n = 0;
  if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); }
  return n > 0 ? n : 148;
}

extern "C" int mmlrec_emb_adam_catch_up(const float* X, int64_t ldx, int32_t B, const int64_t* field_meta, int32_t F_s,
                                        int32_t D, float* emb, float* exp_avg, float* exp_avg_sq, int32_t* row_touch,
                                        const MmlrecHyper* hyper, const float* hist, int32_t cap, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(X && field_meta && emb && exp_avg && exp_avg_sq && row_touch && hyper && hist, "null argument");
  MMLREC_CHECK_ARG(B > 0 && F_s > 0 && D > 0 && (D & 3) == 0 && cap > 1 && (cap & (cap - 1)) == 0, "bad sizes");
  const int64_t n = (int64_t)B * F_s;
  int grid = (int)((n + 255) / 256);
  if (grid > 8 * emb_sm_count()) grid = 8 * emb_sm_count();
  launch_pdl(emb_adam_catch_up_kernel, dim3(grid), dim3(256), 0, stream, X, ldx, B, field_meta, F_s, D, emb, exp_avg, exp_avg_sq,
             row_touch, hyper, reinterpret_cast<const float2*>(hist), cap);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_emb_adam_catch_up_keys(const uint64_t* rq_keys, int32_t B_all, const int64_t* field_meta, int32_t F_s,
                                             int32_t D, float* emb, float* exp_avg, float* exp_avg_sq, int32_t* row_touch,
                                             const MmlrecHyper* hyper, int32_t step_offset, const float* hist, int32_t cap,
                                             void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(rq_keys && field_meta && emb && exp_avg && exp_avg_sq && row_touch && hyper && hist, "null argument");
  MMLREC_CHECK_ARG(B_all > 0 && F_s > 0 && D > 0 && (D & 3) == 0 && cap > 1 && (cap & (cap - 1)) == 0, "bad sizes");
  const int64_t n = (int64_t)B_all * F_s;
  int grid = (int)((n + 255) / 256);
  if (grid > 8 * emb_sm_count()) grid = 8 * emb_sm_count();
  launch_pdl(emb_adam_catch_up_keys_kernel, dim3(grid), dim3(256), 0, stream, rq_keys, B_all, field_meta, F_s, D, emb, exp_avg,
             exp_avg_sq, row_touch, hyper, step_offset, reinterpret_cast<const float2*>(hist), cap);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_emb_adam_flush(float* emb, float* exp_avg, float* exp_avg_sq, int32_t* row_touch, int64_t total_rows,
                                     int32_t D, const MmlrecHyper* hyper, const float* hist, int32_t cap, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(emb && exp_avg && exp_avg_sq && row_touch && hyper && hist, "null argument");
  MMLREC_CHECK_ARG(total_rows >= 0 && D > 0 && (D & 3) == 0 && cap > 1 && (cap & (cap - 1)) == 0, "bad sizes");
  if (total_rows == 0) return 0;
  int grid = (int)((total_rows + 255) / 256);
  if (grid > 8 * emb_sm_count()) grid = 8 * emb_sm_count();
  emb_adam_flush_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(emb, exp_avg, exp_avg_sq, row_touch, total_rows, D, hyper,
                                                               reinterpret_cast<const float2*>(hist), cap);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_sort_field_ids(const float* X, int64_t ldx, int32_t B, const int64_t* field_meta, int32_t F_s,
                                     int32_t* sorted_ids, int32_t* sorted_pos, uint64_t* keys_ws, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(B > 0 && F_s > 0, "bad sizes");
  int n_pad = 32;
  while (n_pad < B) n_pad <<= 1;
  const int chunk = n_pad < kSortChunk ? n_pad : kSortChunk;
  const size_t smem = (size_t)chunk * sizeof(uint64_t);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n_pad / chunk, F_s);
  const bool single = n_pad <= kSortChunk;
  sort_local_kernel<<<grid, kSortThreads, smem, st>>>(X, ldx, B, field_meta, keys_ws, n_pad, chunk, 0, 0,
                                                      single ? 1 : 0, sorted_ids, sorted_pos);
  MMLREC_CHECK_LAUNCH(1);
  for (int k = chunk << 1; k <= n_pad && !single; k <<= 1) {
    for (int j = k >> 1; j >= chunk; j >>= 1) {
      dim3 g2(cdiv(n_pad >> 1, 256) < 1184 ? cdiv(n_pad >> 1, 256) : 1184, F_s);
      sort_global_step_kernel<<<g2, 256, 0, st>>>(keys_ws, n_pad, j, k);
      MMLREC_CHECK_LAUNCH(1);
    }
    sort_local_kernel<<<grid, kSortThreads, smem, st>>>(X, ldx, B, field_meta, keys_ws, n_pad, chunk, 1, k,
                                                        k == n_pad ? 1 : 0, sorted_ids, sorted_pos);
    MMLREC_CHECK_LAUNCH(1);
  }
  return 0;
}

// sort of the keys the peers pushed into this rank's receive buffer (peer.cu); slots nobody wrote hold the
// sentinel ~0 and sort to the end (sorted id -1); the consumed half of the buffer is reset to the sentinel
extern "C" int mmlrec_sort_field_keys(uint64_t* rx_keys, int32_t B_all, int32_t F_s, const MmlrecHyper* hyper,
                                      int32_t* sorted_ids, int32_t* sorted_pos, uint64_t* keys_ws, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(rx_keys && hyper && B_all > 0 && F_s > 0, "bad args");
  int n_pad = 32;
  while (n_pad < B_all) n_pad <<= 1;
  const int chunk = n_pad < kSortChunk ? n_pad : kSortChunk;
  const size_t smem = (size_t)chunk * sizeof(uint64_t);
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(n_pad / chunk, F_s);
  const bool single = n_pad <= kSortChunk;
  sort_local_kernel<<<grid, kSortThreads, smem, st>>>(nullptr, 0, B_all, nullptr, keys_ws, n_pad, chunk, 2, 0,
                                                      single ? 1 : 0, sorted_ids, sorted_pos, rx_keys, hyper);
  MMLREC_CHECK_LAUNCH(1);
  for (int k = chunk << 1; k <= n_pad && !single; k <<= 1) {
    for (int j = k >> 1; j >= chunk; j >>= 1) {
      dim3 g2(cdiv(n_pad >> 1, 256) < 1184 ? cdiv(n_pad >> 1, 256) : 1184, F_s);
      sort_global_step_kernel<<<g2, 256, 0, st>>>(keys_ws, n_pad, j, k);
      MMLREC_CHECK_LAUNCH(1);
    }
    sort_local_kernel<<<grid, kSortThreads, smem, st>>>(nullptr, 0, B_all, nullptr, keys_ws, n_pad, chunk, 1, k,
                                                        k == n_pad ? 1 : 0, sorted_ids, sorted_pos);
    MMLREC_CHECK_LAUNCH(1);
  }
  return 0;
}

extern "C" int mmlrec_emb_backward_update(const float* d_input, int64_t ld, int32_t B, const int32_t* sorted_ids,
                                          const int32_t* sorted_pos, const int64_t* field_meta, int32_t F_s, int32_t D,
                                          float* emb, float* state1, float* state2, int32_t* row_touch,
                                          const MmlrecHyper* hyper, float* grad_rows_out, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(B > 0 && F_s > 0 && D > 0 && (D & 3) == 0, "bad sizes");
  MMLREC_CHECK_ARG((ld & 3) == 0, "d_input row stride must be a multiple of 4 floats");
  MMLREC_CHECK_ARG(hyper != nullptr, "null hyper");
  MMLREC_CHECK_ARG(emb != nullptr || grad_rows_out != nullptr, "nothing to do");
  SegArgs a{d_input, ld, B, sorted_ids, sorted_pos, field_meta, D, emb, state1, state2, row_touch, hyper, grad_rows_out, 0};
  dim3 grid(cdiv(B, kSegThreads), F_s);
  emb_seg_update_kernel<<<grid, kSegThreads, 0, (cudaStream_t)stream>>>(a);
  MMLREC_RETURN_LAUNCH(1);
}

// owner side of the row-sharded tables (peer.cu): same kernel over the receive rows d_rx [B_all][ld] the peers
// filled with gradient rows; slots this rank does not own carry sentinel ids (< 0) and are skipped
extern "C" int mmlrec_emb_backward_update_sharded(const float* d_rx, int64_t ld, int32_t B_all, const int32_t* sorted_ids,
                                                  const int32_t* sorted_pos, const int64_t* field_meta, int32_t F_s,
                                                  int32_t D, float* emb, float* state1, float* state2, int32_t* row_touch,
                                                  const MmlrecHyper* hyper, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(B_all > 0 && F_s > 0 && D > 0 && (D & 3) == 0 && (ld & 3) == 0, "bad sizes");
  MMLREC_CHECK_ARG(d_rx && emb && hyper, "null argument");
  SegArgs a{d_rx, ld, B_all, sorted_ids, sorted_pos, field_meta, D, emb, state1, state2, row_touch, hyper, nullptr, 0};
  dim3 grid(cdiv(B_all, kSegThreads), F_s);
  emb_seg_update_kernel<<<grid, kSegThreads, 0, (cudaStream_t)stream>>>(a);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_emb_stamp_rows(const int32_t* sorted_ids, const int64_t* field_meta, int32_t F_s, int32_t B, int32_t D,
                                     int32_t* row_touch, const MmlrecHyper* hyper, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(F_s > 0 && B > 0 && D > 0 && row_touch && hyper, "bad args");
  const int64_t n = (int64_t)F_s * B;
  int grid = (int)((n + 255) / 256 < 1184 ? (n + 255) / 256 : 1184);
  emb_stamp_rows_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(sorted_ids, field_meta, F_s, B, D, row_touch, hyper);
  MMLREC_RETURN_LAUNCH(1);
}

extern "C" int mmlrec_emb_adam_dense_sweep(float* emb, float* exp_avg, float* exp_avg_sq, const int32_t* row_touch,
                                           int64_t total_rows, int32_t D, const MmlrecHyper* hyper, void* stream) {
  using namespace mmlrec;
  MMLREC_CHECK_ARG(total_rows >= 0 && D > 0 && (D & 3) == 0, "bad sizes");
  if (total_rows == 0) return 0;
  int64_t n4 = total_rows * (D >> 2);
  const int cap_grid = emb_sm_count() * 8;
  int grid = (int)((n4 + 255) / 256 < cap_grid ? (n4 + 255) / 256 : cap_grid);
  emb_adam_sweep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(emb, exp_avg, exp_avg_sq, row_touch, total_rows, D, hyper);
  MMLREC_RETURN_LAUNCH(1);
}
