/*
 * mmlrec_b200.h -- C ABI of the B200-native MMLRec training hot path (libmmlrec_b200.so).
 *
 * The reference (/root/reference) is pure Python: it has no FFI of its own.  Each entry point
 * below replaces the ATen call sequence behind one reference function; the citation names the
 * reference lines it stands in for.  All pointers are DEVICE pointers unless the name ends in
 * _host; `stream` is a cudaStream_t passed as void*; every call is stream-ordered, asynchronous
 * and performs no allocation and no host synchronisation (so a whole step can be captured in a
 * CUDA graph).  Return value: 0 on success, otherwise a cudaError_t (>0) or -1 for a rejected
 * argument; mmlrec_last_error() describes the most recent failure of the calling thread.
 *
 * Matrices are row-major.  "ld" is the row stride in ELEMENTS.  fp32 everywhere except where a
 * parameter is spelled bf16 (raw uint16_t bit patterns).
 */
#ifndef MMLREC_B200_H_
#define MMLREC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMLREC_ABI_VERSION 1
#define MMLREC_MAX_GATE_EXPERTS 32
#define MMLREC_MAX_TASKS 16

/* activation codes (model/utils.py:10-37 activation_layer; pepnet.py:31-32 GateNN's 2*sigmoid) */
enum { MMLREC_ACT_NONE = 0, MMLREC_ACT_RELU = 1, MMLREC_ACT_SIGMOID = 2, MMLREC_ACT_SIGMOID2 = 3 };
/* optimizer codes (model/basemodel.py:569-584 _get_optim) */
enum { MMLREC_OPT_SGD = 0, MMLREC_OPT_ADAGRAD = 1, MMLREC_OPT_ADAM = 2, MMLREC_OPT_RMSPROP = 3 };
/* head kinds (model/utils.py:242-248 PredictionLayer + basemodel.py:595-604 loss) */
enum { MMLREC_HEAD_SIGMOID_BCE = 0, MMLREC_HEAD_IDENTITY_MSE = 1 };

int         mmlrec_abi_version(void);
const char* mmlrec_last_error(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
int64_t     mmlrec_launch_count(void);
/* sizeof of ABI struct `which` (0 Hyper, 1 GemmF32, 2 GemmTcDesc, 3 Gate, 4 ExpertGrad, 5 Head, 6 GateLevel) */
int64_t     mmlrec_struct_size(int32_t which);

/* ---------------------------------------------------------------------------------------------
 * Optimizer clock.  torch.optim keeps `step` per parameter and derives Adam's bias corrections
 * from it on the host (torch/optim/adam.py _single_tensor_adam).  To keep the step graph-
 * capturable the clock lives on the device: one MmlrecHyper per optimizer, advanced by a kernel.
 * ------------------------------------------------------------------------------------------- */
typedef struct MmlrecHyper {
  int32_t step;        /* number of optimizer steps taken (t) */
  int32_t optimizer;   /* MMLREC_OPT_* */
  float   lr, beta1, beta2, eps;     /* eps: 1e-8 Adam, 1e-10 Adagrad, 1e-8 RMSprop */
  float   step_size;   /* Adam: lr / (1 - beta1^t) */
  float   bc2_sqrt;    /* Adam: sqrt(1 - beta2^t) */
  float   alpha;       /* RMSprop smoothing (0.99) */
  float   one_minus_beta1, one_minus_beta2, one_minus_alpha;  /* rounded from double, like torch's python scalars */
  double  lr_d, beta1_d, beta2_d;    /* the python-float hyper-parameters; bias corrections are formed in double */
} MmlrecHyper;
/* step += 1 and refresh step_size / bc2_sqrt (double precision pow, like the host code it replaces) */
int mmlrec_hyper_advance(MmlrecHyper* hyper, void* stream);
/* same, and records {step_size, bc2_sqrt} of the new step t in hist[(t mod cap)] (float pairs, cap a power of two):
 * the history the lazy dense-Adam catch-up replays */
int mmlrec_hyper_advance_hist(MmlrecHyper* hyper, float* hist, int32_t cap, void* stream);
/* The same catch-up on the owner of ROW-SHARDED tables: the rows are named by the request keys of the forward exchange
 * (rq_keys[2][F_s][B_all], key = local_row << 32 | pos, ~0 = not this owner's; parity = (step + step_offset) & 1); runs
 * between the ids barrier and mmlrec_emb_serve_rows.  No reference counterpart (torch.optim.Adam is dense,
 * model/basemodel.py:569-584); see mmlrec_emb_adam_catch_up. */
int mmlrec_emb_adam_catch_up_keys(const uint64_t* rq_keys, int32_t B_all, const int64_t* field_meta, int32_t F_s,
                                  int32_t D, float* emb, float* exp_avg, float* exp_avg_sq, int32_t* row_touch,
                                  const MmlrecHyper* hyper, int32_t step_offset, const float* hist, int32_t cap,
                                  void* stream);
/* Exact LAZY dense Adam on the embedding tables (the reference's nn.Embedding(sparse=False) + torch.optim.Adam moves
 * EVERY row every step, model/utils.py:475-479; a row the batch does not touch still takes a zero-gradient
 * step).  Instead of sweeping the whole table every step, a row remembers the step it is current for (row_touch) and
 * the missed zero-gradient steps are replayed -- same fp32 operations, the recorded per-step factors -- when the row is
 * next needed.  catch_up: at the start of step t (hyper already advanced) every row named by X's id columns is
 * brought to step t-1 (the forward gather then reads exact values; K2 later applies step t and stamps the row).
 * flush: every row is brought to the current step (before predict / state_dict / a checkpoint, and before the history
 * ring wraps).  Bit-identical to mmlrec_emb_adam_dense_sweep after every step (tests/test_kernels_gpu.py). */
int mmlrec_emb_adam_catch_up(const float* X, int64_t ldx, int32_t B, const int64_t* field_meta, int32_t F_s, int32_t D,
                             float* emb, float* exp_avg, float* exp_avg_sq, int32_t* row_touch, const MmlrecHyper* hyper,
                             const float* hist, int32_t cap, void* stream);
int mmlrec_emb_adam_flush(float* emb, float* exp_avg, float* exp_avg_sq, int32_t* row_touch, int64_t total_rows, int32_t D,
                          const MmlrecHyper* hyper, const float* hist, int32_t cap, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1: multi-field gather + concat.
 * Replaces BaseModel.input_from_feature_columns (model/basemodel.py:461-487: per field
 * X[:, j:j+1].long() -> nn.Embedding) followed by combined_dnn_input (model/utils.py:434-446).
 *   X            [B, ldx] fp32; sparse ids are carried as fp32 and truncated toward zero.
 *   emb          flat storage holding every table; field f's table starts at element
 *                field_meta[f*4+0] and has field_meta[f*4+1] rows of D floats.
 *   field_meta   int64 [F_s,4] = {table_offset, vocabulary, x_column, out_column}
 *   dense_xcol   int32 [F_d]: X column of each dense feature; written to out columns
 *                dense_out_col .. dense_out_col+F_d-1.
 *   out_f32      [B, ld_f32] (nullable); out_bf16 [B, ld_bf16] (nullable; columns >= in_dim up to
 *                ld_bf16 are written as zero so the buffer can feed a K-padded tensor-core GEMM).
 *   oob_flag     nullable int32: set to 1 if any id fell outside [0, vocabulary) (the id is
 *                clamped; the reference would raise IndexError).
 * ------------------------------------------------------------------------------------------- */
int mmlrec_gather_concat(const float* X, int64_t ldx, int32_t B,
                         const float* emb, const int64_t* field_meta, int32_t F_s, int32_t D,
                         const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                         float* out_f32, int64_t ld_f32,
                         uint16_t* out_bf16, int64_t ld_bf16,
                         int32_t* oob_flag, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2: embedding backward as sort + segmented warp-shuffle reduce with the optimizer's row
 * update fused in.  Replaces F_s x embedding_dense_backward (autograd of basemodel.py:475-477,
 * which materialises a dense [V,D] gradient per table) plus the torch.optim update of every
 * table (basemodel.py:313).
 *
 * mmlrec_sort_field_ids: per field, stable sort of the batch's ids.
 *   sorted_ids/sorted_pos int32 [F_s, B]; keys_ws: uint64 workspace of F_s * n_pad elements,
 *   n_pad = B rounded up to a power of two (>= 32).
 * mmlrec_emb_backward_update: for every run of equal ids, sum the gradient slices
 *   d_input[pos, out_column : out_column+D] in ascending pos order (deterministic, no atomics)
 *   and apply the optimizer to that row.  state1: Adagrad sum / Adam exp_avg / RMSprop
 *   square_avg; state2: Adam exp_avg_sq.  row_touch (nullable, int32 per table row over the whole
 *   flat storage, i.e. index = element_offset / D) is stamped with hyper->step for Adam's
 *   dense sweep.
 * mmlrec_emb_adam_dense_sweep: torch's Adam is dense (nn.Embedding(sparse=False)): rows with
 *   zero gradient still move by their momentum.  This applies that zero-gradient update to every
 *   row NOT stamped in this step (exact dense-Adam semantics).
 * ------------------------------------------------------------------------------------------- */
int mmlrec_sort_field_ids(const float* X, int64_t ldx, int32_t B, const int64_t* field_meta, int32_t F_s,
                          int32_t* sorted_ids, int32_t* sorted_pos, uint64_t* keys_ws, void* stream);
int mmlrec_emb_backward_update(const float* d_input, int64_t ld, int32_t B,
                               const int32_t* sorted_ids, const int32_t* sorted_pos,
                               const int64_t* field_meta, int32_t F_s, int32_t D,
                               float* emb, float* state1, float* state2, int32_t* row_touch,
                               const MmlrecHyper* hyper, float* grad_rows_out, void* stream);
/* ---------------------------------------------------------------------------------------------
 * Row-sharded tables over NVLink peer memory (BASELINE config 5: 26 x 10M-row tables on
 * 8 GPUs).  owner(id) = id mod R keeps row id at local row id / R; every rank lays its shard out with
 * the same per-field offsets (field_meta[f*4+0], rows = ceil(vocabulary / R)).  The reference has no
 * multi-GPU path: the contract is the single-process step at the global batch.
 * Every exchange is a one-sided store into a peer's buffer at a slot fixed by (rank, sample, field):
 *   mmlrec_peer_*             cudaMalloc + CUDA IPC export / import; mmlrec_peer_barrier = flag exchange
 *                             through peer memory (err_flag := 2 after ~10 s without the peers).
 *   mmlrec_emb_push_ids       [ids all-to-all] key (id / R) << 32 | pos, pos = rank * b + i, into the
 *                             OWNER's rq_keys [2][F_s][B_all] uint64 (sentinel ~0; half = hyper->step & 1).
 *   mmlrec_emb_serve_rows     [rows all-to-all] owner: every received key -> its local row, stored into the
 *                             REQUESTER's staging rows rows_in[r] [b][F_s*D].
 *   mmlrec_gather_concat_staged   K1 assembling dnn_input from the staging rows (+ dense columns).
 *   mmlrec_gather_concat_sharded  alternative forward without barriers: K1 reads every row straight from
 *                             its owner's shard (fast while the shards fit the peer TLB reach).
 *   mmlrec_sort_field_keys    owner: per-field sort of the received keys (sentinels last, id -1),
 *                             resetting the consumed half to the sentinel.
 *   mmlrec_emb_push_grads     [row-grad all-to-all] D gradient floats of every local (sample, field) into
 *                             the owner's rx_grad [B_all][F_s*D] at row pos.
 *   mmlrec_emb_backward_update_sharded   K2 over rx_grad with the sorted received ids (negative skipped).
 * Order per step: push_ids, barrier, serve_rows, barrier, K1 ... push_grads, dense all-reduce (barrier), K2.
 * ------------------------------------------------------------------------------------------- */
int mmlrec_peer_alloc(void** ptr, int64_t bytes);                 /* zero-filled */
int mmlrec_peer_free(void* ptr);
int mmlrec_peer_export(void* ptr, unsigned char* handle64);       /* 64-byte cudaIpcMemHandle_t */
int mmlrec_peer_import(const unsigned char* handle64, void** ptr);
int mmlrec_peer_close(void* ptr);
int mmlrec_peer_fill_u64(uint64_t* p, int64_t n, uint64_t v /* 0 or ~0 */, void* stream);
/* SUM all-reduce of `n` floats (n % 4 == 0) over peer memory: rank r reduces slice r of every rank's `peer_in[p]` in rank
 * order (bit-identical result on every rank) and stores it into slice r of every `peer_out[p]`.  Bracket with
 * mmlrec_peer_barrier on both sides.  The data-parallel towers' dense-gradient collective (no reference counterpart: the
 * reference is single-process; the gradient is what loss.backward() leaves in .grad, model/basemodel.py:309-311). */
int mmlrec_peer_allreduce_f32(const float* const* peer_in, float* const* peer_out, int64_t n, int32_t rank, int32_t R,
                              void* stream);
int mmlrec_peer_barrier(int32_t* const* peer_flags /* [R] -> int32 [R] */, int32_t* local_epoch, int32_t rank,
                        int32_t R, int32_t* err_flag, void* stream);
int mmlrec_emb_push_ids(const float* X, int64_t ldx, int32_t b, const int64_t* field_meta, int32_t F_s,
                        int32_t rank, int32_t R, int32_t B_all, uint64_t* const* rq_keys,
                        const MmlrecHyper* hyper, int32_t step_offset, int32_t* oob_flag, void* stream);
int mmlrec_emb_serve_rows(const uint64_t* rq_keys, const float* emb, const int64_t* field_meta, int32_t F_s,
                          int32_t D, int32_t b, int32_t B_all, float* const* rows_in,
                          const MmlrecHyper* hyper, int32_t step_offset, void* stream);
int mmlrec_gather_concat_staged(const float* X, int64_t ldx, int32_t B, const float* staged_rows,
                                const int64_t* field_meta, int32_t F_s, int32_t D,
                                const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                                float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16, void* stream);
int mmlrec_gather_concat_sharded(const float* X, int64_t ldx, int32_t B,
                                 const float* const* shards, int32_t n_shards,
                                 const int64_t* field_meta, int32_t F_s, int32_t D,
                                 const int32_t* dense_xcol, int32_t F_d, int32_t dense_out_col,
                                 float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                                 int32_t* oob_flag, void* stream);
int mmlrec_emb_push_grads(const float* X, int64_t ldx, int32_t b, const float* d_input, int64_t ld,
                          const int64_t* field_meta, int32_t F_s, int32_t D, int32_t rank, int32_t R,
                          int32_t B_all, float* const* rx_grad, void* stream);
int mmlrec_sort_field_keys(uint64_t* rq_keys, int32_t B_all, int32_t F_s, const MmlrecHyper* hyper,
                           int32_t* sorted_ids, int32_t* sorted_pos, uint64_t* keys_ws, void* stream);
int mmlrec_emb_backward_update_sharded(const float* d_rx, int64_t ld, int32_t B_all,
                                       const int32_t* sorted_ids, const int32_t* sorted_pos,
                                       const int64_t* field_meta, int32_t F_s, int32_t D,
                                       float* emb, float* state1, float* state2, int32_t* row_touch,
                                       const MmlrecHyper* hyper, void* stream);

/* stamp row_touch[row] = hyper->step for every row the sorted batch ids name (lets the sweep below run
 * BEFORE / concurrently with the backward pass: untouched rows need no gradient) */
int mmlrec_emb_stamp_rows(const int32_t* sorted_ids, const int64_t* field_meta, int32_t F_s, int32_t B, int32_t D,
                          int32_t* row_touch, const MmlrecHyper* hyper, void* stream);
int mmlrec_emb_adam_dense_sweep(float* emb, float* exp_avg, float* exp_avg_sq, const int32_t* row_touch,
                                int64_t total_rows, int32_t D, const MmlrecHyper* hyper, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3 (fp32 parity mode): grouped GEMM on the CUDA cores.  One launch runs a list of independent
 * problems C = act(A * B^T + bias) with arbitrary element strides, which covers forward
 * (A = activations [M,K], B = nn.Linear weight [N,K]; model/utils.py:146-161 DNN.forward),
 * dgrad (A = dZ [M,N], B = W^T via strides) and wgrad (A = dZ^T, B = X^T via strides).
 *   mask/ldmask: out = mask(m,n) > 0 ? out : 0  -- the ReLU backward of the layer that produced
 *                this GEMM's input (fused threshold_backward).
 *   rowsum_a:    nullable [M]: sum over k of A(m,k); in a wgrad problem A = dZ^T so this is the
 *                bias gradient.
 *   accumulate:  C += result.
 * The problem table lives in device memory; tile_prefix (int32 [n_problems+1], device) is the
 * exclusive prefix of 64x64 tile counts.
 * ------------------------------------------------------------------------------------------- */
typedef struct MmlrecGemmF32 {
  const float* A; const float* B; float* C;
  int64_t a_rs, a_cs;    /* A(m,k) = A[m*a_rs + k*a_cs] */
  int64_t b_rs, b_cs;    /* B(n,k) = B[n*b_rs + k*b_cs] */
  int64_t ldc;
  const float* bias;     /* [N] or NULL */
  const float* mask; int64_t ldmask;
  float* rowsum_a;       /* [M] or NULL */
  int32_t M, N, K;
  int32_t act;           /* MMLREC_ACT_* */
  int32_t accumulate;
  int32_t reserved;
} MmlrecGemmF32;
int mmlrec_gemm_grouped_f32(const MmlrecGemmF32* problems, const int32_t* tile_prefix,
                            int32_t n_problems, int32_t total_tiles, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3 (bf16 mode): grouped GEMM on the tcgen05 tensor cores, TMA-fed, fp32 accumulation in TMEM.
 * D[M,N] = epilogue(A * B^T): A and B are bf16, each either K-major (row-major [rows,K]) or
 * MN-major ([K,rows], i.e. the transpose of a row-major activation -- what wgrad needs), so no
 * transposed copies are ever materialised.  Tensor maps are encoded on the host at plan time
 * (mmlrec_tc_encode_problem) into an array of device-resident problem records.
 * ------------------------------------------------------------------------------------------- */
typedef struct MmlrecGemmTcDesc {           /* host-side description of one problem */
  const uint16_t* A; const uint16_t* B;     /* bf16 */
  int64_t lda, ldb;                         /* row stride (elements) of the row-major arrays */
  int32_t a_mn_major, b_mn_major;           /* 0: array is [rows,K]; 1: array is [K,rows] */
  int32_t M, N, K;
  float* C_f32; int64_t ldc_f32;            /* nullable outputs, any combination */
  uint16_t* C_bf16; int64_t ldc_bf16;
  const float* bias;                        /* [N] fp32 or NULL */
  const uint16_t* mask; int64_t ldmask;     /* bf16 [M,N]: keep where mask > 0 */
  float* colsum;                            /* nullable [N]: column sums of the fp32 result, atomically accumulated */
  int32_t act, accumulate;                  /* accumulate applies to C_f32 only */
  /* ReLU bit masks (1 bit per element, used by the CTA-pair kernel; the one-CTA kernel ignores them and needs `mask`).
   * Layout of a bit array over a row-major [rows, 32 * chunks] activation: word ((row / 32) * chunks + col / 32) * 32
   * + row % 32 holds the 32 columns [32 * (col / 32), +32) of that row (column j of the chunk at bit j / 2 + 16 * (j % 2):
   * the two halves of a packed bf16 pair sit 16 bits apart), so the 32 rows a warp handles are 128 contiguous bytes.  `relu_bits_out`: the epilogue also stores "result > 0" for its outputs (forward of a ReLU layer);
   * `mask_bits`: applied instead of `mask` (dgrad through that ReLU).  *_chunks = words per row block, *_chunk0 = the
   * chunk of this problem's column 0 (its first column must be a multiple of 32). */
  uint32_t* relu_bits_out; int32_t bits_out_chunks, bits_out_chunk0;
  const uint32_t* mask_bits; int32_t mask_bits_chunks, mask_bits_chunk0;
  /* CTA-pair kernel only: store the fp32 result TRANSPOSED, C_f32[n * ldc_f32 + m] = D[m][n] (no bias / mask / activation /
   * accumulate / colsum).  Lets a weight gradient dW[N_out, K_in] with N_out <= 128 be computed as the product
   * X^T dZ (M = K_in: full 256-row pair tiles) instead of dZ^T X (M = N_out: half of every pair tile empty). */
  int32_t c_transposed, pad1;
  /* with c_transposed (N <= 128): also out[n] = sum_k B[n][k] -- the bias gradient when B = dZ -- from one extra MMA per
   * k-step against an all-ones A tile, accumulated in the spare TMEM columns of the stage */
  float* colsum_b;
} MmlrecGemmTcDesc;
/* size in bytes of one device problem record; the table is `n * mmlrec_tc_record_bytes()` */
int64_t mmlrec_tc_record_bytes(void);
/* encode problem `desc_host` into `record_host` (host staging memory, record_bytes long) */
int mmlrec_tc_encode_problem(const MmlrecGemmTcDesc* desc_host, void* record_host);
int mmlrec_gemm_grouped_tc(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                           int32_t total_tiles, void* stream);
/* same launch with a caller-supplied static schedule: CTA b (of n_ctas) executes tiles
 * tile_order[cta_start[b] .. cta_start[b+1]) in that order (int32 device arrays).  Lets the host balance long
 * (wgrad, K = batch) and short (dgrad, K = layer width) tiles across the SMs; deterministic, no atomics. */
int mmlrec_gemm_grouped_tc_scheduled(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                     int32_t total_tiles, const int32_t* tile_order, const int32_t* cta_start,
                                     int32_t n_ctas, void* stream);
/* SM count of the current device (the scheduler's CTA budget); 0 without a device */
int32_t mmlrec_tc_sm_count(void);
/* same launch (tile_order / cta_start may be NULL = round-robin), plus profiling stamps (int64, device,
 * 1024 + 8 * n_ctas entries): stamps[i*16 + slot] = clock64() of CTA 0 at its i-th tile (i < 64): slots 0-3 TMA
 * producer (tile start, table read, first slot free, last load issued), 4-7 MMA issuer (tile start, accumulator
 * free, first operands landed, last commit), 8-12 epilogue warp 0 (tile start, bias staged, accumulator ready, tile
 * stored, accumulator released); stamps[1024 + cta*8 + k] = %globaltimer (ns) of every CTA: 0 setup done, 1 producer
 * done, 2 MMA issuer done, 3 epilogue done, and [4] = its number of tiles.  Profiling aid. */
int mmlrec_gemm_grouped_tc_debug(const void* records, const int32_t* tile_prefix, int32_t n_problems,
                                 int32_t total_tiles, const int32_t* tile_order, const int32_t* cta_start,
                                 int32_t n_ctas, int64_t* stamps, void* stream);
/* CTA-pair version of the same grouped GEMM (csrc/gemm_tc2.cu): clusters of two CTAs run tcgen05.mma.cta_group::2 on
 * 256 x BN tiles (BN = 256, or 128 for narrow problems / bias-gradient problems wider than 240 columns), which halves
 * the operand bytes each SM pulls from L2 per output element.  Same problem description, its own record format.
 * The schedule arrays are per PAIR: pair p executes tiles tile_order[pair_start[p] .. pair_start[p+1]); NULL =
 * round-robin.  `stamps` (nullable, 1024 + 8 * 2 * n_pairs int64) receives the per-CTA %globaltimer stamps described
 * at mmlrec_gemm_grouped_tc_debug. */
int64_t mmlrec_tc2_record_bytes(void);
int32_t mmlrec_tc2_num_tiles(const MmlrecGemmTcDesc* desc_host);
int mmlrec_tc2_encode_problem(const MmlrecGemmTcDesc* desc_host, void* record_host);
int mmlrec_gemm_grouped_tc2(const void* records, const int32_t* tile_prefix, int32_t n_problems, int32_t total_tiles,
                            const int32_t* tile_order, const int32_t* pair_start, int32_t n_pairs, int64_t* stamps,
                            void* stream);
/* tiles a problem occupies (BLOCK_M=128 x BLOCK_N=128) */
int32_t mmlrec_tc_num_tiles(int32_t M, int32_t N);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm1d in training mode + activation (model/utils.py:132-134, :153-154; momentum 0.1,
 * eps 1e-5, unbiased running variance).  Z is the Linear output [M, ldz]; columns [0,N).
 * Forward writes Y = act(bn(Z)) and saves mean / invstd for backward.  Backward takes dY (already
 * masked by the activation derivative) and returns dZ plus dgamma / dbeta.
 * ------------------------------------------------------------------------------------------- */
int mmlrec_bn_forward(const float* Z, int64_t ldz, int32_t M, int32_t N,
                      const float* gamma, const float* beta, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, int32_t n_tracked,
                      float* save_mean, float* save_invstd,
                      float* Y, int64_t ldy, uint16_t* Y_bf16, int64_t ldy_bf16,
                      int32_t act, int32_t training, void* stream);
int mmlrec_bn_backward(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int32_t M, int32_t N,
                       const float* gamma, const float* save_mean, const float* save_invstd,
                       float* dZ, int64_t lddz, uint16_t* dZ_bf16, int64_t lddz_bf16,
                       float* dgamma, float* dbeta, void* stream);
/* Synchronised BatchNorm for the data-parallel step (R ranks of M rows each: the statistics of the GLOBAL batch, which is
 * what the single-process reference normalises over, model/utils.py:129-134).
 *   forward : mmlrec_bn_stats (stats[0:N] = mean_r, stats[N:2N] = sum (z - mean_r)^2) -> all-gather to [R][2N] ->
 *             mmlrec_bn_combine (global mean / invstd into save_*, running statistics, counters) ->
 *             mmlrec_bn_forward(..., training = 2) (normalise with the given save_mean / save_invstd)
 *   backward: mmlrec_bn_backward_sums (sums[0:N] = sum dy, sums[N:2N] = sum dy * xhat of this rank; also this rank's
 *             dgamma / dbeta, which the dense-gradient all-reduce adds up) -> SUM all-reduce of sums ->
 *             mmlrec_bn_backward_synced (dZ from the global sums, M_total = R * M) */
int mmlrec_bn_stats(const float* Z, int64_t ldz, int32_t M, int32_t N, float* stats, void* stream);
int mmlrec_bn_combine(const float* all_stats, int32_t R, int32_t M, int32_t N, float* running_mean, float* running_var,
                      int64_t* num_batches_tracked, int32_t n_tracked, float* save_mean, float* save_invstd, void* stream);
int mmlrec_bn_backward_sums(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int32_t M, int32_t N,
                            const float* save_mean, const float* save_invstd, float* sums, float* dgamma, float* dbeta,
                            void* stream);
int mmlrec_bn_backward_synced(const float* dY, int64_t lddy, const float* Z, int64_t ldz, int32_t M, int32_t N,
                              const float* gamma, const float* save_mean, const float* save_invstd,
                              float* dZ, int64_t lddz, uint16_t* dZ_bf16, int64_t lddz_bf16,
                              const float* global_sums, int32_t M_total, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Gate head + softmax + expert mixture (model/mmoe.py:80-88, model/ple.py:127-152):
 *   logits = gate_in * Wg^T (bias-free), p = softmax(logits), mix = sum_e p_e * expert_e.
 * One launch serves every gate of a level; records live in device memory.
 * Backward kernel A (per gate): dlogits, d(gate_in) (optionally ReLU-masked by gate_in > 0,
 * optionally accumulated), dWg.  Backward kernel B (per expert): d(expert_e) = sum over the gates
 * that use it of p * d(mix), optionally ReLU-masked by expert_e > 0.
 * ------------------------------------------------------------------------------------------- */
typedef struct MmlrecGate {
  const float* gate_in; int64_t ld_gate_in; int32_t Hg; int32_t n_e;
  const float* Wg; int64_t ld_Wg;                    /* [n_e, Hg], row stride ld_Wg (dWg uses the same stride) */
  const float* expert[MMLREC_MAX_GATE_EXPERTS]; int64_t ld_expert; int32_t H; int32_t pad0;
  float* probs;                                      /* [B, n_e] saved for backward */
  float* mix; int64_t ld_mix;                        /* [B, H] */
  uint16_t* mix_bf16; int64_t ld_mix_bf16;           /* nullable bf16 copy (feeds tensor-core GEMMs) */
  /* backward */
  const float* d_mix; int64_t ld_d_mix;              /* NULL: this gate receives no gradient */
  float* d_gate_in; int64_t ld_d_gate_in; int32_t relu_mask_gate_in; int32_t accumulate_d_gate_in;
  uint16_t* d_gate_in_bf16; int64_t ld_d_gate_in_bf16; /* nullable bf16 copy */
  float* dWg;                                        /* [n_e, Hg] */
} MmlrecGate;
typedef struct MmlrecExpertGrad {
  const float* expert; int64_t ld_expert;            /* forward output (for the ReLU mask) */
  float* d_expert; int64_t ld_d_expert; int32_t H; int32_t n_users;
  const float* user_probs[MMLREC_MAX_TASKS + 1]; int32_t user_prob_ld[MMLREC_MAX_TASKS + 1];
  int32_t user_prob_col[MMLREC_MAX_TASKS + 1];
  const float* user_d_mix[MMLREC_MAX_TASKS + 1]; int64_t user_d_mix_ld[MMLREC_MAX_TASKS + 1];
  int32_t relu_mask; int32_t pad0;
  uint16_t* d_expert_bf16; int64_t ld_d_expert_bf16; /* nullable bf16 copy */
} MmlrecExpertGrad;
int mmlrec_gate_mix_forward(const MmlrecGate* gates, int32_t n_gates, int32_t B, void* stream);
int mmlrec_gate_mix_backward(const MmlrecGate* gates, int32_t n_gates,
                             const MmlrecExpertGrad* experts, int32_t n_experts, int32_t B,
                             int32_t max_ne, int32_t max_hg,
                             int32_t serialize_gates /* !=0 when two gates share a gate_in: each CTA then walks the
                                                        gates in order so d(gate_in) accumulates without a race */,
                             float* scratch, int32_t* counters /* int32 [n_gates], zero-initialised once */,
                             void* stream);
/* scratch floats needed by mmlrec_gate_mix_backward (max_ne / max_hg: maxima over the gate table) */
int64_t mmlrec_gate_mix_backward_scratch(int32_t n_gates, int32_t max_ne, int32_t max_hg, int32_t B);

/* Row-fused variant of the same stage (the fast path): one record describes ALL gates of a level and
 * the level's distinct expert activations; a warp owns a sample, reads every expert row exactly once
 * and produces every gate's mixture (forward) or every d(expert), d(gate_in) and dWg contribution
 * (backward) in one pass.  Limits: n_gates <= 8, n_experts <= 32, H % 4 == 0, and
 * sum_g n_e[g]*Hg[g] <= MMLREC_LEVEL_MAX_WG floats (gate-head weights are staged in shared memory);
 * callers fall back to mmlrec_gate_mix_* otherwise. */
#define MMLREC_LEVEL_MAX_GATES 8
#define MMLREC_LEVEL_MAX_EXPERTS 32
#define MMLREC_LEVEL_MAX_WG 3072
typedef struct MmlrecGateLevel {
  int32_t n_gates, n_experts, H, expert_relu;                     /* expert_relu: mask d(expert) by expert > 0 */
  const float* expert[MMLREC_LEVEL_MAX_EXPERTS]; int64_t ld_expert;
  float* d_expert[MMLREC_LEVEL_MAX_EXPERTS]; int64_t ld_d_expert;            /* backward outputs: fp32 ... */
  uint16_t* d_expert_bf16[MMLREC_LEVEL_MAX_EXPERTS]; int64_t ld_d_expert_bf16; /* ... or bf16 (either may be NULL) */
  int8_t slot[MMLREC_LEVEL_MAX_EXPERTS][MMLREC_LEVEL_MAX_GATES];  /* position of expert u in gate g's softmax, -1 if unused */
  const float* gate_in[MMLREC_LEVEL_MAX_GATES]; int64_t ld_gate_in[MMLREC_LEVEL_MAX_GATES];
  const float* Wg[MMLREC_LEVEL_MAX_GATES]; int64_t ld_Wg[MMLREC_LEVEL_MAX_GATES];
  int32_t Hg[MMLREC_LEVEL_MAX_GATES]; int32_t n_e[MMLREC_LEVEL_MAX_GATES];
  float* probs[MMLREC_LEVEL_MAX_GATES];
  float* mix[MMLREC_LEVEL_MAX_GATES]; int64_t ld_mix[MMLREC_LEVEL_MAX_GATES];
  uint16_t* mix_bf16[MMLREC_LEVEL_MAX_GATES]; int64_t ld_mix_bf16[MMLREC_LEVEL_MAX_GATES];
  const float* d_mix[MMLREC_LEVEL_MAX_GATES]; int64_t ld_d_mix[MMLREC_LEVEL_MAX_GATES];   /* NULL: gate gets no gradient */
  float* d_gate_in[MMLREC_LEVEL_MAX_GATES]; int64_t ld_d_gate_in[MMLREC_LEVEL_MAX_GATES];
  uint16_t* d_gate_in_bf16[MMLREC_LEVEL_MAX_GATES]; int64_t ld_d_gate_in_bf16[MMLREC_LEVEL_MAX_GATES];
  int32_t relu_mask_gate_in[MMLREC_LEVEL_MAX_GATES]; int32_t accumulate_d_gate_in[MMLREC_LEVEL_MAX_GATES];
  float* dWg[MMLREC_LEVEL_MAX_GATES];
  /* bit g set: gate g mixes expert u as a CONSTANT (`x.detach()`, hmoe.py:130): its mixture and its softmax gradient
   * use the expert's value, but no gradient flows from gate g into d_expert[u] (tiled backward only) */
  uint32_t detach_mask[MMLREC_LEVEL_MAX_EXPERTS];
} MmlrecGateLevel;
int mmlrec_gate_level_forward(const MmlrecGateLevel* level, int32_t B, void* stream);
int mmlrec_gate_level_backward(const MmlrecGateLevel* level, int32_t B, int32_t total_wg /* sum_g n_e[g]*Hg[g] */,
                               int32_t total_ne /* sum_g n_e[g] */, int32_t total_hg /* sum_g Hg[g] */,
                               float* scratch, int32_t* counter /* one int32, zero-initialised once */, void* stream);
/* scratch floats for mmlrec_gate_level_backward */
int64_t mmlrec_gate_level_backward_scratch(int32_t total_wg, int32_t B);
/* Tiled backward of the same stage (default when it fits): a CTA stages the rows of 8 samples in shared
 * memory (cp.async) and runs the five phases from there.  Extra requirements: every Hg % 4 == 0, gate_in /
 * d_gate_in / Wg rows 16-byte aligned (pointer and row stride), and the _smem() figure <= 110 KB. */
int mmlrec_gate_level_backward_tiled(const MmlrecGateLevel* level, int32_t B, int32_t n_gates, int32_t n_experts,
                                     int32_t H, int32_t total_wg, int32_t total_ne, int32_t total_hg,
                                     float* scratch, void* stream);
int64_t mmlrec_gate_level_backward_tiled_scratch(int32_t total_wg, int32_t B);
/* Tiled forward (same staging, same extra requirements as the tiled backward) */
int mmlrec_gate_level_forward_tiled(const MmlrecGateLevel* level, int32_t B, int32_t n_experts, int32_t H,
                                    int32_t total_wg, int32_t total_ne, int32_t total_hg, void* stream);
int64_t mmlrec_gate_level_forward_tiled_smem(int32_t n_experts, int32_t H, int32_t total_wg, int32_t total_ne,
                                             int32_t total_hg);
int64_t mmlrec_gate_level_backward_tiled_smem(int32_t n_gates, int32_t n_experts, int32_t H, int32_t total_wg,
                                              int32_t total_ne, int32_t total_hg);

/* ---------------------------------------------------------------------------------------------
 * Heads + loss, forward and backward in one pass (training) or forward only (predict).
 * Replaces tower_dnn_final_layer (Linear(H,1,bias=False), mmoe.py:52-55), PredictionLayer
 * (utils.py:242-248), F.binary_cross_entropy(.., reduction='sum') summed over tasks
 * (basemodel.py:294-296; log terms clamped at -100) and their autograd.
 *   esmm = flags: bit 0: two heads share ONE bias and pred = [p0, p0*p1] (esmm.py:57-62); bit 1: task t's logit carries
 *   the biases of tasks 0..t (mlp.py:45-52); bit 2: every head adds the SAME bias parameter (heads[0].bias): its gradient
 *   is the sum over the heads (escm.py:86-87 without the product head).
 * ------------------------------------------------------------------------------------------- */
typedef struct MmlrecHead {
  const float* h; int64_t ld_h; int32_t H; int32_t kind;     /* tower output [B,H]; MMLREC_HEAD_* */
  const float* w; const float* bias;                          /* [H], [1] */
  float* d_h; int64_t ld_d_h; int32_t relu_mask;              /* d(tower output), nullable */
  int32_t mask_col;                                           /* column of the scenario mask this head is multiplied by */
  float* dw; float* dbias;                                    /* [H], [1] */
  uint16_t* d_h_bf16; int64_t ld_d_h_bf16;                    /* nullable bf16 copy of d_h */
  const float* bias2; float* dbias2;                          /* optional second scalar bias (PEPNet: the final
                                                                 Linear's own bias next to the PredictionLayer's) */
} MmlrecHead;
int mmlrec_heads_forward_backward(const MmlrecHead* heads, int32_t T, int32_t B,
                                  const float* y, int64_t ldy,
                                  float* pred, int64_t ld_pred, float* loss /*[T+1]: per task, total*/,
                                  int32_t esmm, int32_t training,
                                  float* scratch, int64_t scratch_floats, int32_t* counters, void* stream);
/* The same with the scenario mask the reference's model classes and loop are written for but never receive (SURVEY Q4:
 * the loop sets domain_mask = None unconditionally, basemodel.py:265-266): pred[b][t] = head output * mask[b][mask_col_t]
 * (model/mmoe.py:101-106) and loss_t = BCE(pred_t, y_t, weight = mask column, reduction = 'sum')
 * (model/basemodel.py:273-282).  mask: fp32 [B, ld_mask] of 0 / 1.  Binary heads only. */
int mmlrec_heads_forward_backward_masked(const MmlrecHead* heads, int32_t T, int32_t B, const float* y, int64_t ldy,
                                         const float* mask, int64_t ld_mask, float* pred, int64_t ld_pred, float* loss,
                                         int32_t flags, int32_t training, float* scratch, int64_t scratch_floats,
                                         int32_t* counters, void* stream);
/* Backward of the heads for an upstream gradient d_pred = dL/d(pred) [B, ld_d_pred] supplied by the caller instead of
 * the fused BCE / MSE (the differentiable forward() of the model classes, model/mmoe.py:65: any loss built by autograd
 * on the returned probabilities).  Recomputes the logits, writes d(tower output), dw, dbias; `loss` receives zeros. */
int mmlrec_heads_backward_external(const MmlrecHead* heads, int32_t T, int32_t B, const float* d_pred, int64_t ld_d_pred,
                                   float* pred, int64_t ld_pred, float* loss, int32_t esmm,
                                   float* scratch, int64_t scratch_floats, int32_t* counters, void* stream);
int64_t mmlrec_heads_scratch(int32_t T, int32_t max_h, int32_t B);

/* ---------------------------------------------------------------------------------------------
 * Dense optimizer over the flat parameter store (torch.optim.{Adam,Adagrad,SGD,RMSprop} with
 * the reference's defaults, basemodel.py:569-584 / :313), one launch for all dense parameters.
 * Optionally refreshes the bf16 shadow copy the tensor-core GEMMs read.
 * ------------------------------------------------------------------------------------------- */
int mmlrec_dense_optimizer_step(float* param, const float* grad, float* state1, float* state2, int64_t n,
                                const MmlrecHyper* hyper, uint16_t* bf16_shadow, void* stream);
/* Same step, 128-bit accesses (n, slice_stride multiples of 4; buffers 16-byte aligned), with the gradient given as
 * `n_slices` partial buffers `slice_stride` floats apart, added in slice order before the update: split-K wgrad problems
 * write their partial tiles into slices 1..S-1 (slice 0 = the ordinary gradient buffer), so no separate reduction pass. */
int mmlrec_dense_optimizer_step_sliced(float* param, const float* grad, float* state1, float* state2, int64_t n,
                                       const MmlrecHyper* hyper, uint16_t* bf16_shadow, int32_t n_slices,
                                       int64_t slice_stride, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Element-wise stages used by STAR (model/star.py + SharedSpecificLinear, model/utils.py:163-223) and
 * PEPNet (model/pepnet.py).  "dkind" selects the derivative folded into a gradient write:
 * 0 none, 1 ReLU (keep where value > 0), 2 "2*sigmoid" (times y*(1 - y/2), GateNN's output).
 * ------------------------------------------------------------------------------------------- */
/* dst[:, :cols] = src[:, :cols]  (fp32 and/or bf16 destination): the concat of detached inputs */
int mmlrec_copy_cols(const float* src, int64_t ld_src, float* dst_f32, int64_t ld_f32, uint16_t* dst_bf16,
                     int64_t ld_bf16, int32_t rows, int32_t cols, void* stream);
/* out = a * b  (fp32 and/or bf16 copy) */
int mmlrec_mul_forward(const float* a, int64_t lda, const float* b, int64_t ldb, float* out_f32, int64_t ld_f32,
                       uint16_t* out_bf16, int64_t ld_bf16, int32_t rows, int32_t cols, void* stream);
/* da (+)= dkind_a(d_out * b, a) ; db (+)= dkind_b(d_out * a, b); each side optional, fp32 or bf16 */
int mmlrec_mul_backward(const float* d_out, int64_t ld_dout, const float* a, int64_t lda, const float* b, int64_t ldb,
                        float* da_f32, uint16_t* da_bf16, int64_t ld_da, int32_t dkind_a, int32_t acc_a,
                        float* db_f32, uint16_t* db_bf16, int64_t ld_db, int32_t dkind_b, int32_t acc_b,
                        int32_t rows, int32_t cols, void* stream);
/* AITM information transfer between two consecutive tasks (reference model/aitm.py:82-91: cat([p, q], 1) -> h1/h2/h3 ->
 * softmax(sum(K*Q, 2) / sqrt(H), dim=1) * V summed over the two tokens).  vkq: [rows, 6H] = V_p K_p Q_p V_q K_q Q_q (the
 * three projections of token p, then of token q); out[r, :] = a_p V_p + a_q V_q (fp32 and / or bf16); attn [rows, 2]
 * receives (a_p, a_q) for the backward pass (may be NULL for inference). */
int mmlrec_aitm_attention_forward(const float* vkq, int64_t ld, int32_t rows, int32_t H, float* out_f32, int64_t ld_f32,
                                  uint16_t* out_bf16, int64_t ld_bf16, float* attn, void* stream);
/* d_vkq (same 6H layout, fp32 or bf16, assigned) from d_out [rows, H] (fp32) and the saved attn */
int mmlrec_aitm_attention_backward(const float* d_out, int64_t ld_dout, const float* vkq, int64_t ld, const float* attn,
                                   int32_t rows, int32_t H, float* d_vkq_f32, uint16_t* d_vkq_bf16, int64_t ld_d,
                                   void* stream);
/* APG layer (reference model/apg.py:76-78, :96-99: ``torch.matmul(output_nk.unsqueeze(1), specific_weight_kk.view(-1, k, k))
 * .squeeze() + specific_bias_kk``): a per-sample [1, k] x [k, k] product.  nk [B, k]; wkk [B, k*k] (row b = the sample's
 * matrix, row-major [i][j]) and bkk [B, k] are the outputs of the two scene-embedding Linear layers.
 *   out[b, j] = sum_i nk[b, i] * wkk[b, i*k + j] + bkk[b, j]          (fp32 and/or bf16 copy) */
int mmlrec_apg_mix_forward(const float* nk, int64_t ld_nk, const float* wkk, int64_t ld_w, const float* bkk, int64_t ld_b,
                           int32_t B, int32_t k, float* out_f32, int64_t ld_f32, uint16_t* out_bf16, int64_t ld_bf16,
                           void* stream);
/* d(nk), d(wkk), d(bkk) (assigned; each fp32 or bf16) from d(out) [B, k] (fp32) */
int mmlrec_apg_mix_backward(const float* d_kk, int64_t ld_dkk, const float* nk, int64_t ld_nk, const float* wkk, int64_t ld_w,
                            int32_t B, int32_t k, float* d_nk_f32, uint16_t* d_nk_bf16, int64_t ld_dnk, float* d_wkk_f32,
                            uint16_t* d_wkk_bf16, int64_t ld_dw, float* d_bkk_f32, uint16_t* d_bkk_bf16, int64_t ld_db,
                            void* stream);
/* out[n] = sum_b Z[b, n], Z [B, N] fp32 or bf16 (exactly one non-NULL), fixed summation order: the bias gradient of a
 * Linear whose weight is stored [K, N] and applied as x @ W + b (apg.py:92, :99 ``torch.matmul(x, shared_weight) + bias``) */
int mmlrec_colsum(const float* z_f32, const uint16_t* z_bf16, int64_t ld, int32_t B, int32_t N, float* out, void* stream);
/* SNR-trans / MSSM gate (reference model/snr_trans.py:9-50, model/mssm.py:9-60): out_i = sum_j z_ij * (x_j @ M_ij), z_ij the
 * hard-concrete gate of (u_ij, alpha): s = sigmoid(log u - log(1-u) + log(alpha)/0.9), z = clamp(1.2 s - 0.1, 0, 1).
 * zdim = 1 (SNR-trans): u [n_out, n_in], one scalar per connection; zdim = U (MSSM): u [n_out, n_in, U], one gate per
 * output unit v of the connection.  alpha [1], trans [n_out, n_in, U, U] (constants).  Derived weight in nn.Linear layout
 * over the concatenated inputs:
 *   w_eff[i*U + v, j*U + u] = z_ij[v] * trans[i][j][u][v]        (fp32 + optional bf16 shadow, same ld) */
int mmlrec_snr_gate_weights(const float* u, const float* alpha, const float* trans, int32_t n_out, int32_t n_in, int32_t U,
                            int32_t zdim, float* w_eff, int64_t ld_w, uint16_t* w_eff_bf16, void* stream);
/* d_u (assigned; NULL where u is a constant: MSSM, mssm.py:27-29 keeps it in a plain list), d_alpha (assigned) from
 * d(w_eff); dz_scratch: n_out * n_in * zdim floats */
int mmlrec_snr_gate_fold(const float* d_w_eff, int64_t ld_w, const float* trans, const float* u, const float* alpha,
                         int32_t n_out, int32_t n_in, int32_t U, int32_t zdim, float* dz_scratch, float* d_u,
                         float* d_alpha, void* stream);
/* STAR: effective weights of all T domains in nn.Linear layout:
 *   w_eff[t*N + n, k] = spec[t][k, n] * shared[k, n];  b_eff[t*N + n] = spec_b[t][n] + shared_b[n]
 * spec / spec_b: T device pointers (int64 array on device); bf16 shadow optional. */
int mmlrec_star_weights(const int64_t* spec_ptrs, const int64_t* spec_b_ptrs, const float* shared, const float* shared_b,
                        int32_t T, int32_t K, int32_t N, float* w_eff, int64_t ld_w, uint16_t* w_eff_bf16, float* b_eff,
                        void* stream);
/* STAR backward fold: d_shared[k,n] = sum_t d_w_eff[t*N+n,k] * spec[t][k,n]; d_spec_last[k,n] = d_w_eff[(T-1)*N+n,k] *
 * shared[k,n]; d_shared_b[n] = sum_t d_b_eff[t*N+n]; d_spec_b_last[n] = d_b_eff[(T-1)*N+n].  live[t]=0 skips a domain
 * whose effective weight received no gradient. */
int mmlrec_star_fold(const float* d_w_eff, int64_t ld_w, const float* d_b_eff, const int64_t* spec_ptrs,
                     const float* shared, const int32_t* live, int32_t T, int32_t K, int32_t N, float* d_shared,
                     float* d_shared_b, float* d_spec_last, float* d_spec_b_last, void* stream);

/* small utilities */
/* L2 regularisation of the dense parameters (model/basemodel.py:524-540 get_regularization_loss, added to the loss at
 * :303): grad[i] += 2 * l2_coef[i] * param[i]; *reg_out = sum l2_coef[i] * param[i]^2 (deterministic).  l2_coef is 0
 * wherever a parameter is not registered; a NEGATIVE coefficient means "no kernel writes this gradient entry": the
 * entry is assigned 2 * |coef| * param instead of accumulated.  scratch: mmlrec_l2_scratch() floats. */
int mmlrec_l2_regularize(const float* param, float* grad, const float* l2_coef, int64_t n, float* reg_out, float* scratch,
                         void* stream);
int64_t mmlrec_l2_scratch(void);
/* Deterministic split-K for the batch-contraction wgrad GEMMs: the S partial problems write their tiles into S
 * scratch slices; this sums the slices in a fixed order into the gradient buffer.
 * segments: int64 [n_segments][3] = {dst element offset, src element offset inside a slice, element count}. */
int mmlrec_sum_slices(const int64_t* segments, int32_t n_segments, int64_t max_n, float* dst, const float* src,
                      int32_t S, int64_t slice_stride, void* stream);
int mmlrec_fill_f32(float* p, int64_t n, float v, void* stream);
int mmlrec_cast_f32_to_bf16(const float* src, int64_t ld_src, uint16_t* dst, int64_t ld_dst,
                            int32_t rows, int32_t cols, int32_t cols_pad, void* stream);
int mmlrec_cast_bf16_to_f32(const uint16_t* src, int64_t ld_src, float* dst, int64_t ld_dst,
                            int32_t rows, int32_t cols, void* stream);
/* out(m,n) = a(m,n) * b(m,n) (* c) ; used by STAR weight build and PEPNet gating */
int mmlrec_mul_f32(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                   int32_t rows, int32_t cols, int32_t accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMLREC_B200_H_ */
