/*
 * cfm_b200.h — C ABI of libcfm_b200.so, the B200 (sm_100a) implementation of the
 * CEO-Recommender two-tower training + scoring hot path.
 *
 * The reference (SMaric93/CEO-Recommender) has no FFI of its own: its hot path is a
 * chain of PyTorch ATen calls.  Each entry point below names the reference call
 * sequence it replaces (paths relative to the reference repository root).  Every
 * pointer is a raw DEVICE pointer unless marked host; buffers are owned by the caller
 * (torch's caching allocator in the shipped Python host); the library never allocates
 * device memory.  All work is enqueued on `stream` (a cudaStream_t passed as void*)
 * and nothing synchronises.  Return value: 0 on success, a negative CFM_ERR_* code
 * otherwise; cfm_last_error() gives the message of the calling thread's last failure.
 *
 * Structs use only int64_t / double / pointers so that any FFI (ctypes, cffi, cgo,
 * JNI) can mirror them field by field without packing rules.
 */
#ifndef CFM_B200_H
#define CFM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CFM_ABI_VERSION 3
#define CFM_MAX_TABLES 16

#define CFM_OK 0
#define CFM_ERR_INVALID (-1)     /* bad dimension / null pointer / misaligned buffer */
#define CFM_ERR_UNSUPPORTED (-2) /* shape outside what the kernels are built for     */
#define CFM_ERR_CUDA (-3)        /* a CUDA runtime call failed (message has details) */
#define CFM_ERR_BATCHNORM_B1 (-4)/* train-mode BatchNorm with a single row (torch raises too) */

/* bits OR-ed by kernels into the caller's device-side `err_flag` word */
#define CFM_FLAG_INDEX_OOB 1     /* categorical index outside [0, table_rows) — torch raises IndexError */

int cfm_abi_version(void);
const char* cfm_last_error(void);
/* SM count, compute capability and the number of persistent CTAs the tower kernels launch
 * per tower (callers size the per-CTA partial buffers with it). */
int cfm_device_info(int64_t* sm_count, int64_t* cc_major, int64_t* cc_minor, int64_t* tower_ctas);

/* Bench aids.  cfm_launch_count: kernels this library has launched (optionally reset).
 * cfm_profile_enable(1): bracket every kernel family with cudaEvents on its stream; cfm_profile_read
 * synchronises the device and returns summed milliseconds and launch counts per CFM_PROF_* slot. */
#define CFM_PROF_SLOTS 14
int64_t cfm_launch_count(int64_t reset);
int cfm_profile_enable(int64_t on);
int cfm_profile_read(double* ms /* [CFM_PROF_SLOTS] host */, int64_t* counts /* host */, int64_t n_slots);

/* ------------------------------------------------------------------------------------------
 * Towers.  One tower = categorical embedding gather + numeric concat + 3-layer MLP
 *   Linear(in,h1) -> BN -> ReLU -> Dropout(p1) -> Linear(h1,h2) -> [BN] -> ReLU -> [Dropout(p2)] -> Linear(h2,d_out)
 * replaces: ceo_firm_matching/model.py:69-76 (both towers of CEOFirmMatcher, bn2=1) and
 *           ceo_firm_matching/structural_model.py:120-127 (StructuralDistillationNet, bn2=0).
 * Layouts are torch's: row-major, Linear weight [out,in], x_num [B,n_num] f32, x_cat [B,n_tables] i64,
 * table i is [table_rows[i], emb_dim] f32.  Concat order is [x_num | E_0 | E_1 | ...].
 * ------------------------------------------------------------------------------------------ */
typedef struct cfm_tower {
    int64_t n_num, n_tables, emb_dim, h1, h2, d_out;
    int64_t bn2;                 /* 1: BatchNorm1d after Linear 2 */
    double drop1, drop2;         /* dropout probability after activation 1 / 2 (0 = none) */
    int64_t tower_id;            /* distinguishes the dropout streams of the towers in one call */
    int64_t precision;           /* 0: fp32-class (3xTF32 error-compensated tensor-core products), 1: single-pass TF32 */
    /* inputs */
    const float* x_num;
    const int64_t* x_cat;
    const float* tables[CFM_MAX_TABLES];
    int64_t table_rows[CFM_MAX_TABLES];
    /* parameters */
    const float *w1, *b1, *w2, *b2, *w3, *b3;
    const float *bn1_w, *bn1_b, *bn2_w, *bn2_b;
    float *bn1_rm, *bn1_rv, *bn2_rm, *bn2_rv; /* running stats; updated in training mode */
    int64_t *bn1_nbt, *bn2_nbt;                /* num_batches_tracked; +1 in training mode (nullable) */
    /* saved activations (written by fwd, read by bwd) */
    float *h1_raw;               /* [B,h1]  Linear-1 output before BN */
    float *h2_raw;               /* [B,h2]  Linear-2 output before BN/ReLU */
    float *out;                  /* [B,d_out] tower output (pre-normalisation latent / logits) */
    float *bn1_stat, *bn2_stat;  /* [4,h]: batch mean, 1/sqrt(var+eps), and the two BN-backward batch means */
    float *scratch;              /* per-CTA partials: >= tower_ctas * cfm_tower_scratch_floats() floats */
    float *wimg;                 /* cfm_tower_wimg_floats() floats: (hi, lo)-split, swizzled images of the three weight
                                    matrices for the tcgen05 stage kernels, rebuilt by every cfm_towers_fwd call and read
                                    by the cfm_towers_bwd call that follows it.  NULL: the mma.sync stage kernels run. */
    float *a1, *a2;              /* [B,h1], [B,h2]: activation codes of the inputs of Linear 2 / 3, written by the tcgen05
                                    forward stages: x-hat (the raw pre-activation without BatchNorm) where the unit is active
                                    and kept by dropout, NaN elsewhere - the backward recovers the activation, its derivative
                                    and x-hat from this one tensor; required (non-NULL) together with wimg */
    float *xstash;               /* cfm_tower_xstash_floats(t, B) floats, nullable: the tcgen05 stage-1 forward saves its
                                    gathered input tiles here (16 KB blocks of 64 rows x 64 columns) and the stage-1
                                    backward streams them back with bulk copies instead of gathering every embedding
                                    row from HBM a second time */
} cfm_tower_t;

/* floats of `scratch` needed per persistent CTA for this tower shape (fwd and bwd share it) */
int64_t cfm_tower_scratch_floats(const cfm_tower_t* t);
/* floats of `wimg` for this tower shape */
int64_t cfm_tower_wimg_floats(const cfm_tower_t* t);
/* floats of `xstash` for this tower shape and batch */
int64_t cfm_tower_xstash_floats(const cfm_tower_t* t, int64_t B);

/* Forward of `n_towers` towers over the same B rows in one set of launches.
 * training=1: batch-stat BN (running stats updated), dropout drawn from a counter-based
 * Philox stream keyed by (seed, offset); training=0: running-stat BN, no dropout. */
int cfm_towers_fwd(const cfm_tower_t* towers, int64_t n_towers, int64_t B, int64_t training,
                   uint64_t seed, uint64_t offset, const uint64_t* rng_offset_dev /* nullable: device counter
                   added to `offset`, so a captured CUDA graph draws a fresh mask every replay */,
                   int32_t* err_flag, void* stream);
/* *counter_dev += inc on the stream (one graph node per step keeps the dropout stream moving) */
int cfm_counter_advance(uint64_t* counter_dev, uint64_t inc, void* stream);

typedef struct cfm_tower_grads {
    const float* g_out;          /* [B,d_out] gradient w.r.t. tower output */
    float *dw1, *db1, *dw2, *db2, *dw3, *db3;   /* parameter gradients (overwritten) */
    float *dbn1_w, *dbn1_b, *dbn2_w, *dbn2_b;
    float *dy1, *dy2;            /* scratch [B,h1], [B,h2] */
    float *dx_emb;               /* [B, n_tables*emb_dim] per-pair embedding-row gradients (nullable) */
    float *dx_num;               /* [B, n_num] gradient w.r.t. numeric inputs (nullable) */
} cfm_tower_grads_t;

/* Backward of the towers given the same descriptors (same training flag / seed / offset as fwd).
 * replaces: the autograd graph torch builds for model.py:69-76 (training.py:54 `loss.backward()`). */
int cfm_towers_bwd(const cfm_tower_t* towers, const cfm_tower_grads_t* grads, int64_t n_towers, int64_t B,
                   int64_t training, uint64_t seed, uint64_t offset, const uint64_t* rng_offset_dev, void* stream);

/* Keep-mask the towers' dropout uses, materialised for differential tests:
 * mask[r, c] (uint8) for `site` (0 = after activation 1, 1 = after activation 2). */
int cfm_dropout_mask(uint8_t* mask, int64_t B, int64_t width, double p, int64_t tower_id, int64_t site,
                     uint64_t seed, uint64_t offset, void* stream);

/* ------------------------------------------------------------------------------------------
 * Embedding gradient: deterministic sorted-segment reduce (no atomics).
 * replaces: aten::embedding_dense_backward for each nn.Embedding of model.py:24-33.
 * For every table i: grad_table_i[idx] = sum over rows r with x_cat[r,i]==idx of dx_emb[r, i*E:(i+1)*E],
 * summed in increasing r.  Rows of grad tables not touched are left as they are (caller keeps them zero:
 * see cfm_emb_grad_rezero).
 *   keys_tmp/vals_tmp/keys_sorted/vals_sorted: scratch [n_tables*B] (i64 keys, i32 values)
 *   sort_tmp: sort scratch of sort_tmp_bytes (query with cfm_emb_grad_tmp_bytes)
 * After the call keys_sorted lists every (table, index) touched (runs of equal keys); handing it to
 * cfm_emb_grad_rezero before the next step zeroes exactly those rows again, so a persistent dense
 * gradient buffer stays exact without a full-table memset.
 * ------------------------------------------------------------------------------------------ */
int64_t cfm_emb_grad_tmp_bytes(int64_t n_tables, int64_t B);
int cfm_emb_grad_segment_reduce(const int64_t* x_cat, const float* dx_emb, int64_t B, int64_t n_tables,
                                int64_t emb_dim, float* const* grad_tables /* host array of device ptrs */,
                                const int64_t* table_rows /* host */, int64_t* keys_tmp, int32_t* vals_tmp,
                                int64_t* keys_sorted, int32_t* vals_sorted, void* sort_tmp, int64_t sort_tmp_bytes,
                                void* stream);
int cfm_emb_grad_rezero(float* const* grad_tables /* host */, const int64_t* table_rows /* host */, int64_t n_tables,
                        int64_t emb_dim, const int64_t* keys_sorted, int64_t n_items, void* stream);

/* Joint form for the towers of one step (same batch): ONE key build and ONE radix sort over the (table, index)
 * pairs of every tower, then one segment-reduce launch per tower on its contiguous range of the sorted keys.
 * Same result, bit for bit, as one cfm_emb_grad_segment_reduce per tower; scratch sized for
 * n_items = B * (total tables) (cfm_emb_grad_tmp_bytes(total tables, B)).  cfm_emb_grad_joint_rezero zeroes the rows
 * named by keys_sorted of the previous joint reduce (x_cat / dx_emb of the groups are ignored). */
#define CFM_MAX_GROUPS 4
typedef struct cfm_emb_group {
    const int64_t* x_cat;                    /* [B, n_tables] */
    const float* dx_emb;                     /* [B, n_tables*emb_dim] */
    int64_t n_tables;
    int64_t emb_dim;
    float* grad_tables[CFM_MAX_TABLES];      /* [rows_i, emb_dim] dense gradients */
    int64_t table_rows[CFM_MAX_TABLES];
} cfm_emb_group_t;
/* phase 0: everything; 1: key build + sort only (needs x_cat only: can run on a side stream beside the forward);
 * 2: segment reduce only (needs dx_emb; after phase 1 on the same scratch). */
int cfm_emb_grad_joint_reduce(const cfm_emb_group_t* groups /* host */, int64_t n_groups, int64_t B, int64_t phase,
                              int64_t* keys_tmp, int32_t* vals_tmp, int64_t* keys_sorted, int32_t* vals_sorted,
                              void* sort_tmp, int64_t sort_tmp_bytes, void* stream);
int cfm_emb_grad_joint_rezero(const cfm_emb_group_t* groups /* host */, int64_t n_groups, int64_t B,
                              const int64_t* keys_sorted, void* stream);

/* ------------------------------------------------------------------------------------------
 * Table-sharded embeddings over NVLink peer memory (one process per GPU, single node; the reference has no
 * multi-GPU code - SURVEY 8e).  Each table - or each of `pieces` equal column slices of it - is owned by one rank;
 * the other ranks map the owner's tensor (CUDA IPC) and pass the mapped pointer wherever a table pointer is expected.
 *   cfm_enable_peer_access  : cudaDeviceEnablePeerAccess from the current device to `peer_device` (idempotent).
 *   cfm_ipc_export/open/close: see below.
 *   cfm_emb_gather_rows     : stash[k, b, c] = tables[k*pieces + c/(E/pieces)][x_cat[b, k], c]; tables[] holds, per
 *                             (table, piece), the base of the FULL [rows, E] tensor as mapped from that piece's
 *                             owner.  stash is [n_tables, B, E]; the towers then run on it (table k = stash[k],
 *                             index = b), so forward and stage-1 backward touch NVLink once per step.
 *                             Out-of-range index: err_flag |= 1.
 *   cfm_emb_grad_peer_reduce: the owner's sorted-segment reduce over EVERY rank's batch: one key build + one radix
 *                             sort over the owned slices of all groups, one reduce launch per group.  For owned slice j the
 *                             items are (rank r, row b) with index x_cat[r][b*n_cols + col] and gradient columns
 *                             dx_emb[r][b*n_cols*E + col*E + col0 ... + width); both buffers are read in place
 *                             through the peer mappings; grad[idx, col0 ... col0+width) is written.  Runs of equal
 *                             keys are summed in (rank, row) order == the order of the concatenated global batch.
 *                             Scratch as for cfm_emb_grad_segment_reduce with n_items = (total owned)*n_peers*B
 *                             (cfm_emb_grad_tmp_bytes((total owned)*n_peers, B)).
 *   cfm_emb_grad_peer_rezero: zero the slices named by keys_sorted of the previous reduce (x_cat/dx_emb ignored).
 * ------------------------------------------------------------------------------------------ */
#define CFM_MAX_PEERS 8
#define CFM_MAX_SLOTS 64
typedef struct cfm_peer_table {
    int64_t n_cols;                          /* categorical columns of the tower this table belongs to */
    int64_t col;                             /* this table's column in that tower's x_cat / dx_emb     */
    int64_t col0;                            /* first embedding column of the owned slice              */
    int64_t rows;                            /* rows of the table                                       */
    float* grad;                             /* [rows, emb_dim] dense gradient, local                   */
    const int64_t* x_cat[CFM_MAX_PEERS];     /* per rank: [B, n_cols] indices (peer-mapped)             */
    const float* dx_emb[CFM_MAX_PEERS];      /* per rank: [B, n_cols*emb_dim] gradient rows (peer-mapped) */
} cfm_peer_table_t;
int cfm_enable_peer_access(int64_t peer_device);
/* CUDA IPC plumbing for the mappings: the owner exports the cudaMalloc allocation holding `ptr` (64-byte handle +
 * byte offset of ptr inside it); every other process opens the handle ONCE, with ITS OWN device current (peer access
 * to the owner's device is enabled lazily), and addresses the buffer at base + offset. */
int cfm_ipc_export(const void* ptr, uint8_t* handle_out /* 64 bytes */, int64_t* offset_out);
int cfm_ipc_open(const uint8_t* handle /* 64 bytes */, void** base_out);
int cfm_ipc_close(void* base);
int cfm_emb_gather_rows(const int64_t* x_cat, int64_t B, int64_t n_tables, int64_t emb_dim, int64_t pieces,
                        const float* const* tables /* host array of n_tables*pieces device ptrs */,
                        const int64_t* table_rows /* host */, float* stash, int32_t* err_flag, void* stream);
typedef struct cfm_peer_group {              /* the owned slices that share one embedding width (one tower) */
    const cfm_peer_table_t* owned;           /* host array */
    int64_t n_owned;
    int64_t emb_dim;                         /* row width of these tables */
    int64_t width;                           /* columns per owned slice (emb_dim / pieces) */
} cfm_peer_group_t;
/* phase 0: everything; 1: key build + sort only (needs the peers' x_cat: can run on a side stream while the step
 * computes); 2: segment reduce only (needs the peers' dx_emb; after phase 1 on the same scratch). */
int cfm_emb_grad_peer_reduce(const cfm_peer_group_t* groups /* host */, int64_t n_groups, int64_t n_peers, int64_t B,
                             int64_t phase, int64_t* keys_tmp, int32_t* vals_tmp, int64_t* keys_sorted,
                             int32_t* vals_sorted, void* sort_tmp, int64_t sort_tmp_bytes, void* stream);
int cfm_emb_grad_peer_rezero(const cfm_peer_group_t* groups /* host */, int64_t n_groups, int64_t n_peers, int64_t B,
                             const int64_t* keys_sorted, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused dense Adam step over all tensors of a model in one launch.
 * replaces: torch.optim.Adam(capturable=True) as used by training.py:32,55 - same arithmetic, operation by
 * operation (torch/optim/adam.py::_multi_tensor_adam, capturable branch), one pass instead of ~12.
 *   tensors: HOST array of records (they travel as kernel parameters: no descriptor memory, graph-capturable)
 *   step   : device fp32 [n_steps] (torch keeps one state["step"] per parameter, all equal): every element is
 *            incremented by the call, element 0 is the step count used
 *   variant: 0 (contraction pattern bit-equal to torch's kernels); 1..3 exist for the differential test
 * ------------------------------------------------------------------------------------------ */
typedef struct cfm_adam_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t numel;
} cfm_adam_tensor_t;
int cfm_adam_step(const cfm_adam_tensor_t* tensors /* host */, int64_t n_tensors, float* step, int64_t n_steps, double lr,
                  double beta1, double beta2, double eps, int64_t variant, void* stream);

/* ------------------------------------------------------------------------------------------
 * Cosine head: L2-normalise both latents, row-wise dot, times exp(logit_scale).
 * replaces: model.py:79-87 (eps = 0: plain division) and contrastive.py:64,70,92-93 (eps = 1e-12, F.normalize).
 * The weighted MSE of training.py:52 can be fused into both directions.
 * ------------------------------------------------------------------------------------------ */
/* forward: score [B] (+ unit latents if asked); with target/weights/loss also loss = mean(w (s-t)^2) */
int cfm_cosine_head_fwd(const float* u, const float* v, const float* logit_scale, int64_t B, int64_t D,
                        double eps, float* score /* [B] */, float* u_hat /* nullable [B,D] */,
                        float* v_hat /* nullable [B,D] */, const float* target /* nullable [B] */,
                        const float* weights /* nullable [B] */, float* loss /* nullable [1] */,
                        float* partial /* >= 4096 floats, zero on first use; needed with loss */, void* stream);
/* backward: d_score [B] and/or the weighted-MSE gradient 2 w (s-t) g_loss / B (when target/weights given;
 * g_loss = device scalar dL/dloss, nullable = 1), plus optional gradients flowing into the unit latents. */
int cfm_cosine_head_bwd(const float* u, const float* v, const float* logit_scale, const float* d_score /* nullable */,
                        const float* d_uhat /* nullable [B,D] */, const float* d_vhat /* nullable [B,D] */,
                        const float* target /* nullable */, const float* weights /* nullable */,
                        const float* g_loss /* nullable */, int64_t B, int64_t D, double eps, float* du, float* dv,
                        float* d_logit_scale /* [1] */, float* partial /* >= 4096 floats */, void* stream);

/* ------------------------------------------------------------------------------------------
 * Projection heads of the contrastive model:  out = normalize(W2 . relu(W1 . x + b1) + b2).
 * replaces: ceo_firm_matching/contrastive.py:41-50 (firm_projector / ceo_projector, nn.Sequential(Linear, ReLU,
 *           Linear)), :96-97 (F.normalize(..., dim=1), eps = 1e-12) and their autograd (:244-260).
 * fp32 FMA arithmetic, layers up to 64 wide; several heads (both sides of the model) share one launch.
 * ------------------------------------------------------------------------------------------ */
typedef struct cfm_projector {
    int64_t d_in, d_hid, d_out;
    const float *w1, *b1, *w2, *b2;   /* torch layouts: w1 [d_hid,d_in], w2 [d_out,d_hid] */
    const float *x;                   /* [B,d_in] */
    float *hid;                       /* [B,d_hid] relu(W1 x + b1): written by fwd, read by bwd */
    float *raw;                       /* [B,d_out] W2 hid + b2 before normalisation: written by fwd, read by bwd */
    float *out;                       /* [B,d_out] raw / max(||raw||, eps) (unused by bwd) */
} cfm_projector_t;
typedef struct cfm_projector_grads {
    const float* g_out;               /* [B,d_out] gradient w.r.t. out */
    float *dx;                        /* [B,d_in] (nullable) */
    float *dw1, *db1, *dw2, *db2;     /* parameter gradients (overwritten) */
    float *scratch;                   /* cfm_projector_scratch_floats(head, B) floats: per-CTA partials, added in CTA order */
} cfm_projector_grads_t;
int64_t cfm_projector_scratch_floats(const cfm_projector_t* head, int64_t B);
int cfm_projector_fwd(const cfm_projector_t* heads, int64_t n_heads /* 1..4 */, int64_t B, double eps, void* stream);
int cfm_projector_bwd(const cfm_projector_t* heads, const cfm_projector_grads_t* grads, int64_t n_heads, int64_t B,
                      double eps, void* stream);

/* ------------------------------------------------------------------------------------------
 * Structural head: softmax(c_logits) . A . softmax(f_logits) expected match, KL distillation loss and
 * the gradients w.r.t. both logit sets, one kernel.
 * replaces: structural_model.py:130-141 + structural_training.py:75-77 (+ autograd of both).
 *   T == 5 types.  targets nullable (then loss/KL terms are skipped).  d_match nullable (gradient flowing
 *   into expected_match, e.g. IlluminationEngine structural_explain.py:76-79).  kl_scale = dLoss/d(KL sum),
 *   the op divides by B itself ('batchmean').
 * ------------------------------------------------------------------------------------------ */
int cfm_structural_head(const float* c_logits, const float* f_logits, const float* A /* [5,5] */,
                        const float* target_ceo, const float* target_firm, const float* d_match,
                        int64_t B, double kl_scale, float* match /* [B] */, float* loss /* [1] nullable */,
                        float* d_c_logits /* nullable [B,5] */, float* d_f_logits /* nullable [B,5] */,
                        float* partial /* >= 4096 floats */, void* stream);

/* ------------------------------------------------------------------------------------------
 * InfoNCE (tcgen05 / TMEM / TMA).  Operands are bf16 row-major [rows, Dp], Dp in {64, 128} (features zero-padded
 * by cfm_pack_rows_bf16), rows of unit L2 norm (what contrastive.py:96-97 feeds info_nce_loss).
 * replaces: contrastive.py:129-138 (torch.mm + 2x F.cross_entropy) and its autograd; S is never written to HBM.
 *   cfm_infonce_rowsum: rowsum[i] = sum_j exp((x_i . y_j - 1)/T)  (fixed maximum 1/T), diag[i] = x_i . y_(i+diag_offset)
 *   cfm_infonce_loss:   loss = 1/(2 B_total) sum_i [log R_i + log C_i + 2/T - 2 diag_i/T]   over the n local rows
 *   cfm_infonce_grad:   dX[i] = g_loss * ( 1/(2 B T) sum_j E_ij (1/rowsum_x[i] + 1/rowsum_y[j]) y_j - y_(i+off)/(B T) )
 * The symmetric loss needs two rowsum calls and both gradients two grad calls, each with (X, Y) swapped.
 * `part` is scratch of cfm_simtile_chunks(R, C) * R floats (rowsum) or * R * Dp floats (grad).
 * ------------------------------------------------------------------------------------------ */
int64_t cfm_simtile_chunks(int64_t R, int64_t C);
/* test hook: pin the number of 128-row blocks a similarity-kernel CTA owns (0 = automatic, 1, 2) */
int cfm_simtile_set_rb(int64_t rb);
/* Tuning knob: bit 0 / bit 1 = the row-sum / gradient kernels evaluate every fourth exponential as a degree-4 polynomial
   on the FMA pipe instead of MUFU.EX2 (results agree to 3e-6 relative).  Default 1: the row-sum pass is bound by the
   MUFU pipe (XU 85 % busy) and gains 12 %; the gradient pass is bound by instruction issue and the tensor pipe and
   loses 13 % with it (profiles/r02_notes.md). */
int cfm_simtile_set_poly(int64_t mask);
int cfm_pack_rows_bf16(const float* in /* [R,D] */, int64_t R, int64_t D, int64_t Dp, void* out_bf16 /* [R,Dp] */,
                       void* stream);
int cfm_infonce_rowsum(const void* x_bf16, const void* y_bf16, int64_t R, int64_t C, int64_t Dp, double temperature,
                       int64_t diag_offset, float* rowsum /* [R] */, float* diag /* [R] nullable */, float* part,
                       void* stream);
/* Row sums AND column sums of exp((S - 1/T)/T) from ONE pass over S = X . Y^T: the exponentials of every tile are
 * added along the rows (as above) and, summed over the 128 rows of the tile by warp shuffles, along the columns.
 * colsum[j] = sum_i exp((x_i.y_j - 1)/T) is the row sum of S^T, so the forward of the symmetric loss
 * (contrastive.py:129-138) needs one similarity pass instead of two.  col_part: cfm_infonce_colpart_floats(R, C)
 * floats of scratch (per-128-row-block column sums, added in block order: bitwise reproducible). */
int64_t cfm_infonce_colpart_floats(int64_t R, int64_t C);
int cfm_infonce_rowcolsum(const void* x_bf16, const void* y_bf16, int64_t R, int64_t C, int64_t Dp, double temperature,
                          int64_t diag_offset, float* rowsum /* [R] */, float* colsum /* [C] */,
                          float* diag /* [R] nullable */, float* part, float* col_part, void* stream);
int cfm_infonce_loss(const float* rowsum_row, const float* rowsum_col, const float* diag, int64_t n,
                     double temperature, int64_t B_total, float* loss /* [1] */, void* stream);
int cfm_infonce_grad(const void* x_bf16, const void* y_bf16, int64_t R, int64_t C, int64_t D, int64_t Dp,
                     double temperature, int64_t diag_offset, int64_t B_total, const float* rowsum_x /* [R] */,
                     const float* rowsum_y /* [C] */, const float* diag /* [R] from cfm_infonce_rowsum */,
                     const float* g_loss /* device scalar, nullable = 1 */,
                     float* dx /* [R,D] f32 */, float* part, void* stream);
/* parity/debug aid: the raw fp32 score tile S = X . Y^T written out as [R, C] */
int cfm_simtile_scores(const void* x_bf16, const void* y_bf16, int64_t R, int64_t C, int64_t Dp, float* out,
                       void* stream);

/* ------------------------------------------------------------------------------------------
 * All-pairs scoring with streaming top-k.
 * replaces: analytical_extensions.py:471,483 (torch.mm + np.argsort), contrastive.py:307-310 (sim.sort).
 * Pass 1 (tcgen05/TMEM): scores of the bf16 copies, tile by tile; each row keeps the columns whose score beats a
 * running threshold in a TK_CAP-entry buffer that is compacted to its TK_KEEP best when full.  Pass 2: every
 * candidate within `margin` of the k-th best is rescored in fp64 from the fp32 operands and the k best are emitted
 * ordered (score desc, index asc), score = scale * <row, col>.  `margin` must bound twice the error of a
 * 16-bit-operand score (2^-7 * max|row| * max|col| for bf16 copies, 2^-10 for fp16 copies, which are the better
 * choice whenever the operands fit fp16's range, e.g. unit-norm tower outputs); rows whose completeness cannot be proven get row_flag = 1 and
 * must be redone exactly by the caller.  k <= 128.  The [R,C] score matrix never exists in memory.
 *   rows [R,D] / cols [C,D] f32 and bf16 copies [R,Dp] / [C,Dp]; col_offset is added to emitted indices (shards)
 *   scratch: cand [lists*Rpad, CFM_TOPK_CAP] 8-byte (score bits, column) entries, cand_cnt/cand_thr [lists*Rpad],
 *            lists = cfm_simtile_chunks(R, C), Rpad = R rounded up to 256
 * ------------------------------------------------------------------------------------------ */
#define CFM_TOPK_CAP 384
int cfm_pack_rows_f16(const float* in /* [R,D] */, int64_t R, int64_t D, int64_t Dp, void* out_f16 /* [R,Dp] */,
                      void* stream);
int cfm_allpairs_topk(const float* rows_f32, const float* cols_f32, const void* rows_16 /* [R,Dp] */,
                      const void* cols_16 /* [C,Dp] */, int64_t operands_f16 /* 0: bf16 copies, 1: fp16 copies */,
                      int64_t R, int64_t C, int64_t D, int64_t Dp, int64_t k, double scale, double margin,
                      int64_t col_offset, float* out_score /* [R,k] */, double* out_score64 /* [R,k] nullable */,
                      int64_t* out_idx /* [R,k] */, int32_t* row_flag /* [R] */, void* cand, int32_t* cand_cnt,
                      float* cand_thr, void* stream);
/* merge `n_parts` (<= 16) per-shard top-k lists [n_parts,R,k] into the global top-k (score desc, index asc).
 * Pass the fp64 scores of cfm_allpairs_topk (score_is_f64 = 1) to keep the exact cross-shard ordering: two fp64
 * scores may round to the same fp32 value. */
int cfm_topk_merge(const void* part_score, int64_t score_is_f64, const int64_t* part_idx, int64_t n_parts, int64_t R,
                   int64_t k, float* out_score, int64_t* out_idx, void* stream);
/* 1-indexed rank of column target_col[i] in row i: 1 + #{j : s_ij > s_it or (s_ij == s_it and j < t)}
 * (contrastive.py:312-320), scores formed in fp64 from the fp32 operands */
int cfm_allpairs_rank(const float* rows_f32, const float* cols_f32, int64_t R, int64_t C, int64_t D,
                      const int64_t* target_col /* [R] */, int64_t* rank /* [R] */, void* stream);
/* The same rank for the positive pair on a shifted diagonal, target column = i + diag_offset (retrieval metrics of
 * contrastive.py:296-332 / run_deep_extensions.py:470-483), counted in the epilogue of the tcgen05 similarity kernel:
 * every 16-bit-operand score outside the window [d_i - err_bound, d_i + err_bound] around the row's exact positive
 * score d_i decides its comparison on the spot; the (row, column) pairs inside the window are listed and compared
 * exactly (fp64 of the fp32 operands, ties by column index), so `rank` equals cfm_allpairs_rank's.  err_bound must
 * bound |16-bit-operand score - exact score| (2u |row| |col| + the fp32 accumulation error; u = 2^-11 fp16, 2^-9 bf16).
 *   status[0] = 1 when more than amb_cap pairs fell inside the windows (rank is then incomplete: use
 *   cfm_allpairs_rank), status[1] = pairs listed.  Scratch: part [cfm_simtile_chunks(R,C) * Rpad] with Rpad = R rounded
 *   up to 256, extra [R], diag64 [R], window [2R], amb [amb_cap] pairs of uint32, amb_n [1] (64-bit: the count
 *   cannot wrap even when every pair of a degenerate input falls inside the windows). */
int cfm_allpairs_diag_rank(const float* rows_f32, const float* cols_f32, const void* rows_16, const void* cols_16,
                           int64_t operands_f16, int64_t R, int64_t C, int64_t D, int64_t Dp, int64_t diag_offset,
                           double err_bound, int64_t* rank /* [R] */, int32_t* status /* [2] */, int32_t* part,
                           int32_t* extra, double* diag64, float* window, void* amb, int64_t amb_cap, uint64_t* amb_n,
                           void* stream);

/* ------------------------------------------------------------------------------------------
 * Self-test of the tensor-core building blocks of the tower kernels (tcgen05 kind::tf32 on (hi, lo)-split fp32
 * operands in 128-byte-swizzled shared-memory column blocks).  One CTA computes
 *   mode 0: D[M,N] = A[M,K] . B[N,K]^T    mode 1: D[M,N] = A[K,M]^T . B[K,N]    mode 2: D[M,N] = A[M,K] . B[K,N]
 * (A, B row-major fp32 in global memory; M in {64,128}; passes 3 = fp32-class, 1 = single TF32 pass) and dumps the
 * raw TMEM accumulator as out[128 lanes][cols], cols = N rounded up to a power of two >= 32.
 * ------------------------------------------------------------------------------------------ */
/* debug: per-CTA globaltimer trace of the tcgen05 tower kernels (buffer of n_ctas * 128 uint64, NULL = off);
 * code = 10 * (0 forward, 1 backward) + stage selects the launches that record */
int cfm_debug_set_trace(uint64_t* buf, int64_t code);
/* timing probe: reps back-to-back M x N x 8 tcgen05.mma over nacc accumulators; out[0] = cycles to completion, out[1] = issue cycles */
int cfm_tc_mma_probe(long long* out, int64_t M, int64_t N, int64_t reps, int64_t nacc, void* stream);
int cfm_tc_selftest(const float* A, const float* B, float* out, int64_t mode, int64_t M, int64_t N, int64_t K,
                    int64_t passes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CFM_B200_H */
