/*
 * dgll_b200.h — C ABI of the B200-native neighbourhood-aggregation path.
 *
 * This is the drop-in boundary for the hot path of dke-lab/dgll: every entry
 * point takes plain device pointers + sizes + a CUDA stream (as void*), never
 * a torch type, and returns an int status (0 = OK) instead of calling
 * exit(1).  Each declaration cites the reference interface it replaces
 * (paths relative to the reference tree).
 *
 * Conventions (all entry points unless stated otherwise)
 *   - pointers are DEVICE pointers borrowed from the caller (caller owns/frees);
 *   - `stream` is a cudaStream_t cast to void* (NULL = legacy default stream);
 *     calls are stream-ordered and do NOT synchronise the device;
 *   - thread-safe: no global mutable state besides per-device cached properties
 *     and a thread-local error string (dgllb_last_error);
 *   - leading dimensions (ld*) are in ELEMENTS, sizes are element/row counts;
 *   - return codes: DGLLB_OK, DGLLB_ERR_INVALID (bad argument),
 *     DGLLB_ERR_CUDA (a CUDA call failed; message in dgllb_last_error),
 *     DGLLB_ERR_UNSUPPORTED (shape/dtype not supported by this build).
 *
 * The two legacy symbols at the bottom keep the reference's exact C ABI
 * (dgll/FusedKernel/gcn_fused_kernel.cu:190-195,238-244).
 */
#ifndef DGLL_B200_H_
#define DGLL_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGLLB_OK 0
#define DGLLB_ERR_INVALID 1
#define DGLLB_ERR_CUDA 2
#define DGLLB_ERR_UNSUPPORTED 3

/* reduce ops (NeighborAggregator aggr_method, dgll/nn/Convolution/sageconv.py:32-38;
 * Pooling reduce=, dgll/nn/GlobalPooling/Pooling.py:37,59,81) */
#define DGLLB_SUM 0
#define DGLLB_MEAN 1
#define DGLLB_MAX 2

/* feature dtypes */
#define DGLLB_F32 0
#define DGLLB_BF16 1

/* epilogue flags */
#define DGLLB_EPI_RELU 1 /* fmaxf(x,0) as gcn_fused_kernel.cu:67 */
#define DGLLB_EPI_ELU 2  /* F.elu as gatconv.py:143-145 */

/* GAT score variants */
#define DGLLB_GAT_SOFTMAX 0     /* softmax_j(leakyrelu(.)), gatconv.py:30-54 (dense gatConv) */
#define DGLLB_GAT_EXP_NEG 1     /* exp(-leakyrelu(.))/rowsum, gatconv.py:125-139 (sparseGatConv) */

/* ---------------------------------------------------------------- misc -- */

/* ABI version of this library (major*1000+minor). */
int dgllb_version(void);
/* Thread-local, NUL-terminated description of the last error on this thread. */
const char* dgllb_last_error(void);
/* Properties of the CURRENT device (cached per device). Any pointer may be NULL. */
int dgllb_device_info(int* sm_count, int* cc_major, int* cc_minor, int64_t* l2_bytes);
/* Number of kernels this library has launched on this process (all threads).
 * bench.py reads it to report `gpu_launches`. */
int64_t dgllb_launch_count(void);
/*
 * Tuning options (kernel family pins, block sizes, NVTX on/off).  Each option is an integer that starts from the
 * environment variable DGLLB_<NAME> (read ONCE per process) and is changed afterwards only here; no launch path calls
 * getenv().  `value` is a decimal integer or one of the option's words; NULL / "" / "auto" = library default.
 *   spmm_kernel  auto | rowsplit | stream | wholerow        gat_kernel  auto | group | row
 *   gat_bwd_kernel  auto | twopass | fused  (fused = the single pass over the transposed CSR, default where it applies)
 *   spmm_tb, rows_tb, rows_ns, rows_d, rows_stream (1 = off, n >= 2 = n rows per warp), gat_row_warps, gat_bwd_tb,
 *   gat_bwd_depth (gradient rows in flight per lane: 4 | 8), bin_tb                               integers (0 = default)
 *   rows_sharded_bps  resident 64-thread blocks per SM of dgllb_spmm_csr_sharded (5..16; 0 = no limit) — a caller that
 *                     runs it beside other work (the pipelined trainer) keeps it from filling the register file
 *   gemm_kernel  3 = precision-2 one-tile-per-CTA kernel, 4 = precision-2 persistent kernel, 5 = precision 0 always on
 *                the exact FMA kernel (0 = automatic choice)
 *   nvtx         1 = NVTX ranges around the entry points named like the reference's (FeatureCache/storage.py:164-206)
 */
int dgllb_set_option(const char* name, const char* value);
int dgllb_get_option(const char* name, int* value_out);

/* ---------------------------------------------------- CSR aggregation -- */

/*
 * Opaque load-balancing plan for a static CSR (rows longer than a chunk are
 * split into nnz-chunks that are reduced with atomics).  Optional: every
 * aggregation entry point accepts plan == NULL (pure row-split).
 * Replaces nothing in the reference (it launches one block per row,
 * gcn_fused_kernel.cu:211); it exists because Reddit-shaped degree skew needs it.
 * dgllb_csr_plan_create synchronises `stream` once (reads back the item count).
 */
typedef struct dgllb_csr_plan dgllb_csr_plan;
int dgllb_csr_plan_create(const void* row_ptr, int row_ptr_is64, int64_t n_rows,
                          int chunk_edges /* 0 = default */, void* stream,
                          dgllb_csr_plan** plan_out);
int dgllb_csr_plan_info(const dgllb_csr_plan* plan, int64_t* n_heavy_rows,
                        int64_t* n_chunks, int* chunk_edges);
void dgllb_csr_plan_destroy(dgllb_csr_plan* plan);

/*
 * out[i, 0:F] = epi( row_scale[i] * reduce_{e in [row_ptr[i], row_ptr[i+1])}
 *                     ( values[e] * X[col_idx[e], 0:F] )  + addend[i, 0:F] + bias[0:F] )
 *
 * The SpMM behind: torch.spmm(adj, support) dgll/nn/Convolution/gcnconv.py:31;
 * torch.sparse.mm Evaluation/PPI/gcn_model.py:76; the aggregation loop
 * gcn_fused_kernel.cu:41-57; NeighborAggregator mean/sum/max sageconv.py:32-38;
 * DGL GraphConv/SAGEConv update_all as called at GPU Accelerator/CommGNNModel.py:23-28,72-77;
 * GinConv Adj@Feat ginconv.py:27; scatter() pooling Pooling.py:37,59,81
 * (pass col_idx == NULL: identity columns, i.e. a segment reduce).
 *
 *   row_ptr     int32[n_dst+1] or int64[n_dst+1] (row_ptr_is64)
 *   nnz         an UPPER BOUND on row_ptr[n_dst] known to the host (normally the length of col_idx); it
 *               sizes the grid of the streaming (row-aligned nnz-split) kernel without a device read-back.
 *               Pass a negative value when unknown: the one-warp-per-row kernel runs instead.
 *   col_idx     int32[nnz] (rows of X; NULL = identity: column e is row e)
 *   values      float[nnz] or NULL (all ones)
 *   X           x_dtype[n_src, ldx] row-major (DGLLB_F32 / DGLLB_BF16), F <= ldx
 *   out         float[n_dst, ldo]
 *   reduce      DGLLB_SUM / DGLLB_MEAN (divide by row degree; empty row -> 0) /
 *               DGLLB_MAX (empty row -> 0)
 *   row_scale   float[n_dst] or NULL; addend float[n_dst, ld_add] or NULL;
 *   bias        float[F] or NULL; epilogue = bit-or of DGLLB_EPI_*
 *   argmax_out  int32[n_dst, F] or NULL (DGLLB_MAX only): edge index e that won,
 *               -1 for empty rows (needed by the max backward)
 * 128-bit vector loads are used when X is 16-byte aligned and ldx is a multiple
 * of 4 (f32) / 8 (bf16); 128-bit stores when out is 16-byte aligned and ldo a multiple
 * of 4; otherwise scalar accesses.  Accumulation is fp32 in CSR edge order per row
 * (deterministic unless `plan` splits the row).
 */
int dgllb_spmm_csr(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                   const float* values, const void* X, int x_dtype, int64_t ldx,
                   float* out, int64_t ldo, int64_t n_dst, int64_t n_src, int64_t nnz, int F,
                   int reduce, const float* row_scale, const float* addend,
                   int64_t ld_add, const float* bias, int epilogue,
                   int32_t* argmax_out, const dgllb_csr_plan* plan, void* stream);

/*
 * The same sum/mean aggregation with the source table node-range PARTITIONED over the GPUs of one NVSwitch box and read
 * in place: row id c lives in shard c / rows_per_shard at local row c % rows_per_shard; shard_ptrs[n_shards] is a DEVICE
 * array of the shards' base addresses (the local shard and the peers' shards mapped with dgllb_ipc_import).  The halo
 * exchange of SURVEY.md §8(e) — bucket ids by owner, all_to_all ids, owner-side gather, all_to_all rows, un-permute —
 * becomes the aggregation kernel's own 128-bit loads over NVLink; nothing is staged or read back.  With n_shards == 1
 * it is the fused gather+aggregation over a resident table (dgll/data/dgraph.py:105 + the mean of
 * GPU Accelerator/CommGNNModel.py:72-77).  stride_bytes must be a multiple of 16.  No plan: rows are expected short
 * (sampled blocks); no epilogue.
 */
int dgllb_spmm_csr_sharded(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                           const float* values, const void* const* shard_ptrs, int n_shards,
                           int64_t rows_per_shard, int64_t stride_bytes, int x_dtype, float* out,
                           int64_t ldo, int64_t n_dst, int F, int reduce, void* stream);

/*
 * SDDMM: out_e[e] = < A[row(e), 0:F], B[col_idx[e], 0:F] > for every CSR edge.
 * Replaces the dense N x N product + gather in SpecialSpmmFunction.backward,
 * dgll/nn/Convolution/gatconv.py:76-78.
 */
int dgllb_sddmm_csr(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                    const float* A, int64_t lda, const float* B, int64_t ldb,
                    float* out_e, int64_t n_rows, int F, void* stream);

/*
 * Scatter for the MAX backward: grad_X[col_idx[argmax[i,f]], f] += grad_out[i,f].
 * grad_X must be pre-zeroed by the caller.
 */
int dgllb_spmm_max_backward(const int32_t* col_idx, const int32_t* argmax,
                            const float* grad_out, int64_t ldg, float* grad_X,
                            int64_t ldx, int64_t n_dst, int F, void* stream);

/*
 * CSR transpose (stable sort by column): builds the CSR of A^T, carrying an
 * optional per-edge value array and producing the edge permutation
 * perm[e_T] = e (index of the same edge in the input CSR).  Used for
 * grad_X = A^T grad_out (gatconv.py:80 `a.t().matmul(grad_output)`).
 * All outputs are caller-allocated: t_row_ptr (same width as row_ptr)[n_cols+1],
 * t_col_idx int32[nnz], t_values float[nnz] or NULL (ones when values == NULL),
 * perm int32[nnz] or NULL.  Within a transposed row, edges keep source-row
 * order (stable), so the result is deterministic.  nnz must be < 2^31-1.
 * Entries whose column id equals n_cols are padding: they belong to no row of the
 * result (t_row_ptr[n_cols] = number of real entries), so a fixed-capacity edge
 * array padded with n_cols can be transposed without knowing its fill on the host.
 * Workspace comes from the stream-ordered allocator (cudaMallocAsync).
 */
int dgllb_csr_transpose(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                        const float* values, int64_t n_rows, int64_t n_cols, int64_t nnz,
                        void* t_row_ptr, int32_t* t_col_idx, float* t_values,
                        int32_t* perm, void* stream);

/* ------------------------------------------------------ feature gather -- */

/*
 * out[i, 0:row_bytes) = table[ids[i], 0:row_bytes)   (exact byte copy)
 * Replaces features[nodes] dgll/data/dgraph.py:105 and the cache gather
 * gpu_fix_cache[name][cacheid] dgll/FeatureCache/storage.py:185-187,205-209.
 * ids are int64 (ids_is64) or int32.  Rows move as TMA bulk copies
 * (global->shared->global) when table/out base and strides are multiples of
 * 16 bytes; otherwise a vector/scalar LDG path runs.  ids are NOT range-checked.
 */
int dgllb_gather_rows(const void* table, int64_t table_stride_bytes, const void* ids,
                      int ids_is64, void* out, int64_t out_stride_bytes, int64_t n_rows,
                      int64_t row_bytes, void* stream);

/*
 * GraphCacheServer.fetch_data split gather (storage.py:151-198):
 *   hit  (gpu_flag[id] != 0): out[i] = cache_table[localid2cacheid[id]]
 *   miss                    : out[i] = host_table[nid_map ? nid_map[id] : id]
 * host_table must be device-accessible (pinned+mapped host memory or device memory).
 * miss_count (device int64*, may be NULL) is atomically incremented by the number
 * of misses (storage.py:213-215 try_num/miss_num accounting).
 */
int dgllb_gather_rows_cached(const void* cache_table, int64_t cache_stride_bytes,
                             const void* host_table, int64_t host_stride_bytes,
                             const int64_t* ids, const uint8_t* gpu_flag,
                             const int64_t* localid2cacheid, const int64_t* nid_map,
                             void* out, int64_t out_stride_bytes, int64_t n_rows,
                             int64_t row_bytes, int64_t* miss_count, void* stream);

/*
 * Feature table sharded by NODE RANGE over the GPUs of one NVSwitch box (shard s holds rows
 * [s*rows_per_shard, (s+1)*rows_per_shard)), read in place over NVLink:
 *   out[i, 0:row_bytes) = shard_ptrs[id / rows_per_shard][(id % rows_per_shard) * stride_bytes ...]
 * shard_ptrs is a DEVICE array of n_shards device pointers (the local shard + peer shards mapped with
 * dgllb_ipc_import).  Replaces, for the partitioned table, the gather of dgll/data/dgraph.py:105 and the
 * host/RPC feature fetch of FeatureCache/storage.py:101-126 — no collective, no host read-back
 * (SURVEY.md §8 e "P2P-map all shards and let the gather kernel issue remote 128-bit loads").
 * A NEGATIVE id is a padding slot of a fixed-capacity id array (dgllb_build_block_cap): nothing is read, the output
 * row is zero-filled.
 */
int dgllb_gather_rows_sharded(const void* const* shard_ptrs, int n_shards, int64_t rows_per_shard,
                              int64_t stride_bytes, const void* ids, int ids_is64, void* out,
                              int64_t out_stride_bytes, int64_t n_rows, int64_t row_bytes, void* stream);

/*
 * CUDA IPC plumbing for the above.  export: 64-byte handle of the allocation containing dev_ptr + the offset of
 * dev_ptr inside it (caching allocators sub-allocate).  import: map a peer's allocation (peer access enabled
 * lazily) and return the pointer at `offset`.  release: unmap (pass the same offset).
 */
int dgllb_ipc_export(const void* dev_ptr, unsigned char* handle64, int64_t* offset);
int dgllb_ipc_import(const unsigned char* handle64, int64_t offset, void** dev_ptr_out);
int dgllb_ipc_release(void* dev_ptr, int64_t offset);

/* -------------------------------------------------- dense transform X.W -- */

/*
 * C[M,N] = epi( A[M,K] . B[K,N] + bias[N] ), all row-major fp32 in HBM.
 * precision 0: fp32-grade results, the parity path (<=1e-5 rel): products of >= 2.5e7 multiply-adds run as precision 3,
 *              smaller ones (and every one under option gemm_kernel=5) on the exact fp32 FMA kernel (SIMT);
 * precision 1: tcgen05 tensor cores, bf16 operands (packed by a pre-pass),
 *              fp32 accumulate in TMEM (<=1e-2 rel); the faster choice for large square products;
 * precision 2: tcgen05 tensor cores, TF32 operands read by TMA straight from the fp32 tensors (no packing, no
 *              transposition: operands stored the other way round are loaded MN-major), fp32 accumulate in TMEM
 *              (<=2e-3 rel) — the fast path for the layer shapes of this path, which are bound by HBM traffic.
 * precision 3: "3xTF32" — the precision-2 kernel with every fp32 operand word split in shared memory into
 *              hi = tf32(v) and lo = tf32(v - hi) and three MMAs per k-step (lo*hi + hi*lo + hi*hi), fp32
 *              accumulation in TMEM: 1-4e-6 of max|ref|, deterministic (fixed-order split-K).
 * transA / transB: use A^T (A stored [K,M]) / B^T (B stored [N,K]).
 * Replaces torch.mm(x, W) gcnconv.py:30, gcn_model.py:70, gatconv.py:117 and the
 * per-edge recomputed transform of gcn_fused_kernel.cu:46-54.
 */
int dgllb_gemm_f32(const float* A, int64_t lda, int transA, const float* B, int64_t ldb,
                   int transB, float* C, int64_t ldc, int64_t M, int64_t N, int64_t K,
                   const float* bias, int epilogue, int accumulate, int precision,
                   void* stream);

/* --------------------------------------------------------- fused GAT ---- */

/*
 * Fused SDDMM + edge-softmax + aggregation for multi-head GAT.
 *   z_e   = leakyrelu_slope( el[i,h] + er[j,h] )        i = row (dst), j = col_idx[e]
 *   mode DGLLB_GAT_SOFTMAX : alpha_e = softmax over row i of (+z_e)
 *   mode DGLLB_GAT_EXP_NEG : alpha_e = exp(-z_e) / sum_row exp(-z_e)
 *   out[i,h,:] = epi( sum_e alpha_e * Wh[j,h,:] )
 * computed in one pass with an online (running-max) softmax; the [E] score
 * vector is never materialised.  Replaces gatconv.py:117-139 (sparseGatConv,
 * two COO SpMMs + edge cat) and :30-54 (dense N^2 softmax).
 *   Wh  float[n_src, ldw]  heads*D <= ldw, head h occupies columns [h*D,(h+1)*D)
 *   el  float[n_dst, ld_e] (first `heads` columns), er float[n_src, ld_e]
 *       (el_i = a_l . Wh_i, er_j = a_r . Wh_j; with W extended by the columns
 *        W_h a_l,h and W_h a_r,h the dense transform emits Wh, el, er in one GEMM)
 *   out float[n_dst, ldo]
 *   row_max / row_sum: float[n_dst, heads] (compact) or NULL — saved for the backward
 *   (max of signed score, sum of exp(score - max)); rows without edges give out = 0.
 */
int dgllb_gat_forward(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                      const float* Wh, int64_t ldw, const float* el, const float* er,
                      int64_t ld_e, float* out, int64_t ldo, float* row_max, float* row_sum,
                      int64_t n_dst, int64_t n_src, int heads, int D, float slope,
                      int mode, int epilogue,
                      float drop_p /* attention dropout (gatconv.py:37, :132): applied to the normalised attention, the
                                      row sum keeps every edge; 0 = off */,
                      uint64_t drop_seed /* mask = f(seed, edge position in the CSR, head): reproducible in the backward */,
                      const dgllb_csr_plan* plan /* optional: long rows are split into chunks whose partial softmax
                                                    states are merged exactly (no atomics) */,
                      void* stream);

/*
 * Backward of dgllb_gat_forward (epilogue gradient already applied by the caller):
 * given g = dL/dout [n_dst, ldg] computes
 *   d_el[n_dst, ld_de], d_er[n_src, ld_de] (first `heads` columns), d_Wh[n_src, ldd]
 * Pass 1 runs over the forward CSR: recomputes alpha, forms dalpha = <g_i, Wh_j>
 * (the SDDMM) and writes (alpha, dz) per (edge, head) to `edge_ws`
 * float[2*nnz*heads] (8-byte aligned); pass 2 runs over the transposed CSR
 * (t_row_ptr/t_col_idx/perm from dgllb_csr_transpose) and accumulates
 * d_Wh_j = sum_i alpha_ij g_i and d_er_j.  No atomics: deterministic.
 * `plan` (built on row_ptr) / `t_plan` (built on t_row_ptr), both optional, split
 * rows longer than their chunk_edges into (row, chunk) items whose partial sums
 * go to an internal workspace and are merged in a fixed order (still no atomics).
 * `out` is the forward output BEFORE the epilogue activation.
 * Replaces SpecialSpmmFunction.backward gatconv.py:71-81 + autograd of :117-139.
 */
int dgllb_gat_backward(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                       const void* t_row_ptr, const int32_t* t_col_idx, const int32_t* perm,
                       const float* Wh, int64_t ldw, const float* el, const float* er,
                       int64_t ld_e, const float* out, int64_t ldo, const float* row_max,
                       const float* row_sum, const float* g, int64_t ldg, float* d_Wh,
                       int64_t ldd, float* d_el, float* d_er, int64_t ld_de, float* edge_ws,
                       int64_t n_dst, int64_t n_src, int heads, int D, float slope,
                       int mode, float drop_p, uint64_t drop_seed /* the forward's values */,
                       const dgllb_csr_plan* plan, const dgllb_csr_plan* t_plan, void* stream);

/* mask_out[e*heads + h] = 0 (dropped) or 1/(1-drop_p) (kept): the multiplier the GAT kernels apply to edge e, head h
 * for this seed.  Test / debugging aid: lets a host restatement replay the exact mask. */
int dgllb_gat_dropout_mask(uint64_t drop_seed, int64_t nnz, int heads, float drop_p, float* mask_out, void* stream);

/* ------------------------------------------------- binarized aggregation -- */

/*
 * Bit-pack the sign of a feature table: bit f of word (f/32) of row r is
 * (X[r,f] >= 0).  Pad bits are 0.  words_per_row >= ceil(F/32).
 * The reference only names this feature (README.md:11); semantics are
 * defined in SURVEY.md §8 a18.
 */
int dgllb_binarize_pack(const float* X, int64_t ldx, uint32_t* packed,
                        int64_t words_per_row, int64_t n_rows, int F, void* stream);

/*
 * cnt[i,f] = sum_{e in row i} bit f of packed[col_idx[e]]          (int32, exact)
 * out modes: 0 = counts as int32; 1 = float sum of +-1 (2*cnt - deg);
 *            2 = float mean of +-1 ((2*cnt - deg)/deg, empty row -> 0).
 * Bit-sliced formulation: lane l owns packed word l of every neighbour; counts are kept in
 * carry-save bit planes (32 features per bitwise op) and scattered feature-major with shuffles.
 */
int dgllb_bin_spmm_csr(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                       const uint32_t* packed, int64_t words_per_row, void* out,
                       int64_t ldo, int64_t n_dst, int F, int out_mode,
                       const dgllb_csr_plan* plan /* optional nnz-split of long rows; exact (integer atomics) */,
                       void* stream);

/* ------------------------------------------------ sampling / blocks ------ */

/*
 * Uniform neighbour sampling WITHOUT replacement, min(deg, fanout) per seed
 * (the device-side counterpart of Base_sampler.sample_neighbours,
 * dgll/sampling/base_sampler.py:45-58; fanout < 0 keeps all neighbours).
 * Output is a CSR block by destination with GLOBAL source ids:
 *   out_row_ptr int32[n_seeds+1] (exclusive scan of min(deg,fanout)),
 *   out_col     int32[>= n_seeds*fanout] (caller-sized; exact nnz = out_row_ptr[n_seeds]).
 * Neighbour order inside a row is the order of the sampled positions
 * (ascending position), deterministic for a given (seed, row).
 * Bit-exact parity with Python's random.sample is impossible on device; the
 * parity path feeds host-sampled index lists instead.
 */
int dgllb_sample_neighbors(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                           const void* seeds, int seeds_is64, int64_t n_seeds, int fanout,
                           uint64_t rng_seed, int32_t* out_row_ptr, int32_t* out_col,
                           void* stream);

/*
 * Block (MFG) construction: dst-first compaction of a sampled edge list, on the device.
 *   dst_ids     int64[n_dst]  (unique; they become local ids 0..n_dst-1 in order)
 *   row_ptr     int32[n_dst+1], nbr_global int32[>= row_ptr[n_dst]]  (output of dgllb_sample_neighbors);
 *               nnz_cap = allocated length of nbr_global / col_local (the true nnz is read on the device)
 *   src_ids     int64[n_dst + nnz_cap] out: unique ids of cat(dst_ids, neighbours) in first-occurrence order
 *   col_local   int32[nnz_cap] out: neighbours relabelled into src_ids' index space
 *   counts_out  int32[2] out (device): {num_src, nnz}
 * Deterministic (no sort, no order-dependent atomics).  Replaces the relabelling DGL's to_block does for the blocks
 * consumed at GPU Accelerator/MQGCN.py:41-50 and the unique-node bookkeeping of dgll/sampling/base_sampler.py:81.
 */
int dgllb_build_block(const int64_t* dst_ids, int64_t n_dst, const int32_t* row_ptr,
                      const int32_t* nbr_global, int64_t nnz_cap, int64_t* src_ids, int32_t* col_local,
                      int32_t* counts_out, void* stream);

/*
 * Fixed-capacity variants for pipelines that must not read sizes back (CUDA-graph capture of sampler + block builder +
 * training step).  Convention: a NEGATIVE id is a padding slot.
 *   dgllb_sample_neighbors_cap: seeds[n_seeds_cap] may end in -1 entries (degree 0 rows); the effective random seed is
 *     rng_seed + *rng_offset (device memory, may be NULL), so a replayed graph draws new samples when a tiny kernel bumps
 *     the offset.
 *   dgllb_build_block_cap: dst_ids[n_dst_cap] may end in -1 entries; src_ids is pre-filled with -1 (so it can be the next
 *     layer's padded seed array as it is), unused slots of col_local[nnz_cap] are set to col_pad (the padding column id of
 *     the consumer, e.g. the n_cols that dgllb_csr_transpose drops), counts_out int32[3] = {num_src, nnz, n_dst_valid}.
 */
int dgllb_sample_neighbors_cap(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx,
                               const void* seeds, int seeds_is64, int64_t n_seeds_cap, int fanout,
                               uint64_t rng_seed, const uint64_t* rng_offset, int32_t* out_row_ptr,
                               int32_t* out_col, void* stream);
int dgllb_build_block_cap(const int64_t* dst_ids, int64_t n_dst_cap, const int32_t* row_ptr,
                          const int32_t* nbr_global, int64_t nnz_cap, int col_pad, int64_t* src_ids,
                          int32_t* col_local, int32_t* counts_out, void* stream);

/* ------------------------------------------- layer-wise importance sampling -- */

/*
 * FastGCN / LADIES ("flat", "WRS" switches) on the device, replacing the scipy pipeline of
 * GPU Accelerator/MQLadies.py:74-89, MQFastGCN.py:72-88, MQFastGCNFlatWrs.py:76-101 and utils.py:199-213.
 * The Laplacian is a CSR with int32 columns and fp64 values (scipy's dtype).  Capacity convention: arrays sized by an
 * upper bound, true counts written to device memory, so one layer needs two small read-backs.
 *
 * dgllb_csr_slice_rows_ptr / _fill:  Q = M[rows, :]   (out_row_ptr int64[n_rows+1]; read out_row_ptr[n_rows] = nnz(Q)
 *   before allocating out_col / out_values; values may be NULL).
 */
int dgllb_csr_slice_rows_ptr(const void* row_ptr, int row_ptr_is64, const int64_t* rows, int64_t n_rows,
                             int64_t* out_row_ptr, void* stream);
int dgllb_csr_slice_rows_fill(const void* row_ptr, int row_ptr_is64, const int32_t* col_idx, const double* values,
                              const int64_t* rows, int64_t n_rows, const int64_t* out_row_ptr, int32_t* out_col,
                              double* out_values, void* stream);
/*
 * Candidate columns and their probabilities: for every distinct column c of the nnz entries,
 *   prob_i[c] = sum of values^2 (sqrt of it if flat), accumulated in stored order, products and sums rounded
 *   separately (bit-identical to sum(Q.multiply(Q), axis=0)); prob = prob_i / sum(prob_i).
 * cand_cols int32[nnz] / cand_prob double[nnz]: the first n_cand entries are the distinct columns in ascending order
 * and their normalised probabilities, the rest -1 / 0.  stats double[3] (device) = {n_cand, sum(prob_i), #(prob_i > 0)}.
 */
int dgllb_col_sqsum(const int32_t* col_idx, const double* values, int64_t nnz, int64_t n_cols, int flat,
                    int32_t* cand_cols, double* cand_prob, double* stats, void* stream);
/*
 * Draw min(fanout, #(prob > 0)) candidates without replacement with probabilities cand_prob, in drawing order
 * (exponential keys -log(u)/p, smallest first; u from a counter-based generator keyed by (seed, node id)).
 * cand_cols NULL = candidate i is node i.  sel int32[fanout] = candidate indices, picks int64[fanout] = node ids
 * (-1 beyond the count), count int64[1] (device) = number drawn.  Same distribution as
 * np.random.choice(n, s_num, p=prob, replace=False) (utils.py:201, MQFastGCN.py:80), not the same stream.
 */
int dgllb_weighted_choice(const int32_t* cand_cols, const double* cand_prob, int64_t n_cand, int fanout,
                          uint64_t seed, int32_t* sel, int64_t* picks, int64_t* count, void* stream);
/*
 * Column scales for the drawn candidates sel[0..count): mode 0 = 1/p/count (MQFastGCN.py:82), mode 1 = the WRS
 * estimator (utils.py:199-213, same operations in the same order, n_total = len(p) there).  scale double[cap],
 * zero beyond count.
 */
int dgllb_importance_scale(const double* cand_prob, const int32_t* sel, const int64_t* count, int cap,
                           int64_t n_total, int mode, double* scale, void* stream);
/* pos[picks[k]] = k for k < min(cap, *count) (count NULL = cap), or -1 when reset != 0; picks < 0 skipped. */
int dgllb_scatter_pos(int32_t* pos, const int64_t* picks, const int64_t* count, int64_t cap, int reset, void* stream);
/*
 * adj = Q[:, picks].multiply(scale).tocsr():  keeps the entries whose column c has pos[c] >= 0, relabels them to
 * pos[c], scales by scale[pos[c]] (NULL = 1), rows sorted by the new label.  nnz_cap = allocated length of q_col /
 * out_col / out_values (nnz(Q) is read from q_row_ptr[n_rows] on the device); out_row_ptr[n_rows] = nnz(adj).
 * q_values / out_values may be NULL (topology only: what create_block receives at MQLadies.py:85).
 */
int dgllb_csr_select_cols(const int64_t* q_row_ptr, const int32_t* q_col, const double* q_values, int64_t n_rows,
                          int64_t nnz_cap, const int32_t* pos, const double* scale, int64_t* out_row_ptr,
                          int32_t* out_col, double* out_values, void* stream);

/* ------------------------------------------------------------ legacy ABI -- */

/*
 * Same symbols and argument lists as dgll/FusedKernel/gcn_fused_kernel.cu:190-195
 * and :238-244.  Semantics: H = relu(A_hat (X W)) and its TRUE gradients
 * (the reference backward is numerically wrong):
 *   grad_W += (A_hat X)^T (G*mask), grad_X += A_hat^T ((G*mask) W^T), mask = H>0,
 * accumulated into the caller-zeroed buffers as gcn_extension.cpp:84-85 expects.
 * Differences kept on purpose: launches go to the legacy default stream and the
 * call blocks until completion (as the reference does with cudaDeviceSynchronize),
 * but errors are reported on stderr + dgllb_last_error instead of exit(1).
 * row_ptr is taken as authoritative (num_neighbors[i] is honoured as
 * min(num_neighbors[i], row_ptr[i+1]-row_ptr[i]) exactly like the guard at .cu:42).
 */
void launch_gcn_fused_kernel(const int* row_ptr, const int* col_idx, const float* values,
                             const float* X, const float* W, float* H,
                             const int* num_neighbors, int N, int F_padded, int actual_F,
                             int H_dim, int total_nnz);
void launch_gcn_fused_kernel_backward_optimized(
    const int* row_ptr, const int* col_idx, const float* values, const float* X,
    const float* W, const float* grad_output, float* grad_W, float* grad_X,
    const int* num_neighbors, int N, int F_padded, int actual_F, int H_dim, int total_nnz);

/* v2 of the above: stream-ordered, non-blocking, returns a status.
 * `H` (forward output) is required by the backward for the ReLU mask. */
int dgllb_gcn_fused_forward(const int* row_ptr, const int* col_idx, const float* values,
                            const float* X, const float* W, float* H,
                            const int* num_neighbors, int N, int F_padded, int actual_F,
                            int H_dim, int total_nnz, void* stream);
int dgllb_gcn_fused_backward(const int* row_ptr, const int* col_idx, const float* values,
                             const float* X, const float* W, const float* H,
                             const float* grad_output, float* grad_W, float* grad_X,
                             const int* num_neighbors, int N, int F_padded, int actual_F,
                             int H_dim, int total_nnz, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DGLL_B200_H_ */
