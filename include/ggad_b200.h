/*
 * ggad_b200 -- C ABI of the B200-native GGAD message-passing / outlier-synthesis path.
 *
 * The reference (mala-lab/GGAD) is pure Python and has NO FFI / plugin interface for
 * this path; its boundary is the nn.Module surface.  The entry
 * points below are what a binding for that path has to bind: each one names the
 * reference code it replaces (paths relative to the reference checkout).  The Python
 * host side (ggad_b200/*.py) binds them with ctypes and re-creates the reference's
 * module classes on top; INTEGRATION.md shows the stub a reference maintainer adds.
 *
 * Conventions
 *   - plain pointers and sizes only; all pointers are DEVICE pointers unless the
 *     parameter name ends in _host; `stream` is a cudaStream_t passed as void*.
 *   - every function returns 0 (GGAD_OK) or a negative GGAD_ERR_* code and never
 *     throws; ggad_last_error() gives a thread-local message.  Work is enqueued on
 *     `stream` and is asynchronous unless stated otherwise.
 *   - feature matrices are row-major fp32 with a leading dimension (in floats) that is
 *     a multiple of 4 and a 16-byte aligned base; the logical width `d` must also be a
 *     multiple of 4 (pad with zero columns -- sums stay exact).  CSR: rowptr int64,
 *     col int32, val fp32 or NULL (= all ones).
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns GGAD_ERR_CUDA.
 */
#ifndef GGAD_B200_H_
#define GGAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GGAD_OK 0
#define GGAD_ERR_INVALID (-1)     /* bad argument (null pointer, negative size, d % 4 != 0 ...) */
#define GGAD_ERR_ALIGN (-2)       /* pointer / leading dimension not 16-byte aligned */
#define GGAD_ERR_CUDA (-3)        /* CUDA runtime error (message in ggad_last_error) */
#define GGAD_ERR_UNSUPPORTED (-4) /* width or size outside what the kernels cover */

#ifndef GGAD_TILE_ITEMS
#define GGAD_TILE_ITEMS 2048 /* merge-path items (rows + edges) per CTA tile */
#endif
#define GGAD_MAX_WIDTH 768   /* max logical width d per launch (745 pads to 748) */

typedef void* ggad_stream_t;

#if defined(__GNUC__)
#define GGAD_API __attribute__((visibility("default")))
#else
#define GGAD_API
#endif

/* ---- library / device -------------------------------------------------------- */
GGAD_API int ggad_version(void);
GGAD_API const char* ggad_last_error(void);
/* number of CUDA kernels this library has launched since it was loaded */
GGAD_API int64_t ggad_launch_count(void);
/* The K7 index kernels take their temporaries (radix-sort double buffers, scan storage) from a library-owned
 * stream-ordered memory pool that is kept across calls; this returns it to the driver (synchronises the device). */
GGAD_API int ggad_trim_workspace(void);
/* Grow that pool to at least `bytes` once (stream-ordered allocate + free), so that a caller whose temporaries
 * vary from call to call (mini-batch frontiers) does not pay the driver at every new maximum. */
GGAD_API int ggad_reserve_workspace(int64_t bytes, ggad_stream_t stream);
GGAD_API int ggad_device_info(int* sm_count, int64_t* l2_bytes, int* cc_major, int* cc_minor, int64_t* hbm_bytes);

/* ---- K1/K2/K3/K5: CSR neighbor gather-reduce with fused per-row epilogue -------
 *
 *   acc_r  = row_scale[r] * sum_{e in row r} val[e] * col_scale[c_e] * X[xmap[c_e]]      c_e = col[e]
 *   z_r    = acc_r + bias                                    (written to z if non-NULL)
 *   y_r    = prelu_slope ? PReLU(z_r) : relu ? ReLU(z_r) : z_r   (written to y if non-NULL)
 *   sumsq_r   = |y_r|^2                                      (if sumsq non-NULL)
 *   dot_out_r = dot_scale[r] * < y_r , dot_mat[dot_rows ? dot_rows[r] : r] >   (if dot_out non-NULL)
 *
 * Edges whose col_scale is 0 or whose xmap is negative are skipped without loading X.
 *
 * Replaces: torch.bmm / torch.spmm in GCN.forward (model.py:26-35, bias + PReLU fused);
 * their autograd backward (same call on the transposed CSR); adj[0,S,:] @ emb
 * (model.py:151-155, on the row-extracted CSR); sim*raw_adj column sums (run.py:182-188,
 * dot epilogue on CSR(R^T) rows with col_scale = 1/|e_i|); mask.mm(embed_matrix) in
 * MeanAggregator / GCNAggregator (src/graphsage.py:98,326,355) and mask_row.mm(...)
 * (src/graphsage.py:421) on batch-local block CSRs with xmap = frontier node ids.
 *
 * If tile_row/tile_edge/ws are given (see ggad_plan_*), the merge-path tiled kernel is
 * used: every CTA owns GGAD_TILE_ITEMS (rows+edges), stages its col/val slice in
 * shared memory with a TMA bulk copy, splits it evenly over lane groups, and combines
 * partial rows deterministically (no float atomics).  Otherwise a group-per-row kernel
 * is used (fine for small graphs; serialises on hub rows).
 */
typedef struct ggad_gather_desc {
  /* CSR of the aggregation operator */
  const int64_t* rowptr; /* [n_rows + 1] */
  const int32_t* col;    /* [nnz] */
  const float* val;      /* [nnz] or NULL */
  int64_t n_rows;
  int64_t nnz;
  /* gathered operand */
  const float* x;         /* [n_x_rows, ldx] */
  int64_t ldx;            /* floats, multiple of 4 */
  const int32_t* xmap;    /* NULL or [n_cols]: row of x for column c (<0: skip) */
  const float* col_scale; /* NULL or [n_cols] */
  const float* row_scale; /* NULL or [n_rows] */
  int32_t d;              /* logical width, multiple of 4, <= GGAD_MAX_WIDTH */
  int32_t relu;           /* 1: ReLU epilogue (ignored if prelu_slope given) */
  /* epilogue */
  const float* bias;        /* NULL or [d] */
  const float* prelu_slope; /* NULL or [1] (device) */
  float* y;                 /* NULL or [n_rows, ldy] */
  float* z;                 /* NULL or [n_rows, ldy] pre-activation */
  int64_t ldy;
  float* sumsq;            /* NULL or [n_rows] */
  const float* dot_mat;    /* NULL or [*, lddot] */
  int64_t lddot;
  const int32_t* dot_rows; /* NULL or [n_rows] */
  const float* dot_scale;  /* NULL or [n_rows] */
  float* dot_out;          /* NULL or [n_rows] */
  /* merge-path plan (all three or none) */
  const int32_t* tile_row;  /* [n_tiles + 1] */
  const int64_t* tile_edge; /* [n_tiles + 1] */
  int64_t n_tiles;
  float* ws; /* [2 * n_tiles * d] partial-row workspace */
  /* fused exchange (multi-GPU): every finished row of y is ALSO stored to these peer-mapped buffers
   * (NVLink P2P stores issued by the CTA that finished the row, at the end of its tile; e.g. torch
   * symmetric memory), or once to an NVSwitch multicast address (multimem.st).  y must be given.  Pointers address row 0 of THIS launch's rows inside the
   * replicated [N, ldy] matrix of each peer; same ldy as y.  The caller synchronises the ranks
   * (barrier) before anyone reads the replicated matrix. */
  float* y_peer[7];
  int32_t n_peer;
  int32_t tile_epoch; /* chase mode (tile_done below): value published per finished tile, != the buffer's old contents */
  float* y_multicast; /* NULL or multicast address covering all ranks: alone = every row once through the switch;
                         together with y_peer[] + mc_min_peers (below) = only the rows that many peers need */
  /* halo exchange: NULL (every row goes to every peer) or [n_rows] bit masks -- bit p set means y_peer[p]
   * gathers row r in its next pass (the row is a column of that peer's CSR shard), so only those rows
   * cross NVLink.  Rows a peer does not need are left untouched in its replica. */
  const uint32_t* peer_need;
  /* chase mode: NULL, or [n_tiles] flags (needs the plan and y_peer/peer_need as above).  The tiled kernel then
   * does NOT push its rows itself; every CTA publishes tile_done[k] = tile_epoch (release, GPU scope) once the rows
   * its tile finished are in y, and a concurrently running ggad_halo_chase moves them over NVLink from a few SMs
   * of its own.  Rows cut by tile boundaries are still pushed by the fix-up kernel of this launch. */
  int32_t* tile_done;
  /* hybrid exchange: with y_peer[] AND y_multicast given, a row whose mask has at least mc_min_peers bits set is
   * stored once through the multicast address (it lands in every rank's replica) instead of once per peer that
   * needs it; 0 = off (y_multicast alone then means: every row through the multicast address). */
  int32_t mc_min_peers;
  /* number of rows of x, or 0 when the caller does not state it.  Only the TMA row-staging variant of the kernel
   * (environment GGAD_TMA_ROWS, an A/B knob) needs it, to bound its tensor map; it is skipped when this is 0. */
  int32_t n_x_rows;
} ggad_gather_desc_t;

GGAD_API int ggad_gather_reduce(const ggad_gather_desc_t* desc, ggad_stream_t stream);

/* Halo exchange of a row block that was not produced by a gather launch (e.g. a rank's shard of the input
 * features): row r of y[n_rows, ldy] is stored at the same row offset of y_peer_host[p] (HOST array of n_peer
 * peer-mapped DEVICE pointers, each addressing row 0 of this block inside the peer's replicated matrix) for
 * every p whose bit is set in peer_need[r] (NULL = all rows to all peers).  Same masks and semantics as
 * ggad_gather_desc_t.peer_need; the reference has no multi-GPU path -- this is the exchange
 * step of the destination-range sharded layer pass. */
GGAD_API int ggad_halo_push(const float* y, int64_t ldy, int64_t n_rows, int32_t d, const uint32_t* peer_need,
                            float* const* y_peer_host, int32_t n_peer, ggad_stream_t stream);

/* Exchange half of the chase mode: a persistent kernel (n_ctas CTAs, 0 = default) to be enqueued on a SECOND,
 * higher-priority stream right AFTER the ggad_gather_reduce launch that carries the same tile_done / tile_epoch.
 * Warp w waits (acquire, bounded spin) for tile_done[k] == tile_epoch of tiles k = w, w + W, ... and copies the
 * rows that tile finished -- [tile_row[k] (+1 if that row began in an earlier tile), tile_row[k+1]) -- from y
 * into the peers selected by peer_need, exactly as the in-kernel push does.  A row whose mask has at least
 * mc_min_peers bits set is instead stored ONCE through the NVSwitch multicast address y_multicast (if non-NULL;
 * multimem.st lands in every rank's replica, which is harmless for ranks that do not gather the row).
 * Deadlock-free in every launch order: enqueued after the gather it can at worst run after it (no overlap).
 * The reference has no multi-GPU path. */
typedef struct ggad_chase_desc {
  const float* y; /* this rank's rows of the replicated matrix (same pointer as the gather's y) */
  int64_t ldy;
  int32_t d;
  int32_t n_peer;
  const int64_t* rowptr; /* CSR of the gather launch (to tell whether a tile's first row began earlier) */
  int64_t n_rows;
  const int32_t* tile_row;
  const int64_t* tile_edge;
  int64_t n_tiles;
  const int32_t* tile_done;
  int32_t tile_epoch;
  int32_t n_ctas;
  const uint32_t* peer_need; /* NULL = every row to every peer */
  float* y_peer[7];
  float* y_multicast;
  int32_t mc_min_peers; /* <= 0: never multicast */
  int32_t reserved;
} ggad_chase_desc_t;
GGAD_API int ggad_halo_chase(const ggad_chase_desc_t* desc, ggad_stream_t stream);

/* Merge-path plan for a CSR: n_tiles = ceil((n_rows + nnz) / GGAD_TILE_ITEMS). */
GGAD_API int64_t ggad_plan_num_tiles(int64_t n_rows, int64_t nnz);
GGAD_API int ggad_plan_build(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int32_t* tile_row /*[n_tiles+1]*/,
                    int64_t* tile_edge /*[n_tiles+1]*/, ggad_stream_t stream);

/* ---- K6: dense projections on the tensor cores ----------------------------------------
 *   C[M,N] = act( alpha * op(A) * op(B) + beta * C ),  row-major fp32, act = ReLU if relu != 0
 *   op(A) = A[M,K] (lda >= K)  or, trans_a, A stored as [K,M] (lda >= M);
 *   op(B) = B[K,N] (ldb >= N)  or, trans_b, B stored as [N,K] (ldb >= K)  -- a torch nn.Linear weight.
 * Replaces nn.Linear / torch.mm around the aggregation: seq_fts = X W^T (model.py:27), fc1..fc4 (model.py:156,
 * 176-180), W . agg^T / fc(ego) / weight . emb (src/graphsage.py:412,419,430,174) and the two backward GEMMs of each
 * (dX = dY W: no transposes; dW = dY^T X: trans_a).
 * path 0 (auto): problems with 16-byte aligned bases / leading dimensions and enough work run on tcgen05 (operands
 * split into three bf16 terms each on the fly, bf16 MMAs accumulated in fp32 tensor memory: ~2^-22 relative error,
 * i.e. fp32 accuracy; plain TF32 would miss the 1e-4 parity tolerance); the rest (e.g. the h/4 -> 1 score layer,
 * mini-batch blocks) on a register-tiled SIMT FFMA kernel.  path 1 / 2 force SIMT / tensor core (2 fails with
 * GGAD_ERR_UNSUPPORTED if the operands do not qualify). */
GGAD_API int ggad_dense_matmul(int32_t trans_a, int32_t trans_b, int64_t m, int64_t n, int64_t k, const float* a, int64_t lda,
                               const float* b, int64_t ldb, float* c, int64_t ldc, float alpha, float beta, int32_t relu,
                               int32_t path, ggad_stream_t stream);

/* Plan with a FIXED tile count n_tiles >= ggad_plan_num_tiles(n_rows, nnz) (the extra tiles are empty): a CSR whose arrays
 * are static buffers padded to a capacity -- so that a gather launch captured in a CUDA graph with the capacity as its
 * nnz keeps working when the real edge count changes from replay to replay; only this (tiny) plan kernel is re-run.
 * Used by the graphed mini-batch step for the ego-mean operator of src/graphsage.py:421 and its transpose. */
GGAD_API int ggad_plan_build_padded(const int64_t* rowptr, int64_t n_rows, int64_t nnz, int64_t n_tiles, int32_t* tile_row,
                                    int64_t* tile_edge, ggad_stream_t stream);

/* ---- K6 (mini-batch): the dense tail of a GGAD training batch, fused -------------------------------------
 * From combined [B,h] = ReLU(W agg^T)^T (src/graphsage.py:412), ego [B,h] = mask_row . emb_U (:421), fc [h,h] (:430),
 * weight [h] (:174) and labels [B] in {0,1}: outlier generation ReLU(fc ego), the label-0-first column permutation of
 * combined_all (:450), scores + BCE-with-logits against the un-permuted labels (:246), the cosine local affinity of
 * column p against ego row p and its margin (:234-240), the reconstruction term (:197-198) and
 * total = cls + margin + 0.1 rec (:258) -- one CTA per batch position + a one-CTA reduction.  out[0..3] = total, cls,
 * margin, rec.  ggad_minibatch_tail_bwd is the hand-derived backward for d total (grad_total: device float): it fills
 * d_combined [B,ld], d_apre [B,h] (gradient w.r.t. fc ego before the ReLU; d fc = d_apre^T ego is one
 * ggad_dense_matmul), d_ego [B,h] (to be pushed through the transposed ego-mean operator), d_scores [B], d_weight [h].
 * All buffers are caller-owned device memory; the forward outputs (rows, apre_src, apre_own, scores, bce, cos, dist, norms, src, out) are the saved
 * state the backward reads.  Deterministic (single writers or two commutative atomic adds per element). */
typedef struct ggad_tail_desc {
  const float* combined; int64_t ld_combined;
  const float* ego; int64_t ld_ego;
  const float* fc;        /* [h,h] row-major (nn.Linear weight) */
  const float* weight;    /* [h] */
  const int64_t* labels;  /* [B] */
  int32_t batch, h;
  float* rows;            /* [B,h]  R = combined_all^T */
  float* apre_src;        /* [B,h]  fc ego[src[p]] for positions whose source row has label 1 */
  float* apre_own;        /* [B,h]  fc ego[p] for rows with label 1 */
  float* scores;          /* [B] */
  float* bce;             /* [B] */
  float* cos;             /* [B] */
  float* dist;            /* [B] */
  float* norms;           /* [2B] */
  int32_t* src;           /* [B] source row of position p */
  float* out;             /* [8] */
  const float* grad_total; /* backward only: d loss / d total (device scalar) */
  float* d_combined; int64_t ld_d_combined;
  float* d_apre;          /* [B,h] */
  float* d_ego;           /* [B,h] */
  float* d_scores;        /* [B] */
  float* d_weight;        /* [h] */
} ggad_tail_desc_t;
GGAD_API int ggad_minibatch_tail_fwd(const ggad_tail_desc_t* desc, ggad_stream_t stream);
GGAD_API int ggad_minibatch_tail_bwd(const ggad_tail_desc_t* desc, ggad_stream_t stream);

/* ---- K4: backward helper of the local-affinity cosine ---------------------------
 * de_k = ( g_k - e^_k <e^_k, g_k> ) * inv_norm_k   with e^_k = e_k * inv_norm_k, in place on g.
 * (autograd of run.py:177-180.) */
GGAD_API int ggad_normalize_backward(const float* e, int64_t lde, const float* inv_norm, float* g, int64_t ldg,
                            int64_t n_rows, int32_t d, ggad_stream_t stream);

/* Row L2 statistics: inv_norm[r] = 1/|x_r| with 1/0 -> 0 (run.py:177-179); sumsq optional. */
GGAD_API int ggad_row_inv_norm(const float* x, int64_t ldx, int64_t n_rows, int32_t d, float* inv_norm, float* sumsq,
                      ggad_stream_t stream);

/* ---- K7: index / degree work (integer results are bit-exact vs the CPU path) ----- */
/* keys = (row << 32 | col), sorted in place; fills rowptr[n_rows+1] and col[n]. Synchronous w.r.t. stream
 * ordering (uses stream-ordered temporaries). Duplicate edges are kept. */
GGAD_API int ggad_coo_keys_to_csr(uint64_t* keys, int64_t n, int64_t n_rows, int64_t* rowptr, int32_t* col, ggad_stream_t stream);
/* CSR -> CSR of the transpose, stable in source-row order (replaces scipy .transpose().tocsr()).
 * val/valT may be NULL; perm (NULL or [nnz]) receives the source edge index of each transposed edge. */
GGAD_API int ggad_csr_transpose(const int64_t* rowptr, const int32_t* col, const float* val, int64_t n_rows, int64_t n_cols,
                       int64_t nnz, int64_t* rowptrT, int32_t* colT, float* valT, int64_t* perm, ggad_stream_t stream);
/* Sub-CSR of the listed rows (adj[0,S,:] of model.py:151 without densifying):
 * sub_rowptr[n_sel+1] must already hold the exclusive prefix sum of the selected degrees. */
GGAD_API int ggad_csr_extract_rows(const int64_t* rowptr, const int32_t* col, const float* val, const int32_t* rows, int64_t n_sel,
                          const int64_t* sub_rowptr, int32_t* sub_col, float* sub_val, ggad_stream_t stream);
/* integer in-degree histogram of col[] (int32 counts, exact) */
GGAD_API int ggad_col_histogram(const int32_t* col, int64_t nnz, int32_t* counts, int64_t n_cols, ggad_stream_t stream);

/* ---- K7: adjacency preprocessing of program A on the device (utils.py:47-54, run.py:98-101) ----
 * deg[r] = sum of row r in fp64 (val NULL = all ones): the `rowsum` of normalize_adj. */
GGAD_API int ggad_csr_row_sum_f64(const int64_t* rowptr, const float* val, int64_t n_rows, double* deg, ggad_stream_t stream);
/* "+ I" on a square CSR with sorted columns, in two steps (the result has one more entry per row without a
 * diagonal): out_rowptr[n+1] first (*nnz_host = new nnz; synchronises), then the entries
 *     out_val = fp32( (fp64(val) * scale[r]) * scale[c]  + [r == c] )      (scale NULL: fp32(fp64(val) + [r == c]))
 * i.e. on CSR(A^T) with scale = deg^-1/2 this is adj = normalize_adj(A) + I, on CSR(A) without scale it is
 * raw_adj = A + I -- same operation order and the same single rounding as scipy + the fp32 cast of run.py:106-109,
 * so the values are bit-identical to the reference's dense tensors. */
GGAD_API int ggad_csr_add_identity_rowptr(const int64_t* rowptr, const int32_t* col, int64_t n, int64_t* out_rowptr,
                                          int64_t* nnz_host, ggad_stream_t stream);
GGAD_API int ggad_csr_scale_add_identity(const int64_t* rowptr, const int32_t* col, const float* val, const double* scale,
                                         int64_t n, const int64_t* out_rowptr, int32_t* out_col, float* out_val,
                                         ggad_stream_t stream);

/* ---- K7: mini-batch frontier on the device --------------------------------------------------
 * Replaces the Python set unions of GCNAggregator / MeanAggregator (src/graphsage.py:305-311,335-341,
 * 82-88).  adj_rowptr/adj_col is the device CSR of the adjacency lists (neighbor ids sorted, as built
 * from the un-pickled dict of src/utils.py:27-28).  One aggregation hop is
 *   ggad_block_rowptr  -> exact integer row degrees |N(v)| (+ the node itself if add_self and absent),
 *                         exclusive scan into block_rowptr[n_batch+1]; *nnz_host = total (synchronises)
 *   ggad_block_fill    -> block columns as GLOBAL ids
 *   ggad_unique_sorted -> the frontier: sorted unique ids; *n_unique_host = |U| (synchronises)
 *   ggad_block_remap   -> block columns as positions in the frontier + exact batch-local column degrees
 * All outputs are integers and bit-exact with the CPU restatement. */
GGAD_API int ggad_block_rowptr(const int64_t* adj_rowptr, const int32_t* adj_col, int64_t n_nodes, const int32_t* nodes,
                               int64_t n_batch, int32_t add_self, int64_t* block_rowptr, int64_t* nnz_host,
                               ggad_stream_t stream);
GGAD_API int ggad_block_fill(const int64_t* adj_rowptr, const int32_t* adj_col, int64_t n_nodes, const int32_t* nodes,
                             int64_t n_batch, int32_t add_self, const int64_t* block_rowptr, int32_t* block_col,
                             ggad_stream_t stream);
GGAD_API int ggad_unique_sorted(const int32_t* keys, int64_t n, int64_t key_bound, int32_t* uniq /*[n]*/,
                                int64_t* n_unique_host, ggad_stream_t stream);
GGAD_API int ggad_block_remap(const int32_t* cols, int64_t nnz, const int32_t* uniq, int64_t n_unique,
                              int32_t* local /*[nnz]*/, int32_t* cdeg /*[n_unique]*/, ggad_stream_t stream);

/* Hop block that is only GATHERED from the feature table (the hop-2 block of GCNAggregator, src/graphsage.py:335-355):
 * neither the frontier list nor local column ids are needed -- only the batch-local column degree
 * cdeg(u) = number of block rows containing u (src/graphsage.py:347), as the per-edge value 1/sqrt(cdeg(col[e])).
 * counts[n_nodes] is int32 scratch (zeroed here, exact integer histogram); block_col holds GLOBAL ids from
 * ggad_block_fill, so the gather runs straight on the table with per-edge values: no sort, no unique, no remap,
 * no size read-back. */
GGAD_API int ggad_block_col_weights(const int32_t* block_col, int64_t nnz, int32_t* counts, int64_t n_nodes, float* val,
                                    ggad_stream_t stream);

/* ---- synthetic graphs (C5 / S64 generator) -------------------------
 * R-MAT (a,b,c,d) edges for destination shard `shard` of `n_shards` (power of two), each shard
 * owning n_local nodes; emits keys (dst_global << 32 | src_global), dst in the shard's range,
 * src anywhere.  If filter_lo < filter_hi only edges with src in [filter_lo, filter_hi) are kept
 * as (src << 32 | dst) keys (used to build the transposed shard) and *n_out_host receives the count
 * (this variant synchronises the stream). */
GGAD_API int ggad_rmat_keys(uint64_t* keys, int64_t n_edges, int64_t n_local, int32_t n_shards, int32_t shard, uint64_t seed,
                   float a, float b, float c, int64_t filter_lo, int64_t filter_hi, int64_t* n_out_host,
                   ggad_stream_t stream);

/* ---- host-buffer entry point (end-to-end path; copies inside) --------------------
 * One fwd + bwd pass of a plain SpMM layer with HOST feature buffers:
 *   y = A x ;  dx = A^T y       (loss = |y|^2 / 2, so dy = y)
 * x_host is copied in, dx_host (and y_host if non-NULL) copied out, *loss_host = |y|^2/2.
 * dev_x [n_cols*d], dev_y [n_rows*d], dev_dx [n_cols*d] and dev_ws are caller-owned device scratch;
 * dev_ws holds 2*max(a.n_tiles, at.n_tiles)*d + roundup4(n_rows) + 4 floats.
 * The CSRs (A and A^T with their plans) stay resident on the device like model state.
 * Synchronous: returns after the results are in host memory. */
typedef struct ggad_resident_csr {
  const int64_t* rowptr;
  const int32_t* col;
  const float* val;
  const float* row_scale;
  const float* col_scale;
  int64_t n_rows, n_cols, nnz;
  const int32_t* tile_row;
  const int64_t* tile_edge;
  int64_t n_tiles;
} ggad_resident_csr_t;

GGAD_API int ggad_spmm_fwd_bwd_host(const ggad_resident_csr_t* a, const ggad_resident_csr_t* at, const float* x_host,
                           float* y_host, float* dx_host, double* loss_host, int32_t d, float* dev_x, float* dev_y,
                           float* dev_dx, float* dev_ws, ggad_stream_t stream);
/* Same work, enqueued only (no synchronisation): all host buffers incl. loss_host must be pinned.  Calls
 * on different streams with their own dev_* scratch overlap one step's H2D with the previous step's D2H
 * (PCIe is full duplex), which is how a training loop would pipeline host-fed batches. */
GGAD_API int ggad_spmm_fwd_bwd_host_enqueue(const ggad_resident_csr_t* a, const ggad_resident_csr_t* at,
                                            const float* x_host, float* y_host, float* dx_host, double* loss_host,
                                            int32_t d, float* dev_x, float* dev_y, float* dev_dx, float* dev_ws,
                                            ggad_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GGAD_B200_H_ */
