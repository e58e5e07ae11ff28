/*
 * qrs_b200.h — C ABI of libqrs_b200.so: the B200 (sm_100a) implementation of the
 * LSH → top-k collaborative-filtering hot path of
 * wamuumu/query-recommendation-system.
 *
 * The reference has no FFI: its "operator API" for this path is two Python
 * classes (lsh.py:7-54 `LSH`, recommender.py:49-343 `Recommender`).  Each entry
 * point below replaces one step of those classes and cites it; the Python
 * mirror of the classes lives in query_recommendation_system_b200/{lsh,recommender}.py
 * and calls these symbols through ctypes (INTEGRATION.md shows the binding a
 * reference maintainer would add).
 *
 * Conventions
 *   - plain C: pointers, sizes, a CUDA stream passed as void* (cudaStream_t).
 *   - every buffer is CALLER-OWNED DEVICE memory; nothing is allocated, freed or
 *     retained by the library, and there is no global mutable state (fixes the
 *     reference's class-level LSH.buckets, lsh.py:9).  Calls on different
 *     streams/devices may run concurrently.
 *   - return value: QRS_OK (0) or a negative qrs_status; no exceptions cross
 *     the ABI.  Kernel launches are asynchronous on `stream`; a returned QRS_OK
 *     means "enqueued".  Counters the host must read are written to device
 *     memory the caller copies back itself.
 *   - workspaces: every function that needs scratch has a *_workspace_bytes()
 *     twin; `ws` must be 256-byte aligned and at least that large.
 *   - "sets × elements" CSR: rowptr int64[n+1], colidx int32[nnz] ascending and
 *     distinct within a row.  Binding A: set = query, element = dataset tuple
 *     (recommender.py:68-103).  Binding B: set = user, element = rated query.
 */
#ifndef QRS_B200_H
#define QRS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qrs_status {
  QRS_OK = 0,
  QRS_ERR_INVALID_ARG = -1,   /* null pointer, negative size, K % b != 0 (lsh.py:20) ... */
  QRS_ERR_UNSUPPORTED = -2,   /* shape outside what the kernels were built for */
  QRS_ERR_WORKSPACE = -3,     /* workspace too small */
  QRS_ERR_CUDA = -4           /* a CUDA runtime call failed; see qrs_last_cuda_error() */
} qrs_status;

#define QRS_MAX_TOPK 64          /* neighbour-table width limit (reference k = round(log_1.5 n) <= 45) */
#define QRS_MAX_BANDS 256        /* the band index travels in 8 bits of a bucket record */
#define QRS_HASH_BITS 24
#define QRS_SENTINEL_HASH 0xFFFFFFu /* bucket record of an all-(-1) band: never grouped */

int qrs_version(void);                       /* 100 * major + minor */
const char *qrs_status_string(int status);
const char *qrs_last_cuda_error(void);       /* thread-local text of the last CUDA failure */
int qrs_device_sm_count(int device);         /* -1 on error */
/* Diagnostics: number of kernels this library has launched in the process so far
 * (monotonic; the only process-wide state, never read by the kernels). */
unsigned long long qrs_launch_count(void);

/* Bucket record (u64), the unit that is sorted, exchanged between GPUs and
 * turned into candidate pairs:
 *     bits 63..40  hash24 of the band's r int16 lanes (QRS_SENTINEL_HASH when the
 *                  lanes are all -1, lsh.py:44-47)
 *     bits 39..32  band index
 *     bits 31..0   global set id
 * Equal (hash24, band) is only a HINT that two sets share a bucket;
 * qrs_pairs_score re-checks the exact lanes, so hash collisions cost work,
 * never correctness. */

/* ---------------------------------------------------------------- a1: MinHash
 * Replaces Recommender.compute_signatures (recommender.py:105-143):
 *   sig[s][i] = min_{e in set(s)} perm_i[e],  -1 for an empty set,
 * where perm_i is the i-th np.random.permutation(m) (recommender.py:120),
 * supplied by the host as the element-major table perm_tab[e*K + i].
 */

/* Setup: the first R entries of the reference's argsort (recommender.py:121) for
 * every hash, R a multiple of 8, R <= m.  The element e with perm_tab[e*K + i] == r
 * (r < R) is stored chunk-major, eight ranks of one hash in one 32-byte sector:
 *     order_head[((r / 8) * Kp + i) * 8 + r % 8],   Kp = K rounded up to 32
 * (a lane that owns hash i reads its next 8 ranks with two 16-byte loads, a warp reads
 * one contiguous kilobyte).  order_head holds qrs_perm_head_elems(K, R) = R * Kp
 * int32, 16-byte aligned; the padding hashes are zero. */
int64_t qrs_perm_head_elems(int32_t K, int32_t R);
int qrs_perm_head_invert(const int32_t *perm_tab, int32_t m, int32_t K, int32_t R,
                         int32_t *order_head, void *stream);

#define QRS_MINHASH_AUTO 0         /* tile scan when a word map fits shared memory (see TILE);
                                      otherwise per set: first-touch when len*len >= m, else
                                      direct.  Sets with len*len < m always take direct.     */
#define QRS_MINHASH_DIRECT 1       /* per-nonzero gather-min from perm_tab                */
#define QRS_MINHASH_FIRST_TOUCH 2  /* membership-map scan in permuted order (the reference's
                                      own first-touch walk, recommender.py:125-129), one set
                                      per CTA                                             */
#define QRS_MINHASH_TILE 3         /* the same walk for 32 / 16 / 8 consecutive sets at once:
                                      one 32- / 16- / 8-bit membership word per element in
                                      shared memory (32 sets up to m = 32768, 16 sets up to
                                      m ~ 100k, 8 sets up to m ~ 200k), lane = hash, ranks kept
                                      bit-sliced                                          */

/* Signatures + fused epilogue.  Optional outputs (pass NULL to skip):
 *   norm2[s]   = sum_i sig[s][i]^2                       (for a8)
 *   records[band*rec_stride + s] = bucket record of (set id_base + s, band);
 *                requires 0 < b <= QRS_MAX_BANDS and K % b == 0.
 * order_head/R may be NULL/0, which forces the DIRECT kernel; R must be a
 * multiple of 32.  First-touch needs m <= qrs_minhash_max_elements(), the tile
 * scan a word map plus a (sets x K) signature tile in shared memory and R <= 32768
 * (QRS_ERR_UNSUPPORTED when a forced mode does not fit).  colidx needs 4-byte
 * alignment only and is never read beyond rowptr[n] entries. */
int64_t qrs_minhash_max_elements(void);
int qrs_minhash_perm(const int64_t *rowptr, const int32_t *colidx, int64_t n, int32_t m,
                     const int32_t *perm_tab, const int32_t *order_head, int32_t R,
                     int32_t K, int32_t b, int64_t id_base,
                     int32_t *sig, int64_t *norm2, uint64_t *records, int64_t rec_stride,
                     int32_t mode, void *stream);

/* ------------------------------------------------------ a2/a3: banding
 * Replaces LSH.make_subvecs + the key construction of LSH.compute_buckets
 * (lsh.py:17-38).  keys[s*b + band] = the r int16 lanes packed big-endian
 * (lane 0 most significant) — the exact bucket id for r <= 4; for r > 4 (r <= 8)
 * pass keys_hi as well (lanes 0..r-5 in keys_hi, last 4 lanes in keys).
 * skipped[s*b + band] = 1 iff every lane is -1 (lsh.py:47). */
int qrs_band_keys(const int32_t *sig, int64_t n, int32_t K, int32_t b,
                  uint64_t *keys, uint64_t *keys_hi, uint8_t *skipped, void *stream);

/* Bucket records (and optionally norm2) from an existing signature matrix; same
 * layout as qrs_minhash_perm's. */
int qrs_band_records(const int32_t *sig, int64_t n, int32_t K, int32_t b, int64_t id_base,
                     uint64_t *records, int64_t rec_stride, int64_t *norm2, void *stream);

/* ------------------------------------------------------ a3/a4: grouping
 * Replaces the per-band dict insertion (lsh.py:35-37): a stable LSD radix sort
 * of u64 keys on bits [bit_lo, bit_hi), independently for `segs` consecutive
 * segments of `n_per_seg` keys.  On return the sorted keys are in `keys`.
 * `tmp` has the size of `keys`. */
size_t qrs_sort_workspace_bytes(int64_t n_per_seg, int32_t segs);
int qrs_sort_u64(uint64_t *keys, uint64_t *tmp, int64_t n_per_seg, int32_t segs,
                 int32_t bit_lo, int32_t bit_hi, void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------ a5: candidate pairs
 * Replaces LSH.get_candidates (lsh.py:40-54).
 * Input: n_records records sorted on bits [32, 64) (any number of bands, one
 * flat array).  slot_off u64[n_records + 1] receives the exclusive prefix sum of
 * "number of earlier records in the same (hash24, band) run"; slot_off[n_records]
 * is the number of pair slots = sum over runs of h*(h-1)/2, an upper bound on
 * (and normally close to) the number of (pair, band) coincidences. */
size_t qrs_pair_slots_workspace_bytes(int64_t n_records);
int qrs_pair_slots(const uint64_t *records, int64_t n_records, uint64_t *slot_off,
                   void *ws, size_t ws_bytes, void *stream);

/* ------------------------------------------------------ a5 + a8: emit, dedup, score
 * Replaces the global-set dedup of get_candidates (lsh.py:51-53) and the rounded
 * cosine (recommender.py:203-204).  For every two members (i, j) of a run in
 * band B, the pair is verified against the signatures themselves (exact int16
 * lane equality, not-all-(-1)) and emitted iff B is the FIRST band in which the
 * two signatures share a bucket — so every candidate pair is emitted exactly
 * once and hash collisions cannot create false candidates.
 *   pairs[t] = i << 32 | j  (i < j, global ids as stored in the records)
 *   sim_milli[t] = rint(1000 * cos(sig_i, sig_j))   (np.around(.,3) * 1000)
 * sig/norm2 are indexed by (id - sig_id_base).  *n_pairs (device u64, zeroed by
 * the caller) counts emitted pairs; pairs beyond `capacity` are counted but not
 * stored.  Output order is unspecified (set semantics).  sim_milli may be NULL
 * (candidates only). */
int qrs_pairs_score(const uint64_t *records, int64_t n_records, const uint64_t *slot_off,
                    const int32_t *sig, const int64_t *norm2, int64_t sig_id_base,
                    int32_t K, int32_t b,
                    uint64_t *pairs, int32_t *sim_milli, int64_t capacity,
                    unsigned long long *n_pairs, void *stream);

/* Score an explicit pair list (ids index sig directly). */
int qrs_pair_cosine(const int32_t *sig, const int64_t *norm2, int32_t K,
                    const uint64_t *pairs, int64_t n_pairs, int32_t *sim_milli, void *stream);

/* ------------------------------------------------------ a7 + a9: neighbour tables
 * Replaces the adjacency build and top-k (recommender.py:183-210) with the
 * canonical order (similarity descending, neighbour id ascending).
 * Rows [row_base, row_base + n_rows) are produced; a pair end outside that
 * range contributes only to the end inside it.
 *   idx int32[n_rows*k] (-1 padded), val double[n_rows*k] (0 padded), cnt int32[n_rows]
 * adj: u64[2*n_pairs] scratch for the adjacency lists. */
size_t qrs_topk_workspace_bytes(int64_t n_rows);
int qrs_topk_tables(const uint64_t *pairs, const int32_t *sim_milli, int64_t n_pairs,
                    int64_t row_base, int64_t n_rows, int32_t k,
                    uint64_t *adj, void *ws, size_t ws_bytes,
                    int32_t *idx, double *val, int32_t *milli /* optional: int32[n_rows*k], val * 1000 */,
                    int32_t *cnt, void *stream);

/* ------------------------------------------------------ a10 + a11: prediction
 * Replaces weighted_average (recommender.py:36-47) and the per-cell loop +
 * hybrid rule of Recommender.compute_scores (recommender.py:311-331).
 * Ratings: CSR users × queries, rowptr int64[U+1], colidx int32 ascending,
 * val uint8 (0 = unrated; stored zeros are allowed and ignored).  Tables as
 * produced by qrs_topk_tables (or by the host for the reference's PCA+BIRCH
 * user branch, recommender.py:274-286); pass idx == NULL for an absent
 * branch.  fp64, left-to-right in table order, no FMA contraction,
 * round-half-even — bit-identical to the reference's arithmetic.
 * User-table ids are GLOBAL user ids (rows of the CSR); the users predicted
 * are [user_base, user_base + n_users), table row t ↔ user user_base + t.
 *
 * Dense form (drop-in for compute_scores at reference scale):
 *   out int32[n_users*Q]: rated cells keep the rating, unrated cells get the
 *   hybrid prediction (0 = "missed", recommender.py:341). */
int qrs_predict_dense(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val,
                      int64_t U, int32_t Q,
                      const int32_t *q_idx, const double *q_val, const int32_t *q_cnt, int32_t kq,
                      const int32_t *u_idx, const double *u_val, const int32_t *u_cnt, int32_t ku,
                      int64_t user_base, int64_t n_users, int32_t *out, void *stream);

/* Sparse form for the scaled configurations (user branch only, query branch
 * absent ⇒ q = 0): per user the non-zero predictions over unrated queries.
 * The query axis is cut into nt = ceil(Q / tile) windows of `tile` columns
 * (a multiple of 32, <= qrs_predict_max_tile()); one WARP accumulates one
 * (user, window) in its slice of shared memory, so smaller windows put more
 * independent warps on an SM (qrs_predict_warps_per_sm) and larger ones fill
 * the 32-entry steps better.  For user t and window c,
 *   row_start int64[n_users*nt], row_cnt int32[n_users*nt] (index t*nt + c)
 * locate a contiguous slice of (cols int32[], preds uint8[]), ascending in
 * col.  Slices are placed by atomic cursors, so their order in the buffer is
 * unspecified: the buffer of `capacity` entries is cut into n_regions equal
 * regions, each with its own cursor (one cursor would serialise millions of
 * same-address atomics); cursors u64[n_regions + 1] (device, zeroed by the caller)
 * end as the number of entries REQUESTED in each region, writes beyond a
 * region's capacity/n_regions entries are dropped (the count still advances),
 * and row_start is an absolute index into cols/preds.  n_regions = 1 gives one
 * dense buffer.  cols/preds may be NULL with capacity 0 for a counting pass. */
/* Every column a neighbour contributed to (and the user has not rated) gets ONE prediction byte in the
 * slice, ascending in column; with u_milli the byte is never 0, with u_val it is 0 when the similarity
 * sums cancel (weightSum == 0 or prediction 0, recommender.py:44).  row_cnt counts those bytes and
 * cursors[n_regions] (one extra u64) receives the number of non-zero ones.  How the columns are reported:
 *   QRS_COLS_I32 / QRS_COLS_U16: cols is int32[capacity] / uint16[capacity] (Q <= 65536), aligned with preds;
 *   QRS_COLS_BITMAP: cols is uint32[n_users * nt * tile / 32], zeroed by the caller: bit c of word
 *       (t*nt + window) * tile/32 + c/32 marks column window*tile + c as having a byte in preds. */
#define QRS_COLS_I32 0
#define QRS_COLS_U16 1
#define QRS_COLS_BITMAP 2
int32_t qrs_predict_max_tile(int32_t integer_tables /* 1 when u_milli is used */);
/* warps (= (user, window) items in flight) per SM for a window width; 0 if the width is invalid or does not fit */
int32_t qrs_predict_warps_per_sm(int32_t tile, int32_t integer_tables);
size_t qrs_predict_sparse_workspace_bytes(int64_t U, int64_t nnz, int32_t Q, int32_t tile, int64_t capacity);
/* Row preparation, once per ratings CSR and window width: cuts every row at the window bounds and writes the
 * (row, window) sub-segments as 32-bit words (column - window start) << 8 | rating, zero-padded to whole
 * 128-byte steps, into the workspace (at most 4 * (nnz + 32 * U * nt) bytes of it).  The predictor reads only
 * this copy in its loop, so colidx / val need no particular alignment or padding.  The order of the words inside
 * a sub-segment is the library's business (column order, or — where windows hold some 24 entries or more on
 * average — one cell per shared-memory bank class so that the predictor's accumulator updates do not collide;
 * the workspace records which, environment QRS_PREDICT_PACK=class|column forces one).  qrs_predict_sparse calls it
 * itself unless `prepared` != 0 says the workspace already holds the preparation of the same CSR, Q and tile
 * (several user ranges predicted from one preparation). */
int qrs_predict_prepare(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val,
                        int64_t U, int64_t nnz, int32_t Q, int32_t tile, int64_t capacity,
                        void *ws, size_t ws_bytes, void *stream);
/* Rows sharded over several GPUs (one process per GPU): rowptr/colidx/val and U describe THIS rank's rows
 * [my_peer * U, (my_peer + 1) * U) of the matrix (every rank holds U rows), neighbour ids in u_idx are GLOBAL row
 * ids, and peer_ws[r] is rank r's prepared workspace (same U, nnz bound, Q, tile) mapped into this process
 * (qrs_peer_export / qrs_peer_open; peer_ws[my_peer] is ignored in favour of ws).  Neighbour rows that live on
 * other ranks are then read in place over NVLink: every rank must have finished qrs_predict_prepare before any
 * rank predicts, and must not prepare again while a peer still predicts (prepared must be non-zero with peers).
 * peer_ws == NULL: all rows are local (ids in u_idx are rows of this CSR).
 * u_milli (optional): the same similarities as integers, u_val[t] == u_milli[t] / 1000.0 exactly
 * (what np.around(., 3) produces, recommender.py:204, 270; qrs_topk_tables emits it).  When given,
 * the kernel accumulates exact integers (half the shared memory, no fp64 in the inner loop) and
 * falls back to the reference's sequential fp64 arithmetic only for cells whose value is exactly
 * half-way between two integers — the results are identical to the u_val path bit for bit. */
/* Staged peer rows (optional, several GPUs).  Reading a neighbour row over NVLink one 128-byte step at a time
 * costs the predicting warp a link round trip per batch; qrs_peer_rows_stage instead copies every row that the
 * table (u_idx / u_cnt of local users 0 .. n_users-1, global ids) names on another rank into `stage` — whole rows,
 * thousands of lines in flight — and qrs_predict_sparse, given the same buffer as peer_stage (stage_users = that
 * n_users), reads them locally.  cap_steps = 128-byte steps the buffer can hold; the device word stage[0] (u64)
 * receives the steps this table needs, and when that exceeds cap_steps nothing is copied and the predictor reads
 * the rows in place as without staging (the caller reads the word back with its other counters and sizes the
 * next call).  Call after every rank has prepared, before predicting; results are identical either way. */
size_t qrs_peer_rows_stage_bytes(int64_t n_users, int32_t ku, int32_t Q, int32_t tile, int64_t cap_steps);
int qrs_peer_rows_stage(const uint64_t *peer_ws /* HOST array [n_peers] */, int32_t n_peers, int32_t my_peer,
                        int64_t U, int32_t Q, int32_t tile, const int32_t *u_idx, const int32_t *u_cnt, int32_t ku,
                        int64_t n_users, int64_t cap_steps, void *stage, size_t stage_bytes, void *stream);
int qrs_predict_sparse(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val,
                       int64_t U, int64_t nnz, int32_t Q,
                       const int32_t *u_idx, const double *u_val, const int32_t *u_milli,
                       const int32_t *u_cnt, int32_t ku,
                       int64_t user_base, int64_t n_users, int32_t tile,
                       void *ws, size_t ws_bytes, int32_t prepared,
                       const uint64_t *peer_ws /* HOST array [n_peers] or NULL */, int32_t n_peers, int32_t my_peer,
                       const void *peer_stage /* or NULL */, int64_t stage_users,
                       int64_t *row_start, int32_t *row_cnt,
                       void *cols, int32_t cols_mode /* QRS_COLS_* */, uint8_t *preds,
                       int64_t capacity, int32_t n_regions, unsigned long long *cursors, void *stream);

/* ------------------------------------------------------ f-1: set construction
 * Replaces the predicate evaluation of Recommender.compute_shingles (recommender.py:68-103): answer sets of
 * conjunctive `attr = value` queries over the dictionary-encoded table, as the sets x elements CSR that
 * qrs_minhash_perm consumes (set = query, element = tuple index, ascending).
 *   table_cm int32[F * drows] column-major codes; queries int32[n_queries * F], -1 = no condition on that
 *   attribute (the reference's "", recommender.py:83), any other negative = a value absent from the column.
 * qrs_answer_sets_count writes rowptr int64[n_queries + 1]; the caller reads rowptr[n_queries], allocates
 * colidx and calls qrs_answer_sets_fill.  F <= 32. */
size_t qrs_answer_sets_workspace_bytes(int64_t n_queries);
int qrs_answer_sets_count(const int32_t *table_cm, int64_t drows, int32_t F, const int32_t *queries,
                          int64_t n_queries, int64_t *rowptr, void *ws, size_t ws_bytes, void *stream);
int qrs_answer_sets_fill(const int32_t *table_cm, int64_t drows, int32_t F, const int32_t *queries,
                         int64_t n_queries, const int64_t *rowptr, int32_t *colidx, void *stream);

/* ------------------------------------------------------ f-2: per-user top-N recommendations
 * Replaces what Recommender.top_k_queries serves (recommender.py:345-374): per user the N largest non-zero
 * predictions among the cells just predicted, from the sparse output of qrs_predict_sparse (QRS_COLS_I32 or
 * QRS_COLS_U16), ordered prediction descending, column ascending.
 *   out_idx int32[n_users*N] (-1 padded), out_val uint8[n_users*N] (0 padded), out_cnt int32[n_users]; N <= 64. */
int qrs_recommend_topn(const int64_t *row_start, const int32_t *row_cnt, const void *cols, int32_t cols_mode,
                       const uint8_t *preds, int64_t n_users, int32_t nt, int32_t N,
                       int32_t *out_idx, uint8_t *out_val, int32_t *out_cnt, void *stream);

/* ------------------------------------------------------ f-3: user-similarity branch at scale
 * Replaces the scoring half of Recommender.compute_userSimilarities (recommender.py:256-269): the
 * cosine of the mean-centred rating rows, np.around(., 3), negatives -> 0.  The reference centres an
 * INTEGER block (recommender.py:263), so a centred value is trunc(r - mean) = (r * n - S) / n in C
 * integer division (n rated cells summing to S); cells with rating 0 stay 0.  The candidate pairs
 * come from the LSH stages over the users x queries rows instead of PCA + BIRCH clusters.
 *
 * qrs_center_rows: cval int16[nnz] (centred value of every CSR cell), norm2 int64[n_rows].
 * qrs_pair_centered_cosine: sim_milli[p] = max(0, rint(1000 * cos)) for pairs[p] = i << 32 | j (row
 * ids of this CSR); a zero-norm row scores 0.  flags: QRS_PAIRS_GROUPED = equal first ids are
 * adjacent (e.g. after qrs_sort_u64 on bits [32, 64)), which lets row i be staged in shared memory
 * (m <= 204800); otherwise, or with QRS_PAIRS_FORCE_SEARCH, the shorter row is searched in the longer. */
#define QRS_PAIRS_GROUPED 1
#define QRS_PAIRS_FORCE_SEARCH 2
int qrs_center_rows(const int64_t *rowptr, const uint8_t *val, int64_t n_rows, int16_t *cval,
                    int64_t *norm2, void *stream);
int qrs_pair_centered_cosine(const int64_t *rowptr, const int32_t *colidx, const uint8_t *val,
                             const int16_t *cval, const int64_t *norm2, const uint64_t *pairs,
                             int64_t n_pairs, int32_t m, int32_t flags, int32_t *sim_milli, void *stream);

/* ------------------------------------------------------ f-4: CSV ingest of the utility matrix
 * Replaces dt.fread of utility_matrix.csv (main.py:69-75) for the path: the file as the reference's
 * generator writes it (generator.py:59-77, 329-332 - csv.writer: "\r\n" or "\n" rows, first line = the
 * query ids WITHOUT a user field, then `user,r,r,,r,...`; empty field = not rated; integers, here 0..255
 * with 0 = not rated) is parsed on the device into the ratings CSR (rowptr, colidx ascending, val).
 * `text` is the whole file in device memory, 16-byte aligned.
 *
 * qrs_csv_line_ends: *n_lines = number of '\n' bytes; with capacity > 0 also line_end[k] = position of the
 *   k-th one (k < capacity).  Call once with capacity 0 to size line_end.  If the text does not end with
 *   '\n' the caller appends n_bytes as the end of the last line.
 * qrs_csv_ratings_count: rows = lines first_line .. first_line + n_rows - 1 (first_line = 1 skips the
 *   header).  rowptr int64[n_rows + 1] = exclusive scan of the rated cells per row; id_end[r] = position of
 *   the comma closing row r's user id; err[0] = number of malformed rows (field count != Q + 1, a
 *   non-digit in a rating field, a quote, a value > 255), err[1] = the first such row (INT32_MAX if none).
 *   Malformed rows count 0 cells; qrs_csv_ratings_fill must only be called when err[0] == 0.
 * ws: qrs_csv_workspace_bytes(n_bytes, n_rows) (either argument may be 0 for the call that does not use it). */
size_t qrs_csv_workspace_bytes(int64_t n_bytes, int64_t n_rows);
int qrs_csv_line_ends(const uint8_t *text, int64_t n_bytes, int64_t *line_end, int64_t capacity,
                      uint64_t *n_lines, void *ws, size_t ws_bytes, void *stream);
int qrs_csv_ratings_count(const uint8_t *text, int64_t n_bytes, const int64_t *line_end, int64_t first_line,
                          int64_t n_rows, int32_t Q, int64_t *rowptr, int64_t *id_end, int32_t *err, void *ws, size_t ws_bytes,
                          void *stream);
int qrs_csv_ratings_fill(const uint8_t *text, int64_t n_bytes, const int64_t *line_end, int64_t first_line,
                         int64_t n_rows, int32_t Q, const int64_t *rowptr, int32_t *colidx, uint8_t *val, void *stream);

/* ------------------------------------------------------ f-4: CSV export
 * Replaces the export of main.py:97-111 (predictions.to_numpy().tolist() -> generator.write_csv,
 * generator.py:59-77: csv.writer rows `user,v,v,...` + "\r\n"; the caller writes the header line of query
 * ids).  A row is the union of the ratings CSR and, optionally, the sparse predictions of
 * qrs_predict_sparse (row_start / row_cnt int[n_rows, nt], cols QRS_COLS_I32 or QRS_COLS_U16, preds) - the two
 * never share a cell (recommender.py:302).  zero_as_empty = 0: final_predictions.csv (a 0 is printed);
 * 1: the utility-matrix form (generator.py:329-332, empty field = not rated), which qrs_csv_ratings_* reads back.
 * ids / ids_off int64[n_rows + 1]: the user ids as concatenated bytes.  Q <= 196608.
 * count: text_off int64[n_rows + 1] = byte offset of every row (text_off[n_rows] = total size).
 * fill: the text itself (`text` 16-byte aligned). */
size_t qrs_csv_format_workspace_bytes(int64_t n_rows);
int qrs_csv_format_count(const int64_t *ids_off, const int64_t *rowptr, const uint8_t *val,
                         const int64_t *pred_start, const int32_t *pred_cnt, const uint8_t *preds, int32_t nt,
                         int64_t n_rows, int32_t Q, int32_t zero_as_empty, int64_t *text_off, void *ws,
                         size_t ws_bytes, void *stream);
int qrs_csv_format_fill(const uint8_t *ids, const int64_t *ids_off, const int64_t *rowptr, const int32_t *colidx,
                        const uint8_t *val, const int64_t *pred_start, const int32_t *pred_cnt,
                        const void *pred_cols, int32_t cols_mode, const uint8_t *preds, int32_t nt, int64_t n_rows,
                        int32_t Q, int32_t zero_as_empty, const int64_t *text_off, uint8_t *text, void *stream);

/* Compact CSR wire format of a utility matrix with at most 65536 queries: 16-bit column ids, widened on the
 * device to the int32 colidx every stage takes (half the host -> device bytes of the largest input). */
int qrs_cols_widen_u16(const uint16_t *cols16, int64_t n, int32_t *colidx, void *stream);

/* ------------------------------------------------------ multi-GPU plumbing
 * Rows of the matrix are sharded across the GPUs of one box, one process per GPU.
 *
 * Peer buffers: memory other ranks map and read in place (CUDA IPC).  qrs_peer_alloc allocates, qrs_peer_export
 * writes the 64-byte handle another process passes to qrs_peer_open to obtain its own mapping of the same memory
 * (peer access is enabled on first use); the predictor reads neighbour rows of other ranks through such
 * mappings (qrs_predict_sparse, peer_ws). */
int qrs_peer_alloc(size_t bytes, void **ptr);
int qrs_peer_free(void *ptr);
int qrs_peer_export(void *ptr, unsigned char *handle64);
int qrs_peer_open(const unsigned char *handle64, void **ptr);
int qrs_peer_close(void *ptr);
/* Fixed-capacity exchanges: both all-to-alls of a step move n_ranks equal slots of `cap` 64-bit words per rank,
 * unused entries are all-ones (a sentinel bucket hash for records, a row id no rank owns for pairs, ignored by
 * qrs_pair_slots / qrs_topk_tables), so no split size ever has to reach the host between the collectives.
 * counts u64[n_ranks] = entries that wanted each slot; *overflow gets bit 0 (records) / bit 1 (pairs) set when a
 * slot was too small (the caller re-runs the step with a larger cap).
 *   qrs_group_pad: records sorted by their top lg bits (destination rank = bucket hash prefix, 2^lg ranks)
 *   -> out[2^lg * cap].
 *   qrs_route_pairs: scored pairs (first min(*n_pairs_dev, capacity_in) entries) -> one word per destination
 *   (i << 37 | j << 11 | sim + 1000; ids < 2^26), to the owner of i and, if different, of j.
 *   qrs_unpack_routed: received words -> pairs / sim_milli (sentinels stay all-ones pairs). */
int qrs_group_pad(const uint64_t *sorted, int64_t n, int32_t lg, int64_t cap, uint64_t *out,
                  unsigned long long *counts, uint32_t *overflow, void *stream);
int qrs_route_pairs(const uint64_t *pairs, const int32_t *sims, const unsigned long long *n_pairs_dev,
                    int64_t capacity_in, int64_t rows_per_rank, int32_t n_ranks, int64_t cap, uint64_t *out,
                    unsigned long long *counts, uint32_t *overflow, void *stream);
int qrs_unpack_routed(const uint64_t *words, int64_t n, uint64_t *pairs, int32_t *sims, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QRS_B200_H */
