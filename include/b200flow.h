/*
 * b200flow.h — C ABI of libb200flow.so: the B200 (sm_100a) hot path of the
 * flow-classification pipeline that biagiom/spark-network-traffic-classifier
 * drives through pyspark.ml.
 *
 * The reference has no FFI of its own: its operator API is the pyspark.ml
 * Estimator/Transformer contract exercised at
 *   code/network_traffic_classifier_kdd99.py:34-37,45-46,64,79,82,86-91
 *   code/network_traffic_classifier_cicids17.py:41-46,68,83,86,90-95
 * Each entry point below names the MLlib operator whose
 * arithmetic it replaces.  The Python shim in
 * spark-network-traffic-classifier_b200/pyspark binds these through ctypes
 * (see INTEGRATION.md for the stub a maintainer would add).
 *
 * Conventions
 *  - every data pointer is a DEVICE pointer owned by the caller unless the
 *    parameter name ends in _host; the library allocates no persistent memory;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*);
 *    nothing synchronises the host;
 *  - return value: 0 = ok, negative = error (b200flow_last_error() has text);
 *  - rows are `int64_t`; everything else that indexes columns/bins/classes is int32;
 *  - RNG: Philox4x32-10, keyed by (seed, purpose) and counted by GLOBAL row /
 *    (tree,node) so results do not depend on how rows are sharded over GPUs
 *    (spec in DESIGN.md §RNG).
 */
#ifndef B200FLOW_H
#define B200FLOW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200FLOW_OK            0
#define B200FLOW_ERR_ARG      -1
#define B200FLOW_ERR_CUDA     -2
#define B200FLOW_ERR_LIMIT    -3

/* element types of dense matrices handed to / produced by the library */
#define B200FLOW_F32 0
#define B200FLOW_F64 1

/* ---- encode plan: one descriptor per OUTPUT slot of the assembled vector ---- */
#define B200FLOW_SRC_F32    0  /* numeric field stored as float   */
#define B200FLOW_SRC_F64    1  /* numeric field stored as double  */
#define B200FLOW_SRC_I32    2  /* numeric field stored as int32   */
#define B200FLOW_SRC_INDEX  3  /* int32 dictionary code -> StringIndexer rank, emitted as a number */
#define B200FLOW_SRC_ONEHOT 4  /* int32 dictionary code -> rank; emits (rank == hot) ? 1 : 0       */

typedef struct b200flow_slot {
    int32_t kind;      /* B200FLOW_SRC_*                                             */
    int32_t src_off;   /* byte offset of the source field inside one raw record      */
    int32_t lut_off;   /* INDEX/ONEHOT: first entry of this column's code->rank LUT  */
    int32_t lut_len;   /* INDEX/ONEHOT: dictionary size; code outside [0,len) or rank<0 = invalid */
    int32_t hot;       /* ONEHOT: the rank this slot lights up for                   */
    int32_t reserved;
    double  mean;      /* StandardScaler withMean: subtracted first (0.0 = off)      */
    double  scale;     /* StandardScaler withStd: 1/sigma, or 0 when sigma==0 (1.0 = off) */
} b200flow_slot;       /* 40 bytes */

const char* b200flow_last_error(void);
int  b200flow_version(void);

/* ------------------------------------------------------------------ encode ---
 * R1  StringIndexer.fit  (kdd99.py:34-37, cicids17.py:45): per-code counts of one
 * int32 dictionary-code field of the raw records.  counts[K] (int64) is ADDED to
 * (caller zeroes; multi-GPU: allreduce the K counts).  Codes outside [0,K) are ignored. */
int b200flow_category_counts(const void* records, int64_t n_rows, int32_t row_bytes,
                             int32_t src_off, int32_t K, int64_t* counts, void* stream);

/* the same for up to 8 code fields in ONE pass over the records (a Pipeline of StringIndexers, kdd99.py:34-37): counts is the
 * concatenation [K_0 | K_1 | ...] (int64, zero-initialised by the caller); src_offs / Ks are HOST arrays of n_cols entries;
 * sum(K) <= 8192. */
int b200flow_category_counts_multi(const void* records, int64_t n_rows, int32_t row_bytes, int32_t n_cols,
                                   const int32_t* src_offs_host, const int32_t* Ks_host, int64_t* counts, void* stream);

/* R2+R3+R3b+R3c  StringIndexerModel.transform + OneHotEncoder + StandardScaler +
 * VectorAssembler.transform fused (kdd99.py:37,46; cicids17.py:42,46): raw AoS
 * records -> dense row-major [n_rows, n_out] matrix (out_dtype F32/F64), computed
 * in fp64: out = (value - mean) * scale.
 *   plan        device array of n_out b200flow_slot
 *   lut         device int32 LUT pool (rank per code, -1 = unseen), may be NULL
 *   label_*     optional label column: int32 code field -> rank (int32) in label_out
 *               (label_off < 0: no label); an unseen label marks the row invalid
 *   valid_out   optional uint8[n_rows]: 0 when the row has an unseen code, or
 *               (check_nan != 0) a NaN numeric field  (handleInvalid="skip"/"error")
 * Full tiles move global->shared->global with TMA bulk copies; records and the
 * output must be 16-byte aligned. */
int b200flow_encode(const void* records, int64_t n_rows, int32_t row_bytes,
                    const b200flow_slot* plan, int32_t n_out,
                    const int32_t* lut, int32_t lut_total,
                    int32_t label_off, int32_t label_lut_off, int32_t label_lut_len,
                    int32_t check_nan,
                    void* out, int32_t out_dtype, int32_t* label_out, uint8_t* valid_out,
                    void* stream);

/* R2+R3 feeding R4/R5 without the dense matrix (SURVEY.md 8d "Encode -> bins"; the vector VectorAssembler builds at
 * kdd99.py:45-46 / cicids17.py:41-46 is consumed by RandomForest.run at kdd99.py:79 / cicids17.py:83 only through
 * findSplits and TreePoint.convertToTreeRDD):
 * sample_records = b200flow_sample_rows on raw records: the Bernoulli row sample of findSplits, every sampled row
 * evaluated through the encode plan (F slots) in fp64; round_f32 != 0 rounds each value to float first (the value an
 * f32 feature matrix would have held).  sample is COLUMN-major [F][cap] fp64 as for sample_rows. */
int b200flow_sample_records(const void* records, int64_t n_rows, int32_t row_bytes,
                            const b200flow_slot* plan, int32_t F, const int32_t* lut, int32_t round_f32,
                            uint64_t seed, uint64_t keep_threshold, int64_t row_offset,
                            double* sample, int64_t cap, int32_t* n_sampled, void* stream);

/* encode_bins = b200flow_encode + b200flow_bin_rows in one pass: raw records -> uint8 TreePoint records
 * tp[n_rows][tp_stride] (bin per plan slot, the label's StringIndexer rank at byte F, zero padding), 16-byte words out.
 * bad (int32[2], caller zeroes): bad[0] += categorical cells outside [0, arity) / non-integral (binned to a value no
 * left-set contains), bad[1] += NaN numeric cells (when check_nan) + unseen dictionary codes.  label_out optional. */
int b200flow_encode_bins(const void* records, int64_t n_rows, int32_t row_bytes,
                         const b200flow_slot* plan, int32_t F, const int32_t* lut, int32_t lut_total,
                         int32_t label_off, int32_t label_lut_off, int32_t label_lut_len,
                         int32_t check_nan, int32_t round_f32,
                         int32_t thr_f32 /* != 0: every continuous slot's value is exactly a float (f32 field with mean 0 / scale 1, or
                                            round_f32): the search then runs on thresholds rounded DOWN to float — the same bins */,
                         const double* thresholds, const int32_t* n_thr, const int32_t* arity, int32_t max_bins,
                         uint8_t* tp, int32_t tp_stride, int32_t* label_out, int32_t* bad, void* stream);

/* R3c  StandardScaler.fit (ml/feature/StandardScaler.scala [MLlib]; north_star encode, not called by kdd99.py/cicids17.py): per-column shifted power sums of a dense [n, D] matrix
 * (leading dimension ld elements): sum[d] += Σ(x-shift[d]), sumsq[d] += Σ(x-shift[d])².
 * shift may be NULL (=0).  Two calls (shift = 0, then shift = mean) give the
 * corrected two-pass variance; multi-GPU: allreduce the 2·D doubles between them. */
int b200flow_column_moments(const void* x, int32_t dtype, int64_t n_rows, int32_t D, int64_t ld,
                            const double* shift, double* sum, double* sumsq, void* stream);

/* --------------------------------------------------------------- tree prep ---
 * R4  RandomForest.findSplits (inside fit: kdd99.py:79, cicids17.py:83), sampling half: Bernoulli(keep_threshold / 2^32) row
 * sample keyed by (seed, global row); gathers the sampled rows of the dense feature
 * matrix into a COLUMN-major fp64 buffer sample[F][cap]; *n_sampled is advanced
 * atomically (caller zeroes; rows beyond cap are counted but not stored). */
int b200flow_sample_rows(const void* x, int32_t dtype, int64_t n_rows, int32_t F, int64_t ld,
                         uint64_t seed, uint64_t keep_threshold, int64_t row_offset,
                         double* sample, int64_t cap, int32_t* n_sampled, void* stream);

/* R4  findSplitsForContinuousFeature: for every continuous feature (arity[f]==0)
 * sort its n_s samples (in place, sample[f*cap .. +n_s)), run-length the distinct
 * values and walk MLlib's stride rule -> thresholds[f*(max_bins-1) ..], n_thr[f].
 * Categorical features (arity>0) get n_thr = 0.  scratch: F * pow2ceil(n_s) doubles. */
int b200flow_find_splits(double* sample, int64_t cap, int32_t n_s, int32_t F,
                         const int32_t* arity, int32_t max_bins,
                         double* thresholds, int32_t* n_thr,
                         const int32_t* n_s_dev /* NULL, or the device-side sample count (then n_s is a host upper bound) */,
                         void* stream);

/* R5  TreePoint.convertToTreeRDD/findBin: dense features (+ int32 labels, may be NULL)
 * -> binned TreePoint records tp[n][tp_stride] (uint8): bytes [0,F) = bin per feature
 * (continuous: lower_bound over thresholds; categorical: (int)x), byte F = label.
 * bad_rows (int32, caller zeroes) counts rows with a categorical value outside
 * [0,arity) or non-integral (MLlib raises). */
int b200flow_bin_rows(const void* x, int32_t dtype, int64_t n_rows, int32_t F, int64_t ld,
                      const double* thresholds, const int32_t* n_thr, const int32_t* arity,
                      int32_t max_bins, const int32_t* labels,
                      uint8_t* tp, int32_t tp_stride, int32_t* bad_rows, void* stream);

/* helper of R5-R7 inside RandomForestClassifier.fit (kdd99.py:79, cicids17.py:83) and of R9 (kdd99.py:82); no MLlib counterpart — exact because the histograms are integer sums.  Row de-duplication (flow records repeat massively: KDD99 has 4.9 M rows but ~1.07 M distinct ones).  Rows whose TreePoint
 * records (first key_bytes bytes: bins + label) are identical are interchangeable for the trees: uid[row] = index of the
 * row's unique record (numbered in order of each group's first row), tp_unique[U][tp_stride] = the unique records,
 * *n_unique = U (device scalar).  Scratch (caller-owned): table/minrow int32[table_cap] (power of two >= 2*n_rows),
 * slot_of/rep/flag int32[n_rows], pos int64[n_rows+1]. */
int b200flow_dedup_rows(const uint8_t* tp, int64_t n_rows, int32_t tp_stride, int32_t key_bytes,
                        int32_t* table, int32_t* minrow, int64_t table_cap, int32_t* slot_of, int32_t* rep,
                        int32_t* flag, int64_t* pos, int64_t* n_unique, int32_t* uid, uint8_t* tp_unique, void* stream);

/* R6  BaggedPoint.convertToBaggedRDD (inside fit: kdd99.py:79, cicids17.py:83): W[tree][uid[row]] += Poisson weight of (tree, global row).  poisson_cdf: 32 increasing
 * uint32 thresholds, weight = #{k: cdf[k] != 2^32-1 && r >= cdf[k]} with r = word tree%4 of Philox(seed,'BAGG', row, tree/4);
 * NULL = no bagging (weight 1 per row, numTrees==1); poisson_cdf_host = the same 32 values in host memory (the first
 * thresholds travel as kernel arguments).  uid NULL = identity.  perm (optional) = the rows grouped by unique id
 * (b200flow_group_rows); uid is then given in that order (uperm): a duplicate group is a run of adjacent lanes and costs one
 * RED per warp and tree.  W uint32[T][n_unique], caller zeroes. */
int b200flow_bag_weights(uint64_t seed, int32_t T, int64_t row_offset, int64_t n_rows,
                         const uint32_t* poisson_cdf, const uint32_t* poisson_cdf_host,
                         const int32_t* uid, const int32_t* perm, int64_t n_unique, uint32_t* W, void* stream);

/* R6 helper (BaggedPoint.convertToBaggedRDD, inside fit: kdd99.py:79, cicids17.py:83): counting sort of the rows by unique id: perm[p] = row, uperm[p] = uid[perm[p]] (non-decreasing).  Scratch: gsize/cursor
 * int32[n_unique], goff int64[n_unique+1]. */
int b200flow_group_rows(const int32_t* uid, int64_t n_rows, int64_t n_unique, int32_t* gsize, int64_t* goff,
                        int32_t* cursor, int32_t* perm, int32_t* uperm, void* stream);

/* R6 (BaggedPoint.convertToBaggedRDD, inside fit: kdd99.py:79, cicids17.py:83): entries of every tree = its non-zero (unique record, summed weight) pairs in unique-id order.  Pass 1: non-zeros per
 * (tree, block of 1024 uniques) -> blk_cnt[T][n_blocks]. */
int b200flow_bag_count(const uint32_t* W, int32_t T, int64_t n_unique, int32_t* blk_cnt, void* stream);

/* R6 (inside fit: kdd99.py:79), pass 2: given blk_off = exclusive scan of blk_cnt (int64, tree-major) write the entries; one entry = 8 bytes
 * {uint32 unique record index, uint32 weight}. */
int b200flow_bag_fill(const uint32_t* W, int32_t T, int64_t n_unique, const int64_t* blk_off, void* ent, void* stream);

/* (no MLlib counterpart; inside fit: kdd99.py:79, cicids17.py:83) exclusive prefix sum utilities used by the trainer (single launch, any n) */
int b200flow_exclusive_scan_i32_to_i64(const int32_t* in, int64_t n, int64_t* out,
                                       int64_t* total, void* stream);

/* ------------------------------------------------------------ tree growing ---
 * One LEVEL of every tree is processed at once.  An active node is a "slot":
 *   slot_tree[s], slot_nid[s] (MLlib node id: root 1, children 2i/2i+1),
 *   slot_node[s]  index of the node in the forest node pool,
 *   seg_begin/seg_end[s]  its bagged entries inside ent (8-byte {record index, weight} pairs).  */

/* per-node feature subsets (RandomForest.selectNodesToSplit [MLlib]; inside fit: kdd99.py:79, cicids17.py:83): m of F features by a
 * partial Fisher-Yates keyed by (seed, tree, nid), sorted ascending -> subset[s*m..].
 * m == F gives the identity. */
int b200flow_feature_subsets(uint64_t seed, int32_t n_slots, const int32_t* slot_tree,
                             const uint32_t* slot_nid, int32_t F, int32_t m,
                             uint16_t* subset, void* stream);

/* R7  findBestSplits/binSeqOp (ml/tree/impl/RandomForest.scala [MLlib]; inside fit: kdd99.py:79, cicids17.py:83) — HOT LOOP A.  hist[s][j][bin][class] += w for every entry
 * of slot s and every j < m (feature subset[s*m+j]).  hist (uint32) must be zeroed by
 * the caller; layout stride = m * n_bins * C.  chunk_off = exclusive scan over slots of
 * ceil(len/chunk_rows) (int64[n_slots+1]); the grid is one CTA per chunk. */
int b200flow_hist_level(const uint8_t* tp, int32_t tp_stride, int32_t F,
                        const void* ent,
                        int32_t n_slots, const int64_t* seg_begin, const int64_t* seg_end,
                        const int64_t* chunk_off, int64_t n_chunks, int32_t chunk_rows,
                        const uint16_t* subset, int32_t m, int32_t n_bins, int32_t C,
                        uint32_t* hist, void* stream);

/* split record written by score_level, one per slot */
typedef struct b200flow_split {
    int32_t  feat;        /* feature index, -1 = no valid split (leaf)                   */
    int32_t  kind;        /* 0 continuous (left iff bin <= bin_thr), 1 categorical (mask) */
    int32_t  bin_thr;     /* continuous: split index s (threshold = thresholds[f][s])     */
    int32_t  flags;       /* bit0 node is leaf, bit1 left child leaf, bit2 right child leaf */
    double   gain;
    double   impurity;    /* of this node                                                 */
    uint64_t mask[4];     /* categorical: bit c set = category c goes left                */
} b200flow_split;         /* 64 bytes */

/* R8  binsToBestSplit / calculateImpurityStats / Gini — HOT LOOP B.  Reads the (all-reduced)
 * histograms, writes split[s], the node's own class counts and both children's class counts
 * (uint32 [n_slots][C] each).  feat_bins[f] = number of bins of feature f; feat_kind[f]:
 * 0 continuous, 1 ordered categorical, 2 unordered categorical (bins = categories).
 * level/max_depth/min_instances/min_info_gain as in MLlib's Strategy. */
int b200flow_score_level(const uint32_t* hist, int32_t n_slots, const uint16_t* subset,
                         int32_t m, int32_t n_bins, int32_t C,
                         const int32_t* feat_bins, const int32_t* feat_kind,
                         int32_t level, int32_t max_depth, int32_t min_instances,
                         double min_info_gain,
                         b200flow_split* split, uint32_t* node_counts,
                         uint32_t* left_counts, uint32_t* right_counts, void* stream);

/* forest node pool (SoA, all trees in one pool; roots are nodes 0..T-1) */
typedef struct b200flow_node {
    int32_t feat;      /* -1 = leaf                                      */
    int32_t kind_bin;  /* kind<<16 | bin_thr                             */
    int32_t left;      /* pool index of left child; right = left + 1     */
    uint32_t nid;      /* MLlib node id                                  */
} b200flow_node;       /* 16 bytes */

/* R8 driver side (LearningNode growth in RandomForest.findBestSplits, ml/tree/impl/RandomForest.scala [MLlib]; inside fit:
 * kdd99.py:79, cicids17.py:83).  Grows the pool by one level: for each slot writes its node record (+ mask, counts), creates
 * two children per split (counts from left/right_counts), and emits the next level's slots for
 * the non-leaf children (next_* arrays, capacity 2*n_slots; next_parent = parent slot*2+side;
 * child_slot[2*s+side] = index of that child among the next slots, -1 when it is a leaf; may be NULL).
 * counters: int64[8] header {node pool size (in/out), number of next slots (out), overflow flag (out: 1 =
 * pool_capacity too small, nothing written), pool size before the call, 4 entries free for the caller}
 * followed by 2*ceil(n_slots/256) int32 of scratch. */
int b200flow_grow_level(int32_t n_slots, const int32_t* slot_tree, const uint32_t* slot_nid,
                        const int32_t* slot_node, const b200flow_split* split,
                        const uint32_t* node_counts, const uint32_t* left_counts,
                        const uint32_t* right_counts, int32_t C,
                        b200flow_node* nodes, uint64_t* node_mask, uint32_t* pool_counts,
                        int32_t* node_tree, int64_t pool_capacity,
                        int32_t* next_tree, uint32_t* next_nid, int32_t* next_node,
                        int32_t* next_parent, int32_t* child_slot, int64_t* counters, void* stream);

/* R7 unfused fallback (the row -> node relation MLlib recomputes with predictImpl in findBestSplits; inside fit: kdd99.py:79, cicids17.py:83): routes every entry of every split slot to its child: left entries grow up from seg_begin,
 * right entries grow down from seg_end inside the same range of the destination buffers;
 * cursors[s*2+{0,1}] (int32, caller zeroes) end as (#left, #right).  Entries of children
 * that are leaves are dropped. */
int b200flow_partition_level(const uint8_t* tp, int32_t tp_stride,
                             const void* ent, void* ent_out,
                             int32_t n_slots, const int64_t* seg_begin, const int64_t* seg_end,
                             const int64_t* chunk_off, int64_t n_chunks, int32_t chunk_rows,
                             const b200flow_split* split, int32_t* cursors, void* stream);

/* R7 (inside fit: kdd99.py:79, cicids17.py:83) fused with the row routing: partition_level(L) + hist_level(L+1) in ONE pass — every entry's TreePoint
 * record is gathered once, routed by its parent's split and accumulated into its CHILD's histogram
 * (hist_next[child_slot][j][bin][class], child feature subsets in subset_next, caller zeroes hist_next and
 * cursors).  The kernel is persistent (148 x k CTAs); each warp gathers the records of its entries with
 * asynchronous copies into a private shared-memory tile and there is no CTA barrier except when the parent slot
 * changes.
 * b200flow_route_hist_config() picks the launch shape for a level shape (host-only, no device needed): it returns
 * 1 and writes chunk_rows (entries per routing chunk = warps x entries per warp step; pass it to plan_route and to
 * route_hist_level) and m_pass (subset features whose two child histograms share shared memory; m_pass < m means
 * ceil(m / m_pass) feature passes over the entries, only the first of which routes — DecisionTree nodes, whose
 * histograms cover every feature), or returns 0 when even one feature's pair of child histograms does not fit
 * (then use partition_level followed by hist_level).
 * flags bit 0: route — write the kept entries to ent_out and count them in cursors (8-byte aligned: a slot's pair is
 * advanced by one 64-bit atomic); without it only the child
 * histograms are built (ent_out / cursors may be NULL): the level-0 pass, whose segments do not change, and the
 * pass that builds the deepest scored level, whose entries are never read again. */
int b200flow_route_hist_config(int32_t F, int32_t m, int32_t n_bins, int32_t C, int32_t* chunk_rows, int32_t* m_pass);
int b200flow_route_hist_level(const uint8_t* tp, int32_t tp_stride, int32_t F,
                              const void* ent, void* ent_out,
                              int32_t n_slots, const int64_t* seg_begin, const int64_t* seg_end,
                              const int64_t* chunk_off, const int64_t* n_chunks_dev /* = chunk_off[n_slots], on the device */,
                              int64_t n_chunks_max /* host upper bound: sizes the scratch and the table launch */,
                              int32_t chunk_rows,
                              const b200flow_split* split, const int32_t* child_slot, int32_t* cursors,
                              void* chunk_scratch /* 16 bytes per chunk (n_chunks_max), 16-byte aligned */,
                              const uint16_t* subset_next, int32_t m, int32_t n_bins, int32_t C,
                              uint32_t* hist_next, int32_t flags, void* stream);

/* routing plan of a scored level: n_chunks[s] = ceil(len(s) / chunk_rows) for a split parent with at least one non-leaf
 * child, else 0 (its exclusive scan is route_hist_level's chunk_off); also scatters split[s].gain into node_gain[slot_node[s]]
 * when node_gain != NULL (TreeEnsembleModel.featureImportances needs the gains; MLlib keeps them in the Node objects). */
int b200flow_plan_route(int32_t n_slots, const b200flow_split* split, const int64_t* seg_begin, const int64_t* seg_end,
                        int32_t chunk_rows, const int32_t* slot_node, double* node_gain, int32_t* n_chunks,
                        int32_t* cursors /* NULL, or int32[2 * n_slots] zeroed here for the routing pass */, void* stream);

/* R7/R8 bookkeeping (inside fit: kdd99.py:79, cicids17.py:83): segment table of the next level from the parents' ranges and the partition cursors */
int b200flow_next_segments(int32_t n_next /* or an upper bound */, const int64_t* n_next_dev /* NULL or the device-side count */,
                           const int32_t* next_parent,
                           const int64_t* seg_begin, const int64_t* seg_end,
                           const int32_t* cursors, int64_t* next_begin, int64_t* next_end,
                           void* stream);

/* R9 preparation (LeafNode / ImpurityCalculator.prob in ml/tree/Node.scala [MLlib]; fit at kdd99.py:79): leaf payloads: prob[node][k] = counts[k] / Σcounts (fp64 true division), 0 if Σ == 0 */
int b200flow_finalize_forest(int64_t n_nodes, const uint32_t* pool_counts, int32_t C,
                             double* leaf_prob, void* stream);

/* ----------------------------------------------------------------- predict ---
 * R9  RandomForestClassificationModel.transform (kdd99.py:82, cicids17.py:86) — HOT LOOP C.  Walks all T trees for every
 * binned row; raw[n][C] = Σ_t leaf_prob (tree order, fp64), prob = raw/Σraw, pred = first argmax.
 * dt_mode != 0 (DecisionTreeClassifier): raw = leaf class counts.  raw/prob may be NULL. */
int b200flow_predict(const uint8_t* tp, int32_t tp_stride, int64_t n_rows,
                     const b200flow_node* nodes, const uint64_t* node_mask,
                     const double* leaf_prob, const uint32_t* pool_counts,
                     int32_t T, int32_t C, int32_t dt_mode,
                     const void* top_nodes /* NULL or the table of b200flow_build_top_nodes */, int32_t top_levels,
                     double* raw, double* prob, double* pred, void* stream);

/* shared-memory table for predict: top[tree][nid] (16-byte b200flow_node, [T][2^top_levels], entry 0 unused) = the tree's
 * node with MLlib node id nid < 2^top_levels.  Built once per model; the walk of the first top_levels levels then reads
 * shared memory instead of issuing one L1 request per lane and level. */
int b200flow_build_top_nodes(const b200flow_node* nodes, const int32_t* node_tree, int64_t n_nodes, int32_t T,
                             int32_t top_levels, void* top, void* stream);

/* R9 helper (model.transform, kdd99.py:82, cicids17.py:86): out[i] = src[idx[i]] for rows of row_bytes (multiple of 4): spreads the predictions computed once per UNIQUE test record
 * (b200flow_dedup_rows) back to the rows. */
int b200flow_gather_rows(const void* src, int32_t row_bytes, const int32_t* idx, int64_t n_rows, void* out, void* stream);

/* R10 MulticlassMetrics (evaluator.evaluate: kdd99.py:86-91, cicids17.py:90-95): confusion matrix cm[label*C + pred] += 1 (int64, caller zeroes).
 * pred / label are fp64 columns (as in the prediction DataFrame). */
int b200flow_confusion(const double* pred, const double* label, int64_t n_rows, int32_t C,
                       int64_t* cm, void* stream);

/* -------------------------------------------------- either side of the path ---
 * DataFrame.randomSplit (kdd99.py:52, cicids17.py:56): split id per row from a uniform keyed
 * by (seed, global row): first k with u < cum_bounds[k] (n_splits <= 8).  out uint8[n]. */
int b200flow_random_split(uint64_t seed, int64_t row_offset, int64_t n_rows,
                          const double* cum_bounds_host, int32_t n_splits, uint8_t* split_id,
                          void* stream);

/* SURVEY 8f rank 1 (Dataset.randomSplit kdd99.py:52, where cicids17.py:30-35, handleInvalid=skip cicids17.py:41): stable row compaction (where / handleInvalid="skip" / one randomSplit part):
 * keeps rows with flag[i] == want; out_rows gets the kept rows' row_bytes-sized records in
 * order.  Counting per block and the scan happen inside; scratch: (n_blocks+1) int64 followed by
 * n_blocks int32, n_blocks = ceil(n/1024); *n_kept (device int64) receives the count. */
int b200flow_compact_rows(const void* rows, int64_t n_rows, int32_t row_bytes,
                          const uint8_t* flag, int32_t want, void* out_rows,
                          int64_t* scratch, int64_t* n_kept, void* stream);

/* -------------------------------------------------- CSV text -> flow records ---
 * replaces `spark.read.csv(path, inferSchema=True, header=...)` at kdd99.py:25 and
 * cicids17.py:19-20.  The caller copies the file's bytes to the device (16-byte aligned) and owns every buffer.
 * Unquoted fields, ',' delimiter, LF or CRLF line ends, blank lines skipped (univocity's default).
 * Column classes follow Spark's inference order; parsing is exact (Java parseInt / parseDouble semantics, see
 * csrc/csv_number.h) or the field is counted in bad[] — never approximated. */
#define B200FLOW_CSV_NULL   0
#define B200FLOW_CSV_INT32  1
#define B200FLOW_CSV_INT64  2   /* inference only: such a column is read as DOUBLE (the record layout has no int64) */
#define B200FLOW_CSV_DOUBLE 3
#define B200FLOW_CSV_STRING 4   /* stored as an int32 dictionary code; -1 = null */
typedef struct b200flow_csv_col {
    int32_t type;       /* B200FLOW_CSV_INT32 / DOUBLE / STRING */
    int32_t rec_off;    /* byte offset of the field inside the output record (4-byte aligned) */
    int32_t str_index;  /* STRING: which dictionary table (0..n_string_columns-1) */
    int32_t reserved;
} b200flow_csv_col;
/* bad[8] (device, zero-initialised except [1] and [5] = ~0): [0] rows whose field count != n_cols, [1] first such row,
 * [2] rows longer than 4096 bytes, [3] fields that do not parse as their column's type, [4] numeric literals outside the
 * exact range (more than 19 digits that straddle a rounding boundary, |exponent| > 27), [5] first (row << 16 | col) of
 * [3]/[4], [6] dictionary table full, [7] dictionary lookups that failed (hash collision). */

/* line index, step 1: counts[b] = non-empty lines starting in text block b (4096 bytes each); flags[0] bit 0 = a '"' was seen */
int b200flow_csv_count_lines(const uint8_t* text, int64_t n_bytes, int32_t* counts, unsigned long long* flags, void* stream);
/* line index, step 2: bases = exclusive prefix sum of counts (int64); row_starts[i] = byte offset of non-empty line i */
int b200flow_csv_line_starts(const uint8_t* text, int64_t n_bytes, const int64_t* bases, int64_t* row_starts, void* stream);
/* inferSchema: col_class[c] = max over rows of the field's class (atomicMax into zeroed int32[n_cols]), col_null[c] = 1 if
 * any field of the column is empty.  flags: bit 0 ignoreLeadingWhiteSpace, bit 1 ignoreTrailingWhiteSpace. */
int b200flow_csv_infer(const uint8_t* text, int64_t n_bytes, const int64_t* row_starts, int64_t n_rows, int32_t n_cols,
                       int32_t flags, int32_t* col_class, int32_t* col_null, unsigned long long* bad, void* stream);
/* string columns: keys [n_str][2^cap_log2] (zeroed) receive the 64-bit FNV-1a hash of every distinct value, pos_len
 * (filled with INT64_MAX) the smallest (byte offset << 16 | length) at which it occurs — order of first appearance */
int b200flow_csv_dictionary(const uint8_t* text, int64_t n_bytes, const int64_t* row_starts, int64_t n_rows, int32_t n_cols,
                            int32_t flags, const b200flow_csv_col* cols, unsigned long long* keys, long long* pos_len,
                            int32_t cap_log2, unsigned long long* bad, void* stream);
/* fields -> records[n_rows][row_bytes]; slot_code[n_str][2^cap_log2] = dictionary code of each occupied slot (the host
 * assigns codes after reading the tables); a STRING field must byte-equal its slot's first occurrence or bad[7] counts it */
int b200flow_csv_parse(const uint8_t* text, int64_t n_bytes, const int64_t* row_starts, int64_t n_rows, int32_t n_cols,
                       int32_t flags, const b200flow_csv_col* cols, const unsigned long long* keys, const long long* pos_len,
                       const int32_t* slot_code, int32_t cap_log2, void* records, int32_t row_bytes, unsigned long long* bad,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200FLOW_H */
