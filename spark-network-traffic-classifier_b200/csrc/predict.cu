// predict.cu — batch prediction (HOT LOOP C), the confusion-matrix kernel of
// MulticlassMetrics (R10) and the two relational steps either side of the path (§8f rank 1):
// randomSplit and stable row compaction.
// Reference call sites: model.transform(test_set) kdd99.py:82 / cicids17.py:86; evaluator.evaluate
// kdd99.py:86-91; randomSplit kdd99.py:52; where / handleInvalid="skip" cicids17.py:30-35,41.
#include <stdlib.h>

#include "common.cuh"

namespace b200flow {

// ------------------------------------------------------------------ R9 predict
// A thread owns kPredRows rows and walks them through each tree TOGETHER: the walk is a chain of dependent 16-byte node
// loads (L1/L2 hits: the pool of a 100-tree depth-16 forest is a few MB), so two independent chains per thread double the
// memory-level parallelism.  A row's bins live in transposed smem (word k of thread t at [k*blockDim + t], conflict-free);
// votes accumulate in fp64, in tree order, in smem ([class*blockDim + t]).
//
// What bounds the walk is the L1: below the first few levels every lane of a warp is at a different node, so each level costs
// 32 separate sector requests per warp and row (ncu: issue 36 %, LSU 29 %, long-scoreboard stalls — the t-stage serialises
// them).  The top `top_levels` levels of the CURRENT tree (2^K - 1 nodes, heap-indexed by MLlib's node id) are therefore
// staged in shared memory, double-buffered with cp.async one tree ahead: a scattered LDS.128 costs a handful of bank
// wavefronts instead of 32 tag lookups, and only the levels below K go to the L1/L2.

// top[tree][nid] = the tree's node with MLlib id nid, for nid < 2^K (entry 0 unused)
__global__ void __launch_bounds__(256) build_top_kernel(const int4* __restrict__ nodes, const int32_t* __restrict__ node_tree,
                                                        int64_t n_nodes, int K, int4* top) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    const int4 nd = nodes[i];
    if ((uint32_t)nd.w < (1u << K)) top[((int64_t)node_tree[i] << K) + nd.w] = nd;
}

template <int kPredRows>
__global__ void __launch_bounds__(128) predict_kernel(const uint8_t* __restrict__ tp, int stride, int64_t n,
                                                      const b200flow_node* __restrict__ nodes,
                                                      const unsigned long long* __restrict__ node_mask,
                                                      const double* __restrict__ leaf_prob,
                                                      const uint32_t* __restrict__ pool_counts, int T, int C, int dt_mode,
                                                      const int4* __restrict__ top, int K,
                                                      double* raw, double* prob, double* pred) {
    extern __shared__ __align__(16) uint8_t sm[];
    const int bd = blockDim.x, tid = threadIdx.x;
    const int words = stride / 4;
    uint32_t* binw = (uint32_t*)sm;                                           // [kPredRows][words][bd]
    double* votes = (double*)(sm + (size_t)kPredRows * words * bd * 4);       // [kPredRows][C][bd]
    const int topn = top ? (1 << K) : 0;
    int4* topbuf = (int4*)(votes + (size_t)kPredRows * C * bd);               // [2][topn] when the top table is given
    auto stage_top = [&](int t, int buf) {                                    // asynchronous copy of tree t's table
        if (t < T) for (int i = tid; i < topn; i += bd) cp_async16(topbuf + (size_t)buf * topn + i, top + ((int64_t)t << K) + i);
        cp_async_commit();
    };
    for (int64_t base = (int64_t)blockIdx.x * bd * kPredRows; base < n; base += (int64_t)gridDim.x * bd * kPredRows) {
        int64_t row[kPredRows]; bool live[kPredRows];
#pragma unroll
        for (int r = 0; r < kPredRows; ++r) {
            row[r] = base + (int64_t)r * bd + tid; live[r] = row[r] < n;
            uint32_t* bw = binw + (size_t)r * words * bd;
            if (live[r]) {
                const uint4* src = (const uint4*)(tp + row[r] * stride);
                for (int q = 0; q < words / 4; ++q) {
                    const uint4 v = ld_stream_u4(src + q);
                    bw[(4 * q + 0) * bd + tid] = v.x; bw[(4 * q + 1) * bd + tid] = v.y;
                    bw[(4 * q + 2) * bd + tid] = v.z; bw[(4 * q + 3) * bd + tid] = v.w;
                }
            }
            for (int k = 0; k < C; ++k) votes[((size_t)r * C + k) * bd + tid] = 0.0;
        }
        const uint32_t* bw0 = binw + tid;                                   // this thread's column of the transposed bins
        const int4* nodes4 = (const int4*)nodes;                            // {feat, kind<<16|bin, left, nid}
        if (top) stage_top(0, 0);
        for (int t = 0; t < T; ++t) {
            const int4* tb = topbuf + (size_t)(t & 1) * topn;
            if (top) {
                cp_async_wait_all();
                __syncthreads();                                            // table of tree t landed; everybody left tree t-1
                stage_top(t + 1, (t + 1) & 1);
            }
            int idx[kPredRows]; uint32_t nid[kPredRows]; int4 nd[kPredRows];
#pragma unroll
            for (int r = 0; r < kPredRows; ++r) { idx[r] = t; nid[r] = 1u; nd[r] = top ? tb[1] : __ldg(nodes4 + t); if (!live[r]) nd[r].x = -1; }
            while (true) {
                bool any = false;
#pragma unroll
                for (int r = 0; r < kPredRows; ++r) {
                    if (nd[r].x >= 0) {
                        const int f = nd[r].x;
                        const int bin = (bw0[(r * words + (f >> 2)) * bd] >> ((f & 3) * 8)) & 0xff;
                        const int right = nd[r].y < 65536 ? (bin > nd[r].y)
                                                          : !((node_mask[(int64_t)idx[r] * 4 + (bin >> 6)] >> (bin & 63)) & 1ull);
                        idx[r] = nd[r].z + right;
                        nid[r] = 2u * nid[r] + (uint32_t)right;
                        any = true;
                    }
                }
                if (!any) break;
#pragma unroll
                for (int r = 0; r < kPredRows; ++r)
                    if (nd[r].x >= 0) nd[r] = nid[r] < (uint32_t)topn ? tb[nid[r]] : __ldg(nodes4 + idx[r]);
            }
#pragma unroll
            for (int r = 0; r < kPredRows; ++r) {
                if (!live[r]) continue;
                double* vt = votes + (size_t)r * C * bd + tid;
                if (dt_mode) { for (int k = 0; k < C; ++k) vt[k * bd] += (double)pool_counts[(int64_t)idx[r] * C + k]; }
                else { for (int k = 0; k < C; ++k) vt[k * bd] += leaf_prob[(int64_t)idx[r] * C + k]; }
            }
        }
#pragma unroll
        for (int r = 0; r < kPredRows; ++r) {
            if (!live[r]) continue;
            const double* vt = votes + (size_t)r * C * bd + tid;
            double s = 0.0; int arg = 0; double best = vt[0];
            for (int k = 0; k < C; ++k) { const double v = vt[k * bd]; s += v; if (v > best) { best = v; arg = k; } }
            for (int k = 0; k < C; ++k) {
                const double v = vt[k * bd];
                if (raw) raw[row[r] * C + k] = v;
                if (prob) prob[row[r] * C + k] = s != 0.0 ? v / s : 0.0;
            }
            pred[row[r]] = (double)arg;
        }
        if (top) { cp_async_wait_all(); __syncthreads(); }                  // the look-ahead copy of the last tree is a no-op commit
    }
}

// ------------------------------------------------------------------ row gather (predictions of unique records -> rows)
__global__ void __launch_bounds__(256) gather_rows_kernel(const uint32_t* __restrict__ src, int words, const int32_t* __restrict__ idx,
                                                          int64_t n, uint32_t* out) {
    const int64_t total = n * words;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / words; const int w = (int)(e - r * words);
        out[e] = __ldg(src + (int64_t)idx[r] * words + w);
    }
}

// ------------------------------------------------------------------ R10 confusion matrix
__global__ void __launch_bounds__(256) confusion_kernel(const double* __restrict__ pred, const double* __restrict__ label,
                                                        int64_t n, int C, unsigned long long* cm, int use_smem) {
    extern __shared__ uint32_t sh_cm[];
    if (use_smem) { for (int i = threadIdx.x; i < C * C; i += blockDim.x) sh_cm[i] = 0; __syncthreads(); }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = (int)label[i], p = (int)pred[i];
        if (l >= 0 && l < C && p >= 0 && p < C) {
            if (use_smem) atomicAdd(&sh_cm[l * C + p], 1u); else atomicAdd(&cm[l * C + p], 1ull);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < C * C; i += blockDim.x) if (sh_cm[i]) atomicAdd(&cm[i], (unsigned long long)sh_cm[i]);
    }
}

// ------------------------------------------------------------------ randomSplit
struct SplitBounds { double cum[8]; int n; };

__global__ void __launch_bounds__(256) random_split_kernel(uint64_t seed, int64_t row_offset, int64_t n, SplitBounds bnd,
                                                           uint8_t* out) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t g = (uint64_t)(row_offset + i);
        const uint4 r = philox_keyed(seed, PURPOSE_RSPLIT, (uint32_t)g, (uint32_t)(g >> 32), 0u, 0u);
        const double u = (double)r.x * 2.3283064365386963e-10;     // 2^-32
        int k = 0;
        while (k < bnd.n - 1 && !(u < bnd.cum[k])) ++k;
        out[i] = (uint8_t)k;
    }
}

// ------------------------------------------------------------------ stable row compaction
constexpr int kCompactRows = 1024;

__global__ void __launch_bounds__(256) compact_count_kernel(const uint8_t* __restrict__ flag, int64_t n, int want, int32_t* blk_cnt) {
    __shared__ int cnt;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    int c = 0;
    const int64_t rb = (int64_t)blockIdx.x * kCompactRows + threadIdx.x * 4;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (rb + k < n && (flag[rb + k] != 0) == (want != 0)) ++c;
    c = warp_sum(c);
    if (lane_id() == 0 && c) atomicAdd(&cnt, c);
    __syncthreads();
    if (threadIdx.x == 0) blk_cnt[blockIdx.x] = cnt;
}

__global__ void __launch_bounds__(256) compact_scatter_kernel(const uint8_t* __restrict__ rows, int64_t n, int row_bytes,
                                                              const uint8_t* __restrict__ flag, int want,
                                                              const int64_t* __restrict__ blk_off, uint8_t* out) {
    __shared__ int sh[33];
    __shared__ int src_of[kCompactRows];                 // kept rows of this block, in order
    const int64_t rb0 = (int64_t)blockIdx.x * kCompactRows;
    const int64_t rb = rb0 + threadIdx.x * 4;
    int keep[4], c = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { keep[k] = (rb + k < n && (flag[rb + k] != 0) == (want != 0)) ? 1 : 0; c += keep[k]; }
    int tot;
    int pos = block_exclusive_scan(c, sh, &tot);
#pragma unroll
    for (int k = 0; k < 4; ++k) if (keep[k]) src_of[pos++] = threadIdx.x * 4 + k;
    __syncthreads();
    const int words = row_bytes / 4;
    const uint32_t* src = (const uint32_t*)(rows + rb0 * row_bytes);
    uint32_t* dst = (uint32_t*)(out + blk_off[blockIdx.x] * row_bytes);
    for (int64_t i = threadIdx.x; i < (int64_t)tot * words; i += blockDim.x) {
        int r = (int)(i / words), wd = (int)(i - (int64_t)r * words);
        dst[i] = src[(int64_t)src_of[r] * words + wd];
    }
}

}  // namespace b200flow

using namespace b200flow;

extern "C" int b200flow_build_top_nodes(const b200flow_node* nodes, const int32_t* node_tree, int64_t n_nodes, int32_t T,
                                        int32_t top_levels, void* top, void* stream) {
    B2F_REQUIRE(nodes && node_tree && top && T > 0 && top_levels >= 1 && top_levels <= 10 && ((uintptr_t)top & 15) == 0, "build_top_nodes: bad arguments");
    if (n_nodes <= 0) return B200FLOW_OK;
    build_top_kernel<<<(unsigned)((n_nodes + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const int4*)nodes, node_tree, n_nodes, top_levels, (int4*)top);
    return check_launch("build_top_nodes");
}

extern "C" int b200flow_predict(const uint8_t* tp, int32_t tp_stride, int64_t n_rows, const b200flow_node* nodes,
                                const uint64_t* node_mask, const double* leaf_prob, const uint32_t* pool_counts, int32_t T,
                                int32_t C, int32_t dt_mode, const void* top_nodes, int32_t top_levels,
                                double* raw, double* prob, double* pred, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;            // empty batch: nothing to do (pointers may be NULL)
    B2F_REQUIRE(tp && nodes && pred && T > 0 && C > 0 && (tp_stride & 15) == 0, "predict: bad arguments");
    B2F_REQUIRE(dt_mode ? pool_counts != nullptr : leaf_prob != nullptr, "predict: missing leaf payload");
    B2F_REQUIRE(((uintptr_t)tp & 15) == 0, "predict: tp must be 16-byte aligned");
    B2F_REQUIRE(!top_nodes || (top_levels >= 1 && top_levels <= 10 && ((uintptr_t)top_nodes & 15) == 0), "predict: bad top table");
    static int rows_per_thread = -1;                      // tuning knob: independent tree walks per thread (memory-level parallelism)
    if (rows_per_thread < 0) { const char* e = getenv("B200FLOW_PRED_ROWS"); rows_per_thread = (e && atoi(e) == 4) ? 4 : 2; }
    const int kPredRows = rows_per_thread;
    int bd = 128;
    size_t per_thread = ((size_t)tp_stride + (size_t)C * 8) * kPredRows;
    while (bd > 32 && per_thread * bd > 96 * 1024) bd >>= 1;
    size_t smem = per_thread * bd + (top_nodes ? (size_t)2 * 16 * ((size_t)1 << top_levels) : 0);
    B2F_REQUIRE(smem <= 200 * 1024, "predict: too many classes/features for shared memory");
    int grid = grid_for(n_rows, bd * kPredRows, kNumSMs * 16);
    cudaError_t e;
    if (kPredRows == 4) {
        e = cudaFuncSetAttribute(predict_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) predict_kernel<4><<<grid, bd, smem, (cudaStream_t)stream>>>(tp, tp_stride, n_rows, nodes, (const unsigned long long*)node_mask, leaf_prob,
                                                                                          pool_counts, T, C, dt_mode, (const int4*)top_nodes, top_levels, raw, prob, pred);
    } else {
        e = cudaFuncSetAttribute(predict_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) predict_kernel<2><<<grid, bd, smem, (cudaStream_t)stream>>>(tp, tp_stride, n_rows, nodes, (const unsigned long long*)node_mask, leaf_prob,
                                                                                          pool_counts, T, C, dt_mode, (const int4*)top_nodes, top_levels, raw, prob, pred);
    }
    if (e != cudaSuccess) { set_error("predict: %s", cudaGetErrorString(e)); return B200FLOW_ERR_CUDA; }
    return check_launch("predict");
}

extern "C" int b200flow_gather_rows(const void* src, int32_t row_bytes, const int32_t* idx, int64_t n_rows, void* out, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;
    B2F_REQUIRE(src && idx && out && row_bytes > 0 && (row_bytes & 3) == 0, "gather_rows: bad arguments");
    const int words = row_bytes / 4;
    gather_rows_kernel<<<grid_for(n_rows * words, 256 * 4, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>((const uint32_t*)src, words, idx, n_rows, (uint32_t*)out);
    return check_launch("gather_rows");
}

extern "C" int b200flow_confusion(const double* pred, const double* label, int64_t n_rows, int32_t C, int64_t* cm, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;            // empty batch: nothing to do (pointers may be NULL)
    B2F_REQUIRE(pred && label && cm && C > 0 && C <= 1024, "confusion: bad arguments");
    int use_smem = (size_t)C * C * 4 <= 64 * 1024;
    size_t smem = use_smem ? (size_t)C * C * 4 : 0;
    if (smem > 48 * 1024) cudaFuncSetAttribute(confusion_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int grid = grid_for(n_rows, 256 * 8, kNumSMs * 4);
    confusion_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(pred, label, n_rows, C, (unsigned long long*)cm, use_smem);
    return check_launch("confusion");
}

extern "C" int b200flow_random_split(uint64_t seed, int64_t row_offset, int64_t n_rows, const double* cum_bounds_host,
                                     int32_t n_splits, uint8_t* split_id, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;            // empty batch: nothing to do (pointers may be NULL)
    B2F_REQUIRE(cum_bounds_host && split_id && n_splits >= 1 && n_splits <= 8, "random_split: bad arguments");
    SplitBounds b; b.n = n_splits;
    for (int i = 0; i < 8; ++i) b.cum[i] = i < n_splits ? cum_bounds_host[i] : 2.0;
    int grid = grid_for(n_rows, 256 * 4, kNumSMs * 8);
    random_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(seed, row_offset, n_rows, b, split_id);
    return check_launch("random_split");
}

extern "C" int b200flow_compact_rows(const void* rows, int64_t n_rows, int32_t row_bytes, const uint8_t* flag, int32_t want,
                                     void* out_rows, int64_t* scratch, int64_t* n_kept, void* stream) {
    B2F_REQUIRE(rows && flag && out_rows && scratch && n_kept && row_bytes > 0 && (row_bytes & 3) == 0, "compact_rows: bad arguments");
    B2F_REQUIRE(((uintptr_t)rows & 3) == 0 && ((uintptr_t)out_rows & 3) == 0, "compact_rows: buffers must be 4-byte aligned");
    if (n_rows <= 0) { cudaMemsetAsync(n_kept, 0, 8, (cudaStream_t)stream); return check_launch("compact_rows"); }
    const int64_t nb = (n_rows + kCompactRows - 1) / kCompactRows;
    // scratch: int64 off[nb+1] first (8-aligned), then int32 cnt[nb]
    int64_t* off = scratch;
    int32_t* cnt = (int32_t*)(scratch + nb + 1);
    compact_count_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>(flag, n_rows, want, cnt);
    int rc = b200flow_exclusive_scan_i32_to_i64(cnt, nb, off, n_kept, stream);
    if (rc) return rc;
    compact_scatter_kernel<<<(unsigned)nb, 256, 0, (cudaStream_t)stream>>>((const uint8_t*)rows, n_rows, row_bytes, flag, want, off, (uint8_t*)out_rows);
    return check_launch("compact_rows");
}
