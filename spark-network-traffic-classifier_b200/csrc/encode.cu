// encode.cu — the fused row-parallel encode path.
//
// Replaces, for the reference call sites kdd99.py:34-37,45-46 and cicids17.py:41-46, MLlib's
// StringIndexer.fit (category counts), StringIndexerModel.transform (code -> rank lookup),
// OneHotEncoder (expand), StandardScaler (fit moments + scale) and VectorAssembler (concat)
// with ONE pass over the raw AoS flow records.
//
// encode kernel, data movement (HBM-bound; algorithmic bytes/row = row_bytes + n_out*sizeof(out) + 4):
//   HBM --cp.async.bulk (TMA, UBLKCP) + mbarrier--> smem record tile [R x row_bytes], NS-deep ring
//   threads: one fixed OUTPUT slot per thread, rows strided -> smem out tile [R x n_out] (bank-conflict free)
//   smem out tile --cp.async.bulk store (bulk_group)--> HBM, triple buffered; ONE __syncthreads per tile
// Arithmetic is fp64 ((v - mean) * scale, no FMA contraction) and rounded once to the output type.
#include <stdlib.h>

#include "common.cuh"

namespace b200flow {

// ------------------------------------------------------------------ R1 category counts
__global__ void __launch_bounds__(256) category_counts_kernel(const uint8_t* __restrict__ rec, int64_t n, int row_bytes,
                                                              int src_off, int K, unsigned long long* counts,
                                                              int use_smem) {
    extern __shared__ uint32_t sh_cnt[];
    if (use_smem) {
        for (int i = threadIdx.x; i < K; i += blockDim.x) sh_cnt[i] = 0;
        __syncthreads();
    }
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int code = __ldg((const int*)(rec + i * row_bytes + src_off));
        if (code >= 0 && code < K) {
            if (use_smem) atomicAdd(&sh_cnt[code], 1u);
            else atomicAdd(&counts[code], 1ull);
        }
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < K; i += blockDim.x)
            if (sh_cnt[i]) atomicAdd(&counts[i], (unsigned long long)sh_cnt[i]);
    }
}

// several columns in ONE pass over the records: the code fields of a flow record sit in the same one or two 32-byte sectors,
// so four StringIndexer fits cost one read of the batch instead of four (kdd99.py:34-37 fits four indexers back to back)
struct CountCols { int n; int off[8]; int K[8]; int base[8]; };

__global__ void __launch_bounds__(256) category_counts_multi_kernel(const uint8_t* __restrict__ rec, int64_t n, int row_bytes,
                                                                    const CountCols cc, int total, unsigned long long* counts) {
    extern __shared__ uint32_t sh_cnt[];
    for (int i = threadIdx.x; i < total; i += blockDim.x) sh_cnt[i] = 0;
    __syncthreads();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t* r = rec + i * row_bytes;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            if (c < cc.n) {
                const int code = __ldg((const int*)(r + cc.off[c]));
                if (code >= 0 && code < cc.K[c]) atomicAdd(&sh_cnt[cc.base[c] + code], 1u);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < total; i += blockDim.x)
        if (sh_cnt[i]) atomicAdd(&counts[i], (unsigned long long)sh_cnt[i]);
}

// ------------------------------------------------------------------ fused encode
struct EncodeArgs {
    const uint8_t* records; int64_t n_rows; int row_bytes;
    const b200flow_slot* plan; int n_out;
    const int32_t* lut; int lut_total; int lut_in_smem;
    int label_off, label_lut_off, label_lut_len, check_nan;
    void* out; int32_t* label_out; uint8_t* valid_out;
    int R;            // rows per tile (multiple of 4)
    int in_stride;    // bytes per input stage (128-aligned)
    int out_stride;   // bytes per output stage (128-aligned)
};

constexpr int kEncStages = 3;     // TMA load ring depth
constexpr int kEncOutBufs = 3;    // output tiles: two bulk stores may be in flight while the third tile is computed
constexpr int kEncThreads = 256;
constexpr int kEncMaxCat = 8;     // distinct categorical sources whose rank is shared through shared memory

template <typename OUT>
__global__ void __launch_bounds__(kEncThreads) encode_kernel(const EncodeArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    // layout: [in ring][out x3][mbar][badtag 2*R][plan][lut]
    uint8_t* in_base = smem;
    uint8_t* out_base = in_base + (size_t)kEncStages * a.in_stride;
    uint64_t* mbar = (uint64_t*)(out_base + kEncOutBufs * (size_t)a.out_stride);
    int32_t* badtag = (int32_t*)(mbar + kEncStages);
    b200flow_slot* plan_sh = (b200flow_slot*)(badtag + 2 * a.R + ((2 * a.R) & 1));   // keep 8-byte alignment
    int32_t* cat_tab = (int32_t*)(plan_sh + a.n_out);          // [3][kEncMaxCat + 1]: field offset, LUT offset, LUT length
    int32_t* rank_sh = cat_tab + 3 * (kEncMaxCat + 1);         // [R][kEncMaxCat]: StringIndexer rank per (row, categorical source)
    int16_t* slot_cat = (int16_t*)(rank_sh + a.R * kEncMaxCat);   // [n_out]: categorical source of a slot, -1 numeric, -2 inline
    int16_t* slot_rep = slot_cat + a.n_out + (a.n_out & 1);
    int32_t* lut_sh = (int32_t*)(slot_rep + a.n_out + (a.n_out & 1));
    __shared__ int sh_ncat;

    const int tid = threadIdx.x, bd = blockDim.x;
    const int R = a.R, n_out = a.n_out, row_bytes = a.row_bytes;
    const int64_t n_tiles = (a.n_rows + R - 1) / R;
    const uint32_t tile_in_bytes = (uint32_t)R * row_bytes;
    const uint32_t tile_out_bytes = (uint32_t)R * n_out * sizeof(OUT);

    if (tid == 0) {
        for (int s = 0; s < kEncStages; ++s) mbar_init(&mbar[s], 1);
        fence_mbar_init();
    }
    for (int i = tid; i < n_out * (int)(sizeof(b200flow_slot) / 4); i += bd) ((uint32_t*)plan_sh)[i] = ((const uint32_t*)a.plan)[i];
    if (a.lut_in_smem) for (int i = tid; i < a.lut_total; i += bd) lut_sh[i] = a.lut[i];
    for (int i = tid; i < 2 * R; i += bd) badtag[i] = 0;
    __syncthreads();
    const int32_t* lut = a.lut_in_smem ? lut_sh : a.lut;
    // distinct categorical sources (field, LUT): the code -> rank lookup is done ONCE per (row, source) in a pre-pass, not
    // once per one-hot slot (70 slots share KDD's `service` lookup)
    for (int d = tid; d < n_out; d += bd) {
        int rep = -1;
        if (plan_sh[d].kind >= B200FLOW_SRC_INDEX) {
            rep = d;
            if (n_out <= 1024)
                for (int e = 0; e < d; ++e)
                    if (plan_sh[e].kind >= B200FLOW_SRC_INDEX && plan_sh[e].src_off == plan_sh[d].src_off &&
                        plan_sh[e].lut_off == plan_sh[d].lut_off && plan_sh[e].lut_len == plan_sh[d].lut_len) { rep = e; break; }
        }
        slot_rep[d] = (int16_t)rep;
    }
    __syncthreads();
    if (tid == 0) {
        int nc = 0;
        for (int d = 0; d < n_out; ++d) {
            int id = -1;
            if (slot_rep[d] == d) {
                if (nc < kEncMaxCat && n_out <= 1024) {
                    id = nc; cat_tab[nc] = plan_sh[d].src_off; cat_tab[(kEncMaxCat + 1) + nc] = plan_sh[d].lut_off;
                    cat_tab[2 * (kEncMaxCat + 1) + nc] = plan_sh[d].lut_len; ++nc;
                } else id = -2;
            }
            slot_cat[d] = (int16_t)id;
        }
        int n_cat_slots = 0;
        for (int d = 0; d < n_out; ++d) n_cat_slots += slot_rep[d] >= 0 ? 1 : 0;
        if (n_cat_slots < 3 * nc) {                          // few slots per source (plain StringIndexer columns): the shared
            nc = 0;                                          // pre-pass and its barrier cost more than the inline lookups
            for (int d = 0; d < n_out; ++d) if (slot_rep[d] >= 0) slot_cat[d] = -2;
        }
        if (a.label_off >= 0) {                              // the label column is one more source (index nc)
            cat_tab[nc] = a.label_off; cat_tab[(kEncMaxCat + 1) + nc] = a.label_lut_off; cat_tab[2 * (kEncMaxCat + 1) + nc] = a.label_lut_len;
        }
        sh_ncat = nc;
    }
    __syncthreads();
    for (int d = tid; d < n_out; d += bd) { const int rep = slot_rep[d]; if (rep >= 0 && rep != d) slot_cat[d] = slot_cat[rep]; }
    __syncthreads();
    const int ncat = sh_ncat, nsrc = ncat + (a.label_off >= 0 ? 1 : 0);

    auto issue_load = [&](int64_t tile, int s) {
        int64_t rows = a.n_rows - tile * R;
        if (rows >= R) {
            mbar_arrive_expect_tx(&mbar[s], tile_in_bytes);
            bulk_g2s(in_base + (size_t)s * a.in_stride, a.records + tile * (int64_t)tile_in_bytes, tile_in_bytes, &mbar[s]);
        } else {
            mbar_arrive(&mbar[s]);      // ragged last tile: loaded cooperatively below
        }
    };
    if (tid == 0)
        for (int s = 0; s < kEncStages; ++s) {
            int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
            if (t < n_tiles) issue_load(t, s);
        }

    // thread -> output slot mapping: one fixed slot per thread when n_out <= blockDim, rows fastest, so that the lanes of
    // a warp share their slot's source kind (no divergence between the numeric and the lookup loops)
    const bool fixed = n_out <= bd;
    const int rp = fixed ? bd / n_out : 1;
    const int d0 = fixed ? tid / rp : tid;
    const int r0 = fixed ? tid % rp : 0;
    const int dstep = fixed ? n_out : bd;
    const bool active = fixed ? (tid < rp * n_out) : true;

    int it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int s = it % kEncStages, o = it % kEncOutBufs, ob = it & 1;
        const uint32_t ph = (uint32_t)(it / kEncStages) & 1u;
        const int64_t row_base = tile * R;
        const int rows = (int)min((int64_t)R, a.n_rows - row_base);
        const bool full = rows == R;
        uint8_t* in_t = in_base + (size_t)s * a.in_stride;
        OUT* out_t = (OUT*)(out_base + (size_t)o * a.out_stride);
        int32_t* bad = badtag + ob * R;
        const int tag = it + 1;

        mbar_wait(&mbar[s], ph);
        if (!full) {
            const uint32_t* src = (const uint32_t*)(a.records + row_base * row_bytes);
            for (int i = tid; i < rows * row_bytes / 4; i += bd) ((uint32_t*)in_t)[i] = __ldg(src + i);
            __syncthreads();
        }

        if (nsrc > 0) {
            for (int idx = tid; idx < rows * nsrc; idx += bd) {
                const int r = idx / nsrc, c = idx - r * nsrc;
                const int code = *(const int32_t*)(in_t + r * row_bytes + cat_tab[c]);
                const int rank = (code >= 0 && code < cat_tab[2 * (kEncMaxCat + 1) + c]) ? lut[cat_tab[(kEncMaxCat + 1) + c] + code] : -1;
                if (rank < 0) bad[r] = tag;
                if (c < ncat) rank_sh[r * kEncMaxCat + c] = rank;
                else if (a.label_out) a.label_out[row_base + r] = rank;
            }
            if (ncat > 0) __syncthreads();          // ranks visible to the slot loops (label-only: nothing to wait for)
        }
        if (active) {
            // one tight loop per source kind: the slot (and so the kind) is fixed per thread, rows are strided by rp
            for (int d = d0; d < n_out; d += dstep) {
                const b200flow_slot sl = plan_sh[d];
                const uint8_t* p = in_t + r0 * row_bytes + sl.src_off;
                OUT* op = out_t + r0 * n_out + d;
                const int pstep = rp * row_bytes, ostep = rp * n_out;
                const bool ident = sl.mean == 0.0 && sl.scale == 1.0;
                const double mean = sl.mean, scale = sl.scale;
                if (sl.kind == B200FLOW_SRC_F32) {
                    if (ident && sizeof(OUT) == 4 && !a.check_nan) {
#pragma unroll 4
                        for (int r = r0; r < rows; r += rp, p += pstep, op += ostep) *(uint32_t*)op = *(const uint32_t*)p;   // assemble = move
                    } else {
#pragma unroll 4
                        for (int r = r0; r < rows; r += rp, p += pstep, op += ostep) {
                            const double v = (double)(*(const float*)p);
                            if (a.check_nan && v != v) bad[r] = tag;
                            *op = ident ? (OUT)v : (OUT)((v - mean) * scale);
                        }
                    }
                } else if (sl.kind == B200FLOW_SRC_F64) {
#pragma unroll 4
                    for (int r = r0; r < rows; r += rp, p += pstep, op += ostep) {
                        const uint32_t* q = (const uint32_t*)p;              // 4-byte aligned reads: fields need not be 8-aligned
                        const double v = __hiloint2double((int)q[1], (int)q[0]);
                        if (a.check_nan && v != v) bad[r] = tag;
                        *op = ident ? (OUT)v : (OUT)((v - mean) * scale);
                    }
                } else if (sl.kind == B200FLOW_SRC_I32) {
#pragma unroll 4
                    for (int r = r0; r < rows; r += rp, p += pstep, op += ostep) {
                        const double v = (double)(*(const int32_t*)p);
                        *op = ident ? (OUT)v : (OUT)((v - mean) * scale);
                    }
                } else {
                    const bool index = sl.kind == B200FLOW_SRC_INDEX;
                    const OUT hot = (OUT)((1.0 - mean) * scale), cold = (OUT)((0.0 - mean) * scale);
                    const int cat = slot_cat[d];
                    if (cat >= 0) {                  // rank looked up once per (row, source) by the pre-pass
                        const int32_t* rk = rank_sh + r0 * kEncMaxCat + cat;
                        const int rstep = rp * kEncMaxCat;
                        if (index) {
#pragma unroll 4
                            for (int r = r0; r < rows; r += rp, rk += rstep, op += ostep) *op = (OUT)(((double)*rk - mean) * scale);
                        } else {
                            const int hotrank = sl.hot;
#pragma unroll 4
                            for (int r = r0; r < rows; r += rp, rk += rstep, op += ostep) *op = (*rk == hotrank) ? hot : cold;
                        }
                    } else {                         // more than kEncMaxCat distinct sources: inline lookup
                        const int32_t* lt = lut + sl.lut_off;
                        for (int r = r0; r < rows; r += rp, p += pstep, op += ostep) {
                            const int code = *(const int32_t*)p;
                            const int rank = (code >= 0 && code < sl.lut_len) ? lt[code] : -1;
                            if (rank < 0) bad[r] = tag;
                            *op = index ? (OUT)(((double)rank - mean) * scale) : (rank == sl.hot ? hot : cold);
                        }
                    }
                }
            }
        }
        if (tid == 0) bulk_wait_read<1>();           // all but the latest store have drained: the next tile's out buffer is free
        fence_proxy_async();
        __syncthreads();                            // the only barrier per tile: tile computed, input stage s consumed
        if (tid == 0) {
            if (full) { bulk_s2g((OUT*)a.out + row_base * n_out, out_t, tile_out_bytes); bulk_commit(); }
            int64_t next = tile + (int64_t)kEncStages * gridDim.x;
            if (next < n_tiles) issue_load(next, s);
        }
        if (!full) {
            OUT* dst = (OUT*)a.out + row_base * n_out;
            for (int i = tid; i < rows * n_out; i += bd) dst[i] = out_t[i];
        }
        if (a.valid_out) for (int r = tid; r < rows; r += bd) a.valid_out[row_base + r] = (bad[r] != tag) ? 1 : 0;
    }
    if (tid == 0) bulk_wait_all<0>();
}

// ------------------------------------------------------------------ R3c column moments
template <typename T>
__global__ void __launch_bounds__(256) column_moments_kernel(const T* __restrict__ x, int64_t n, int D, int64_t ld,
                                                             const double* __restrict__ shift, double* sum, double* sumsq) {
    extern __shared__ double sh_m[];     // [2][blockDim]
    const int tid = threadIdx.x, bd = blockDim.x;
    const int64_t rows_per_block = (n + gridDim.x - 1) / gridDim.x;
    const int64_t rb = (int64_t)blockIdx.x * rows_per_block, re = min(n, rb + rows_per_block);
    for (int dbase = 0; dbase < D; dbase += bd) {
        const int dn = min(bd, D - dbase);
        const int rp = bd / dn;
        double s = 0.0, q = 0.0;
        if (tid < rp * dn) {
            const int d = dbase + tid % dn;
            const double sft = shift ? shift[d] : 0.0;
            for (int64_t r = rb + tid / dn; r < re; r += rp) {
                double v = (double)x[r * ld + d] - sft;
                s += v; q += v * v;
            }
        }
        sh_m[tid] = s; sh_m[bd + tid] = q;
        __syncthreads();
        if (tid < dn) {
            for (int k = 1; k < rp; ++k) { s += sh_m[tid + k * dn]; q += sh_m[bd + tid + k * dn]; }
            atomicAdd(&sum[dbase + tid], s);
            atomicAdd(&sumsq[dbase + tid], q);
        }
        __syncthreads();
    }
}

}  // namespace b200flow

using namespace b200flow;

extern "C" int b200flow_category_counts(const void* records, int64_t n_rows, int32_t row_bytes, int32_t src_off,
                                        int32_t K, int64_t* counts, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;            // empty batch: nothing to do (pointers may be NULL)
    B2F_REQUIRE(records && counts && K > 0 && row_bytes >= 4 && src_off >= 0 && src_off + 4 <= row_bytes && (src_off & 3) == 0 &&
                    (row_bytes & 3) == 0, "category_counts: bad arguments");
    int use_smem = K <= 8192;
    int grid = grid_for(n_rows, 256 * 8, kNumSMs * 8);
    category_counts_kernel<<<grid, 256, use_smem ? K * 4 : 0, (cudaStream_t)stream>>>(
        (const uint8_t*)records, n_rows, row_bytes, src_off, K, (unsigned long long*)counts, use_smem);
    return check_launch("category_counts");
}

extern "C" int b200flow_category_counts_multi(const void* records, int64_t n_rows, int32_t row_bytes, int32_t n_cols,
                                              const int32_t* src_offs_host, const int32_t* Ks_host, int64_t* counts, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;
    B2F_REQUIRE(records && counts && src_offs_host && Ks_host && n_cols >= 1 && n_cols <= 8 && row_bytes >= 4 && (row_bytes & 3) == 0,
                "category_counts_multi: bad arguments");
    CountCols cc; cc.n = n_cols;
    int total = 0;
    for (int c = 0; c < 8; ++c) {
        cc.off[c] = c < n_cols ? src_offs_host[c] : 0; cc.K[c] = c < n_cols ? Ks_host[c] : 0; cc.base[c] = total;
        if (c < n_cols) {
            B2F_REQUIRE(cc.K[c] > 0 && cc.off[c] >= 0 && cc.off[c] + 4 <= row_bytes && (cc.off[c] & 3) == 0, "category_counts_multi: bad column %d", c);
            total += cc.K[c];
        }
    }
    B2F_REQUIRE(total <= 8192, "category_counts_multi: more than 8192 categories in total (count the columns one by one)");
    int grid = grid_for(n_rows, 256 * 8, kNumSMs * 8);
    category_counts_multi_kernel<<<grid, 256, total * 4, (cudaStream_t)stream>>>((const uint8_t*)records, n_rows, row_bytes, cc, total,
                                                                                (unsigned long long*)counts);
    return check_launch("category_counts_multi");
}

extern "C" int b200flow_encode(const void* records, int64_t n_rows, int32_t row_bytes, const b200flow_slot* plan,
                               int32_t n_out, const int32_t* lut, int32_t lut_total, int32_t label_off,
                               int32_t label_lut_off, int32_t label_lut_len, int32_t check_nan, void* out,
                               int32_t out_dtype, int32_t* label_out, uint8_t* valid_out, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;            // empty batch: nothing to do (pointers may be NULL)
    B2F_REQUIRE(records && plan && out, "encode: null pointer");
    B2F_REQUIRE(n_out > 0 && n_out <= 4096 && row_bytes >= 4 && (row_bytes & 3) == 0, "encode: bad n_out/row_bytes");
    B2F_REQUIRE(out_dtype == B200FLOW_F32 || out_dtype == B200FLOW_F64, "encode: bad out_dtype");
    B2F_REQUIRE(((uintptr_t)records & 15) == 0 && ((uintptr_t)out & 15) == 0, "encode: records/out must be 16-byte aligned");
    B2F_REQUIRE(label_off < 0 || (label_off + 4 <= row_bytes && (label_off & 3) == 0 && lut), "encode: bad label_off");
    const int osz = out_dtype == B200FLOW_F32 ? 4 : 8;
    EncodeArgs a;
    a.records = (const uint8_t*)records; a.n_rows = n_rows; a.row_bytes = row_bytes; a.plan = plan; a.n_out = n_out;
    a.lut = lut; a.lut_total = lut ? lut_total : 0; a.lut_in_smem = (lut && lut_total > 0 && lut_total <= 4096) ? 1 : 0;
    a.label_off = label_off; a.label_lut_off = label_lut_off; a.label_lut_len = label_lut_len; a.check_nan = check_nan;
    a.out = out; a.label_out = label_out; a.valid_out = valid_out;
    // tile rows: keep one CTA near 52 KB of smem so four CTAs share an SM (>= 64 KB of loads in flight per SM)
    const int fixed_bytes = kEncStages * 8 + n_out * (int)sizeof(b200flow_slot) + (a.lut_in_smem ? a.lut_total * 4 : 0) + 1024 +
                            3 * (kEncMaxCat + 1) * 4 + 4 * (n_out + 1);
    const int per_row = row_bytes * kEncStages + n_out * osz * kEncOutBufs + 8 + kEncMaxCat * 4;
    static int budget_kb = -1;                             // tuning knob: per-CTA shared memory target
    if (budget_kb < 0) { const char* e = getenv("B200FLOW_ENC_SMEM_KB"); budget_kb = e ? atoi(e) : 52; }
    int budget = budget_kb * 1024 - fixed_bytes;
    int R = budget > 0 ? budget / per_row : 0;
    if (R < 4) R = 4;                    // very wide rows: fewer CTAs per SM
    if (R > 512) R = 512;
    R &= ~3;
    if ((int64_t)R > ((n_rows + 3) & ~(int64_t)3)) R = (int)((n_rows + 3) & ~(int64_t)3);
    a.R = R;
    a.in_stride = (R * row_bytes + 127) & ~127;
    a.out_stride = (R * n_out * osz + 127) & ~127;
    size_t smem = (size_t)kEncStages * a.in_stride + kEncOutBufs * (size_t)a.out_stride + kEncStages * 8 + (2 * R + 2) * 4 +
                  (size_t)n_out * sizeof(b200flow_slot) + (a.lut_in_smem ? (size_t)a.lut_total * 4 : 0) + 16 +
                  3 * (kEncMaxCat + 1) * 4 + (size_t)R * kEncMaxCat * 4 + 4 * ((size_t)n_out + 1);
    B2F_REQUIRE(smem <= 227 * 1024, "encode: record too wide for shared memory (row_bytes=%d n_out=%d)", row_bytes, n_out);
    const int64_t n_tiles = (n_rows + R - 1) / R;
    int ctas_per_sm = (int)((220 * 1024) / (smem + 1024));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    int grid = (int)(n_tiles < (int64_t)kNumSMs * ctas_per_sm ? n_tiles : (int64_t)kNumSMs * ctas_per_sm);
    cudaError_t e;
    if (out_dtype == B200FLOW_F32) {
        e = cudaFuncSetAttribute(encode_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) encode_kernel<float><<<grid, kEncThreads, smem, (cudaStream_t)stream>>>(a);
    } else {
        e = cudaFuncSetAttribute(encode_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) encode_kernel<double><<<grid, kEncThreads, smem, (cudaStream_t)stream>>>(a);
    }
    if (e != cudaSuccess) { set_error("encode: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B200FLOW_ERR_CUDA; }
    return check_launch("encode");
}

extern "C" int b200flow_column_moments(const void* x, int32_t dtype, int64_t n_rows, int32_t D, int64_t ld,
                                       const double* shift, double* sum, double* sumsq, void* stream) {
    if (n_rows <= 0) return B200FLOW_OK;            // empty batch: nothing to do (pointers may be NULL)
    B2F_REQUIRE(x && sum && sumsq && D > 0 && ld >= D, "column_moments: bad arguments");
    B2F_REQUIRE(dtype == B200FLOW_F32 || dtype == B200FLOW_F64, "column_moments: bad dtype");
    int grid = grid_for(n_rows, 256, kNumSMs * 4);
    size_t smem = 2 * 256 * sizeof(double);
    if (dtype == B200FLOW_F32)
        column_moments_kernel<float><<<grid, 256, smem, (cudaStream_t)stream>>>((const float*)x, n_rows, D, ld, shift, sum, sumsq);
    else
        column_moments_kernel<double><<<grid, 256, smem, (cudaStream_t)stream>>>((const double*)x, n_rows, D, ld, shift, sum, sumsq);
    return check_launch("column_moments");
}
