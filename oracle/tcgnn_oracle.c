/* CPU oracle for the TC-GNN aggregation path -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference algorithm (no code shared with the product in
 * tc-gnn_atc23_b200/).  Used by tests/ (checker), __graft_entry__.smoke() (checker) and
 * bench.py's cpu_baseline / --impl reference leg (the timed CPU arm, OpenMP over rows).
 * Parity pinning: tests/test_oracle.py checks every function here against the NumPy
 * restatement (oracle/tcgnn_oracle.py) and against the golden SGT arrays produced by the
 * reference's own compiled TCGNN.preprocess (tests/golden/).
 *
 * build: gcc -O3 -fopenmp -fPIC -shared oracle/tcgnn_oracle.c -o oracle/_build/libtcgnn_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* cvt.rna.tf32.f32 (TCGNN_kernel.cu:436-444 via wmma::__float_to_tf32, crt/mma.h:96-103) */
static inline float tf32_rna(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    if ((u & 0x7F800000u) != 0x7F800000u) u = (u + 0x00001000u) & 0xFFFFE000u;
    memcpy(&x, &u, 4);
    return x;
}

void oracle_tf32_rna(const float *in, float *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) out[i] = tf32_rna(in[i]);
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return (x > y) - (x < y);
}

/* SGT, TCGNN.cpp:172-226.  Single-threaded like the reference (its omp pragmas are ignored
 * by its build, TCGNN_conv/debug.log:28,31).  Returns the TC_Blocks total the reference
 * prints (TCGNN.cpp:225), including the +1 of the extra loop trip when N % blk_h == 0. */
int64_t oracle_sgt(const int32_t *row_ptr, const int32_t *col_idx, int32_t num_nodes,
                   int32_t blk_h, int32_t blk_w, int32_t *block_partition,
                   int32_t *edge_to_col, int32_t *edge_to_row) {
    int64_t total = 0;
    for (int32_t r = 0; r < num_nodes; ++r)                      /* TCGNN.cpp:194-197 */
        for (int32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) edge_to_row[e] = r;
    int32_t nwin = (num_nodes + blk_h - 1) / blk_h;
    for (int32_t w = 0; w < nwin; ++w) {                         /* TCGNN.cpp:200-224 */
        int32_t s = row_ptr[w * blk_h];
        int32_t hi = w * blk_h + blk_h;
        int32_t t = row_ptr[hi < num_nodes ? hi : num_nodes];
        int32_t len = t - s;
        uint32_t *buf = (uint32_t *)malloc((size_t)(len > 0 ? len : 1) * sizeof(uint32_t));
        memcpy(buf, col_idx + s, (size_t)len * sizeof(uint32_t));
        qsort(buf, (size_t)len, sizeof(uint32_t), cmp_u32);       /* :209 */
        int32_t nu = 0;                                           /* :157-170 */
        for (int32_t i = 0; i < len; ++i)
            if (i == 0 || buf[i] != buf[i - 1]) buf[nu++] = buf[i];
        int32_t cnt = nu > 0 ? nu : 1;                            /* empty window -> map of size 1 */
        block_partition[w] = (cnt + blk_w - 1) / blk_w;           /* :216 */
        total += block_partition[w];
        for (int32_t e = s; e < t; ++e) {                         /* :220-223 */
            uint32_t key = (uint32_t)col_idx[e];
            int32_t lo = 0, hi2 = nu;
            while (lo < hi2) { int32_t m = (lo + hi2) >> 1; if (buf[m] < key) lo = m + 1; else hi2 = m; }
            edge_to_col[e] = lo;
        }
        free(buf);
    }
    if (num_nodes % blk_h == 0) total += 1;                       /* extra trip of :200 */
    return total;
}

/* SpMM, TCGNN_kernel.cu:336-454 (weights == NULL) and :459-578 (weights = edgeAttention row 0).
 * Y[i,:] = sum_e w_e * tf32(X[col[e],:]); unweighted uses the 0/1 pattern (:405): a repeated
 * (row, col) pair counts once (rows are scanned for repeats; CSR from dataset.py has none).
 * fp32 accumulate, OpenMP over rows for the CPU-baseline timing. */
void oracle_spmm(const int32_t *row_ptr, const int32_t *col_idx, const float *weights,
                 const float *x, int64_t ldx, float *y, int64_t ldy, int32_t num_nodes,
                 int32_t dim, int32_t round_tf32) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = 0; r < num_nodes; ++r) {
        float *yr = y + (int64_t)r * ldy;
        for (int32_t d = 0; d < dim; ++d) yr[d] = 0.0f;
        for (int32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            int32_t c = col_idx[e];
            float w = 1.0f;
            if (weights) {
                w = round_tf32 ? tf32_rna(weights[e]) : weights[e];
            } else {
                int dup = 0;
                for (int32_t p = row_ptr[r]; p < e; ++p) if (col_idx[p] == c) { dup = 1; break; }
                if (dup) continue;
            }
            const float *xr = x + (int64_t)c * ldx;
            if (round_tf32) for (int32_t d = 0; d < dim; ++d) yr[d] += w * tf32_rna(xr[d]);
            else            for (int32_t d = 0; d < dim; ++d) yr[d] += w * xr[d];
        }
    }
}

/* Fast CPU baseline variant: same maths, assumes unique columns per row (true for CSR built by
 * dataset.py:94-104) so the duplicate scan is skipped.  Timed by bench.py. */
void oracle_spmm_fast(const int32_t *row_ptr, const int32_t *col_idx, const float *x, int64_t ldx,
                      float *y, int64_t ldy, int32_t row_begin, int32_t row_end, int32_t dim) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = row_begin; r < row_end; ++r) {
        float *yr = y + (int64_t)r * ldy;
        for (int32_t d = 0; d < dim; ++d) yr[d] = 0.0f;
        for (int32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            const float *xr = x + (int64_t)col_idx[e] * ldx;
            for (int32_t d = 0; d < dim; ++d) yr[d] += tf32_rna(xr[d]);
        }
    }
}

/* SDDMM, TCGNN_kernel.cu:584-728: out[e] = sum_k tf32(X[row(e),k]) * tf32(X[col(e),k]). */
void oracle_sddmm(const int32_t *row_ptr, const int32_t *col_idx, const float *x, int64_t ldx,
                  float *out, int32_t num_nodes, int32_t dim, int32_t round_tf32) {
#pragma omp parallel for schedule(dynamic, 64)
    for (int32_t r = 0; r < num_nodes; ++r) {
        const float *xr = x + (int64_t)r * ldx;
        for (int32_t e = row_ptr[r]; e < row_ptr[r + 1]; ++e) {
            const float *xc = x + (int64_t)col_idx[e] * ldx;
            float acc = 0.0f;
            if (round_tf32) for (int32_t d = 0; d < dim; ++d) acc += tf32_rna(xr[d]) * tf32_rna(xc[d]);
            else            for (int32_t d = 0; d < dim; ++d) acc += xr[d] * xc[d];
            out[e] = acc;
        }
    }
}
