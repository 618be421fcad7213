/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Plain-C restatement of the reference's contact-map
 * kernels.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load this; the product path never does.
 *
 * Parity is PINNED: tests/test_oracle_cmap.py checks every function below bit-for-bit
 * against the reference itself (mDeepFRI/contact_map_utils.pyx compiled unchanged into
 * oracle/_ref/ by oracle/Makefile) and against the committed fixtures in tests/golden/.
 *
 * Build: gcc -O3 -fopenmp -fPIC -shared -ffp-contract=off  (no FMA, like the reference's
 * x86-64 baseline build, setup.py:241).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* contact_map_utils.pyx:17-37 — D[i,j] = D[j,i] = sum_k (X[i,k]-X[j,k])^2, accumulated
 * from d = 0.0f in k order, unfused; diagonal untouched (0). */
void mdf_oracle_pairwise_sqeuclidean(const float *X, int n, int m, float *D, int threads)
{
    memset(D, 0, (size_t)n * (size_t)n * sizeof(float));
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static, 1) num_threads(threads)
    for (int i = 0; i < n; ++i) {
        for (int j = i + 1; j < n; ++j) {
            float d = 0.0f;
            for (int k = 0; k < m; ++k) {
                float diff = X[(size_t)i * m + k] - X[(size_t)j * m + k];
                d = d + (diff * diff);
            }
            D[(size_t)i * n + j] = d;
            D[(size_t)j * n + i] = d;
        }
    }
}

/* bio_utils.py:214-220 / contact_map.py:74 — cmap = (D < float32(thr**2)).astype(int32),
 * strict '<' evaluated in float32 (NumPy-2 weak-scalar promotion). */
void mdf_oracle_threshold(const float *D, int64_t count, float thr2, int32_t *cmap)
{
    for (int64_t i = 0; i < count; ++i) cmap[i] = D[i] < thr2 ? 1 : 0;
}

/* bio_utils.py:222-223 / contact_map.py:88-95 — np.argwhere(cmap == 1) in row-major order.
 * `pairs` may be NULL to only count. */
int64_t mdf_oracle_sparsify(const int32_t *cmap, int n, int32_t *pairs)
{
    int64_t k = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j)
            if (cmap[(size_t)i * n + j] == 1) {
                if (pairs) { pairs[2 * k] = i; pairs[2 * k + 1] = j; }
                ++k;
            }
    return k;
}

/* contact_map_utils.pyx:64-80 — number of non-gap query columns (output side length). */
int mdf_oracle_query_length(const char *q, int aln_len)
{
    int n = 0;
    for (int i = 0; i < aln_len; ++i) n += q[i] != '-';
    return n;
}

/* contact_map_utils.pyx:44-117.  `out` is int32[Lq*Lq]; returns Lq, or -1 on allocation
 * failure.  Target indices are compared against the map size as UNSIGNED (the reference
 * compares a C int against vector::size(), :109), so negative indices are skipped. */
int mdf_oracle_align_contact_map(const char *q, const char *t, int aln_len,
                                 const int32_t *sparse, int64_t nnz, int gen,
                                 int32_t *out, int threads)
{
    int *t2q = (int *)malloc(sizeof(int) * (size_t)(aln_len > 0 ? aln_len : 1));
    size_t gcap = (size_t)(aln_len > 0 ? aln_len : 1) * (size_t)(gen > 0 ? gen : 0) * 4 + 4;
    int *genp = (int *)malloc(sizeof(int) * gcap);
    if (!t2q || !genp) { free(t2q); free(genp); return -1; }
    size_t nt = 0, ng = 0;
    int qi = 0;
    for (int i = 0; i < aln_len; ++i) {                       /* :64-80 column walk */
        if (q[i] == 45) {
            t2q[nt++] = -1;
        } else if (t[i] == 45) {
            for (int j = 1; j <= gen; ++j) {
                genp[ng++] = qi + j; genp[ng++] = qi;
                genp[ng++] = qi - j; genp[ng++] = qi;
            }
            ++qi;
        } else {
            t2q[nt++] = qi;
            ++qi;
        }
    }
    const int Lq = qi;
    memset(out, 0, sizeof(int32_t) * (size_t)Lq * (size_t)Lq);  /* :82 */
    for (int i = 0; i < Lq; ++i) out[(size_t)i * Lq + i] = 1;   /* :85-86 */
    for (size_t i = 0; i < ng; i += 2) {                        /* :91-97 */
        int p1 = genp[i], p2 = genp[i + 1];
        if (0 <= p1 && p1 < Lq && 0 <= p2 && p2 < Lq) {
            out[(size_t)p1 * Lq + p2] = 1;
            out[(size_t)p2 * Lq + p1] = 1;
        }
    }
    if (threads < 1) threads = 1;
#pragma omp parallel for schedule(static) num_threads(threads)
    for (int64_t r = 0; r < nnz; ++r) {                         /* :105-115, one direction */
        int ti = sparse[2 * r], tj = sparse[2 * r + 1];
        if ((size_t)(int64_t)ti < nt && (size_t)(int64_t)tj < nt) {
            int a = t2q[ti], b = t2q[tj];
            if (a != -1 && b != -1) out[(size_t)a * Lq + b] = 1;
        }
    }
    free(t2q); free(genp);
    return Lq;
}
