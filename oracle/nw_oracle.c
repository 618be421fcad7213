/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * CPU restatement of the global alignment the reference requests from PyOpal (mDeepFRI/alignment.py:196-221 align_pairwise:
 * pyopal.Aligner(scoring_matrix, gap_open, gap_extend).align(query, database, algorithm="nw", mode="full"); :163-194
 * best_hit_database: the same in mode="score").  The algorithm lives in a third-party dependency absent from this image and from
 * /root/reference: PyOpal (pyproject.toml: pyopal, un-pinned), a wrapper of the Opal library (Šošić, "An SIMD dynamic
 * programming C/C++ library", 2015).  Opal's published recurrences for algorithm NW, restated here as plain Gotoh dynamic
 * programming over full matrices:
 *      a gap of length k costs gap_open + (k - 1) * gap_extend (Opal: "gapOpen = penalty for the first residue of a gap,
 *      gapExt = for every further one"); both sequences are aligned end to end, end gaps are charged;
 *      H[i][j] = max(H[i-1][j-1] + S(q_i, t_j), E[i][j], F[i][j])
 *      E[i][j] = max(H[i-1][j] - open, E[i-1][j] - ext)      query residue against a gap in the target  -> 'D'
 *      F[i][j] = max(H[i][j-1] - open, F[i][j-1] - ext)      target residue against a gap in the query  -> 'I'
 *      column letters as mDeepFRI.alignment.insert_gaps reads them (alignment.py:38-62): 'M' equal residues, 'X' mismatch,
 *      'I' = '-' in the query, 'D' = '-' in the target.
 * The optimal SCORE is unique.  Which optimal alignment is reported when several exist is Opal's traceback order, which is not
 * documented: PARITY UNPINNED for tied alignments.  The order fixed here: diagonal before 'D' before 'I'; a gap is closed rather
 * than extended when both are optimal.  The reference's own known-answer tests (tests/test_alignment.py:22-34) have no ties
 * and are reproduced (tests/test_alignment.py in this repo).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define NEG (-1000000000)

/* q, t: residue codes < A; S: A x A int8 row-major.  ops: capacity lq + lt.  Returns the score; *n_ops = alignment columns. */
int32_t mdf_oracle_nw(const uint8_t *q, int lq, const uint8_t *t, int lt, const int8_t *S, int A, int open, int ext, char *ops, int *n_ops)
{
    const size_t W = (size_t)lt + 1;
    int32_t *H = malloc(sizeof(int32_t) * (lq + 1) * W), *E = malloc(sizeof(int32_t) * (lq + 1) * W), *F = malloc(sizeof(int32_t) * (lq + 1) * W);
    H[0] = 0; E[0] = F[0] = NEG;
    for (int j = 1; j <= lt; ++j) { H[j] = -(open + (j - 1) * ext); E[j] = NEG; F[j] = H[j]; }
    for (int i = 1; i <= lq; ++i) {
        H[i * W] = -(open + (i - 1) * ext); E[i * W] = H[i * W]; F[i * W] = NEG;
        for (int j = 1; j <= lt; ++j) {
            const int32_t e1 = H[(i - 1) * W + j] - open, e2 = E[(i - 1) * W + j] - ext;
            const int32_t f1 = H[i * W + j - 1] - open, f2 = F[i * W + j - 1] - ext;
            const int32_t e = e2 > e1 ? e2 : e1, f = f2 > f1 ? f2 : f1;
            int32_t h = H[(i - 1) * W + j - 1] + S[q[i - 1] * A + t[j - 1]];
            if (e > h) h = e;
            if (f > h) h = f;
            H[i * W + j] = h; E[i * W + j] = e; F[i * W + j] = f;
        }
    }
    const int32_t score = H[(size_t)lq * W + lt];
    if (ops) {
        int i = lq, j = lt, n = 0, state = 0;          /* 0 = H, 1 = E ('D'), 2 = F ('I') */
        while (i > 0 || j > 0) {
            if (i == 0) { ops[n++] = 'I'; --j; continue; }
            if (j == 0) { ops[n++] = 'D'; --i; continue; }
            if (state == 0) {
                const int32_t h = H[i * W + j];
                if (h == H[(i - 1) * W + j - 1] + S[q[i - 1] * A + t[j - 1]]) { ops[n++] = q[i - 1] == t[j - 1] ? 'M' : 'X'; --i; --j; }
                else if (h == E[i * W + j]) state = 1;
                else state = 2;
            } else if (state == 1) {
                const int extend = E[(i - 1) * W + j] - ext > H[(i - 1) * W + j] - open;
                ops[n++] = 'D'; --i;
                state = extend ? 1 : 0;
            } else {
                const int extend = F[i * W + j - 1] - ext > H[i * W + j - 1] - open;
                ops[n++] = 'I'; --j;
                state = extend ? 2 : 0;
            }
        }
        for (int a = 0, b = n - 1; a < b; ++a, --b) { const char c = ops[a]; ops[a] = ops[b]; ops[b] = c; }
        *n_ops = n;
    }
    free(H); free(E); free(F);
    return score;
}
