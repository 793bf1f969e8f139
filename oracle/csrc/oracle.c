/* Oracle (test infrastructure, never on the product path): C restatement of the scoring + exact masked top-N of
 * base/IterativeRecommender.py:58-60 (predict = Q.dot(P[u])), 102-106 (drop the user's training tracks) and the
 * "find the K biggest scores" intent of 107-145, with the build's canonical score: the float32 FMA chain
 * acc = fmaf(P[u][k], Q[t][k], acc), k = 0..d-1.  oracle/topn.py emulates that chain in numpy
 * (float64 product + float64 sum rounded to float32: one rounding more than a true FMA); this file uses fmaf()
 * itself and is the tie-breaker; tests/test_oracle_golden.py compares the two.
 * Build: make -C oracle   ->  oracle/_build/liboracle.so */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

void oracle_scores_fma32(const float* P_rows, const float* Q, int64_t B, int64_t n, int d, float* out) {
    for (int64_t b = 0; b < B; ++b)
        for (int64_t t = 0; t < n; ++t) {
            float acc = 0.0f;
            for (int k = 0; k < d; ++k) acc = fmaf(P_rows[b * d + k], Q[t * d + k], acc);
            out[b * n + t] = acc;
        }
}

/* ids[B,N] (-1 padded), scores[B,N] (-inf padded): per user the N unmasked tracks with the largest score, order
 * (score desc, track id asc).  uq_* = sorted unique training tracks per user. */
void oracle_topn_exact(const float* P, const float* Q, int64_t n, int d, const int32_t* users, int64_t B, int N,
                       const int64_t* uq_indptr, const int32_t* uq_items, int32_t* ids, float* scores) {
    float* best_s = (float*)malloc(sizeof(float) * (size_t)N);
    int32_t* best_i = (int32_t*)malloc(sizeof(int32_t) * (size_t)N);
    for (int64_t b = 0; b < B; ++b) {
        const int64_t u = users[b];
        int64_t m = uq_indptr[u];
        const int64_t mend = uq_indptr[u + 1];
        int have = 0;
        for (int64_t t = 0; t < n; ++t) {
            while (m < mend && uq_items[m] < t) ++m;
            if (m < mend && uq_items[m] == t) continue;
            float acc = 0.0f;
            for (int k = 0; k < d; ++k) acc = fmaf(P[u * d + k], Q[t * d + k], acc);
            acc += 0.0f;                                  /* -0 -> +0, like the kernels' sort key */
            /* insert: tracks arrive in increasing id, so an equal score goes AFTER the ones already kept */
            int pos = have;
            while (pos > 0 && best_s[pos - 1] < acc) --pos;
            if (pos >= N) continue;
            const int last = have < N ? have : N - 1;
            for (int x = last; x > pos; --x) { best_s[x] = best_s[x - 1]; best_i[x] = best_i[x - 1]; }
            best_s[pos] = acc; best_i[pos] = (int32_t)t;
            if (have < N) ++have;
        }
        for (int x = 0; x < N; ++x) {
            ids[b * N + x] = x < have ? best_i[x] : -1;
            scores[b * N + x] = x < have ? best_s[x] : -INFINITY;
        }
    }
    free(best_s);
    free(best_i);
}
