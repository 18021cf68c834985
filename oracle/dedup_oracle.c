/* dedup_oracle.c -- CPU restatement of the content-duplicate scan. TEST INFRASTRUCTURE ONLY (same rules as
 * aid_oracle.c: tests/, __graft_entry__.smoke() and bench_dedup.py's CPU-baseline legs may load it, as the checker
 * or the timed baseline; the product path never does).
 *
 * PARITY PINNED: unlike the fingerprint stages, the reference implements this arithmetic itself, in Python:
 *   audio-ident-service/app/audio/dedup.py:127-167  _fingerprint_similarity
 *   audio-ident-service/app/audio/dedup.py:170-222  check_content_duplicate
 * tests/golden/dedup_contract.json holds outputs of those two functions (imported from /root/reference by
 * tests/golden/make_dedup_golden.py, including the reference's own test cases tests/test_audio_dedup.py:137-246);
 * tests/test_dedup_oracle.py checks this file against them bit for bit.
 */
#include <stdint.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* dedup.py:127-167 on parsed fingerprints (words = the integers modulo 2^32, which is all :158 looks at).
 * Same operations, same order, IEEE double: (matching_bits / total_bits) * (min_len / max_len). */
double aid_oracle_fp_similarity(const uint32_t *a, int64_t na, const uint32_t *b, int64_t nb) {
    if (na <= 0 || nb <= 0) return 0.0;                         /* :148-149 */
    const int64_t min_len = na < nb ? na : nb;                  /* :152 */
    const int64_t max_len = na < nb ? nb : na;                  /* :164 */
    int64_t matching_bits = 0;
    const int64_t total_bits = min_len * 32;                    /* :157 */
    for (int64_t i = 0; i < min_len; i++)                       /* :159-163 */
        matching_bits += 32 - __builtin_popcount(a[i] ^ b[i]);
    const double length_penalty = (double)min_len / (double)max_len;         /* :165 */
    return ((double)matching_bits / (double)total_bits) * length_penalty;    /* :167 */
}

/* dedup.py:170-222: candidate rows are those with q_lo <= duration <= q_hi (the SQL WHERE, :192-197, with
 * q_lo = duration*0.9 and q_hi = duration*1.1 computed by the caller as :189-190 does); the running best is replaced
 * only by a strictly greater similarity (:206-209), starting from 0.0 / None (:201-202). The threshold test (:214)
 * is left to the caller. Rows are scanned in order inside each thread's contiguous block and the blocks are folded in
 * order, so the first row among equals wins exactly as in the sequential loop. */
void aid_oracle_dedup_scan(const uint32_t *words, const int64_t *off, const double *dur, int64_t n_rows,
                           const uint32_t *q_words, const int64_t *q_off, const double *q_lo, const double *q_hi,
                           int nq, int64_t *best_row, double *best_sim, int n_threads) {
    for (int q = 0; q < nq; q++) {
        const uint32_t *qw = q_words + q_off[q];
        const int64_t qn = q_off[q + 1] - q_off[q];
        double gb = 0.0;
        int64_t gr = -1;
#ifdef _OPENMP
        const int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#pragma omp parallel num_threads(nt)
#endif
        {
            double b = 0.0;
            int64_t r = -1;
#ifdef _OPENMP
#pragma omp for schedule(static) nowait
#endif
            for (int64_t i = 0; i < n_rows; i++) {
                if (!(dur[i] >= q_lo[q] && dur[i] <= q_hi[q])) continue;
                const double s = aid_oracle_fp_similarity(qw, qn, words + off[i], off[i + 1] - off[i]);
                if (s > b) { b = s; r = i; }
            }
#ifdef _OPENMP
#pragma omp critical
#endif
            { if (b > gb || (b == gb && b > 0.0 && r < gr)) { gb = b; gr = r; } }
        }
        best_row[q] = gr;
        best_sim[q] = gb;
    }
    (void)n_threads;
}
