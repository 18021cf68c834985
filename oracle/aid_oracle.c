/* aid_oracle.c -- CPU restatement of the fingerprint-and-match path. TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library, and only as the checker or the timed CPU baseline. The product
 * path (audio_ident_b200/) never calls it and has no CPU fallback.
 *
 * PARITY UNPINNED for stages 1-5: the reference implements none of this arithmetic. Its
 * app/audio/fingerprint.py (reference audio-ident-service/app/audio/fingerprint.py:87-219)
 * writes the PCM to a temp file and executes the third-party binary `olaf_c`
 * (github.com/JorenSix/Olaf, version "latest" = unpinned,
 * docs/research/01-initial-research/07-deliverables.md:41), which is not vendored, not
 * installed here and mocked in every reference test (tests/test_audio_fingerprint.py:141-150).
 * There is no golden vector to pin against. This file is therefore the normative definition
 * of spectrogram, peaks, hashes, index order and votes for this repo; it follows the stage
 * list of BASELINE.json north_star, the I/O contract of fingerprint.py (16 kHz mono f32le in,
 * OlafMatch rows out: fingerprint.py:30-50) and the conceptual description in
 * docs/research/01-initial-research/01-problem-definition.md:22-27 (peaks -> hashed pairs with
 * an offset -> matching by consistent time alignment). Constants: include/aid_params.h.
 * What IS pinned by the reference (the CSV row grammar and the exact-lane post-processing)
 * is checked separately in tests/test_boundary_contract.py against committed golden vectors.
 *
 * Build: make -C oracle   ->  oracle/libaid_oracle.so  (gcc -O3 -march=x86-64-v3 -fopenmp)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../include/aid_params.h"

#define NF AID_NFFT
#define NB AID_NBINS
#define HALF (NF / 2)

/* ------------------------------------------------------------------ tables */
static float  g_win[NF];
static double g_tw_re[HALF / 2], g_tw_im[HALF / 2];   /* e^{-2 pi i k / 512}, k < 256 */
static double g_pw_re[HALF], g_pw_im[HALF];           /* e^{-2 pi i k / 1024}, k < 512 */
static uint16_t g_rev[HALF];
static int g_ready = 0;

static void tables_init(void) {
    if (g_ready) return;
    #pragma omp critical(aid_tables)
    {
        if (!g_ready) {
            for (int n = 0; n < NF; n++)
                g_win[n] = (float)(AID_WIN_A0 - AID_WIN_A1 * cos(2.0 * M_PI * (double)n / (double)(NF - 1)));
            for (int k = 0; k < HALF / 2; k++) {
                g_tw_re[k] = cos(-2.0 * M_PI * k / HALF);
                g_tw_im[k] = sin(-2.0 * M_PI * k / HALF);
            }
            for (int k = 0; k < HALF; k++) {
                g_pw_re[k] = cos(-2.0 * M_PI * k / NF);
                g_pw_im[k] = sin(-2.0 * M_PI * k / NF);
            }
            for (int i = 0; i < HALF; i++) {
                int r = 0;
                for (int b = 0; b < 9; b++) if (i & (1 << b)) r |= 1 << (8 - b);
                g_rev[i] = (uint16_t)r;
            }
            g_ready = 1;
        }
    }
}

/* the float32 window table both sides use */
void aid_oracle_window(float *w) {
    tables_init();
    memcpy(w, g_win, sizeof(g_win));
}

int64_t aid_oracle_num_frames(int64_t n_samples) {
    return n_samples < NF ? 0 : (n_samples - NF) / AID_HOP + 1;
}

/* ------------------------------------------------------------------ stage 1: STFT
 * One frame: 1024 windowed real samples -> 512 bins via a 512-point complex FFT of
 * z[n] = x[2n] + i x[2n+1] (radix-2 decimation in time, double precision) and the usual
 * even/odd split. S[k] = (float) log1p(|X[k]|^2). */
static void frame_spectrum(const float *x, float *S) {
    double re[HALF], im[HALF];
    for (int i = 0; i < HALF; i++) {
        int j = g_rev[i];
        re[j] = (double)x[2 * i] * (double)g_win[2 * i];
        im[j] = (double)x[2 * i + 1] * (double)g_win[2 * i + 1];
    }
    for (int len = 2; len <= HALF; len <<= 1) {
        int half = len >> 1, step = HALF / len;
        for (int s = 0; s < HALF; s += len) {
            for (int k = 0; k < half; k++) {
                double wr = g_tw_re[k * step], wi = g_tw_im[k * step];
                int a = s + k, b = a + half;
                double tr = re[b] * wr - im[b] * wi;
                double ti = re[b] * wi + im[b] * wr;
                re[b] = re[a] - tr; im[b] = im[a] - ti;
                re[a] += tr;        im[a] += ti;
            }
        }
    }
    for (int k = 0; k < NB; k++) {
        int m = (HALF - k) & (HALF - 1);
        double er = 0.5 * (re[k] + re[m]), ei = 0.5 * (im[k] - im[m]);      /* even part  */
        double or_ = 0.5 * (im[k] + im[m]), oi = -0.5 * (re[k] - re[m]);    /* odd part   */
        double xr = er + or_ * g_pw_re[k] - oi * g_pw_im[k];
        double xi = ei + or_ * g_pw_im[k] + oi * g_pw_re[k];
        S[k] = (float)log1p(xr * xr + xi * xi);
    }
}

/* pcm[n_samples] -> S[frames][512]; returns frames */
int64_t aid_oracle_stft(const float *pcm, int64_t n_samples, float *S) {
    tables_init();
    int64_t T = aid_oracle_num_frames(n_samples);
    for (int64_t t = 0; t < T; t++) frame_spectrum(pcm + t * AID_HOP, S + t * NB);
    return T;
}

/* ------------------------------------------------------------------ stage 2: peaks
 * Sliding maximum over a clipped window [i-h, i+h] of a strided sequence (monotonic deque). */
static void sliding_max(const float *in, float *out, int64_t n, int64_t stride, int h, int64_t *dq) {
    int64_t head = 0, tail = 0, next = 0;
    for (int64_t i = 0; i < n; i++) {
        int64_t hi = i + h < n - 1 ? i + h : n - 1;
        for (; next <= hi; next++) {
            float v = in[next * stride];
            while (tail > head && in[dq[tail - 1] * stride] <= v) tail--;
            dq[tail++] = next;
        }
        while (dq[head] < i - h) head++;
        out[i * stride] = in[dq[head] * stride];
    }
}

/* S[T][512] -> keys[] = (t << 9) | f in (t, f) order. Returns the peak count, or -1 if a
 * aligned 256-frame block holds more than AID_PEAK_BLOCK_CAP peaks (capacity rule of aid_params.h).
 * keys must hold AID_PEAK_CAP(T) entries. */
int64_t aid_oracle_peaks(const float *S, int64_t T, uint32_t *keys) {
    if (T <= 0) return 0;
    float *m1 = (float *)malloc((size_t)T * NB * sizeof(float));
    float *m2 = (float *)malloc((size_t)T * NB * sizeof(float));
    int64_t *dq = (int64_t *)malloc((size_t)(T > NB ? T : NB) * sizeof(int64_t));
    for (int64_t t = 0; t < T; t++) sliding_max(S + t * NB, m1 + t * NB, NB, 1, AID_PEAK_HALF_F, dq);
    for (int f = 0; f < NB; f++) sliding_max(m1 + f, m2 + f, T, NB, AID_PEAK_HALF_T, dq);
    int64_t n = 0, in_block = 0;
    for (int64_t t = 0; t < T && n >= 0; t++) {
        if (t % AID_PEAK_BLOCK_FRAMES == 0) in_block = 0;
        for (int f = AID_PEAK_MIN_BIN; f < NB; f++) {
            float v = S[t * NB + f];
            if (v > AID_PEAK_MIN_S && v == m2[t * NB + f]) {
                if (++in_block > AID_PEAK_BLOCK_CAP) { n = -1; break; }
                keys[n++] = ((uint32_t)t << AID_PEAK_F_BITS) | (uint32_t)f;
            }
        }
    }
    free(m1); free(m2); free(dq);
    return n;
}

/* ------------------------------------------------------------------ stage 3: landmark hashes
 * peaks in (t, f) order -> (hash, t_anchor) in (anchor, target) order. Returns the count
 * (never more than n_peaks * AID_FANOUT, which is the capacity the caller provides). */
int64_t aid_oracle_hashes(const uint32_t *keys, int64_t n_peaks, uint32_t *hash, uint32_t *t_anchor) {
    int64_t n = 0;
    for (int64_t i = 0; i < n_peaks; i++) {
        int t1 = (int)(keys[i] >> AID_PEAK_F_BITS), f1 = (int)(keys[i] & (NB - 1));
        int taken = 0;
        for (int64_t j = i + 1; j < n_peaks && taken < AID_FANOUT; j++) {
            int t2 = (int)(keys[j] >> AID_PEAK_F_BITS), f2 = (int)(keys[j] & (NB - 1));
            int dt = t2 - t1, df = f2 > f1 ? f2 - f1 : f1 - f2;
            if (dt > AID_DT_MAX) break;
            if (dt < AID_DT_MIN || df < AID_DF_MIN || df > AID_DF_MAX) continue;
            hash[n] = AID_HASH(f1, f2, dt);
            t_anchor[n] = (uint32_t)t1;
            n++; taken++;
        }
    }
    return n;
}

/* pcm -> hashes. Returns the hash count, -1 on peak overflow. Optional outputs (may be NULL):
 * S_out[frames*512], peaks_out[AID_PEAK_CAP(frames)], n_peaks_out. hash/t_anchor need
 * AID_PEAK_CAP(frames) * AID_FANOUT entries. */
int64_t aid_oracle_fingerprint(const float *pcm, int64_t n_samples, uint32_t *hash, uint32_t *t_anchor,
                               float *S_out, uint32_t *peaks_out, int64_t *n_peaks_out) {
    int64_t T = aid_oracle_num_frames(n_samples);
    if (n_peaks_out) *n_peaks_out = 0;
    if (T == 0) return 0;
    float *S = S_out ? S_out : (float *)malloc((size_t)T * NB * sizeof(float));
    int64_t cap = AID_PEAK_CAP(T);
    uint32_t *pk = peaks_out ? peaks_out : (uint32_t *)malloc((size_t)cap * sizeof(uint32_t));
    aid_oracle_stft(pcm, n_samples, S);
    int64_t np = aid_oracle_peaks(S, T, pk);
    int64_t nh = np < 0 ? -1 : aid_oracle_hashes(pk, np, hash, t_anchor);
    if (n_peaks_out) *n_peaks_out = np;
    if (!S_out) free(S);
    if (!peaks_out) free(pk);
    return nh;
}

/* Ragged batch, one track per OpenMP task. hash_off[i] = i-th track's first slot in
 * hash/t_anchor (caller-sized), n_hash[i] receives its count (-1 on overflow).
 * Returns the number of threads used. This is the timed CPU baseline. */
int aid_oracle_fingerprint_batch(const float *pcm, const int64_t *sample_off, int n_tracks,
                                 uint32_t *hash, uint32_t *t_anchor, const int64_t *hash_off,
                                 int64_t *n_hash, int64_t *n_peaks, int threads) {
    tables_init();
    int used = 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    used = omp_get_max_threads();
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n_tracks; i++) {
        int64_t np = 0;
        n_hash[i] = aid_oracle_fingerprint(pcm + sample_off[i], sample_off[i + 1] - sample_off[i],
                                           hash + hash_off[i], t_anchor + hash_off[i], NULL, NULL, &np);
        if (n_peaks) n_peaks[i] = np;
    }
    return used;
}

/* ------------------------------------------------------------------ stage 4: index
 * entries (hash, track, t_anchor) -> sorted by (hash, track, t_anchor), plus
 * bucket[h] = first entry with hash >= h, h = 0..2^24 (so bucket has 2^24 + 1 slots). */
typedef struct { uint32_t hash, track, t; } entry_t;

static int entry_cmp(const void *a, const void *b) {
    const entry_t *x = (const entry_t *)a, *y = (const entry_t *)b;
    if (x->hash != y->hash) return x->hash < y->hash ? -1 : 1;
    if (x->track != y->track) return x->track < y->track ? -1 : 1;
    if (x->t != y->t) return x->t < y->t ? -1 : 1;
    return 0;
}

void aid_oracle_index_build(uint32_t *hash, uint32_t *track, uint32_t *t, int64_t n, uint64_t *bucket) {
    entry_t *e = (entry_t *)malloc((size_t)(n > 0 ? n : 1) * sizeof(entry_t));
    for (int64_t i = 0; i < n; i++) { e[i].hash = hash[i]; e[i].track = track[i]; e[i].t = t[i]; }
    qsort(e, (size_t)n, sizeof(entry_t), entry_cmp);
    int64_t nb = (int64_t)1 << AID_HASH_BITS, pos = 0;
    for (int64_t h = 0; h <= nb; h++) {
        while (pos < n && (int64_t)e[pos].hash < h) pos++;
        bucket[h] = (uint64_t)pos;
    }
    for (int64_t i = 0; i < n; i++) { hash[i] = e[i].hash; track[i] = e[i].track; t[i] = e[i].t; }
    free(e);
}

/* ------------------------------------------------------------------ stage 5: probe + vote
 * Row layout shared with the C ABI (include/audio_ident_b200.h aid_match_row). */
typedef struct {
    int32_t count;       /* aligned hashes for (track, offset) */
    uint32_t track;      /* global track number */
    int32_t offset;      /* t_ref - t_query, frames */
    int32_t q_first;     /* smallest / largest query anchor frame among the votes */
    int32_t q_last;
} orow_t;

typedef struct { uint64_t key; int32_t tq; } vote_t;

static int vote_cmp(const void *a, const void *b) {
    const vote_t *x = (const vote_t *)a, *y = (const vote_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->tq < y->tq ? -1 : (x->tq > y->tq);
}

static int row_cmp(const void *a, const void *b) {
    const orow_t *x = (const orow_t *)a, *y = (const orow_t *)b;
    if (x->count != y->count) return x->count > y->count ? -1 : 1;
    if (x->track != y->track) return x->track < y->track ? -1 : 1;
    if (x->offset != y->offset) return x->offset < y->offset ? -1 : 1;
    return 0;
}

/* One vote window. tombstone (may be NULL) has one byte per track, non-zero = deleted.
 * Returns the number of rows written (<= max_rows), ordered (count desc, track, offset). */
int aid_oracle_match(const uint32_t *ix_track, const uint32_t *ix_t, const uint64_t *bucket,
                     const uint8_t *tombstone,
                     const uint32_t *q_hash, const uint32_t *q_t, int64_t nq,
                     orow_t *rows, int max_rows) {
    int64_t total = 0;
    for (int64_t i = 0; i < nq; i++) total += (int64_t)(bucket[q_hash[i] + 1] - bucket[q_hash[i]]);
    vote_t *v = (vote_t *)malloc((size_t)(total > 0 ? total : 1) * sizeof(vote_t));
    int64_t nv = 0;
    for (int64_t i = 0; i < nq; i++)
        for (uint64_t p = bucket[q_hash[i]]; p < bucket[q_hash[i] + 1]; p++) {
            if (tombstone && tombstone[ix_track[p]]) continue;
            int64_t off = (int64_t)ix_t[p] - (int64_t)q_t[i] + AID_QUERY_MAX_FRAMES;
            v[nv].key = ((uint64_t)ix_track[p] << 32) | (uint64_t)off;
            v[nv].tq = (int32_t)q_t[i];
            nv++;
        }
    qsort(v, (size_t)nv, sizeof(vote_t), vote_cmp);
    int64_t ncand = 0, capc = 64;
    orow_t *cand = (orow_t *)malloc((size_t)capc * sizeof(orow_t));
    for (int64_t i = 0; i < nv;) {
        int64_t j = i;
        while (j < nv && v[j].key == v[i].key) j++;
        if (j - i >= AID_MIN_VOTES) {
            if (ncand == capc) { capc *= 2; cand = (orow_t *)realloc(cand, (size_t)capc * sizeof(orow_t)); }
            cand[ncand].count = (int32_t)(j - i);
            cand[ncand].track = (uint32_t)(v[i].key >> 32);
            cand[ncand].offset = (int32_t)((int64_t)(v[i].key & 0xffffffffu) - AID_QUERY_MAX_FRAMES);
            cand[ncand].q_first = v[i].tq;
            cand[ncand].q_last = v[j - 1].tq;
            ncand++;
        }
        i = j;
    }
    qsort(cand, (size_t)ncand, sizeof(orow_t), row_cmp);
    int n = ncand < max_rows ? (int)ncand : max_rows;
    memcpy(rows, cand, (size_t)n * sizeof(orow_t));
    free(cand); free(v);
    return n;
}

/* the constants, for Python-side cross-checks: fills out[16] */
void aid_oracle_params(int32_t *out) {
    out[0] = AID_SAMPLE_RATE; out[1] = AID_NFFT; out[2] = AID_HOP; out[3] = AID_NBINS;
    out[4] = AID_PEAK_HALF_F; out[5] = AID_PEAK_HALF_T; out[6] = AID_PEAK_MIN_BIN;
    out[7] = AID_DT_MIN; out[8] = AID_DT_MAX; out[9] = AID_DF_MIN; out[10] = AID_DF_MAX;
    out[11] = AID_FANOUT; out[12] = AID_MIN_VOTES; out[13] = AID_MAX_ROWS;
    out[14] = AID_QUERY_MAX_FRAMES; out[15] = AID_SEG_TRACKS;
}
