/* aid_cpu_f32.c -- the STATED CPU baseline of bench.py: the same specification as aid_oracle.c (include/aid_params.h),
 * written the way a production CPU engine would be -- single precision, SIMD across frames, one streaming pass per
 * track, no full-spectrogram intermediates. TEST / BENCH INFRASTRUCTURE ONLY (same rule as aid_oracle.c: only tests/,
 * smoke() and bench.py's cpu_baseline / --impl reference legs may load it; the product path never does).
 *
 * Why it exists ("reference arm hygiene"): aid_oracle.c is the CHECKER -- double-precision radix-2 FFT
 * and deque maxima over a materialised spectrogram, written for clarity, ~20 us per frame per core. The reference's
 * real engine (olaf_c, un-vendored: reference audio-ident-service/app/audio/fingerprint.py:117-125) bundles pffft, a
 * single-precision SIMD FFT (docs/plans/01-initial-implementation/00-plan-overview.md:237), and streams. A speed-up
 * quoted against the checker would flatter the GPU, so the number bench.py states next to it is this file's:
 *   - 16 frames at a time, structure-of-arrays, so every butterfly is one SIMD operation over 16 frames (gcc
 *     vectorises the inner loops; AVX-512 / AVX2 clones are chosen at load time);
 *   - the 1024-point real transform as a 512-point complex one + split, float throughout, logf from libmvec;
 *   - peaks in the same pass: van Herk row maxima over the 103-bin window, a 25-row ring of them, and only the row
 *     candidates (a handful per frame) are checked down their column 12 frames later;
 *   - hashes with aid_oracle_hashes (integer work, shared).
 * Results follow the specification, not the checker's rounding: the spectrogram is within AID_SPEC_TOL of the
 * double-precision one and the peaks are exactly aid_oracle_peaks of THIS spectrogram (tests/test_oracle.py).
 * PARITY UNPINNED against the reference for the same reason as aid_oracle.c.
 *
 * Build: oracle/Makefile (gcc -O3 -ffast-math -march=x86-64-v3 -fopenmp, target_clones for avx512f).
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
#define V 16                      /* frames per SIMD batch */
#define RING 32                   /* >= 2 * AID_PEAK_HALF_T + 1 rows, power of two */
#define WF (2 * AID_PEAK_HALF_F + 1)

#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define CLONES __attribute__((target_clones("avx512f", "default")))
#else
#define CLONES
#endif

int64_t aid_oracle_hashes(const uint32_t *keys, int64_t n_peaks, uint32_t *hash, uint32_t *t_anchor);

static float f_win[NF];
static float f_tw_re[HALF / 2], f_tw_im[HALF / 2];    /* e^{-2 pi i k / 512} */
static float f_pw_re[HALF], f_pw_im[HALF];            /* e^{-2 pi i k / 1024} */
static uint16_t f_rev[HALF];
static int f_ready = 0;

static void f32_tables(void) {
    if (f_ready) return;
    #pragma omp critical(aid_f32_tables)
    {
        if (!f_ready) {
            for (int n = 0; n < NF; n++)
                f_win[n] = (float)(AID_WIN_A0 - AID_WIN_A1 * cos(2.0 * M_PI * (double)n / (double)(NF - 1)));
            for (int k = 0; k < HALF / 2; k++) { f_tw_re[k] = (float)cos(-2.0 * M_PI * k / HALF); f_tw_im[k] = (float)sin(-2.0 * M_PI * k / HALF); }
            for (int k = 0; k < HALF; k++) { f_pw_re[k] = (float)cos(-2.0 * M_PI * k / NF); f_pw_im[k] = (float)sin(-2.0 * M_PI * k / NF); }
            for (int i = 0; i < HALF; i++) {
                int r = 0;
                for (int b = 0; b < 9; b++) if (i & (1 << b)) r |= 1 << (8 - b);
                f_rev[i] = (uint16_t)r;
            }
            f_ready = 1;
        }
    }
}

/* 16 floats = one frame batch; gcc lowers the arithmetic to one AVX-512 or two AVX2 operations per statement */
typedef float vf __attribute__((vector_size(4 * V), aligned(64)));
typedef struct { vf re[HALF], im[HALF]; float pre[NB][V] __attribute__((aligned(64))), suf[NB][V] __attribute__((aligned(64))); } fft_ws;
/* 16 consecutive frames, bin-major: s[k][v] = S of frame t0 + v at bin k, rm[k][v] = its 103-bin row maximum */
typedef struct { float s[NB][V] __attribute__((aligned(64))), rm[NB][V] __attribute__((aligned(64))); } batch_t;
#define NBATCH 4                  /* ring of batches: a decision at frame d looks at rows d - 12 .. d + 12 */

/* V frames starting at x (hop apart) -> out->s. n_valid <= V frames are real; the others repeat frame 0. */
CLONES static void spectra16(const float *x, int n_valid, fft_ws *ws, batch_t *out) {
    vf *re = ws->re, *im = ws->im;
    const float *xv[V];
    for (int v = 0; v < V; v++) xv[v] = x + (size_t)(v < n_valid ? v : 0) * AID_HOP;
    for (int i = 0; i < HALF; i++) {
        const int j = f_rev[i];
        vf a, b;
        for (int v = 0; v < V; v++) { a[v] = xv[v][2 * i]; b[v] = xv[v][2 * i + 1]; }
        re[j] = a * f_win[2 * i];
        im[j] = b * f_win[2 * i + 1];
    }
    /* stage 1 (w = 1) and stage 2 (w = 1, -i) without multiplies, then general radix-2 stages */
    for (int s = 0; s < HALF; s += 2) {
        const vf ar = re[s], ai = im[s], br = re[s + 1], bi = im[s + 1];
        re[s] = ar + br; im[s] = ai + bi; re[s + 1] = ar - br; im[s + 1] = ai - bi;
    }
    for (int s = 0; s < HALF; s += 4) {
        vf ar = re[s], ai = im[s], br = re[s + 2], bi = im[s + 2];
        re[s] = ar + br; im[s] = ai + bi; re[s + 2] = ar - br; im[s + 2] = ai - bi;
        ar = re[s + 1]; ai = im[s + 1]; br = re[s + 3]; bi = im[s + 3];       /* w = -i: (br, bi) -> (bi, -br) */
        re[s + 1] = ar + bi; im[s + 1] = ai - br; re[s + 3] = ar - bi; im[s + 3] = ai + br;
    }
    for (int len = 8; len <= HALF; len <<= 1) {
        const int half = len >> 1, step = HALF / len;
        for (int s = 0; s < HALF; s += len)
            for (int k = 0; k < half; k++) {
                const float wr = f_tw_re[k * step], wi = f_tw_im[k * step];
                const vf rb = re[s + k + half], ib = im[s + k + half], ra = re[s + k], ia = im[s + k];
                const vf tr = rb * wr - ib * wi, ti = rb * wi + ib * wr;
                re[s + k + half] = ra - tr; im[s + k + half] = ia - ti;
                re[s + k] = ra + tr; im[s + k] = ia + ti;
            }
    }
    for (int k = 0; k < NB; k++) {
        const int m = (HALF - k) & (HALF - 1);
        const float cr = f_pw_re[k], ci = f_pw_im[k];
        const vf er = 0.5f * (re[k] + re[m]), ei = 0.5f * (im[k] - im[m]);
        const vf or_ = 0.5f * (im[k] + im[m]), oi = -0.5f * (re[k] - re[m]);
        const vf xr = er + or_ * cr - oi * ci, xi = ei + or_ * ci + oi * cr;
        const vf pw = 1.0f + (xr * xr + xi * xi);
        for (int v = 0; v < V; v++) out->s[k][v] = pw[v];
    }
    {   /* one flat loop gcc hands to libmvec's SIMD logf */
        float *p = &out->s[0][0];
        for (int i = 0; i < NB * V; i++) p[i] = logf(p[i]);
    }
}

/* clipped sliding maximum over [k - 51, k + 51] along the bins, 16 frames at once (van Herk / Gil-Werman: block prefix
 * and suffix maxima; a window spans at most two blocks of 103) */
CLONES static void row_max16(fft_ws *ws, batch_t *b) {
    float (*pre)[V] = ws->pre, (*suf)[V] = ws->suf;
    for (int b0 = 0; b0 < NB; b0 += WF) {
        const int e = b0 + WF < NB ? b0 + WF : NB;
        float m[V] __attribute__((aligned(64)));
        for (int v = 0; v < V; v++) m[v] = -1.0f;                 /* S >= 0: -1 never wins */
        for (int i = b0; i < e; i++)
            for (int v = 0; v < V; v++) { m[v] = b->s[i][v] > m[v] ? b->s[i][v] : m[v]; pre[i][v] = m[v]; }
        for (int v = 0; v < V; v++) m[v] = -1.0f;
        for (int i = e - 1; i >= b0; i--)
            for (int v = 0; v < V; v++) { m[v] = b->s[i][v] > m[v] ? b->s[i][v] : m[v]; suf[i][v] = m[v]; }
    }
    for (int k = 0; k < NB; k++) {
        const int lo = k - AID_PEAK_HALF_F < 0 ? 0 : k - AID_PEAK_HALF_F;
        const int hi = k + AID_PEAK_HALF_F > NB - 1 ? NB - 1 : k + AID_PEAK_HALF_F;
        if (lo / WF != hi / WF) { for (int v = 0; v < V; v++) b->rm[k][v] = suf[lo][v] > pre[hi][v] ? suf[lo][v] : pre[hi][v]; }
        else if (lo % WF == 0) { for (int v = 0; v < V; v++) b->rm[k][v] = pre[hi][v]; }                            /* clipped low */
        else if (hi == NB - 1 || hi % WF == WF - 1) { for (int v = 0; v < V; v++) b->rm[k][v] = suf[lo][v]; }       /* clipped high */
        else {
            for (int v = 0; v < V; v++) b->rm[k][v] = -1.0f;
            for (int i = lo; i <= hi; i++)
                for (int v = 0; v < V; v++) b->rm[k][v] = b->s[i][v] > b->rm[k][v] ? b->s[i][v] : b->rm[k][v];
        }
    }
}

typedef struct { uint16_t f; float v; } cand_t;

/* One track, streaming in batches of 16 frames. keys (capacity AID_PEAK_CAP(T)) in (t, f) order; returns the peak
 * count or -1 (capacity rule). S_out (optional, tests) receives the spectrogram. */
static int64_t f32_track_peaks(const float *pcm, int64_t n_samples, uint32_t *keys, float *S_out) {
    const int64_t T = n_samples < NF ? 0 : (n_samples - NF) / AID_HOP + 1;
    if (T == 0) return 0;
    fft_ws *ws = (fft_ws *)aligned_alloc(64, sizeof(fft_ws));
    batch_t *ring = (batch_t *)aligned_alloc(64, NBATCH * sizeof(batch_t));
    cand_t (*cand)[NB] = (cand_t (*)[NB])malloc((size_t)NBATCH * V * sizeof(cand_t[NB]));   /* row candidates per frame slot */
    int *ncand = (int *)calloc(NBATCH * V, sizeof(int));
    int64_t n = 0, in_block = 0, decided = 0;      /* frames < decided have been settled */
    int failed = 0;
    for (int64_t t0 = 0; t0 < T && (!failed || S_out); t0 += V) {       /* a failed track still yields its spectrogram */
        const int nv = (int)(T - t0 < V ? T - t0 : V);
        batch_t *bt = ring + (t0 / V) % NBATCH;
        spectra16(pcm + t0 * AID_HOP, nv, ws, bt);
        row_max16(ws, bt);
        if (S_out)
            for (int v = 0; v < nv; v++)
                for (int k = 0; k < NB; k++) S_out[(size_t)(t0 + v) * NB + k] = bt->s[k][v];
        if (failed) continue;
        /* row candidates: value passes the gate and equals its row maximum (a handful per frame) */
        for (int v = 0; v < nv; v++) ncand[(t0 + v) % (NBATCH * V)] = 0;
        for (int k = AID_PEAK_MIN_BIN; k < NB; k++) {
            int any = 0;
            for (int v = 0; v < V; v++) any |= (bt->s[k][v] == bt->rm[k][v]) & (bt->s[k][v] > AID_PEAK_MIN_S);
            if (!any) continue;
            for (int v = 0; v < nv; v++)
                if (bt->s[k][v] == bt->rm[k][v] && bt->s[k][v] > AID_PEAK_MIN_S) {
                    const int slot = (int)((t0 + v) % (NBATCH * V));
                    cand[slot][ncand[slot]].f = (uint16_t)k; cand[slot][ncand[slot]].v = bt->s[k][v];
                    ncand[slot]++;
                }
        }
        /* frame d has its whole (clipped) 25-row window once row min(d + 12, T - 1) exists */
        const int64_t last = t0 + nv - 1;
        const int64_t d_hi = last == T - 1 ? T - 1 : last - AID_PEAK_HALF_T;
        for (; decided <= d_hi; decided++) {
            const int64_t d = decided;
            if (d % AID_PEAK_BLOCK_FRAMES == 0) in_block = 0;
            const int64_t lo = d - AID_PEAK_HALF_T < 0 ? 0 : d - AID_PEAK_HALF_T;
            const int64_t hi = d + AID_PEAK_HALF_T > T - 1 ? T - 1 : d + AID_PEAK_HALF_T;
            const int slot = (int)(d % (NBATCH * V));
            for (int c = 0; c < ncand[slot]; c++) {
                const int f = cand[slot][c].f;
                const float val = cand[slot][c].v;
                int ok = 1;
                for (int64_t r = lo; r <= hi && ok; r++) ok = ring[(r / V) % NBATCH].rm[f][r % V] <= val;
                if (ok) {
                    if (++in_block > AID_PEAK_BLOCK_CAP) { failed = 1; break; }
                    keys[n++] = ((uint32_t)d << AID_PEAK_F_BITS) | (uint32_t)f;
                }
            }
            if (failed) break;
        }
    }
    free(ws); free(ring); free(cand); free(ncand);
    return failed ? -1 : n;
}

int64_t aid_cpu_f32_stft(const float *pcm, int64_t n_samples, float *S) {
    f32_tables();
    const int64_t T = n_samples < NF ? 0 : (n_samples - NF) / AID_HOP + 1;
    if (T == 0) return 0;
    uint32_t *keys = (uint32_t *)malloc((size_t)AID_PEAK_CAP(T) * sizeof(uint32_t));
    f32_track_peaks(pcm, n_samples, keys, S);
    free(keys);
    return T;
}

int64_t aid_cpu_f32_peaks_from_pcm(const float *pcm, int64_t n_samples, uint32_t *keys) {
    f32_tables();
    return f32_track_peaks(pcm, n_samples, keys, NULL);
}

/* Same contract as aid_oracle_fingerprint_batch: one track per OpenMP task, hash_off[i] = track i's first slot,
 * n_hash[i] = its count (-1 on a capacity failure). Returns the number of threads used. */
int aid_cpu_f32_fingerprint_batch(const float *pcm, const int64_t *sample_off, int n_tracks,
                                  uint32_t *hash, uint32_t *t_anchor, const int64_t *hash_off,
                                  int64_t *n_hash, int64_t *n_peaks, int threads) {
    f32_tables();
    int used = 1;
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
    used = omp_get_max_threads();
#endif
    #pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n_tracks; i++) {
        const int64_t ns = sample_off[i + 1] - sample_off[i];
        const int64_t T = ns < NF ? 0 : (ns - NF) / AID_HOP + 1;
        n_hash[i] = 0;
        if (n_peaks) n_peaks[i] = 0;
        if (T == 0) continue;
        uint32_t *keys = (uint32_t *)malloc((size_t)AID_PEAK_CAP(T) * sizeof(uint32_t));
        const int64_t np = f32_track_peaks(pcm + sample_off[i], ns, keys, NULL);
        if (n_peaks) n_peaks[i] = np;
        n_hash[i] = np < 0 ? -1 : aid_oracle_hashes(keys, np, hash + hash_off[i], t_anchor + hash_off[i]);
        free(keys);
    }
    return used;
}
