/* aid_params.h -- the one place the fingerprint engine's algorithm constants live.
 *
 * Shared by the CPU oracle (oracle/aid_oracle.c), the sm_100a kernels
 * (audio_ident_b200/csrc/) and, through aid_get_params(), the Python host code.
 *
 * Provenance: the reference service sets none of these. It shells out to an
 * un-vendored `olaf_c` binary (reference audio-ident-service/app/audio/fingerprint.py:117-125,
 * :185-193) that runs on compiled-in defaults, so the only constants the reference
 * itself confirms are the input format (fingerprint.py:10: 16 kHz mono f32le) and
 * the time unit implied by exact.py:58-62. Everything else below is this repo's
 * own specification ("working parameters"); parity for the
 * stages that use them is therefore "unpinned by the reference" and is judged
 * against oracle/aid_oracle.c.
 */
#ifndef AID_PARAMS_H
#define AID_PARAMS_H

/* ---- stage 1: framing + STFT ------------------------------------------------ */
#define AID_SAMPLE_RATE   16000   /* Hz; reference fingerprint.py:10, decode.py:83-86 */
#define AID_NFFT          1024    /* samples per frame */
#define AID_HOP           128     /* samples between frame starts -> 125 frames/s, 8 ms */
#define AID_NBINS         512     /* stored bins k = 0..511 (Nyquist bin dropped) */
/* window: symmetric Hamming, w[n] = 0.54 - 0.46*cos(2*pi*n/(AID_NFFT-1)), evaluated in
 * double and rounded to float32 once; both oracle and kernels consume that float table. */
#define AID_WIN_A0        0.54
#define AID_WIN_A1        0.46
/* stored value: S = log(1 + re^2 + im^2) as float32 ("log-magnitude", conditioned so an
 * fp32 FFT meets |dS| <= AID_SPEC_TOL * max(|S|, 1) against the double-precision oracle). */
#define AID_SPEC_TOL      1e-4f

/* ---- stage 2: constellation peaks ------------------------------------------- */
#define AID_PEAK_HALF_F   51      /* neighbourhood = 103 bins ...                     */
#define AID_PEAK_HALF_T   12      /* ... x 25 frames, clipped at the spectrogram edge */
#define AID_PEAK_MIN_BIN  9       /* bins below this never become peaks               */
#define AID_PEAK_MIN_S    0.001f  /* S must be strictly greater than this             */
/* a point is a peak iff it passes the two gates above and S equals the maximum of its
 * clipped neighbourhood (all members of an exact tie are peaks).
 * Capacity rule (it only binds on tie-heavy degenerate input, and fails the track): every aligned block of
 * AID_PEAK_BLOCK_FRAMES frames [256b, 256b+256) may hold at most AID_PEAK_BLOCK_CAP peaks. */
#define AID_PEAK_BLOCK_FRAMES  256
#define AID_PEAK_BLOCK_CAP     2048
#define AID_PEAK_CAP(frames) ((((long long)(frames) + AID_PEAK_BLOCK_FRAMES - 1) / AID_PEAK_BLOCK_FRAMES) * AID_PEAK_BLOCK_CAP)
/* peaks are ordered by (t, f); packed as key = (t << 9) | f */
#define AID_PEAK_F_BITS   9
#define AID_MAX_FRAMES    (1 << 22)   /* 4.19 M frames = 9.3 h per track or query */

/* ---- stage 3: landmark pairs -------------------------------------------------- */
#define AID_DT_MIN        2       /* frames between anchor and target ...  */
#define AID_DT_MAX        33
#define AID_DF_MIN        1       /* |f_target - f_anchor| in bins ...      */
#define AID_DF_MAX        128
#define AID_FANOUT        8       /* first AID_FANOUT qualifying targets in (t, f) order */
/* hash = (f_anchor << 15) | (f_target << 6) | dt   -> 24 bits */
#define AID_HASH_BITS     24
#define AID_HASH(f1, f2, dt) ((((unsigned)(f1)) << 15) | (((unsigned)(f2)) << 6) | ((unsigned)(dt)))

/* ---- stage 4/5: index + vote --------------------------------------------------- */
/* The index is a list of segments of at most AID_SEG_TRACKS tracks. Inside a segment a
 * posting is one u32: (local_track << AID_POST_T_BITS) | t_anchor; postings are ordered
 * by (hash, local_track, t_anchor). A vote key is posting + (AID_QUERY_MAX_FRAMES - t_query),
 * i.e. (local_track, t_ref - t_query + AID_QUERY_MAX_FRAMES) in one u32 with no carry. */
#define AID_SEG_TRACK_BITS 14
#define AID_SEG_TRACKS     (1 << AID_SEG_TRACK_BITS)
#define AID_POST_T_BITS    18
#define AID_QUERY_MAX_FRAMES 32768                 /* one vote window: <= 262 s of query audio */
#define AID_INDEX_MAX_FRAMES ((1 << AID_POST_T_BITS) - AID_QUERY_MAX_FRAMES)  /* 229,376 frames = 30.6 min;
                                                       the reference refuses ingest above 30 min (pipeline.py:41) */
#define AID_MIN_VOTES      6      /* a (track, offset) row needs at least this many aligned hashes */
#define AID_MAX_ROWS       50     /* rows returned per query, ordered by (count desc, track asc, offset asc) */

#define AID_FRAME_SECONDS  ((double)AID_HOP / (double)AID_SAMPLE_RATE)   /* 0.008 */

#endif /* AID_PARAMS_H */
