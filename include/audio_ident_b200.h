/* audio_ident_b200.h -- C ABI of the B200 fingerprint-and-match engine (libaudioident_b200.so).
 *
 * This is the drop-in boundary for the reference's hot path. The reference has no FFI for it:
 * audio-ident-service/app/audio/fingerprint.py reaches the engine by writing a temp file and
 * executing `olaf_c {store|query|del}` (fingerprint.py:117-125, :185-193, :239-246) with the index
 * directory in env OLAF_DB (:79-84). Each entry point below names the call it replaces; the Python
 * binding a maintainer would add is audio_ident_b200/_lib.py (ctypes) and is shown in INTEGRATION.md.
 *
 * Conventions: plain pointers and sizes only; every function returns AID_OK (0) or a negative
 * aid_status and never throws; host output buffers are caller-allocated with an explicit capacity;
 * `stream` arguments are a cudaStream_t passed as void* (NULL = the engine's own stream).
 * An engine is bound to one CUDA device; calls on one engine must be serialised by the caller
 * (the reference serialises writers the same way: routers/ingest.py:52, pipeline.py:294).
 * There is no CPU fallback: without a usable CUDA device aid_engine_create fails.
 */
#ifndef AUDIO_IDENT_B200_H
#define AUDIO_IDENT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AID_ABI_VERSION 1

typedef enum {
    AID_OK = 0,
    AID_E_CUDA = -1,        /* a CUDA call failed; aid_last_error() has the text */
    AID_E_ARG = -2,         /* bad argument (null pointer, negative size, unsorted offsets ...) */
    AID_E_CAPACITY = -3,    /* a caller-provided output buffer is too small */
    AID_E_TOO_LONG = -4,    /* a track or query exceeds the frame limits of aid_params.h */
    AID_E_NOT_FOUND = -5,   /* unknown track name */
    AID_E_IO = -6,          /* index directory could not be read or written */
    AID_E_FORMAT = -7,      /* index files are not ours / wrong version */
    AID_E_FULL = -8,        /* index limits reached */
    AID_E_TIMEOUT = -9      /* sharded identification: a peer rank did not deliver its rows in time */
} aid_status;

/* per-track status bits written by the fingerprint stages */
#define AID_TRACK_OK             0
#define AID_TRACK_PEAK_OVERFLOW  1   /* a capacity rule of aid_params.h was broken (tie-heavy input) */
#define AID_TRACK_TOO_LONG       2
#define AID_TRACK_EMPTY          4   /* shorter than one frame: no fingerprints (not an error) */

typedef struct aid_engine aid_engine;

/* One (track, offset) alignment found by a query; what one CSV line of `olaf_c query` carries
 * (reference fingerprint.py:273-277: count, q_start, q_stop, name, id, ref_start, ref_stop).
 * Times are frames of AID_FRAME_SECONDS; ref_start = q_first + offset, ref_stop = q_last + offset. */
typedef struct {
    int32_t  count;     /* aligned hashes */
    uint32_t track;     /* engine-wide track number (the reference's "reference_id") */
    int32_t  offset;    /* t_ref - t_query in frames */
    int32_t  q_first;   /* first / last query anchor frame among the aligned hashes */
    int32_t  q_last;
} aid_match_row;

/* ---- library ------------------------------------------------------------------------------- */
int         aid_abi_version(void);
const char* aid_strerror(int status);
/* constants of include/aid_params.h, in the order of oracle/oracle.py PARAM_NAMES; out[16] */
void        aid_get_params(int32_t* out);
/* number of CUDA devices visible, or a negative aid_status */
int         aid_device_count(void);

/* ---- engine -------------------------------------------------------------------------------- */
int         aid_engine_create(int device, aid_engine** out);
void        aid_engine_destroy(aid_engine* e);
const char* aid_last_error(const aid_engine* e);
/* kernels launched by this engine since creation (bench.py's gpu_launches) */
int64_t     aid_launch_count(const aid_engine* e);
/* blocks until the engine's stream is idle */
int         aid_engine_sync(aid_engine* e);
/* upper bound on frames processed per internal sub-batch (workspace is about 2.4 KB per frame) */
int         aid_engine_set_max_batch_frames(aid_engine* e, int64_t frames);
/* Per-stage device timing with CUDA events on the launching stream (bench.py's roofline numbers).
 * Stages: 0 = STFT kernel, 1 = peak kernel, 2 = peak compaction (scan + copy), 3 = hasher (count + scan +
 * write), 4 = matcher (k_match), 5 = ranking (k_rank, with its peer stores when sharded), 6 = index build,
 * 7 = merge of the ranks' row blocks (k_merge_blocks, including the wait for the slowest rank).
 * aid_engine_stage_times waits for the recorded work, adds the elapsed milliseconds and the number
 * of timed launches per stage into ms[8] / launches[8], and forgets the records. */
int         aid_engine_set_stage_timing(aid_engine* e, int on);
int         aid_engine_stage_times(aid_engine* e, double* ms, int64_t* launches);
/* Kernel selection for tests and A/B measurements (results are bit-identical for every choice; the defaults are the
 * product path). stft_variant: 0 = the scalar FP32 STFT kernel of round 1, 5 = the packed (f32x2) kernel (default;
 * 7 = the same with the separation software-pipelined into the next trip: faster without the group maxima, slower with them; 16 = producer / consumer warps, measured slower).
 * peak_summary != 0 (default): the STFT kernel also emits the maxima of the 32 aligned 16-bin groups of every row
 * and the peak kernel streams those 128 B per row instead of re-reading the 2 KB spectrogram row (packed kernel only). */
int         aid_engine_set_kernels(aid_engine* e, int stft_variant, int peak_summary);

/* ---- fingerprinting: PCM -> (hash, t_anchor) ----------------------------------------------------
 * Replaces the analysis half of `olaf_c store` / `olaf_c query` (fingerprint.py:117-125, :185-193).
 * A batch is ragged: track i is pcm[sample_off[i] .. sample_off[i+1]) (16 kHz mono float32).
 * Outputs are dense: track i owns hash/t_anchor[hash_off[i] .. hash_off[i+1]) in (anchor, target)
 * order; status[i] carries AID_TRACK_* bits (a failed track has zero hashes). */

/* host buffers in, host buffers out (H2D and D2H inside; pinned memory makes them asynchronous) */
int aid_fingerprint_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                         uint32_t* hash, uint32_t* t_anchor, int64_t hash_cap,
                         int64_t* hash_off /* [n_tracks+1] */, int32_t* status /* [n_tracks] */);

/* device-resident PCM in; results stay on the device in engine-owned buffers that are valid until
 * the next call on this engine. Asynchronous on `stream`. Whole batch must fit one sub-batch. */
typedef struct {
    const uint32_t* d_hash;       /* [n_hash] */
    const uint32_t* d_t_anchor;   /* [n_hash] */
    const uint32_t* d_hash_off;   /* [n_tracks+1] */
    const uint32_t* d_peaks;      /* [n_peaks] keys (t << 9 | f), (track, t, f) order */
    const uint32_t* d_peak_off;   /* [n_tracks+1] */
    const int32_t*  d_status;     /* [n_tracks] */
    const float*    d_spec;       /* [total_frames][512] */
    int64_t         total_frames;
} aid_fp_device_result;

int aid_fingerprint_dev(aid_engine* e, const float* d_pcm, const int64_t* sample_off /* host */,
                        int n_tracks, aid_fp_device_result* out, void* stream);
/* Window form: item i is d_pcm[win_begin[i] .. win_end[i]) -- the ranges may overlap or leave gaps, so the three
 * sub-windows of a 5 s query clip (exact.py:48-52) or the sliding windows of a long recording are fingerprinted
 * where they lie instead of being copied out first. */
int aid_fingerprint_windows_dev(aid_engine* e, const float* d_pcm, const int64_t* win_begin /* host */,
                                const int64_t* win_end /* host */, int n_windows, aid_fp_device_result* out, void* stream);

/* single stages on host buffers (parity tests and tools; same kernels as above) */
int aid_stft_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                  float* spec /* [sum frames][512] */);
int aid_peaks_host(aid_engine* e, const float* spec, const int64_t* frame_off /* [n+1] */, int n_tracks,
                   uint32_t* peaks, int64_t peak_cap, int64_t* peak_off /* [n+1] */, int32_t* status);
int aid_hashes_host(aid_engine* e, const uint32_t* peaks, const int64_t* peak_off /* [n+1] */, int n_tracks,
                    uint32_t* hash, uint32_t* t_anchor, int64_t hash_cap, int64_t* hash_off /* [n+1] */);
/* frames a clip of n_samples yields */
int64_t aid_num_frames(int64_t n_samples);

/* ---- index: `olaf_c store` / `olaf_c del` (fingerprint.py:117-125, :239-246) -------------------- */
/* Fingerprints the batch and adds each track under names[i] (the reference passes str(uuid),
 * fingerprint.py:108). ok[i] = 1 if stored. A name that is already present is replaced. */
int aid_index_add_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                       const char* const* names, uint8_t* ok);
/* Same, and the fingerprints that were stored come back in host buffers (track i owns hash/t_anchor[hash_off[i] ..
 * hash_off[i+1]); a track with ok[i] = 0 may own entries, ignore them): the service journals them for persistence
 * without fingerprinting twice or sending them back up (audio_ident_b200/fingerprint.py). */
int aid_index_add_host_fp(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                          const char* const* names, uint8_t* ok, uint32_t* hash, uint32_t* t_anchor,
                          int64_t hash_cap, int64_t* hash_off /* [n_tracks+1] */);
/* Same as aid_index_add_host, PCM already on the device (bulk ingest keeps the PCIe copy out of the way). */
int aid_index_add_dev(aid_engine* e, const float* d_pcm, const int64_t* sample_off, int n_tracks,
                      const char* const* names, uint8_t* ok);
/* Adds precomputed fingerprints (hash, t_anchor per track, dense with hash_off) -- used to merge
 * what other ranks fingerprinted and by tests. */
int aid_index_add_hashes(aid_engine* e, const uint32_t* hash, const uint32_t* t_anchor,
                         const int64_t* hash_off, const int64_t* n_frames, int n_tracks,
                         const char* const* names, uint8_t* ok);
int aid_index_delete(aid_engine* e, const char* name);
/* makes every added track searchable now (otherwise done lazily by the next query) */
int aid_index_commit(aid_engine* e);
int aid_index_clear(aid_engine* e);
/* Tracks live in segments of AID_SEG_TRACKS (16,384), the unit of incremental build and of vote parallelism. Full
 * segments are grouped eight at a time under one hash directory (one 32-byte entry per hash: start + eight 16-bit run
 * lengths) so that the eight CTAs probing them for a window share every directory sector and posting sector. Rows do
 * not depend on the grouping; on = 1 is the default, 0 keeps one table per segment (tests, A/B measurements). */
int aid_index_set_grouping(aid_engine* e, int on);
/* out[8]: [0] live tracks, [1] postings, [2] segments, [3] tracks incl. deleted, [4] device bytes held by the index,
 * [5] segments that share a group directory */
int aid_index_stats(aid_engine* e, int64_t* out);
/* name of track number `track`; returns its length or a negative aid_status */
int aid_index_track_name(aid_engine* e, uint32_t track, char* buf, int buf_len);
/* persistence in the directory the reference calls settings.olaf_lmdb_path (fingerprint.py:79-84) */
int aid_index_save(aid_engine* e, const char* dir);
int aid_index_load(aid_engine* e, const char* dir);

/* ---- identification: `olaf_c query` (fingerprint.py:185-193) ---------------------------------------
 * Each query i is one vote window: pcm[sample_off[i] .. sample_off[i+1]). rows holds max_rows entries
 * per query (query i at rows + i*max_rows), n_rows[i] of them valid, ordered by
 * (count desc, track asc, offset asc). max_rows <= AID_MAX_ROWS. */
int aid_query_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_queries,
                   aid_match_row* rows, int max_rows, int32_t* n_rows);
int aid_query_dev(aid_engine* e, const float* d_pcm, const int64_t* sample_off, int n_queries,
                  aid_match_row* rows, int max_rows, int32_t* n_rows);
/* Window form of aid_query_host: window i is pcm[win_begin[i] .. win_end[i]); the ranges may overlap -- the exact
 * lane's three 3.5 s sub-windows of a 5 s clip (exact.py:48-52, :150-171) are described by offsets into the clip, which is
 * uploaded once. */
int aid_query_windows_host(aid_engine* e, const float* pcm, const int64_t* win_begin, const int64_t* win_end,
                           int n_windows, aid_match_row* rows, int max_rows, int32_t* n_rows);
/* query with precomputed fingerprints (dense, hash_off per query) */
int aid_query_hashes(aid_engine* e, const uint32_t* hash, const uint32_t* t_anchor, const int64_t* hash_off,
                     int n_queries, aid_match_row* rows, int max_rows, int32_t* n_rows);

/* device in, device out, asynchronous on `stream`: window q owns d_hash/d_t_anchor[d_hash_off[q] .. +len) with
 * len = d_hash_len[q] if d_hash_len is given, else d_hash_off[q+1] - d_hash_off[q] (u32 offsets, as produced by
 * aid_fingerprint_dev); d_status may be NULL; d_rows[n_queries][max_rows], d_n_rows[n_queries] are caller-allocated
 * device buffers. This is what the multi-GPU path uses between the NCCL exchanges (no host round trip). */
int aid_match_dev(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t_anchor, const uint32_t* d_hash_off,
                  const uint32_t* d_hash_len, const int32_t* d_status, int n_queries,
                  aid_match_row* d_rows, int max_rows, int32_t* d_n_rows, void* stream);
/* Precondition of aid_match_dev / aid_match_exchange_dev (the fingerprints are on the device, so it cannot be
 * checked here): every t_anchor < AID_QUERY_MAX_FRAMES and every hash < 2^AID_HASH_BITS, i.e. the fingerprints
 * of vote windows of at most AID_QUERY_MAX_FRAMES frames. The entry points that see the PCM or host fingerprints
 * (aid_query_host/dev, aid_query_hashes, aid_identify_exchange_dev/host) check it and return AID_E_TOO_LONG / AID_E_ARG. */

/* Probe statistics of the matcher kernel since the last call, collected while stage timing is on:
 * out[0] = query hashes looked up (one directory / table entry each, per segment), out[1] = postings touched.
 * bench.py's roofline numerator for k_match. Synchronises the device. */
int aid_match_stats(aid_engine* e, int64_t* out /* [2] */);

/* ---- sharded identification: ranking fused with the row exchange over peer memory ----
 * The reference has one index, so one `olaf_c query` (fingerprint.py:185-193) sees every track. With the index
 * sharded over the GPUs of a box (one process per GPU, each with its own engine) every rank probes its shard for
 * the same windows and the rows have to meet. An aid_exchange is this rank's receive window in its own HBM;
 * after the ranks have swapped the 64-byte handles (any transport: torch.distributed, a pipe ...) each rank's
 * ranking kernel stores its rows, renumbered to global track numbers, straight into every rank's window over
 * NVLink and a merge kernel on each rank waits for all blocks on the device and orders their union by
 * (count desc, global track asc, offset asc), keeping max_rows: the same rows on every rank, identical to one
 * unsharded index. No collective call and no host synchronisation.
 * All ranks must call aid_match_exchange_dev with the same window batch, the same number of times, and each rank
 * always on the same stream. AID_MAX_RANKS = GPUs of one NVSwitch box. */
#define AID_MAX_RANKS 8
#define AID_IPC_HANDLE_BYTES 64
typedef struct aid_exchange aid_exchange;
/* max_queries: windows per call; max_hashes_per_rank: room for the query fingerprints one rank contributes per call
 * (0 = 1024 per window of a rank's slice; a 3.5 s window yields about 300) */
int  aid_exchange_create(aid_engine* e, int rank, int world, int max_queries, int64_t max_hashes_per_rank,
                         aid_exchange** out);
/* waits for the device, unmaps the peers' windows and frees this rank's; call it before aid_engine_destroy of its engine
 * and only after every rank has finished its last exchange call (peers may still be storing into the window) */
void aid_exchange_destroy(aid_exchange* x);
/* handle[64] of this rank's window for ranks in other processes (a cudaIpcMemHandle_t) */
int  aid_exchange_handle(aid_exchange* x, uint8_t* handle);
/* handles[world][64] of all ranks in rank order (this rank's own entry is ignored) */
int  aid_exchange_connect(aid_exchange* x, const uint8_t* handles);
/* ranks living in this process (one process driving several engines; tests): peers[world] */
int  aid_exchange_connect_local(aid_exchange* x, aid_exchange* const* peers);
/* how long the merge kernel waits for the slowest rank before it gives up (default 20 s) */
int  aid_exchange_set_timeout_ms(aid_exchange* x, int64_t ms);
/* AID_OK; AID_E_TIMEOUT if a kernel gave up waiting for a peer (n_rows of the affected windows are -1; after a
 * failed wait for the peers' fingerprints nothing is probed and every later step reports -1 rows until this call);
 * AID_E_CAPACITY if a rank's fingerprints did not fit max_hashes_per_rank (its windows matched nothing).
 * Synchronises; reports a failure once and clears it. */
int  aid_exchange_status(aid_exchange* x);
/* aid_match_dev + exchange + merge. d_track_map[n_map] (device, may be NULL) maps this engine's track numbers to
 * global ones; d_rows[n_queries][max_rows] / d_n_rows[n_queries] receive the merged rows. Asynchronous on stream. */
int  aid_match_exchange_dev(aid_engine* e, aid_exchange* x, const uint32_t* d_hash, const uint32_t* d_t_anchor,
                            const uint32_t* d_hash_off, const uint32_t* d_hash_len, const int32_t* d_status,
                            int n_queries, const uint32_t* d_track_map, int64_t n_map, aid_match_row* d_rows,
                            int max_rows, int32_t* d_n_rows, void* stream);

/* The whole sharded step with device-resident query PCM: the batch is sample_off[0..n_windows] (host, identical on
 * every rank); rank r fingerprints windows [r*n/world, (r+1)*n/world), stores the fingerprints into every rank's
 * window, waits on the device for the other slices, probes its shard for ALL windows, and exchanges + merges the rows
 * as aid_match_exchange_dev does. Replaces two NCCL all-gathers, the host synchronisation that sized them and the
 * torch merge of the earlier path. */
int  aid_identify_exchange_dev(aid_engine* e, aid_exchange* x, const float* d_pcm, const int64_t* sample_off,
                               int n_windows, const uint32_t* d_track_map, int64_t n_map, aid_match_row* d_rows,
                               int max_rows, int32_t* d_n_rows, void* stream);

/* Window forms of the two calls above: window i is pcm[win_begin[i] .. win_end[i]), overlaps allowed. The host form
 * copies the span its slice of windows covers ONCE (a 5 s clip instead of three 3.5 s windows: 2.1x less PCIe traffic). */
int  aid_identify_exchange_windows_dev(aid_engine* e, aid_exchange* x, const float* d_pcm, const int64_t* win_begin,
                                       const int64_t* win_end, int n_windows, const uint32_t* d_track_map, int64_t n_map,
                                       aid_match_row* d_rows, int max_rows, int32_t* d_n_rows, void* stream);
int  aid_identify_exchange_windows_host(aid_engine* e, aid_exchange* x, const float* pcm, const int64_t* win_begin,
                                        const int64_t* win_end, int n_windows, const uint32_t* d_track_map, int64_t n_map,
                                        int rows_first, int rows_count, aid_match_row* rows, int max_rows, int32_t* n_rows);

/* The same step from HOST buffers (what a service process holds: the windows' PCM as passed to olaf_query,
 * fingerprint.py:158-183; pinned memory makes the copies asynchronous): only this rank's slice of the batch crosses
 * PCIe (rank r fingerprints windows [r*n/world, (r+1)*n/world)), the merged rows of windows
 * [rows_first, rows_first + rows_count) are copied back into rows[rows_count][max_rows] / n_rows[rows_count].
 * Synchronous. Returns AID_E_TIMEOUT if a peer did not deliver (n_rows of the affected windows are -1).
 * With world == 1 this is aid_query_host with the exchange's buffers. */
int  aid_identify_exchange_host(aid_engine* e, aid_exchange* x, const float* pcm, const int64_t* sample_off,
                                int n_windows, const uint32_t* d_track_map, int64_t n_map,
                                int rows_first, int rows_count, aid_match_row* rows, int max_rows, int32_t* n_rows);

/* ---- content-duplicate scan ----------------------------------------------
 * Replaces the per-row Python loop of audio-ident-service/app/audio/dedup.py:170-222 (check_content_duplicate)
 * and its inner _fingerprint_similarity (:127-167). A store keeps the raw Chromaprint fingerprints of every
 * ingested track resident in HBM: row r = words[off[r] .. off[r+1]) (each word one 32-bit sub-fingerprint,
 * the integers of `fpcalc -raw` taken modulo 2^32 as dedup.py:158 does) with its chromaprint_duration.
 * A scan answers, per query, the reference's loop: among rows with q_lo <= duration <= q_hi (the caller
 * passes duration*0.9 and duration*1.1 computed in double, dedup.py:189-190) the row with the greatest
 * similarity, the first such row on ties; best_row = -1 and best_sim = 0.0 if no row has similarity > 0.
 * best_sim is the IEEE double the reference's Python computes, bit for bit. The threshold test
 * (dedup.py:214) stays with the caller. Calls on one store must be serialised by the caller. */
typedef struct aid_dedup aid_dedup;
int         aid_dedup_create(int device, aid_dedup** out);
void        aid_dedup_destroy(aid_dedup* d);
const char* aid_dedup_last_error(const aid_dedup* d);
int64_t     aid_dedup_size(const aid_dedup* d);
int64_t     aid_dedup_launch_count(const aid_dedup* d);
/* appends n rows; word_off[0] = 0; *first_row receives the row number of the first one */
int aid_dedup_add(aid_dedup* d, const uint32_t* words, const int64_t* word_off /* [n+1] */,
                  const double* duration /* [n] */, int n, int64_t* first_row);
int aid_dedup_scan(aid_dedup* d, const uint32_t* q_words, const int64_t* q_off /* [nq+1] */,
                   const double* q_lo, const double* q_hi, int nq,
                   int64_t* best_row /* [nq] */, double* best_sim /* [nq] */);
/* device time of the last scan's kernels (CUDA events on the store's stream), milliseconds */
double      aid_dedup_last_scan_ms(const aid_dedup* d);

/* ---- decode feed -------------------------------------------------------------
 * The reference's decode_dual_rate starts two ffmpeg children per file, one per output rate
 * (audio-ident-service/app/audio/decode.py:74-87, :37-58). With these entry points one 48 kHz decode is enough:
 * the 16 kHz stream of the fingerprint path is derived on the GPU by a 61-tap zero-phase polyphase decimator,
 * y[j] = sum_{m=-30..30} h[m+30] x[3j+m] (x = 0 outside the clip), h = firwin(61, 1/3, kaiser 5.0) -- the design of
 * scipy.signal.resample_poly(x, 1, 3), which is the oracle. Not bit-compatible with ffmpeg's own resampler. */
int64_t aid_resample_out_len(int64_t n_in);                 /* ceil(n_in / 3) */
void    aid_resample_taps(float* taps /* [61] */);          /* the float32 filter both sides use */
int     aid_resample_48k_to_16k_host(aid_engine* e, const float* pcm48, int64_t n_in, float* pcm16);
int     aid_resample_48k_to_16k_dev(aid_engine* e, const float* d_in, int64_t n_in, float* d_out, void* stream);

/* ---- helpers for bindings that do not link the CUDA runtime themselves ---------------------- */
int aid_device_alloc(aid_engine* e, int64_t bytes, void** d_ptr);
int aid_device_free(aid_engine* e, void* d_ptr);
int aid_copy_to_device(aid_engine* e, void* d_dst, const void* h_src, int64_t bytes);
int aid_copy_to_host(aid_engine* e, void* h_dst, const void* d_src, int64_t bytes);
/* device-to-device, asynchronous on `stream` (e.g. engine-owned results into a torch tensor) */
int aid_copy_device(aid_engine* e, void* d_dst, const void* d_src, int64_t bytes, void* stream);
/* fills d_pcm with the deterministic device-side synthetic corpus used by bench.py: track k of the
 * batch is global track number first_track + k; all tracks have samples_per_track samples. */
int aid_synth_tracks_dev(aid_engine* e, float* d_pcm, int64_t first_track, int n_tracks,
                         int64_t samples_per_track, uint64_t seed, void* stream);
/* same, track k of the batch is global track first_track + k * track_stride (a rank's shard of a round-robin corpus) */
int aid_synth_tracks_strided_dev(aid_engine* e, float* d_pcm, int64_t first_track, int64_t track_stride, int n_tracks,
                                 int64_t samples_per_track, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AUDIO_IDENT_B200_H */
