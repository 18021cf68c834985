// common.cuh -- shared declarations for the sm_100a fingerprint engine kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#include "../../include/aid_params.h"
#include "../../include/audio_ident_b200.h"

#define AID_FULL_MASK 0xffffffffu

// ---- work units ------------------------------------------------------------------
// A ragged batch is cut into fixed-size units so that every kernel's grid is a flat list
// and a CTA (or warp) finds its unit with one lookup.
struct aid_stft_unit {     // one warp: frames [frame0, frame0 + n_frames) of one track
    int64_t pcm_begin;     // first sample of the track in the batch PCM buffer
    int64_t n_samples;     // samples in the track
    int64_t spec_row;      // row of (track, frame0) in the batch spectrogram
    int32_t frame0;
    int32_t n_frames;
};

struct aid_peak_unit {     // one CTA: output rows [row0, row0 + n_rows) of one track
    int64_t spec_row0;     // row of (track, frame 0) in the batch spectrogram
    int32_t track;         // index in the batch
    int32_t n_frames;      // frames in the track
    int32_t row0;
    int32_t n_rows;
};

struct aid_peak_run {      // one warp of the peak kernel: `n_blocks` consecutive peak units (256-frame blocks) of one track
    int32_t first_unit;
    int32_t n_blocks;
};

#define AID_PEAK_RUN_BLOCKS    4       // blocks streamed back to back by one warp (halo re-read: 24 rows per run)
#ifndef AID_STFT_UNIT_FRAMES
#define AID_STFT_UNIT_FRAMES   64      // frames per warp-unit (even)
#endif
#define AID_PEAK_BLOCK_FRAMES  256     // must equal the spec's aligned block (aid_params.h)

// ---- kernel launchers (defined in the .cu files, called by engine.cu) --------------
struct aid_tables {              // device-resident constant tables, built once per engine
    const float* window;         // [1024] float32 Hamming
    const float* twist;          // [AID_TWIST_FLOATS] twiddles between / inside the STFT transforms, see aid_fill_stft_tables
};
constexpr int AID_TWIST_FOLDED = 32 * 32;              // floats of the folded table
constexpr int AID_TWIST_IMAGE = 32 * 32 + 32 * 64;     // + the plain table W_1024^(n1*k1); then the shared-memory image of k_stft_packed
constexpr int AID_TWIST_IMAGE_FLOATS = 2 * 32 * 36;    // [32][36] window / 2, lane-major, then [32][36] packed twiddles (stft.cu)
constexpr int AID_TWIST_FLOATS = AID_TWIST_IMAGE + AID_TWIST_IMAGE_FLOATS;

// Host-side definition of the two tables (double precision, rounded to float once).
//  window[n] : symmetric Hamming, the formula of oracle/aid_oracle.c tables_init.
//  twist[k1][2*i], [2*i+1] = (c, s) of the twiddle w = c - i s = g^m * W_(2h)^k with g = W_1024^k1, for
//    i = 0: stage 0 (m 16, h 1, k 0)      i = 1: stage 1 (m 8, h 2, k 0)      i = 2..3: stage 2 (m 4, h 4, k 0..1)
//    i = 4..7: stage 3 (m 2, h 8, k 0..3)  i = 8..15: stage 4 (m 1, h 16, k 0..7)
//  twist[1024 + k1*64 + 2*n1], [+1] = (c, s) of W_1024^(n1*k1) = c - i s: the inter-transform twiddle itself.
//  twist[AID_TWIST_IMAGE ...]: the two tables as k_stft_packed keeps them in shared memory (copied with 128-bit loads):
//    [lane][j], j < 32 (row stride 36): window[lane + 32 j] / 2 (exact);
//    [k1][4 e + w] (row stride 36): the folded twiddles of the same values, two butterflies per quad: e = 0: stage 1 as
//    (c, -s, s, c); e >= 1: twiddles i = 2e, 2e + 1 of the list above as (c_i, c_i+1, s_i, s_i+1); [k1][32..33]: stage 0 (c, s).
inline void aid_fill_stft_tables(float* window, float* twist) {
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < AID_NFFT; i++)
        window[i] = (float)(AID_WIN_A0 - AID_WIN_A1 * cos(two_pi * (double)i / (double)(AID_NFFT - 1)));
    for (int k1 = 0; k1 < 32; k1++) {
        int i = 0;
        for (int stage = 0; stage < 5; stage++) {
            const int half = 1 << stage, m = 16 >> stage, nk = half >= 2 ? half / 2 : 1;
            for (int k = 0; k < nk; k++, i++) {
                const double a = two_pi * ((double)(k1 * m) / (double)AID_NFFT + (double)k / (double)(2 * half));
                twist[k1 * 32 + 2 * i] = (float)cos(a);
                twist[k1 * 32 + 2 * i + 1] = (float)sin(a);
            }
        }
        for (int n1 = 0; n1 < 32; n1++) {
            const double a = two_pi * (double)(k1 * n1) / (double)AID_NFFT;
            twist[AID_TWIST_FOLDED + k1 * 64 + 2 * n1] = (float)cos(a);
            twist[AID_TWIST_FOLDED + k1 * 64 + 2 * n1 + 1] = (float)sin(a);
        }
    }
    float* img_win = twist + AID_TWIST_IMAGE;
    float* img_tw = img_win + 32 * 36;
    for (int i = 0; i < 2 * 32 * 36; i++) img_win[i] = 0.0f;
    for (int i = 0; i < AID_NFFT; i++) img_win[(i & 31) * 36 + (i >> 5)] = 0.5f * window[i];
    for (int k1 = 0; k1 < 32; k1++) {
        const float* t = twist + k1 * 32;
        float* o = img_tw + k1 * 36;
        o[0] = t[2]; o[1] = -t[3]; o[2] = t[3]; o[3] = t[2];
        for (int e = 1; e < 8; e++)
            for (int w = 0; w < 4; w++) o[4 * e + w] = w < 2 ? t[2 * (2 * e + w)] : t[2 * (2 * e + w - 2) + 1];
        o[32] = t[0]; o[33] = t[1];
    }
}

// variant: 0 = scalar FP32 kernel (round 1), 5 = packed f32x2 kernel with L1 prefetch (default), 7 = the same with the
// separation software-pipelined into the next trip, 16 = the warp-specialised form (other values are A/B shapes, see stft.cu). d_gmax (nullable, packed variants only): [rows][32] group maxima for the peak kernel.
cudaError_t aid_launch_stft_variant(int variant, const aid_tables& tb, const float* d_pcm, const aid_stft_unit* d_units,
                                    int n_units, float* d_spec, float* d_gmax, cudaStream_t st);
int aid_stft_default_variant();      // 5, or the environment's AID_STFT_VARIANT (measurement runs)

// d_gmax (nullable): the STFT's group maxima; with it the peak kernel streams 128 B per row instead of the 2 KB row
cudaError_t aid_launch_peaks(const float* d_spec, const float* d_gmax, const aid_peak_unit* d_units, const aid_peak_run* d_runs,
                             int n_runs, uint32_t* d_slots, uint32_t* d_unit_count, int32_t* d_track_status,
                             cudaStream_t st);

cudaError_t aid_launch_peak_compact(const uint32_t* d_slots, const uint32_t* d_unit_count,
                                    const uint32_t* d_unit_pos, const aid_peak_unit* d_units, int n_units,
                                    uint32_t* d_peaks, uint32_t* d_peak_track, cudaStream_t st);

// exclusive scan of n uint32 values, in place is allowed (out == in); d_tmp holds
// aid_scan_tmp_elems(n) uint32; the grand total is also written to *d_total if non-null.
// If d_n is non-null only the first min(n, *d_n + 1) values are scanned (length known on the device).
size_t aid_scan_tmp_elems(int64_t n);
cudaError_t aid_launch_scan_u32(const uint32_t* d_in, uint32_t* d_out, int64_t n, uint32_t* d_tmp,
                                uint32_t* d_total, const uint32_t* d_n, cudaStream_t st);

cudaError_t aid_launch_hash_count(const uint32_t* d_peaks, const uint32_t* d_peak_track,
                                  const uint32_t* d_peak_off, const uint32_t* d_n_peaks_total,
                                  int64_t max_peaks, uint32_t* d_cnt, cudaStream_t st);
cudaError_t aid_launch_hash_write(const uint32_t* d_peaks, const uint32_t* d_peak_track,
                                  const uint32_t* d_peak_off, const uint32_t* d_n_peaks_total,
                                  int64_t max_peaks, const uint32_t* d_pos,
                                  uint32_t* d_hash, uint32_t* d_t, int64_t hash_cap, int32_t* d_overflow,
                                  cudaStream_t st);

cudaError_t aid_launch_gather_u32(const uint32_t* d_src, const uint32_t* d_idx, uint32_t* d_dst, int n,
                                  cudaStream_t st);
cudaError_t aid_launch_synth(float* d_pcm, int64_t first_track, int n_tracks, int64_t samples_per_track,
                             uint64_t seed, cudaStream_t st, int64_t track_stride = 1);
