// index.h -- private: the in-HBM index structures shared by index.cu and match.cu.
#pragma once
#include "engine.h"

struct TrackInfo {
    std::string name;
    int64_t n_frames = 0;
    uint32_t n_hashes = 0;
    bool deleted = false;
};

struct Segment {
    uint32_t first_track = 0;      // engine-wide number of local track 0 (= segment number * AID_SEG_TRACKS)
    uint32_t n_tracks = 0;
    int64_t n_entries = 0;
    bool dirty = false;            // entries were added since the last build
    DevBuf st_hash, st_post;       // entries in arrival order
    DevBuf bucket, postings;       // sorted form
    DevBuf tomb;                   // u32[AID_SEG_TRACKS/32] deleted-track bits
    std::vector<uint32_t> h_tomb;
    void release();
};

// what the matcher needs of a segment (device copy in Index::d_segdesc)
struct aid_seg_desc {
    const uint32_t* bucket;
    const uint32_t* postings;
    const uint32_t* tomb;
    uint32_t first_track;
    uint32_t n_tracks;
};

struct Index {
    std::vector<TrackInfo> tracks;                       // by engine-wide track number
    std::unordered_map<std::string, uint32_t> by_name;   // live tracks only
    std::vector<Segment*> segs;
    int64_t live_tracks = 0, n_postings = 0;
    DevBuf cursor, scan_tmp, d_jobs, d_segdesc;
    bool segdesc_dirty = true;
    // matcher workspace
    DevBuf cand, cand_n, rows, rows_n;
};

// ---- sharded identification: where k_rank delivers a window's rows (exchange.cu) -------------------------------
// With world == 0 the rows go to the caller's buffers. Otherwise every rank owns a receive window (one device
// allocation, mapped into its peers through CUDA IPC or used directly inside one process):
//   flags  u32[AID_MAX_RANKS]                          epoch of the last complete block received from rank r
//   counts i32[2][world][max_q]                        rows per window        (double-buffered by epoch parity)
//   rows   aid_match_row[2][world][max_q][AID_MAX_ROWS]
// and k_rank stores each window's rows, already in global track numbers, into slot [epoch & 1][rank] of EVERY window
// over NVLink; the last CTA to finish publishes the epoch in every window's flag. The merge kernel of the receiving
// rank waits for the world's flags and merges the blocks. No NCCL call, no host round trip.
constexpr size_t kXchgFlagBytes = 256;
struct RowSink {
    int world = 0, rank = 0, max_q = 0;
    uint32_t epoch = 0;
    unsigned char* window[AID_MAX_RANKS] = {};   // peers' receive windows (window[rank] is this rank's own)
    uint32_t* done = nullptr;                   // CTA completion counter (local)
    const uint32_t* track_map = nullptr;        // engine track number -> global track number (or null)
    uint32_t n_map = 0;
};
__host__ __device__ inline int32_t* xchg_counts(unsigned char* w, int world, int max_q, int parity, int src) {
    return reinterpret_cast<int32_t*>(w + kXchgFlagBytes) + ((size_t)parity * world + src) * max_q;
}
__host__ __device__ inline aid_match_row* xchg_rows(unsigned char* w, int world, int max_q, int parity, int src) {
    const size_t counts_bytes = ((size_t)2 * world * max_q * 4 + 255) / 256 * 256;
    return reinterpret_cast<aid_match_row*>(w + kXchgFlagBytes + counts_bytes) + ((size_t)parity * world + src) * max_q * AID_MAX_ROWS;
}
inline size_t xchg_window_bytes(int world, int max_q) {
    const size_t counts_bytes = ((size_t)2 * world * max_q * 4 + 255) / 256 * 256;
    return kXchgFlagBytes + counts_bytes + (size_t)2 * world * max_q * AID_MAX_ROWS * sizeof(aid_match_row);
}
// match.cu: probe + vote + rank for n_q windows; rows go to (d_rows, d_n_rows) or, with sink.world > 0, to the peers
int aid_match_device_out(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* d_hash_off,
                         const uint32_t* d_hash_len, const int32_t* d_status, int n_q, aid_match_row* d_rows,
                         int max_rows, int32_t* d_n_rows, const RowSink& sink, cudaStream_t st);

int aid_index_commit_on(aid_engine* e, cudaStream_t st);
int aid_index_append(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* h_off,
                     const int32_t* h_status, const int64_t* n_frames, int n, const char* const* names,
                     uint8_t* ok, cudaStream_t st);
