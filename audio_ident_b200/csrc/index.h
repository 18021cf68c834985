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

int aid_index_commit_on(aid_engine* e, cudaStream_t st);
int aid_index_append(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* h_off,
                     const int32_t* h_status, const int64_t* n_frames, int n, const char* const* names,
                     uint8_t* ok, cudaStream_t st);
