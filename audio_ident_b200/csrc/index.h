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
    uint32_t n_deleted = 0;        // tombstones set
    int group = -1, sub = 0;       // member `sub` of Index::groups[group] once sealed and grouped (then bucket/postings are released)
    void release();
};

// Up to kGroupSegs sealed (full) segments share one hash directory: per hash one 32-byte entry
//   u32 start | u16 count[8] | 12 B pad
// and one posting array ordered by (hash, member segment, local track, t_anchor). A (window, segment) CTA of the
// matcher reads ONE sector per query hash -- the same sector its 7 sibling CTAs read, so 7 of 8 lookups are L2 hits --
// and the posting runs of the 8 members of a hash are adjacent (they share sectors too). Votes, keys, tombstones and
// results are exactly those of the plain segments; only where a segment's run of a hash is found changes.
constexpr int kGroupSegs = 8;
struct SegGroup {
    uint32_t first_seg = 0, n_segs = 0;
    int64_t n_entries = 0;
    DevBuf dir, postings;          // u32[2^24][8], u32[n_entries]
    void release() { dir.release(); postings.release(); }
};

// what the matcher needs of a segment (device copy in Index::d_segdesc)
struct aid_seg_desc {
    const uint32_t* bucket;        // plain segment: u32[2^24+1]; null for a grouped one
    const uint32_t* postings;      // the segment's or its group's posting array
    const uint32_t* tomb;
    const uint32_t* dir;           // grouped segment: the group's directory
    uint32_t sub;                  // ... and which member this segment is
    uint32_t first_track;
    uint32_t n_tracks;
    uint32_t n_deleted;            // 0: the matcher skips the tombstone lookup
};

struct Index {
    std::vector<TrackInfo> tracks;                       // by engine-wide track number
    std::unordered_map<std::string, uint32_t> by_name;   // live tracks only
    std::vector<Segment*> segs;
    std::vector<SegGroup*> groups;                       // groups[k] covers the sealed segments among [8k, 8k + 8)
    bool grouping = true;                                // aid_index_set_grouping (tests and A/B measurements)
    int64_t live_tracks = 0, n_postings = 0;
    DevBuf cursor, scan_tmp, d_jobs, d_segdesc, group_start;
    bool segdesc_dirty = true;
    // matcher workspace
    DevBuf cand, cand_n, rows, rows_n;
    DevBuf vote_stats;                                   // u64[1024][2]: hashes probed, postings touched (aid_match_stats)
};

// ---- sharded identification over peer memory (exchange.cu) ---------------------------------------------------
// Every rank owns a receive window (one device allocation, mapped into its peers through CUDA IPC or used directly
// inside one process). Everything that changes per call is double-buffered by epoch parity:
//   flag page (256 B)  u32 row_flag[8] | u32 hash_flag[8] (at +64) | u32 hash_status[8] (at +128):
//                      epoch of the last complete row block / hash block received from rank r
//   counts  i32[2][world][max_q]                        rows per window and source rank
//   rows    aid_match_row[2][world][max_q][AID_MAX_ROWS]
//   begins  u32[2][max_q], lens u32[2][max_q]           where window q's hashes sit in `hash`/`t` (whole batch)
//   hash    u32[2][world][hcap], t u32[2][world][hcap]  the query fingerprints rank r computed for its slice of the batch
// k_publish_hashes stores a rank's fingerprints into every window, k_rank (match.cu) stores each window's rows,
// already in global track numbers, into slot [epoch & 1][rank] of every window; in both the last CTA to finish
// publishes the epoch in every window's flag, and the consumer kernel of the receiving rank waits for the world's
// flags on the device. No NCCL call, no host round trip. With world == 0 a RowSink means "rows go to the caller".
constexpr size_t kXchgFlagBytes = 256;
struct XchgLayout {
    int world = 0, max_q = 0;
    uint32_t hcap = 0;
    size_t off_counts = 0, off_rows = 0, off_begins = 0, off_lens = 0, off_hash = 0, off_t = 0, bytes = 0;
    static XchgLayout make(int world, int max_q, uint32_t hcap) {
        XchgLayout L; L.world = world; L.max_q = max_q; L.hcap = hcap;
        auto up = [](size_t v) { return (v + 255) / 256 * 256; };
        size_t o = kXchgFlagBytes;
        L.off_counts = o; o = up(o + (size_t)2 * world * max_q * 4);
        L.off_rows = o;   o = up(o + (size_t)2 * world * max_q * AID_MAX_ROWS * sizeof(aid_match_row));
        L.off_begins = o; o = up(o + (size_t)2 * max_q * 4);
        L.off_lens = o;   o = up(o + (size_t)2 * max_q * 4);
        L.off_hash = o;   o = up(o + (size_t)2 * world * hcap * 4);
        L.off_t = o;      o = up(o + (size_t)2 * world * hcap * 4);
        L.bytes = o;
        return L;
    }
    __host__ __device__ int32_t* counts(unsigned char* w, int parity, int src) const {
        return reinterpret_cast<int32_t*>(w + off_counts) + ((size_t)parity * world + src) * max_q;
    }
    __host__ __device__ aid_match_row* rows(unsigned char* w, int parity, int src) const {
        return reinterpret_cast<aid_match_row*>(w + off_rows) + ((size_t)parity * world + src) * max_q * AID_MAX_ROWS;
    }
    __host__ __device__ uint32_t* begins(unsigned char* w, int parity) const { return reinterpret_cast<uint32_t*>(w + off_begins) + (size_t)parity * max_q; }
    __host__ __device__ uint32_t* lens(unsigned char* w, int parity) const { return reinterpret_cast<uint32_t*>(w + off_lens) + (size_t)parity * max_q; }
    __host__ __device__ uint32_t* hash(unsigned char* w, int parity) const { return reinterpret_cast<uint32_t*>(w + off_hash) + (size_t)parity * world * hcap; }
    __host__ __device__ uint32_t* t(unsigned char* w, int parity) const { return reinterpret_cast<uint32_t*>(w + off_t) + (size_t)parity * world * hcap; }
    __host__ __device__ static uint32_t* row_flag(unsigned char* w) { return reinterpret_cast<uint32_t*>(w); }
    __host__ __device__ static uint32_t* hash_flag(unsigned char* w) { return reinterpret_cast<uint32_t*>(w) + 16; }
    __host__ __device__ static uint32_t* hash_status(unsigned char* w) { return reinterpret_cast<uint32_t*>(w) + 32; }
};
struct RowSink {
    XchgLayout lay;                             // lay.world == 0: rows go to the caller's buffers
    int rank = 0;
    uint32_t epoch = 0;
    unsigned char* window[AID_MAX_RANKS] = {};   // the ranks' receive windows (window[rank] is this rank's own)
    uint32_t* done = nullptr;                   // CTA completion counter (local)
    const uint32_t* track_map = nullptr;        // engine track number -> global track number (or null)
    uint32_t n_map = 0;
};
// match.cu: probe + vote + rank for n_q windows; rows go to (d_rows, d_n_rows) or, with sink.world > 0, to the peers
int aid_match_device_out(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* d_hash_off,
                         const uint32_t* d_hash_len, const int32_t* d_status, int n_q, aid_match_row* d_rows,
                         int max_rows, int32_t* d_n_rows, const RowSink& sink, cudaStream_t st,
                         const uint32_t* d_abort /* device flag: non-zero = probe nothing (may be null) */);

int aid_index_commit_on(aid_engine* e, cudaStream_t st);
int aid_index_append(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* h_off,
                     const int32_t* h_status, const int64_t* n_frames, int n, const char* const* names,
                     uint8_t* ok, cudaStream_t st);
