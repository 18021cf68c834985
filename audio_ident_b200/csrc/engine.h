// engine.h -- private host-side structures of libaudioident_b200.so (not part of the C ABI).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>
#include "common.cuh"

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        const size_t want = bytes + bytes / 4 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) { p = nullptr; return e; }
        cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

// Host description of one ragged sub-batch.
struct Plan {
    int n_tracks = 0;
    std::vector<int64_t> frames;        // per track
    std::vector<int64_t> frame_off;     // [n+1]
    std::vector<int32_t> host_status;   // AID_TRACK_EMPTY / AID_TRACK_TOO_LONG
    std::vector<aid_stft_unit> sunits;
    std::vector<aid_peak_unit> punits;
    std::vector<aid_peak_run> pruns;
    std::vector<uint32_t> first_punit;  // [n+1]
    int64_t total_frames = 0;
    int64_t peak_cap = 0;               // n_punits * AID_PEAK_BLOCK_CAP
    int64_t hash_cap = 0;               // peak_cap * AID_FANOUT
};

// One stream with its private workspace (two of them double-buffer the host-buffer entry points).
struct Slot {
    cudaStream_t st = nullptr;
    cudaEvent_t done = nullptr;
    DevBuf pcm, desc, spec, gmax, slots, unit_pos, peaks, peak_track, peak_off, pos, hash, t, hash_off, status,
           scan_tmp, misc;
    PinBuf h_desc, h_small;
    // views into `desc` (one upload per sub-batch)
    aid_stft_unit* d_sunits = nullptr;
    aid_peak_unit* d_punits = nullptr;
    aid_peak_run* d_pruns = nullptr;
    uint32_t* d_first_punit = nullptr;
    void release();
};

struct Index;     // index.h

struct aid_engine {
    int device = 0;
    std::string err;
    int64_t launches = 0;
    int64_t max_batch_frames = 8 * 1024 * 1024;
    Slot slot[2];
    DevBuf d_window, d_twiddle;
    aid_tables tables{};
    Index* index = nullptr;
    // kernel selection (aid_engine_set_kernels; the defaults are the product path, the others exist for tests and A/B runs)
    int stft_variant = 5;        // stft.cu aid_launch_stft_variant: 0 = scalar FP32 kernel, 5 = packed f32x2 kernel (7: + pipelined separation)
    bool peak_summary = true;    // the STFT also writes the 16-bin group maxima and the peak kernel streams those
    // optional per-stage timing (aid_engine_set_stage_timing)
    bool timing = false;
    struct StageRec { int stage; cudaEvent_t a, b; };
    std::vector<StageRec> stage_recs;
    std::vector<cudaEvent_t> event_pool;
};

// records a CUDA-event pair around a group of launches when stage timing is on
struct StageTimer {          // records an event pair around a group of launches when timing is on
    aid_engine* e; cudaStream_t st; int stage; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t take(aid_engine* e) {
        cudaEvent_t ev = nullptr;
        if (!e->event_pool.empty()) { ev = e->event_pool.back(); e->event_pool.pop_back(); }
        else if (cudaEventCreate(&ev) != cudaSuccess) { cudaGetLastError(); ev = nullptr; }
        return ev;
    }
    StageTimer(aid_engine* e_, cudaStream_t st_, int stage_) : e(e_), st(st_), stage(stage_) {
        if (!e->timing) return;
        a = take(e); b = take(e);
        if (a) cudaEventRecord(a, st);
    }
    ~StageTimer() {
        if (!e->timing || !a || !b) return;
        cudaEventRecord(b, st);
        e->stage_recs.push_back({stage, a, b});
    }
};


// engine.cu internals used by index.cu / match.cu
int aid_fail_cuda(aid_engine* e, cudaError_t ce, const char* what);
#define AID_CUDA(e, call) do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) return aid_fail_cuda((e), ce_, #call); } while (0)

int aid_build_plan(const int64_t* sample_off, int first, int count, int64_t frame_limit, Plan& plan);
int aid_build_plan_windows(const int64_t* begin_of, const int64_t* end_of, int count, int64_t base, int64_t frame_limit,
                           Plan& plan);
// uploads descriptors, runs stft .. hashes on slot s for `plan`; d_pcm is the sub-batch's first sample.
int aid_run_fingerprint(aid_engine* e, Slot& s, const Plan& plan, const float* d_pcm, cudaStream_t st);
int aid_slot_prepare(aid_engine* e, Slot& s, const Plan& plan, bool need_pcm, int64_t pcm_samples);

Index* aid_index_new();
void aid_index_free(aid_engine* e, Index* ix);
