// index.cu -- the hash index resident in HBM: segments, build (counting sort by hash), persistence.
//
// Replaces stage a7/a8 of SURVEY.md section 8(a): the LMDB store behind `olaf_c store` / `olaf_c del`
// (reference audio-ident-service/app/audio/fingerprint.py:117-125, :239-246; directory = settings.olaf_lmdb_path,
// fingerprint.py:79-84). Definition of the sorted order: oracle/aid_oracle.c aid_oracle_index_build().
//
// Layout (DESIGN.md "Index"): the index is a list of segments of at most AID_SEG_TRACKS (16384) tracks.
// A segment keeps
//   entries  : (hash u32, posting u32) in arrival order, posting = (local_track << 18) | t_anchor
//   bucket   : u32[2^24 + 1], bucket[h] = first sorted posting of hash h
//   postings : u32[n], ordered by (hash, local_track, t_anchor)
// Build = histogram over hashes (global atomics) -> exclusive scan (the bucket table itself) -> scatter with
// per-bucket cursors -> per-bucket sort (buckets hold a handful of postings). Only the last segment is open;
// adding tracks marks it dirty and the next query (or aid_index_commit) rebuilds just that segment.
// Deleting a track sets a tombstone bit that the matcher checks; its postings stay until the segment is rebuilt.
// Full segments give up their own table and share a hash directory eight at a time (index.h SegGroup; build_group
// below): the matcher's eight CTAs of a window then read the same directory sector per hash and adjacent posting runs.
#include <errno.h>
#include <stdio.h>
#include <string.h>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include "engine.h"
#include "index.h"

namespace {

constexpr int64_t kBuckets = (int64_t)1 << AID_HASH_BITS;

// one CTA per track of the sub-batch: copy its hashes into the segment's entry arrays
struct AppendJob { const uint32_t* src_hash; const uint32_t* src_t; uint32_t* dst_hash; uint32_t* dst_post; uint32_t n; uint32_t local; };

__global__ void k_append(const AppendJob* __restrict__ jobs) {
    const AppendJob j = jobs[blockIdx.x];
    for (uint32_t i = threadIdx.x; i < j.n; i += blockDim.x) {
        j.dst_hash[i] = j.src_hash[i];
        j.dst_post[i] = (j.local << AID_POST_T_BITS) | j.src_t[i];
    }
}

__global__ void k_hist(const uint32_t* __restrict__ hash, int64_t n, uint32_t* __restrict__ bucket) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        atomicAdd(bucket + hash[i], 1u);
}

__global__ void k_scatter(const uint32_t* __restrict__ hash, const uint32_t* __restrict__ post, int64_t n,
                          uint32_t* __restrict__ cursor, uint32_t* __restrict__ postings) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        postings[atomicAdd(cursor + hash[i], 1u)] = post[i];
}

// ascending in-place sort of one short run of postings
__device__ void sort_run(uint32_t* a, uint32_t n) {
    if (n < 2) return;
    if (n <= 32) {                                   // insertion sort
        for (uint32_t i = 1; i < n; i++) {
            const uint32_t v = a[i];
            uint32_t j = i;
            while (j > 0 && a[j - 1] > v) { a[j] = a[j - 1]; j--; }
            a[j] = v;
        }
        return;
    }
    // heapsort, in place
    auto sift = [&](uint32_t start, uint32_t end) {
        uint32_t root = start;
        while (2 * root + 1 <= end) {
            uint32_t child = 2 * root + 1, sw = root;
            if (a[sw] < a[child]) sw = child;
            if (child + 1 <= end && a[sw] < a[child + 1]) sw = child + 1;
            if (sw == root) return;
            const uint32_t t = a[root]; a[root] = a[sw]; a[sw] = t;
            root = sw;
        }
    };
    for (int64_t s = (int64_t)(n - 2) / 2; s >= 0; s--) sift((uint32_t)s, n - 1);
    for (uint32_t end = n - 1; end > 0; end--) {
        const uint32_t t = a[end]; a[end] = a[0]; a[0] = t;
        sift(0, end - 1);
    }
}

// every bucket of a segment (one thread per bucket; buckets average a few postings)
__global__ void k_bucket_sort(const uint32_t* __restrict__ bucket, uint32_t* __restrict__ postings) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= kBuckets) return;
    const uint32_t b = bucket[h];
    sort_run(postings + b, bucket[h + 1] - b);
}

// ---- segment groups (index.h SegGroup): directory entry of hash h = dir[8 h .. 8 h + 8): start, 4 words of 16-bit counts
__device__ __forceinline__ uint32_t dir_count(const uint32_t* __restrict__ ent, int sub) {
    return (ent[1 + (sub >> 1)] >> (16 * (sub & 1))) & 0xffffu;
}

__global__ void k_group_hist(const uint32_t* __restrict__ hash, int64_t n, int sub, uint32_t* __restrict__ dir,
                             uint32_t* __restrict__ overflow) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t old = atomicAdd(dir + (size_t)hash[i] * 8 + 1 + (sub >> 1), 1u << (16 * (sub & 1)));
        if (((old >> (16 * (sub & 1))) & 0xffffu) == 0xffffu) *overflow = 1;     // a 16-bit count wrapped: the group is abandoned
    }
}

__global__ void k_group_total(const uint32_t* __restrict__ dir, uint32_t* __restrict__ total) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h > kBuckets) return;
    uint32_t t = 0;
    if (h < kBuckets)
        for (int s = 0; s < kGroupSegs; s++) t += dir_count(dir + (size_t)h * 8, s);
    total[h] = t;
}

// after the scan: start of every hash into the directory, and the scatter cursor of member `sub`
__global__ void k_group_cursor(uint32_t* __restrict__ dir, const uint32_t* __restrict__ start, int sub, int write_start,
                               uint32_t* __restrict__ cursor) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= kBuckets) return;
    uint32_t* ent = dir + (size_t)h * 8;
    uint32_t c = start[h];
    if (write_start) ent[0] = c;
    for (int s = 0; s < sub; s++) c += dir_count(ent, s);
    cursor[h] = c;
}

__global__ void k_group_sort(const uint32_t* __restrict__ dir, int n_segs, uint32_t* __restrict__ postings) {
    const int64_t h = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (h >= kBuckets) return;
    const uint32_t* ent = dir + (size_t)h * 8;
    uint32_t b = ent[0];
    for (int s = 0; s < n_segs; s++) {
        const uint32_t n = dir_count(ent, s);
        sort_run(postings + b, n);
        b += n;
    }
}

cudaError_t grow_copy(DevBuf& b, size_t used_bytes, size_t need_bytes) {
    if (need_bytes <= b.cap) return cudaSuccess;
    size_t want = std::max(need_bytes, b.cap * 2);
    void* np = nullptr;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) { want = need_bytes; e = cudaMalloc(&np, want); }
    if (e != cudaSuccess) return e;
    if (b.p && used_bytes) e = cudaMemcpy(np, b.p, used_bytes, cudaMemcpyDeviceToDevice);
    if (b.p) cudaFree(b.p);
    b.p = np; b.cap = want;
    return e;
}

}  // namespace

Index* aid_index_new() { return new Index(); }

void Segment::release() { st_hash.release(); st_post.release(); bucket.release(); postings.release(); tomb.release(); }

void aid_index_free(aid_engine*, Index* ix) {
    if (!ix) return;
    for (Segment* s : ix->segs) { s->release(); delete s; }
    for (SegGroup* g : ix->groups) { if (g) { g->release(); delete g; } }
    ix->cursor.release(); ix->scan_tmp.release(); ix->d_jobs.release(); ix->d_segdesc.release(); ix->group_start.release();
    ix->cand.release(); ix->cand_n.release(); ix->rows.release(); ix->rows_n.release(); ix->vote_stats.release();
    delete ix;
}

static Segment* open_segment(aid_engine* e, Index* ix) {
    if (!ix->segs.empty() && ix->segs.back()->n_tracks < AID_SEG_TRACKS) return ix->segs.back();
    Segment* s = new Segment();
    s->first_track = (uint32_t)ix->segs.size() * AID_SEG_TRACKS;
    if (s->tomb.ensure(AID_SEG_TRACKS / 8) != cudaSuccess || cudaMemset(s->tomb.p, 0, AID_SEG_TRACKS / 8) != cudaSuccess) {
        cudaGetLastError(); delete s; return nullptr;
    }
    s->h_tomb.assign(AID_SEG_TRACKS / 32, 0);
    ix->segs.push_back(s);
    ix->segdesc_dirty = true;
    (void)e;
    return s;
}

static int set_tombstone(aid_engine* e, Index* ix, uint32_t track) {
    Segment* s = ix->segs[track / AID_SEG_TRACKS];
    const uint32_t local = track % AID_SEG_TRACKS;
    if (!(s->h_tomb[local / 32] & (1u << (local % 32)))) { s->n_deleted++; ix->segdesc_dirty = true; }
    s->h_tomb[local / 32] |= 1u << (local % 32);
    AID_CUDA(e, cudaMemcpy(s->tomb.as<uint32_t>() + local / 32, &s->h_tomb[local / 32], 4, cudaMemcpyHostToDevice));
    TrackInfo& ti = ix->tracks[track];
    if (!ti.deleted) { ti.deleted = true; ix->live_tracks--; }
    return AID_OK;
}

// Registers the tracks of one fingerprinted sub-batch (results in host arrays h_off/h_status, device arrays
// d_hash/d_t) and appends their entries to the open segment(s).
int aid_index_append(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* h_off,
                     const int32_t* h_status, const int64_t* n_frames, int n, const char* const* names,
                     uint8_t* ok, cudaStream_t st) {
    Index* ix = e->index;
    struct Placed { Segment* seg; int64_t at; uint32_t src, cnt, local; };
    std::vector<Placed> placed;
    std::vector<std::pair<Segment*, int64_t>> touched;      // segment, entries it held before this call
    placed.reserve(n);
    for (int i = 0; i < n; i++) {
        ok[i] = 0;
        if (!names[i]) return AID_E_ARG;
        const int64_t cnt = (int64_t)h_off[i + 1] - h_off[i];
        if ((h_status[i] & (AID_TRACK_PEAK_OVERFLOW | AID_TRACK_TOO_LONG)) || n_frames[i] > AID_INDEX_MAX_FRAMES) continue;
        if (ix->tracks.size() >= ((size_t)1 << 25)) return AID_E_FULL;
        Segment* s = open_segment(e, ix);
        if (!s) return aid_fail_cuda(e, cudaErrorMemoryAllocation, "segment allocation");
        auto old = ix->by_name.find(names[i]);
        if (old != ix->by_name.end()) { int rc = set_tombstone(e, ix, old->second); if (rc) return rc; }
        if (touched.empty() || touched.back().first != s) touched.push_back({s, s->n_entries});
        const uint32_t local = s->n_tracks++;
        TrackInfo ti; ti.name = names[i]; ti.n_frames = n_frames[i]; ti.n_hashes = (uint32_t)cnt; ti.deleted = false;
        ix->tracks.push_back(ti);
        ix->by_name[ti.name] = s->first_track + local;
        ix->live_tracks++;
        if (cnt > 0) {
            placed.push_back({s, s->n_entries, h_off[i], (uint32_t)cnt, local});
            s->n_entries += cnt;
            s->dirty = true;
            ix->n_postings += cnt;
        }
        ok[i] = 1;
    }
    for (auto& t : touched) {
        AID_CUDA(e, grow_copy(t.first->st_hash, (size_t)t.second * 4, (size_t)t.first->n_entries * 4));
        AID_CUDA(e, grow_copy(t.first->st_post, (size_t)t.second * 4, (size_t)t.first->n_entries * 4));
    }
    if (!placed.empty()) {
        std::vector<AppendJob> jobs(placed.size());
        for (size_t k = 0; k < placed.size(); k++) {
            const Placed& p = placed[k];
            jobs[k].src_hash = d_hash + p.src; jobs[k].src_t = d_t + p.src;
            jobs[k].dst_hash = p.seg->st_hash.as<uint32_t>() + p.at; jobs[k].dst_post = p.seg->st_post.as<uint32_t>() + p.at;
            jobs[k].n = p.cnt; jobs[k].local = p.local;
        }
        AID_CUDA(e, ix->d_jobs.ensure(jobs.size() * sizeof(AppendJob)));
        AID_CUDA(e, cudaMemcpyAsync(ix->d_jobs.p, jobs.data(), jobs.size() * sizeof(AppendJob), cudaMemcpyHostToDevice, st));
        k_append<<<(unsigned)jobs.size(), 128, 0, st>>>(ix->d_jobs.as<AppendJob>());
        AID_CUDA(e, cudaGetLastError());
        AID_CUDA(e, cudaStreamSynchronize(st));
        e->launches += 1;
    }
    return AID_OK;
}

static int build_segment(aid_engine* e, Index* ix, Segment* s, cudaStream_t st) {
    const int64_t n = s->n_entries;
    AID_CUDA(e, s->bucket.ensure((size_t)(kBuckets + 1) * 4));
    AID_CUDA(e, ix->cursor.ensure((size_t)(kBuckets + 1) * 4));
    AID_CUDA(e, ix->scan_tmp.ensure(aid_scan_tmp_elems(kBuckets + 1) * 4));
    AID_CUDA(e, s->postings.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    uint32_t* bucket = s->bucket.as<uint32_t>();
    AID_CUDA(e, cudaMemsetAsync(bucket, 0, (size_t)(kBuckets + 1) * 4, st));
    if (n > 0) {
        StageTimer tm(e, st, 6);
        const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 16);
        k_hist<<<grid, 256, 0, st>>>(s->st_hash.as<uint32_t>(), n, bucket);
        AID_CUDA(e, aid_launch_scan_u32(bucket, bucket, kBuckets + 1, ix->scan_tmp.as<uint32_t>(), nullptr, nullptr, st));
        AID_CUDA(e, cudaMemcpyAsync(ix->cursor.p, bucket, (size_t)(kBuckets + 1) * 4, cudaMemcpyDeviceToDevice, st));
        k_scatter<<<grid, 256, 0, st>>>(s->st_hash.as<uint32_t>(), s->st_post.as<uint32_t>(), n, ix->cursor.as<uint32_t>(),
                                        s->postings.as<uint32_t>());
        k_bucket_sort<<<(unsigned)((kBuckets + 255) / 256), 256, 0, st>>>(bucket, s->postings.as<uint32_t>());
        AID_CUDA(e, cudaGetLastError());
        e->launches += 6;
    }
    s->dirty = false;
    ix->segdesc_dirty = true;
    return AID_OK;
}

// (Re)builds group k over its sealed member segments [first, first + n). Returns AID_OK with *built = false when a
// 16-bit count would overflow (degenerate corpus): the members then stay plain segments.
static int build_group(aid_engine* e, Index* ix, SegGroup* g, cudaStream_t st, bool* built) {
    *built = false;
    int64_t n = 0;
    for (uint32_t i = 0; i < g->n_segs; i++) n += ix->segs[g->first_seg + i]->n_entries;
    if (n >= ((int64_t)1 << 32) - 1) return AID_OK;
    AID_CUDA(e, g->dir.ensure((size_t)kBuckets * 32));
    AID_CUDA(e, g->postings.ensure((size_t)std::max<int64_t>(n, 1) * 4));
    AID_CUDA(e, ix->cursor.ensure((size_t)(kBuckets + 1) * 4));
    AID_CUDA(e, ix->group_start.ensure((size_t)(kBuckets + 1) * 4 + 256));
    AID_CUDA(e, ix->scan_tmp.ensure(aid_scan_tmp_elems(kBuckets + 1) * 4));
    uint32_t* dir = g->dir.as<uint32_t>();
    uint32_t* start = ix->group_start.as<uint32_t>();
    uint32_t* overflow = start + kBuckets + 1;
    AID_CUDA(e, cudaMemsetAsync(dir, 0, (size_t)kBuckets * 32, st));
    AID_CUDA(e, cudaMemsetAsync(overflow, 0, 4, st));
    { StageTimer tm(e, st, 6);
    const unsigned per_hash = (unsigned)((kBuckets + 1 + 255) / 256);
    for (uint32_t i = 0; i < g->n_segs; i++) {
        const Segment* s = ix->segs[g->first_seg + i];
        if (s->n_entries == 0) continue;
        const int grid = (int)std::min<int64_t>((s->n_entries + 255) / 256, 148 * 16);
        k_group_hist<<<grid, 256, 0, st>>>(s->st_hash.as<uint32_t>(), s->n_entries, (int)i, dir, overflow);
        e->launches += 1;
    }
    uint32_t h_overflow = 0;
    AID_CUDA(e, cudaMemcpyAsync(&h_overflow, overflow, 4, cudaMemcpyDeviceToHost, st));
    AID_CUDA(e, cudaStreamSynchronize(st));
    if (h_overflow) { g->release(); return AID_OK; }
    k_group_total<<<per_hash, 256, 0, st>>>(dir, start);
    AID_CUDA(e, aid_launch_scan_u32(start, start, kBuckets + 1, ix->scan_tmp.as<uint32_t>(), nullptr, nullptr, st));
    for (uint32_t i = 0; i < g->n_segs; i++) {
        const Segment* s = ix->segs[g->first_seg + i];
        k_group_cursor<<<per_hash, 256, 0, st>>>(dir, start, (int)i, i == 0, ix->cursor.as<uint32_t>());
        if (s->n_entries > 0) {
            const int grid = (int)std::min<int64_t>((s->n_entries + 255) / 256, 148 * 16);
            k_scatter<<<grid, 256, 0, st>>>(s->st_hash.as<uint32_t>(), s->st_post.as<uint32_t>(), s->n_entries,
                                            ix->cursor.as<uint32_t>(), g->postings.as<uint32_t>());
        }
        e->launches += 2;
    }
    k_group_sort<<<(unsigned)((kBuckets + 255) / 256), 256, 0, st>>>(dir, (int)g->n_segs, g->postings.as<uint32_t>());
    e->launches += 4; }
    AID_CUDA(e, cudaGetLastError());
    g->n_entries = n;
    *built = true;
    return AID_OK;
}

int aid_index_commit_on(aid_engine* e, cudaStream_t st) {
    Index* ix = e->index;
    // 1. sealed (full) segments join the group of their octet; a group is rebuilt when it gains a member
    if (ix->grouping) {
        const size_t n_groups = (ix->segs.size() + kGroupSegs - 1) / kGroupSegs;
        if (ix->groups.size() < n_groups) ix->groups.resize(n_groups, nullptr);
        for (size_t k = 0; k < n_groups; k++) {
            uint32_t sealed = 0;
            while (sealed < (uint32_t)kGroupSegs && k * kGroupSegs + sealed < ix->segs.size() &&
                   ix->segs[k * kGroupSegs + sealed]->n_tracks == AID_SEG_TRACKS) sealed++;
            SegGroup* g = ix->groups[k];
            if (sealed == 0 || (g && (g->n_segs == sealed || g->n_segs == 0xffffffffu))) continue;
            if (!g) { g = new SegGroup(); ix->groups[k] = g; }
            g->first_seg = (uint32_t)(k * kGroupSegs); g->n_segs = sealed;
            bool built = false;
            int rc = build_group(e, ix, g, st, &built);
            if (rc) return rc;
            if (!built) { g->n_segs = 0xffffffffu; continue; }          // not groupable: do not try again
            AID_CUDA(e, cudaStreamSynchronize(st));
            for (uint32_t i = 0; i < sealed; i++) {
                Segment* s = ix->segs[g->first_seg + i];
                s->group = (int)k; s->sub = (int)i; s->dirty = false;
                s->bucket.release(); s->postings.release();
            }
            ix->segdesc_dirty = true;
        }
    }
    // 2. everything else keeps its own table
    for (Segment* s : ix->segs)
        if (s->group < 0 && (s->dirty || !s->bucket.p)) { int rc = build_segment(e, ix, s, st); if (rc) return rc; }
    if (ix->segdesc_dirty) {
        std::vector<aid_seg_desc> d(ix->segs.size());
        for (size_t i = 0; i < d.size(); i++) {
            const Segment* s = ix->segs[i];
            const SegGroup* g = s->group >= 0 ? ix->groups[s->group] : nullptr;
            d[i].bucket = g ? nullptr : s->bucket.as<uint32_t>();
            d[i].postings = g ? g->postings.as<uint32_t>() : s->postings.as<uint32_t>();
            d[i].dir = g ? g->dir.as<uint32_t>() : nullptr;
            d[i].sub = (uint32_t)s->sub;
            d[i].tomb = s->tomb.as<uint32_t>();
            d[i].first_track = s->first_track;
            d[i].n_tracks = s->n_tracks;
            d[i].n_deleted = s->n_deleted;
        }
        AID_CUDA(e, ix->d_segdesc.ensure(std::max<size_t>(d.size(), 1) * sizeof(aid_seg_desc)));
        AID_CUDA(e, cudaStreamSynchronize(st));
        if (!d.empty()) AID_CUDA(e, cudaMemcpy(ix->d_segdesc.p, d.data(), d.size() * sizeof(aid_seg_desc), cudaMemcpyHostToDevice));
        ix->segdesc_dirty = false;
    }
    return AID_OK;
}

// ----------------------------------------------------------------------------------------- C ABI
// fp_*: optional host copies of the fingerprints that were stored (aid_index_add_host_fp; the caller journals them)
static int add_pcm(aid_engine* e, const float* pcm, bool on_device, const int64_t* sample_off, int n_tracks,
                   const char* const* names, uint8_t* ok, uint32_t* fp_hash = nullptr, uint32_t* fp_t = nullptr,
                   int64_t fp_cap = 0, int64_t* fp_off = nullptr) {
    if (!e || !sample_off || !names || !ok || n_tracks < 0) return AID_E_ARG;
    if (n_tracks > 0 && !pcm && sample_off[n_tracks] > sample_off[0]) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    int64_t fp_pos = 0;
    if (fp_off) fp_off[0] = 0;
    for (int first = 0; first < n_tracks;) {
        int64_t frames = 0; int count = 0;
        while (first + count < n_tracks) {
            const int64_t T = aid_num_frames(sample_off[first + count + 1] - sample_off[first + count]);
            if (count > 0 && frames + T > e->max_batch_frames) break;
            frames += T; count++;
        }
        Plan plan;
        int rc = aid_build_plan(sample_off, first, count, AID_INDEX_MAX_FRAMES, plan);
        if (rc) return rc;
        const int64_t samples = sample_off[first + count] - sample_off[first];
        if ((rc = aid_slot_prepare(e, s, plan, !on_device, samples))) return rc;
        const float* d_pcm = pcm + sample_off[first];
        if (!on_device) {
            if (samples > 0) AID_CUDA(e, cudaMemcpyAsync(s.pcm.p, pcm + sample_off[first], (size_t)samples * 4, cudaMemcpyHostToDevice, s.st));
            d_pcm = s.pcm.as<float>();
        }
        if ((rc = aid_run_fingerprint(e, s, plan, d_pcm, s.st))) return rc;
        std::vector<uint32_t> h_off(count + 1);
        std::vector<int32_t> h_st(count);
        AID_CUDA(e, cudaMemcpyAsync(h_off.data(), s.hash_off.p, (size_t)(count + 1) * 4, cudaMemcpyDeviceToHost, s.st));
        AID_CUDA(e, cudaMemcpyAsync(h_st.data(), s.status.p, (size_t)count * 4, cudaMemcpyDeviceToHost, s.st));
        AID_CUDA(e, cudaStreamSynchronize(s.st));
        for (int i = 0; i < count; i++) h_st[i] |= plan.host_status[i];
        // frames of too-long tracks were planned as 0; report their true length so they are refused
        std::vector<int64_t> nf(count);
        for (int i = 0; i < count; i++) nf[i] = aid_num_frames(sample_off[first + i + 1] - sample_off[first + i]);
        if ((rc = aid_index_append(e, s.hash.as<uint32_t>(), s.t.as<uint32_t>(), h_off.data(), h_st.data(), nf.data(),
                                   count, names + first, ok + first, s.st))) return rc;
        if (fp_off) {                                  // the sub-batch's dense fingerprints, device order, one copy each
            const int64_t total = h_off[count];
            if (fp_pos + total > fp_cap) return AID_E_CAPACITY;
            if (total > 0) {
                AID_CUDA(e, cudaMemcpyAsync(fp_hash + fp_pos, s.hash.p, (size_t)total * 4, cudaMemcpyDeviceToHost, s.st));
                AID_CUDA(e, cudaMemcpyAsync(fp_t + fp_pos, s.t.p, (size_t)total * 4, cudaMemcpyDeviceToHost, s.st));
                AID_CUDA(e, cudaStreamSynchronize(s.st));
            }
            for (int i = 0; i < count; i++) fp_off[first + i + 1] = fp_pos + h_off[i + 1];
            fp_pos += total;
        }
        first += count;
    }
    return AID_OK;
}

extern "C" int aid_index_add_host_fp(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                                     const char* const* names, uint8_t* ok, uint32_t* hash, uint32_t* t_anchor,
                                     int64_t hash_cap, int64_t* hash_off) {
    if (!hash_off || hash_cap < 0 || (hash_cap > 0 && (!hash || !t_anchor))) return AID_E_ARG;
    return add_pcm(e, pcm, false, sample_off, n_tracks, names, ok, hash, t_anchor, hash_cap, hash_off);
}

extern "C" int aid_index_add_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                                  const char* const* names, uint8_t* ok) {
    return add_pcm(e, pcm, false, sample_off, n_tracks, names, ok);
}
extern "C" int aid_index_add_dev(aid_engine* e, const float* d_pcm, const int64_t* sample_off, int n_tracks,
                                 const char* const* names, uint8_t* ok) {
    return add_pcm(e, d_pcm, true, sample_off, n_tracks, names, ok);
}

extern "C" int aid_index_add_hashes(aid_engine* e, const uint32_t* hash, const uint32_t* t_anchor,
                                    const int64_t* hash_off, const int64_t* n_frames, int n_tracks,
                                    const char* const* names, uint8_t* ok) {
    if (!e || !hash_off || !n_frames || !names || !ok || n_tracks < 0) return AID_E_ARG;
    // This entry point is fed from disk (journal replay) and from other ranks: nothing unchecked reaches the device.
    // A hash outside AID_HASH_BITS would index past the bucket table / directory in k_hist, a t_anchor outside the
    // track would spill into the local-track bits of the posting.
    for (int i = 0; i < n_tracks; i++) {
        if (hash_off[i + 1] < hash_off[i]) return AID_E_ARG;
        if (hash_off[i + 1] > hash_off[i] && (!hash || !t_anchor)) return AID_E_ARG;
        if (n_frames[i] > AID_INDEX_MAX_FRAMES) continue;      // refused by aid_index_append (ok[i] = 0), never copied
        const int64_t t_lim = std::max<int64_t>(n_frames[i], 0);
        for (int64_t j = hash_off[i]; j < hash_off[i + 1]; j++)
            if ((hash[j] >> AID_HASH_BITS) != 0 || (int64_t)t_anchor[j] >= t_lim) return AID_E_ARG;
    }
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    const int chunk = 4096;
    for (int first = 0; first < n_tracks; first += chunk) {
        const int count = std::min(chunk, n_tracks - first);
        const int64_t h0 = hash_off[first], total = hash_off[first + count] - h0;
        if (total < 0 || total >= ((int64_t)1 << 32)) return AID_E_ARG;
        if (total > 0 && (!hash || !t_anchor)) return AID_E_ARG;
        AID_CUDA(e, s.hash.ensure((size_t)std::max<int64_t>(total, 1) * 4));
        AID_CUDA(e, s.t.ensure((size_t)std::max<int64_t>(total, 1) * 4));
        if (total > 0) {
            AID_CUDA(e, cudaMemcpyAsync(s.hash.p, hash + h0, (size_t)total * 4, cudaMemcpyHostToDevice, s.st));
            AID_CUDA(e, cudaMemcpyAsync(s.t.p, t_anchor + h0, (size_t)total * 4, cudaMemcpyHostToDevice, s.st));
        }
        std::vector<uint32_t> h_off(count + 1);
        std::vector<int32_t> h_st(count, 0);
        for (int i = 0; i <= count; i++) {
            if (hash_off[first + i] < h0) return AID_E_ARG;
            h_off[i] = (uint32_t)(hash_off[first + i] - h0);
        }
        int rc = aid_index_append(e, s.hash.as<uint32_t>(), s.t.as<uint32_t>(), h_off.data(), h_st.data(), n_frames + first,
                                  count, names + first, ok + first, s.st);
        if (rc) return rc;
    }
    return AID_OK;
}

extern "C" int aid_index_delete(aid_engine* e, const char* name) {
    if (!e || !name) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Index* ix = e->index;
    auto it = ix->by_name.find(name);
    if (it == ix->by_name.end()) return AID_E_NOT_FOUND;
    const uint32_t track = it->second;
    ix->by_name.erase(it);
    return set_tombstone(e, ix, track);
}

extern "C" int aid_index_commit(aid_engine* e) {
    if (!e) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    int rc = aid_index_commit_on(e, e->slot[0].st);
    if (rc) return rc;
    AID_CUDA(e, cudaStreamSynchronize(e->slot[0].st));
    return AID_OK;
}

extern "C" int aid_index_clear(aid_engine* e) {
    if (!e) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    AID_CUDA(e, cudaDeviceSynchronize());
    const bool grouping = e->index->grouping;
    aid_index_free(e, e->index);
    e->index = aid_index_new();
    e->index->grouping = grouping;
    return AID_OK;
}

extern "C" int aid_index_set_grouping(aid_engine* e, int on) {
    if (!e) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Index* ix = e->index;
    if (ix->grouping == (on != 0)) return AID_OK;
    AID_CUDA(e, cudaDeviceSynchronize());
    ix->grouping = on != 0;
    if (!on) {                                   // back to plain segments: their tables are rebuilt by the next commit
        for (SegGroup* g : ix->groups) if (g) { g->release(); delete g; }
        ix->groups.clear();
        for (Segment* s : ix->segs) if (s->group >= 0) { s->group = -1; s->sub = 0; s->dirty = true; }
    }
    ix->segdesc_dirty = true;
    return AID_OK;
}

extern "C" int aid_index_stats(aid_engine* e, int64_t* out) {
    if (!e || !out) return AID_E_ARG;
    Index* ix = e->index;
    out[0] = ix->live_tracks; out[1] = ix->n_postings; out[2] = (int64_t)ix->segs.size(); out[3] = (int64_t)ix->tracks.size();
    int64_t bytes = (int64_t)ix->cursor.cap + ix->scan_tmp.cap;
    for (Segment* s : ix->segs) bytes += s->st_hash.cap + s->st_post.cap + s->bucket.cap + s->postings.cap + s->tomb.cap;
    int64_t grouped = 0;
    for (SegGroup* g : ix->groups) if (g && g->dir.p) { bytes += g->dir.cap + g->postings.cap; grouped += g->n_segs; }
    bytes += ix->group_start.cap;
    out[5] = grouped;
    out[4] = bytes;
    return AID_OK;
}

extern "C" int aid_index_track_name(aid_engine* e, uint32_t track, char* buf, int buf_len) {
    if (!e || !buf || buf_len < 1) return AID_E_ARG;
    Index* ix = e->index;
    if (track >= ix->tracks.size()) return AID_E_NOT_FOUND;
    const std::string& s = ix->tracks[track].name;
    if ((int)s.size() + 1 > buf_len) return AID_E_CAPACITY;
    memcpy(buf, s.c_str(), s.size() + 1);
    return (int)s.size();
}

// ------------------------------------------------------------------------------------ persistence
// One file, <dir>/aidx_b200.bin: header, track table, then per segment its (hash, posting) entries.
// The sorted form is rebuilt on load. Written to a temp name and renamed, so a crash leaves the old file.
namespace {
struct FileHeader { char magic[8]; uint32_t version, n_segments; uint64_t n_tracks, n_postings; };
const char kMagic[8] = {'A', 'I', 'D', 'X', 'B', '2', '0', '0'};
bool wr(FILE* f, const void* p, size_t n) { return n == 0 || fwrite(p, 1, n, f) == n; }
bool rd(FILE* f, void* p, size_t n) { return n == 0 || fread(p, 1, n, f) == n; }
}

extern "C" int aid_index_save(aid_engine* e, const char* dir) {
    if (!e || !dir) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    AID_CUDA(e, cudaDeviceSynchronize());
    Index* ix = e->index;
    mkdir(dir, 0777);
    const std::string tmp = std::string(dir) + "/aidx_b200.bin.tmp", fin = std::string(dir) + "/aidx_b200.bin";
    FILE* f = fopen(tmp.c_str(), "wb");
    if (!f) { e->err = std::string("cannot write ") + tmp + ": " + strerror(errno); return AID_E_IO; }
    FileHeader h; memcpy(h.magic, kMagic, 8); h.version = 1; h.n_segments = (uint32_t)ix->segs.size();
    h.n_tracks = ix->tracks.size(); h.n_postings = (uint64_t)ix->n_postings;
    bool good = wr(f, &h, sizeof h);
    for (const TrackInfo& t : ix->tracks) {
        const uint32_t len = (uint32_t)t.name.size(); const uint8_t del = t.deleted;
        good = good && wr(f, &len, 4) && wr(f, t.name.data(), len) && wr(f, &t.n_frames, 8) && wr(f, &t.n_hashes, 4) && wr(f, &del, 1);
    }
    // entries of deleted tracks are dropped on the way out (the track slots stay, so local numbers keep their meaning):
    // a snapshot never carries tombstoned postings forward
    std::vector<uint32_t> hb, pb;
    uint64_t kept_total = 0;
    for (Segment* s : ix->segs) {
        const uint64_t n = (uint64_t)s->n_entries; const uint32_t nt = s->n_tracks;
        hb.resize((size_t)n); pb.resize((size_t)n);
        if (n && (cudaMemcpy(hb.data(), s->st_hash.p, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
                  cudaMemcpy(pb.data(), s->st_post.p, (size_t)n * 4, cudaMemcpyDeviceToHost) != cudaSuccess)) {
            fclose(f); return aid_fail_cuda(e, cudaGetLastError(), "save D2H");
        }
        uint64_t kept = n;
        if (s->n_deleted) {
            kept = 0;
            for (uint64_t i = 0; i < n; i++) {
                const uint32_t local = pb[(size_t)i] >> AID_POST_T_BITS;
                if (s->h_tomb[local >> 5] & (1u << (local & 31))) continue;
                hb[(size_t)kept] = hb[(size_t)i]; pb[(size_t)kept] = pb[(size_t)i]; kept++;
            }
        }
        good = good && wr(f, &nt, 4) && wr(f, &kept, 8) && wr(f, hb.data(), (size_t)kept * 4) && wr(f, pb.data(), (size_t)kept * 4);
        kept_total += kept;
    }
    if (good && kept_total != (uint64_t)ix->n_postings) {       // the header was written with the uncompacted count
        h.n_postings = kept_total;
        good = fseek(f, 0, SEEK_SET) == 0 && wr(f, &h, sizeof h) && fseek(f, 0, SEEK_END) == 0;
    }
    // durable before it replaces the old snapshot (the caller drops its journal right after): data, then the rename
    good = good && fflush(f) == 0 && fsync(fileno(f)) == 0;
    good = (fclose(f) == 0) && good;
    if (!good || rename(tmp.c_str(), fin.c_str()) != 0) { e->err = std::string("cannot write ") + fin + ": " + strerror(errno); remove(tmp.c_str()); return AID_E_IO; }
    const int dfd = open(dir, O_RDONLY | O_DIRECTORY);
    if (dfd >= 0) { fsync(dfd); close(dfd); }
    return AID_OK;
}

extern "C" int aid_index_load(aid_engine* e, const char* dir) {
    if (!e || !dir) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    const std::string fin = std::string(dir) + "/aidx_b200.bin";
    FILE* f = fopen(fin.c_str(), "rb");
    if (!f) {
        if (errno == ENOENT) return aid_index_clear(e);          // emptied directory == empty index (Makefile:84-93)
        e->err = std::string("cannot read ") + fin + ": " + strerror(errno);
        return AID_E_IO;
    }
    FileHeader h;
    if (!rd(f, &h, sizeof h) || memcmp(h.magic, kMagic, 8) != 0 || h.version != 1) { fclose(f); return AID_E_FORMAT; }
    int rc = aid_index_clear(e);
    if (rc) { fclose(f); return rc; }
    Index* ix = e->index;
    bool good = true;
    ix->tracks.resize((size_t)h.n_tracks);
    for (size_t i = 0; i < ix->tracks.size() && good; i++) {
        uint32_t len = 0; uint8_t del = 0; TrackInfo& t = ix->tracks[i];
        good = rd(f, &len, 4) && len < 4096;
        if (!good) break;
        t.name.resize(len);
        good = rd(f, &t.name[0], len) && rd(f, &t.n_frames, 8) && rd(f, &t.n_hashes, 4) && rd(f, &del, 1);
        t.deleted = del != 0;
        if (!t.deleted) { ix->by_name[t.name] = (uint32_t)i; ix->live_tracks++; }
    }
    std::vector<uint32_t> buf;
    for (uint32_t si = 0; si < h.n_segments && good; si++) {
        uint32_t nt = 0; uint64_t n = 0;
        good = rd(f, &nt, 4) && rd(f, &n, 8) && nt <= AID_SEG_TRACKS;
        if (!good) break;
        Segment* s = new Segment();
        s->first_track = si * AID_SEG_TRACKS; s->n_tracks = nt; s->n_entries = (int64_t)n; s->dirty = true;
        s->h_tomb.assign(AID_SEG_TRACKS / 32, 0);
        ix->segs.push_back(s);
        for (uint32_t l = 0; l < nt; l++)
            if (s->first_track + l < ix->tracks.size() && ix->tracks[s->first_track + l].deleted) { s->h_tomb[l / 32] |= 1u << (l % 32); s->n_deleted++; }
        cudaError_t ce = s->tomb.ensure(AID_SEG_TRACKS / 8);
        if (ce == cudaSuccess) ce = cudaMemcpy(s->tomb.p, s->h_tomb.data(), AID_SEG_TRACKS / 8, cudaMemcpyHostToDevice);
        buf.resize((size_t)n);
        for (DevBuf* dst : {&s->st_hash, &s->st_post}) {
            good = good && rd(f, buf.data(), (size_t)n * 4);
            // a damaged file must not reach the device: hashes index the 2^24-entry tables, postings name local tracks
            if (good) {
                const uint64_t lim = dst == &s->st_hash ? ((uint64_t)1 << AID_HASH_BITS) : ((uint64_t)nt << AID_POST_T_BITS);
                for (uint64_t i = 0; i < n && good; i++) good = (uint64_t)buf[(size_t)i] < lim;
            }
            if (ce == cudaSuccess) ce = dst->ensure(std::max<size_t>((size_t)n, 1) * 4);
            if (ce == cudaSuccess && n && good) ce = cudaMemcpy(dst->p, buf.data(), (size_t)n * 4, cudaMemcpyHostToDevice);
        }
        if (ce != cudaSuccess) { fclose(f); aid_index_clear(e); return aid_fail_cuda(e, ce, "load H2D"); }
        ix->n_postings += (int64_t)n;
    }
    fclose(f);
    if (!good) { aid_index_clear(e); return AID_E_FORMAT; }
    ix->segdesc_dirty = true;
    return aid_index_commit(e);
}
