// match.cu -- hash-index probe + per-track time-offset histogram vote + top-k, sm_100a.
//
// Replaces stage a9 of SURVEY.md section 8(a): the lookup and tally inside `olaf_c query`
// (reference audio-ident-service/app/audio/fingerprint.py:185-193); rows carry what one CSV line of its
// output carries (fingerprint.py:273-277). Definition of the result: oracle/aid_oracle.c aid_oracle_match();
// bit-exact rows (count, track, offset, q_first, q_last) in the same order.
//
// k_match: one CTA per (query window, index segment).
//   0. stage the query's hashes: run begin / length per hash (from the segment's own table or from the directory it
//      shares with up to seven sibling segments, index.h SegGroup), prefix sum of the lengths -> a flat list of votes,
//      so every thread handles the same number of postings whatever the bucket sizes are; a u16 vote -> hash map in
//      the idle sort buffers saves a binary search per vote;
//   1. pass 1 adds every vote key (local_track << 18 | t_ref - t_query + bias: one integer add on the posting)
//      into a 2048-counter shared-memory sketch with atomics (a CTA sees ~1,000 votes; <= 192 go straight to 2.);
//   2. pass 2 -- skipped when no counter reached AID_MIN_VOTES, which is the usual case -- re-reads the postings
//      (now in L1/L2) and inserts only keys whose sketch counter reached AID_MIN_VOTES into an exact shared-memory
//      hash table (atomicCAS on the key, atomicAdd on the count, atomicMin/Max on the query time);
//   3. exact counts >= AID_MIN_VOTES are merged into the CTA's running top-50 with a bitonic network.
//   Windows with more votes than the sketch can filter are processed in R rounds over a hash partition of
//   the keys; if the exact table still fills up the CTA doubles R and starts over. Exactness never depends on
//   the sketch: it only rejects keys that cannot reach the threshold.
// k_rank: one CTA per query merges the per-segment top-50 lists (a track lives in exactly one segment, so no
//   partial counts ever need adding) and writes the final rows.
// HBM traffic per query window: one table / directory entry per hash and 4 B per posting touched, random access.
#include <algorithm>
#include <climits>
#include <cstdint>
#include "engine.h"
#include "index.h"

namespace {

constexpr int kThreads = 128;
constexpr int kSketch = 2048;
constexpr int kTable = 128;
constexpr int kTableMaxLoad = 96;
constexpr int kQChunk = 256;               // query hashes staged at a time (a 3.5 s window has ~110)
constexpr int kVotesPerRound = 2048;       // = kSketch: ~1 vote per counter keeps chance counts >= AID_MIN_VOTES at ~1 per round; a (window, segment) CTA
                                            // sees ~1,000 votes, so the usual CTA needs one round, not two (each round reads every posting twice)
constexpr int kBest = 64;                 // >= AID_MAX_ROWS, power of two
constexpr int kSortN = 256;               // kTable + kBest <= kSortN
constexpr uint32_t kEmpty = 0xffffffffu;
constexpr uint64_t kPad = ~0ull;

struct CandEntry { uint32_t inv_count; uint32_t key; uint32_t tq; };   // tq = q_first | q_last << 16

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}

template <int N, int T>
__device__ __forceinline__ void bitonic_sort(uint64_t* key, uint32_t* val, int tid) {
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < N; i += T) {
                const int p = i ^ j;
                if (p > i) {
                    const uint64_t a = key[i], b = key[p];
                    if ((a > b) == ((i & k) == 0)) {
                        key[i] = b; key[p] = a;
                        const uint32_t t = val[i]; val[i] = val[p]; val[p] = t;
                    }
                }
            }
            __syncthreads();
        }
}

// same network on the first n slots only (n a power of two <= the array size; slots >= n are not touched)
template <int T>
__device__ __forceinline__ void bitonic_sort_n(uint64_t* key, uint32_t* val, int n, int tid) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n; i += T) {
                const int p = i ^ j;
                if (p > i) {
                    const uint64_t a = key[i], b = key[p];
                    if ((a > b) == ((i & k) == 0)) {
                        key[i] = b; key[p] = a;
                        const uint32_t t = val[i]; val[i] = val[p]; val[p] = t;
                    }
                }
            }
            __syncthreads();
        }
}

constexpr int kOwnerCap = kSortN * 6;       // u16 entries that fit the sort buffers (skey + sval), which are idle during the vote loops
struct MatchSmem {
    alignas(16) uint32_t sketch[kSketch / 2];      // 16-bit counters, two per word: a round never holds more than 2048 votes
    alignas(16) uint32_t tkey[kTable], tcnt[kTable], tmin[kTable], tmax[kTable];
    uint32_t qbeg[kQChunk], qadd[kQChunk], qstart[kQChunk + 1];
    alignas(16) uint64_t skey[kSortN];     // skey + sval double as the vote -> hash map `owner` (u16[kOwnerCap])
    uint32_t sval[kSortN];
    uint64_t bkey[kBest];
    uint32_t bval[kBest];
    uint32_t wsum[kThreads / 32];
    uint32_t total, used, overflow, nbest, hit, maybe;
};

// 17 KB of shared memory and <= 40 registers: 12 CTAs (48 warps) per SM. The kernel waits on chains of dependent random
// loads (hash -> directory entry -> posting run), so resident CTAs are what hides them.
__global__ void __launch_bounds__(kThreads, 16)
k_match(const uint32_t* __restrict__ q_hash, const uint32_t* __restrict__ q_t, const uint32_t* __restrict__ hash_off,
        const uint32_t* __restrict__ hash_len, const int32_t* __restrict__ q_status,
        const aid_seg_desc* __restrict__ segs, int n_seg,
        CandEntry* __restrict__ cand, uint32_t* __restrict__ cand_n,
        const uint32_t* __restrict__ abort_flag, unsigned long long* __restrict__ vote_stats) {
    __shared__ MatchSmem sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int q = blockIdx.x / n_seg, sg = blockIdx.x % n_seg;
    // sharded identification: the wait for the peers' query fingerprints gave up (k_wait_hashes) -- the hash window is
    // only partly written, so nothing is probed; k_rank still publishes (empty) blocks and the merge reports -1 rows
    if (abort_flag && *abort_flag) {
        if (tid == 0) cand_n[blockIdx.x] = 0;
        return;
    }
    const aid_seg_desc seg = segs[sg];
    const uint32_t h0 = hash_off[q];
    const uint32_t nh = (q_status && q_status[q] != 0) ? 0u : (hash_len ? hash_len[q] : hash_off[q + 1] - h0);
    const uint32_t* __restrict__ bucket = seg.bucket;
    const uint32_t* __restrict__ postings = seg.postings;
    const bool any_deleted = seg.n_deleted != 0;          // no tombstone lookup (a dependent load per vote) in clean segments
    // Where this segment keeps the postings of hash h. A plain segment has its own table of 2^24 + 1 offsets; a grouped
    // one (index.h SegGroup) shares a directory with up to 7 siblings: one 32-byte entry per hash, the same sector for
    // the sibling CTAs of this window (L2 hits), and sibling runs of a hash are adjacent in the posting array.
    const uint32_t* __restrict__ dir = seg.dir;
    const uint32_t sub = seg.sub;
    auto run_of = [&](uint32_t h, uint32_t& lo) -> uint32_t {
        if (dir) {
            const uint4 a = *reinterpret_cast<const uint4*>(dir + (size_t)h * 8);
            const uint32_t c67 = dir[(size_t)h * 8 + 4];
            uint32_t at = a.x, len = 0;
#pragma unroll
            for (uint32_t j = 0; j < (uint32_t)kGroupSegs; j++) {      // unrolled and predicated: the words stay in registers
                const uint32_t word = j < 2 ? a.y : j < 4 ? a.z : j < 6 ? a.w : c67;
                const uint32_t c = (word >> (16 * (j & 1))) & 0xffffu;
                at += j < sub ? c : 0u;
                len = j == sub ? c : len;
            }
            lo = at;
            return len;
        }
        lo = bucket[h];
        return bucket[h + 1] - lo;
    };

    // Stages hashes [c0, c0 + nc) of the window: bucket begin, time bias and the exclusive prefix of the bucket lengths,
    // so that the votes of the chunk form one flat list; returns their number (the same in every thread).
    auto stage = [&](uint32_t c0, uint32_t nc) -> uint32_t {
        uint32_t run = 0;                       // votes of earlier strides of this chunk
        for (uint32_t b = 0; b < nc; b += kThreads) {
            const uint32_t i = b + tid;
            uint32_t len = 0;
            if (i < nc) {
                uint32_t lo;
                len = run_of(q_hash[h0 + c0 + i], lo);
                sm.qbeg[i] = lo;
                sm.qadd[i] = AID_QUERY_MAX_FRAMES - q_t[h0 + c0 + i];
            }
            uint32_t incl = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t o = __shfl_up_sync(AID_FULL_MASK, incl, d);
                if (lane >= d) incl += o;
            }
            if (lane == 31) sm.wsum[warp] = incl;
            __syncthreads();
            uint32_t before = 0, all = 0;
#pragma unroll
            for (int w = 0; w < kThreads / 32; w++) { const uint32_t c = sm.wsum[w]; all += c; before += w < warp ? c : 0; }
            if (i < nc) sm.qstart[i] = run + before + incl - len;
            run += all;
            __syncthreads();
        }
        if (tid == 0) sm.qstart[nc] = run;
        __syncthreads();
        return run;
    };

    // ---- how many votes does this (window, segment) produce? A window that fits one chunk (the usual case: a 3.5 s
    // window has ~300 hashes) is staged once, here, and its bucket table entries are read once for all passes.
    const bool single = nh <= (uint32_t)kQChunk;
    uint32_t total;
    if (single) {
        total = stage(0, nh);
    } else {
        uint32_t mine = 0;
        for (uint32_t i = tid; i < nh; i += kThreads) { uint32_t lo; mine += run_of(q_hash[h0 + i], lo); }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) mine += __shfl_xor_sync(AID_FULL_MASK, mine, d);
        if (lane == 0) sm.wsum[warp] = mine;
        __syncthreads();
        if (tid == 0) {
            uint32_t t = 0;
            for (int w = 0; w < kThreads / 32; w++) t += sm.wsum[w];
            sm.total = t;
        }
        __syncthreads();
        total = sm.total;
    }
    // bench.py's roofline numerator: hashes probed and postings touched, spread over 1024 counters
    if (vote_stats && tid == 0) {
        atomicAdd(vote_stats + 2 * (blockIdx.x & 1023u), (unsigned long long)nh);
        atomicAdd(vote_stats + 2 * (blockIdx.x & 1023u) + 1, (unsigned long long)total);
    }
    if (total < AID_MIN_VOTES) {
        if (tid == 0) cand_n[blockIdx.x] = 0;
        return;
    }
    uint32_t R = (total + kVotesPerRound - 1) / kVotesPerRound;
    // Few votes (the usual case for one window against one segment): they all fit the exact table, so the sketch
    // and its second pass over the postings are skipped. If a restart is needed the sketch comes back.
    bool direct = total <= kTableMaxLoad;

    for (;;) {                                        // restarted with a finer partition if the exact table fills
        if (tid == 0) { sm.nbest = 0; sm.overflow = 0; }
        for (uint32_t r = 0; r < R; r++) {
            if (!direct) for (int i = tid; i < kSketch / 8; i += kThreads) reinterpret_cast<uint4*>(sm.sketch)[i] = make_uint4(0, 0, 0, 0);
            for (int i = tid; i < kTable / 4; i += kThreads) {
                reinterpret_cast<uint4*>(sm.tkey)[i] = make_uint4(kEmpty, kEmpty, kEmpty, kEmpty);
                reinterpret_cast<uint4*>(sm.tcnt)[i] = make_uint4(0, 0, 0, 0);
                reinterpret_cast<uint4*>(sm.tmin)[i] = make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
                reinterpret_cast<uint4*>(sm.tmax)[i] = make_uint4(0, 0, 0, 0);
            }
            if (tid == 0) { sm.used = 0; sm.hit = 0; sm.maybe = 0; }
            __syncthreads();
            for (int pass = direct ? 1 : 0; pass < 2; pass++) {
                // no counter of the sketch reached AID_MIN_VOTES: no key can have, and the second pass over the postings
                // is skipped (the usual case: one segment in 62 holds the track a window comes from)
                if (pass == 1 && !direct && !sm.maybe) break;
                for (uint32_t c0 = 0; c0 < nh; c0 += kQChunk) {
                    const uint32_t nc = min((uint32_t)kQChunk, nh - c0);
                    const uint32_t nv = single ? total : stage(c0, nc);
                    // vote v belongs to the hash i with qstart[i] <= v < qstart[i + 1]: one table lookup when the chunk's
                    // votes fit the (idle) sort buffers, a binary search otherwise. The table is rebuilt per pass
                    // because the top-k merge at the end of a round sorts in the same memory.
                    uint16_t* owner = reinterpret_cast<uint16_t*>(sm.skey);
                    const bool mapped = nv <= (uint32_t)kOwnerCap;
                    if (mapped && (!single || pass == (direct ? 1 : 0))) {       // a single-chunk window keeps it for both passes
                        for (uint32_t i = tid; i < nc; i += kThreads)
                            for (uint32_t v = sm.qstart[i], ve = sm.qstart[i + 1]; v < ve; v++) owner[v] = (uint16_t)i;
                        __syncthreads();
                    }
                    // Four votes per thread and trip, their posting loads issued before any of them is counted: the
                    // kernel waits on these random loads (ncu: 6.8 long-scoreboard stalls per issued instruction at 39 %
                    // issue utilisation), so memory-level parallelism is what it needs, not fewer instructions.
                    constexpr int kUnroll = 4;
                    for (uint32_t v0 = tid; v0 < nv; v0 += kUnroll * kThreads) {
                        uint32_t los[kUnroll], posts[kUnroll];
#pragma unroll
                        for (int j = 0; j < kUnroll; j++) {
                            const uint32_t v = v0 + j * kThreads;
                            uint32_t lo = 0;
                            if (v < nv) {
                                if (mapped) lo = owner[v];
                                else { uint32_t hi = nc; while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sm.qstart[mid] <= v) lo = mid; else hi = mid; } }
                                posts[j] = postings[sm.qbeg[lo] + (v - sm.qstart[lo])];
                            }
                            los[j] = lo;
                        }
#pragma unroll
                        for (int j = 0; j < kUnroll; j++) {
                            if (v0 + j * kThreads >= nv) continue;
                            const uint32_t lo = los[j], post = posts[j];
                            const uint32_t local = post >> AID_POST_T_BITS;
                            if (any_deleted && (seg.tomb[local >> 5] & (1u << (local & 31)))) continue;
                            const uint32_t key = post + sm.qadd[lo];
                            const uint32_t m = mix32(key);
                            if (R > 1 && (m >> 12) % R != r) continue;
                            const uint32_t idx = m & (kSketch - 1), sh = 16 * (idx & 1);
                            if (pass == 0) {
                                if (((atomicAdd(&sm.sketch[idx >> 1], 1u << sh) >> sh) & 0xffffu) + 1 == AID_MIN_VOTES) sm.maybe = 1;
                                continue;
                            }
                            if (!direct && ((sm.sketch[idx >> 1] >> sh) & 0xffffu) < AID_MIN_VOTES) continue;
                            const uint32_t tq = AID_QUERY_MAX_FRAMES - sm.qadd[lo];
                            uint32_t slot = (m >> 20) & (kTable - 1);
                            for (int probe = 0; probe < kTable; probe++) {
                                const uint32_t old = atomicCAS(&sm.tkey[slot], kEmpty, key);
                                if (old == kEmpty) { if (atomicAdd(&sm.used, 1u) >= kTableMaxLoad) sm.overflow = 1; }
                                if (old == kEmpty || old == key) {
                                    if (atomicAdd(&sm.tcnt[slot], 1u) + 1 == AID_MIN_VOTES) sm.hit = 1;   // a row is born
                                    atomicMin(&sm.tmin[slot], tq);
                                    atomicMax(&sm.tmax[slot], tq);
                                    break;
                                }
                                slot = (slot + 1) & (kTable - 1);
                                if (probe == kTable - 1) sm.overflow = 1;
                            }
                        }
                    }
                    __syncthreads();
                }
            }
            if (sm.overflow) break;
            if (!sm.hit) continue;                           // nothing reached AID_MIN_VOTES this round (the usual case)
            // ---- merge this round's exact counts into the running top-kBest
            const uint32_t nb = sm.nbest;
            for (int i = tid; i < kSortN; i += kThreads) {
                uint64_t k = kPad; uint32_t v = 0;
                if (i < kTable) {
                    if (sm.tkey[i] != kEmpty && sm.tcnt[i] >= AID_MIN_VOTES) {
                        k = ((uint64_t)(0xffffffffu - sm.tcnt[i]) << 32) | sm.tkey[i];
                        v = (sm.tmin[i] & 0xffffu) | (sm.tmax[i] << 16);
                    }
                } else if (i - kTable < (int)nb) { k = sm.bkey[i - kTable]; v = sm.bval[i - kTable]; }
                sm.skey[i] = k; sm.sval[i] = v;
            }
            __syncthreads();
            bitonic_sort<kSortN, kThreads>(sm.skey, sm.sval, tid);
            if (tid < kBest) { sm.bkey[tid] = sm.skey[tid]; sm.bval[tid] = sm.sval[tid]; }
            if (tid == 0) {
                uint32_t n = 0;
                while (n < kBest && sm.skey[n] != kPad) n++;
                sm.nbest = n;
            }
            __syncthreads();
        }
        __syncthreads();
        if (!sm.overflow) break;
        __syncthreads();
        if (direct) direct = false; else R *= 2;
    }

    const uint32_t n = min(sm.nbest, (uint32_t)AID_MAX_ROWS);
    CandEntry* out = cand + (int64_t)blockIdx.x * AID_MAX_ROWS;
    if (tid < (int)n) {
        CandEntry c;
        c.inv_count = (uint32_t)(sm.bkey[tid] >> 32);
        c.key = (uint32_t)sm.bkey[tid];
        c.tq = sm.bval[tid];
        out[tid] = c;
    }
    if (tid == 0) cand_n[blockIdx.x] = n;
}

// ---- per query: merge the segments' lists, order by (count desc, track asc, offset asc), write rows
constexpr int kRankN = 1024;

__global__ void __launch_bounds__(kThreads)
k_rank(const CandEntry* __restrict__ cand, const uint32_t* __restrict__ cand_n, const aid_seg_desc* __restrict__ segs,
       int n_seg, int max_rows, aid_match_row* __restrict__ rows, int32_t* __restrict__ n_rows, const RowSink sink) {
    __shared__ uint64_t skey[kRankN];
    __shared__ uint32_t sval[kRankN];       // index of the entry in cand
    __shared__ uint32_t s_fill, s_best;
    const int tid = threadIdx.x, q = blockIdx.x;
    if (tid == 0) s_best = 0;
    // Most windows have no candidate row in most shards (a query's track lives in one segment of one rank): count first and
    // skip the fill / sort rounds -- two barriers per segment -- when there is nothing to rank.
    uint32_t any = 0;
    for (int s2 = tid; s2 < n_seg; s2 += kThreads) any |= cand_n[(int64_t)q * n_seg + s2];
    const bool nothing = __syncthreads_or((int)any) == 0;
    int sg = nothing ? n_seg : 0;
    while (sg < n_seg) {
        // keep the best kBest so far in slots [0, kBest), fill the rest from the next segments
        const uint32_t nbest = s_best;
        for (int i = tid; i < kRankN; i += kThreads) if (i >= (int)nbest) { skey[i] = kPad; sval[i] = 0; }
        if (tid == 0) s_fill = nbest;
        __syncthreads();
        while (sg < n_seg) {
            const uint32_t n = cand_n[(int64_t)q * n_seg + sg];
            if (s_fill + n > kRankN) break;
            const uint32_t base = s_fill;
            if (tid < (int)n) {
                const int64_t ci = ((int64_t)q * n_seg + sg) * AID_MAX_ROWS + tid;
                const CandEntry c = cand[ci];
                const uint64_t inv = c.inv_count - (0xffffffffu - 0xfffffu);          // 20-bit inverted count
                const uint64_t track = (uint64_t)segs[sg].first_track + (c.key >> AID_POST_T_BITS);
                const uint64_t off = c.key & ((1u << AID_POST_T_BITS) - 1);
                skey[base + tid] = ((c.inv_count <= 0xffffffffu - 0xfffffu ? 0ull : inv) << 44) | (track << 19) | off;
                sval[base + tid] = (uint32_t)ci;
            }
            __syncthreads();
            if (tid == 0) s_fill = base + n;
            __syncthreads();
            sg++;
        }
        // most windows have a handful of candidate rows in one segment: sort only as many slots as are filled
        int width = 32;
        while (width < (int)s_fill) width <<= 1;
        bitonic_sort_n<kThreads>(skey, sval, width, tid);
        if (tid == 0) {
            uint32_t n = 0;
            while (n < kBest && n < (uint32_t)width && skey[n] != kPad) n++;
            s_best = n;
        }
        __syncthreads();
    }
    const uint32_t n = min(s_best, (uint32_t)max_rows);
    aid_match_row r;
    if (tid < (int)n) {
        const CandEntry c = cand[sval[tid]];
        const uint32_t sgi = (uint32_t)((sval[tid] / AID_MAX_ROWS) % n_seg);
        r.count = (int32_t)(0xffffffffu - c.inv_count);
        r.track = segs[sgi].first_track + (c.key >> AID_POST_T_BITS);
        r.offset = (int32_t)(c.key & ((1u << AID_POST_T_BITS) - 1)) - AID_QUERY_MAX_FRAMES;
        r.q_first = (int32_t)(c.tq & 0xffffu);
        r.q_last = (int32_t)(c.tq >> 16);
    }
    if (sink.lay.world == 0) {
        if (tid < (int)n) rows[(int64_t)q * max_rows + tid] = r;
        if (tid == 0) n_rows[q] = (int32_t)n;
        return;
    }
    // ---- sharded: deliver the rows, in global track numbers, to every rank's receive window (peer stores over NVLink).
    // The order inside a rank's block is by local track number; the receiver orders the union by global number.
    if (tid < (int)n && sink.track_map) r.track = r.track < sink.n_map ? sink.track_map[r.track] : 0xffffffffu;
    const int parity = (int)(sink.epoch & 1u);
    for (int p = 0; p < sink.lay.world; p++) {
        unsigned char* w = sink.window[p];
        if (tid < (int)n) sink.lay.rows(w, parity, sink.rank)[(int64_t)q * AID_MAX_ROWS + tid] = r;
        if (tid == 0) sink.lay.counts(w, parity, sink.rank)[q] = (int32_t)n;
    }
    if (tid < (int)n || tid == 0) __threadfence_system();     // the threads that stored to peers: their stores are visible system-wide ...
    __syncthreads();
    if (tid == 0) {
        const uint32_t prev = atomicAdd(sink.done, 1u);
        if (prev == gridDim.x - 1) {              // ... so when the last CTA gets here, the whole block is
            *sink.done = 0;
            __threadfence_system();
            for (int p = 0; p < sink.lay.world; p++) {
                uint32_t* flag = XchgLayout::row_flag(sink.window[p]) + sink.rank;
                asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(flag), "r"(sink.epoch) : "memory");
            }
        }
    }
}

}  // namespace

// Runs the matcher for n_q windows whose fingerprints are on the device: window q owns
// hash/t[hash_off[q] .. +len), len = hash_len[q] if given, else hash_off[q+1] - hash_off[q].
// Rows and counts are written to device memory (d_rows[n_q][max_rows], d_n_rows[n_q]); asynchronous on st.
int aid_match_device_out(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* d_hash_off,
                         const uint32_t* d_hash_len, const int32_t* d_status, int n_q, aid_match_row* d_rows,
                         int max_rows, int32_t* d_n_rows, const RowSink& sink, cudaStream_t st, const uint32_t* d_abort) {
    Index* ix = e->index;
    int rc = aid_index_commit_on(e, st);
    if (rc) return rc;
    const int n_seg = (int)ix->segs.size();
    if (n_q == 0) return AID_OK;
    if (n_seg == 0) {
        if (sink.lay.world == 0) { AID_CUDA(e, cudaMemsetAsync(d_n_rows, 0, (size_t)n_q * 4, st)); return AID_OK; }
        // an empty shard still has to publish its (empty) block: the peers wait for it
        k_rank<<<n_q, kThreads, 0, st>>>(nullptr, nullptr, nullptr, 0, max_rows, nullptr, nullptr, sink);
        AID_CUDA(e, cudaGetLastError());
        e->launches += 1;
        return AID_OK;
    }
    const int64_t n_cta = (int64_t)n_q * n_seg;
    if (n_cta >= ((int64_t)1 << 31)) return AID_E_ARG;
    AID_CUDA(e, ix->cand.ensure((size_t)n_cta * AID_MAX_ROWS * sizeof(CandEntry)));
    AID_CUDA(e, ix->cand_n.ensure((size_t)n_cta * 4));
    unsigned long long* stats = nullptr;
    if (e->timing) {                                  // probe statistics ride along with the stage timing (bench.py)
        if (!ix->vote_stats.p) {
            AID_CUDA(e, ix->vote_stats.ensure(2 * 1024 * sizeof(unsigned long long)));
            AID_CUDA(e, cudaMemsetAsync(ix->vote_stats.p, 0, 2 * 1024 * sizeof(unsigned long long), st));
        }
        stats = ix->vote_stats.as<unsigned long long>();
    }
    { StageTimer tm(e, st, 4);
    k_match<<<(unsigned)n_cta, kThreads, 0, st>>>(d_hash, d_t, d_hash_off, d_hash_len, d_status, ix->d_segdesc.as<aid_seg_desc>(), n_seg,
                                                  ix->cand.as<CandEntry>(), ix->cand_n.as<uint32_t>(), d_abort, stats); }
    { StageTimer tm(e, st, 5);
    k_rank<<<n_q, kThreads, 0, st>>>(ix->cand.as<CandEntry>(), ix->cand_n.as<uint32_t>(), ix->d_segdesc.as<aid_seg_desc>(), n_seg,
                                     max_rows, d_rows, d_n_rows, sink); }
    AID_CUDA(e, cudaGetLastError());
    e->launches += 2;
    return AID_OK;
}

// Same, rows copied to host buffers (synchronous).
static int match_device(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t, const uint32_t* d_hash_off,
                        const int32_t* d_status, int n_q, aid_match_row* rows, int max_rows, int32_t* n_rows,
                        cudaStream_t st) {
    Index* ix = e->index;
    if (n_q == 0) return AID_OK;
    AID_CUDA(e, ix->rows.ensure((size_t)n_q * max_rows * sizeof(aid_match_row)));
    AID_CUDA(e, ix->rows_n.ensure((size_t)n_q * 4));
    int rc = aid_match_device_out(e, d_hash, d_t, d_hash_off, nullptr, d_status, n_q, ix->rows.as<aid_match_row>(), max_rows,
                                  ix->rows_n.as<int32_t>(), RowSink{}, st, nullptr);
    if (rc) return rc;
    AID_CUDA(e, cudaMemcpyAsync(rows, ix->rows.p, (size_t)n_q * max_rows * sizeof(aid_match_row), cudaMemcpyDeviceToHost, st));
    AID_CUDA(e, cudaMemcpyAsync(n_rows, ix->rows_n.p, (size_t)n_q * 4, cudaMemcpyDeviceToHost, st));
    AID_CUDA(e, cudaStreamSynchronize(st));
    return AID_OK;
}

extern "C" int aid_match_dev(aid_engine* e, const uint32_t* d_hash, const uint32_t* d_t_anchor, const uint32_t* d_hash_off,
                             const uint32_t* d_hash_len, const int32_t* d_status, int n_queries,
                             aid_match_row* d_rows, int max_rows, int32_t* d_n_rows, void* stream) {
    if (!e || n_queries < 0 || max_rows < 1 || max_rows > AID_MAX_ROWS) return AID_E_ARG;
    if (n_queries > 0 && (!d_hash_off || !d_rows || !d_n_rows)) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    return aid_match_device_out(e, d_hash, d_t_anchor, d_hash_off, d_hash_len, d_status, n_queries, d_rows, max_rows, d_n_rows,
                                RowSink{}, stream ? (cudaStream_t)stream : e->slot[0].st, nullptr);
}

extern "C" int aid_match_stats(aid_engine* e, int64_t* out) {
    if (!e || !out) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Index* ix = e->index;
    out[0] = out[1] = 0;
    if (!ix->vote_stats.p) return AID_OK;
    AID_CUDA(e, cudaDeviceSynchronize());
    std::vector<unsigned long long> h(2 * 1024);
    AID_CUDA(e, cudaMemcpy(h.data(), ix->vote_stats.p, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 1024; i++) { out[0] += (int64_t)h[2 * i]; out[1] += (int64_t)h[2 * i + 1]; }
    AID_CUDA(e, cudaMemset(ix->vote_stats.p, 0, h.size() * sizeof(unsigned long long)));
    return AID_OK;
}

extern "C" int aid_copy_device(aid_engine* e, void* d_dst, const void* d_src, int64_t bytes, void* stream) {
    if (!e || bytes < 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    if (bytes > 0) AID_CUDA(e, cudaMemcpyAsync(d_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToDevice, stream ? (cudaStream_t)stream : e->slot[0].st));
    return AID_OK;
}

// window q is pcm[win_begin[q] .. win_end[q]); windows may overlap (then a sub-batch uploads the span they cover once)
static int query_pcm(aid_engine* e, const float* pcm, bool on_device, const int64_t* win_begin, const int64_t* win_end, int n_q,
                     aid_match_row* rows, int max_rows, int32_t* n_rows) {
    if (!e || !win_begin || !win_end || !rows || !n_rows || n_q < 0 || max_rows < 1 || max_rows > AID_MAX_ROWS) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    for (int i = 0; i < n_q; i++) {
        if (win_begin[i] < 0 || win_end[i] < win_begin[i] || (!pcm && win_end[i] > win_begin[i])) return AID_E_ARG;
        if (aid_num_frames(win_end[i] - win_begin[i]) > AID_QUERY_MAX_FRAMES) return AID_E_TOO_LONG;
    }
    Slot& s = e->slot[0];
    for (int first = 0; first < n_q;) {
        int64_t frames = 0; int count = 0;
        int64_t s0 = INT64_MAX, s1 = INT64_MIN;
        while (first + count < n_q) {
            const int64_t T = aid_num_frames(win_end[first + count] - win_begin[first + count]);
            if (count > 0 && frames + T > e->max_batch_frames) break;
            frames += T;
            s0 = std::min(s0, win_begin[first + count]); s1 = std::max(s1, win_end[first + count]);
            count++;
        }
        Plan plan;
        int rc = aid_build_plan_windows(win_begin + first, win_end + first, count, s0, AID_QUERY_MAX_FRAMES, plan);
        if (rc) return rc;
        const int64_t samples = s1 - s0;
        if ((rc = aid_slot_prepare(e, s, plan, !on_device, samples))) return rc;
        const float* d_pcm = pcm + s0;
        if (!on_device) {
            if (samples > 0) AID_CUDA(e, cudaMemcpyAsync(s.pcm.p, pcm + s0, (size_t)samples * 4, cudaMemcpyHostToDevice, s.st));
            d_pcm = s.pcm.as<float>();
        }
        if ((rc = aid_run_fingerprint(e, s, plan, d_pcm, s.st))) return rc;
        if ((rc = match_device(e, s.hash.as<uint32_t>(), s.t.as<uint32_t>(), s.hash_off.as<uint32_t>(), s.status.as<int32_t>(),
                               count, rows + (int64_t)first * max_rows, max_rows, n_rows + first, s.st))) return rc;
        first += count;
    }
    return AID_OK;
}

extern "C" int aid_query_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_queries,
                              aid_match_row* rows, int max_rows, int32_t* n_rows) {
    if (!sample_off) return AID_E_ARG;
    return query_pcm(e, pcm, false, sample_off, sample_off + 1, n_queries, rows, max_rows, n_rows);
}
extern "C" int aid_query_dev(aid_engine* e, const float* d_pcm, const int64_t* sample_off, int n_queries,
                             aid_match_row* rows, int max_rows, int32_t* n_rows) {
    if (!sample_off) return AID_E_ARG;
    return query_pcm(e, d_pcm, true, sample_off, sample_off + 1, n_queries, rows, max_rows, n_rows);
}
extern "C" int aid_query_windows_host(aid_engine* e, const float* pcm, const int64_t* win_begin, const int64_t* win_end,
                                      int n_windows, aid_match_row* rows, int max_rows, int32_t* n_rows) {
    return query_pcm(e, pcm, false, win_begin, win_end, n_windows, rows, max_rows, n_rows);
}

extern "C" int aid_query_hashes(aid_engine* e, const uint32_t* hash, const uint32_t* t_anchor, const int64_t* hash_off,
                                int n_queries, aid_match_row* rows, int max_rows, int32_t* n_rows) {
    if (!e || !hash_off || !rows || !n_rows || n_queries < 0 || max_rows < 1 || max_rows > AID_MAX_ROWS) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    const int64_t h0 = hash_off[0], total = hash_off[n_queries] - h0;
    if (total < 0 || total >= ((int64_t)1 << 32)) return AID_E_ARG;
    if (total > 0 && (!hash || !t_anchor)) return AID_E_ARG;
    std::vector<uint32_t> off32(n_queries + 1);
    for (int i = 0; i <= n_queries; i++) {
        if (hash_off[i] < h0 || (i && hash_off[i] < hash_off[i - 1])) return AID_E_ARG;
        off32[i] = (uint32_t)(hash_off[i] - h0);
    }
    for (int64_t i = 0; i < total; i++) if (t_anchor[h0 + i] >= AID_QUERY_MAX_FRAMES || hash[h0 + i] >> AID_HASH_BITS) return AID_E_ARG;
    AID_CUDA(e, s.hash.ensure((size_t)std::max<int64_t>(total, 1) * 4));
    AID_CUDA(e, s.t.ensure((size_t)std::max<int64_t>(total, 1) * 4));
    AID_CUDA(e, s.hash_off.ensure((size_t)(n_queries + 1) * 4));
    if (total > 0) {
        AID_CUDA(e, cudaMemcpyAsync(s.hash.p, hash + h0, (size_t)total * 4, cudaMemcpyHostToDevice, s.st));
        AID_CUDA(e, cudaMemcpyAsync(s.t.p, t_anchor + h0, (size_t)total * 4, cudaMemcpyHostToDevice, s.st));
    }
    AID_CUDA(e, cudaMemcpyAsync(s.hash_off.p, off32.data(), (size_t)(n_queries + 1) * 4, cudaMemcpyHostToDevice, s.st));
    return match_device(e, s.hash.as<uint32_t>(), s.t.as<uint32_t>(), s.hash_off.as<uint32_t>(), nullptr, n_queries, rows, max_rows,
                        n_rows, s.st);
}
