// peaks.cu -- 2-D local-maximum constellation peak picker (103 bins x 25 frames), sm_100a.
//
// Replaces stage a5 of SURVEY.md section 8(a) (event-point picking inside the external `olaf_c`,
// reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the result:
// oracle/aid_oracle.c aid_oracle_peaks(); bit-exact on the same spectrogram.
//
// Design (DESIGN.md "Peak kernel"): ONE WARP streams the rows of a run of up to 4 consecutive aligned 256-frame
// blocks of one track plus a 12-row halo on each side; warps never wait for each other (no block barriers). A lane owns one aligned
// 16-bin group of the row (32 groups = 512 bins) and the warp keeps two tiny rings in shared memory --
// never the spectrogram itself:
//    A [row][g]   = maximum of group g                       (32 rows x 32 floats)
//    C5[row][g]   = max(A[g-2 .. g+2])
//    sign of A    = set if A passes the gate and A == C5 ("candidate group": a point can only be the maximum of
//                   its 103-bin window if it is the maximum of such a group, because groups g-2..g+2 lie inside
//                   every window of group g)
//  * Row pass: 4 x LDG.128 per lane, each instruction 512 contiguous bytes (the next two rows are already in flight),
//    a 4 x 4 transpose-reduce inside every quad of lanes and one shuffle to hand group g to lane g, four shuffles for C5.
//  * Column pass, 12 rows behind: the maximum of a lane's C5 column over the 25 rows of the window comes from the
//    van Herk / Gil-Werman decomposition -- a running prefix maximum over the current block of 25 rows (one
//    register) and suffix maxima of the previous block (written over the C5 ring once per block): one shared
//    load and two FMNMX per row instead of 25. A candidate survives if that maximum does not exceed its value
//    and -- when its bin sits at the edge of its group, so that group g-3 or g+3 lies wholly inside the
//    window too -- neither does the A column maximum of that neighbour (rare: direct sweep + one shuffle).
//    About 0.3 points per row survive.
//  * Survivors: one lane per row of the window re-reads the at most 30 window bins outside the whole groups
//    from global memory (streamed moments ago: L2 hits), all loads in flight together, and tests them exactly.
//  * Peaks collect in a small per-warp buffer and are put in (t, f) order by a warp bitonic network.
// Every comparison is <= against the candidate's own value, so exact ties behave as the specification says.
// HBM traffic: every spectrogram row is read once per block (+24/256 halo rows, L2 hits); peaks out are noise.
// Two forms of the row pass (template parameter SUMMARY of k_peaks), identical peaks:
//  * rows:    the warp streams the spectrogram rows themselves (2 KB per row; HBM bound, 0.91-0.94 of the roofline);
//  * summary: the STFT kernel has already reduced every row to its 32 group maxima (stft.cu GroupMax) and the warp streams
//             those 128 B per row through a cp.async ring; the spectrogram is read only around surviving candidates
//             (the product path; bound by instruction issue and latency; profiles/r02_peaks_summary.md).
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kRing = 32;                   // rows kept (>= 25: the column pass for row c runs when row c + 12 is in)
constexpr int kBlock = 2 * AID_PEAK_HALF_T + 1;   // 25: van Herk block = window height
constexpr int kBuf = 64;                    // peaks buffered in shared memory before spilling to the slot list
constexpr int kHalfF = AID_PEAK_HALF_F, kHalfT = AID_PEAK_HALF_T;

// Every lane reads and writes only its own column of these rings (lane = group), so they need no warp
// synchronisation at all; the layout [row][lane] makes every access conflict-free.
struct WarpSmem {
    float A[kRing][32];        // group maximum; stored NEGATED when the group is a candidate (S >= 0, so the sign is free)
                               // (the settlement reads other lanes' columns of it, behind a __syncwarp)
    float C5[kRing][32];
    uint32_t buf[kBuf];
};

constexpr int kStage = 16;                  // summary rows in flight per warp (SUMMARY kernels only)
template <bool SUMMARY> struct WarpSmemT : WarpSmem { float stage[SUMMARY ? kStage : 1][32]; };

// The window bins of (row, f) that lie outside the whole groups, tested exactly against v.
// i = f & 15, g = f >> 4. Whole groups inside the window: g-2..g+2, plus g-3 if i <= 3, plus g+3 if i >= 12;
// that leaves at most 15 bins on each side, and each side lies inside ONE aligned group (the left edge ends, the right
// edge starts, at a group boundary). So the lane reads those two groups with 2 x 4 LDG.128 -- 16 B per lane and
// instruction: one sector per row -- and masks the bins outside the window, instead of 30 scalar loads that each touch
// a sector per row (the scalar version made the settlement two thirds of the kernel's L1 wavefronts). All loads are
// issued before the first compare (one L2 round trip).
__device__ __forceinline__ bool edges_le(const float* __restrict__ row, int f, float v) {
    const int i = f & 15, g = f >> 4;
    const int gl = (i <= 3 ? g - 3 : g - 2) - 1;           // group holding the left edge bins (none if < 0)
    const int gr = i >= 12 ? g + 4 : g + 3;                // group holding the right edge bins (none if > 31)
    const int lo = f - kHalfF, hi = f + kHalfF;            // clipping at the row ends is implied by gl >= 0 / gr <= 31
    float4 xl[4], xr[4];
    const float4 none = make_float4(-1.0f, -1.0f, -1.0f, -1.0f);
    const float4* pl = reinterpret_cast<const float4*>(row + 16 * max(gl, 0));
    const float4* pr = reinterpret_cast<const float4*>(row + 16 * min(gr, 31));
#pragma unroll
    for (int q = 0; q < 4; q++) {
        xl[q] = gl >= 0 ? __ldg(pl + q) : none;
        xr[q] = gr <= 31 ? __ldg(pr + q) : none;
    }
    float m = -1.0f;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int bl = 16 * gl + 4 * q, br = 16 * gr + 4 * q;
        m = fmaxf(m, bl + 0 >= lo ? xl[q].x : -1.0f); m = fmaxf(m, bl + 1 >= lo ? xl[q].y : -1.0f);
        m = fmaxf(m, bl + 2 >= lo ? xl[q].z : -1.0f); m = fmaxf(m, bl + 3 >= lo ? xl[q].w : -1.0f);
        m = fmaxf(m, br + 0 <= hi ? xr[q].x : -1.0f); m = fmaxf(m, br + 1 <= hi ? xr[q].y : -1.0f);
        m = fmaxf(m, br + 2 <= hi ? xr[q].z : -1.0f); m = fmaxf(m, br + 3 <= hi ? xr[q].w : -1.0f);
    }
    return m <= v;
}

__device__ __forceinline__ void load_row(float4 (&x)[4], const float* __restrict__ srow, int lane) {
    const float4* s = reinterpret_cast<const float4*>(srow) + lane * 4;
#pragma unroll
    for (int q = 0; q < 4; q++) x[q] = __ldg(s + q);
}

// Streaming form of the row load: instruction q reads float4 number 32 q + lane, so a warp instruction covers 512
// contiguous bytes (4 lines, 16 sectors) instead of 32 chunks 64 B apart (16 lines, every sector touched by two
// instructions). The lane then holds four float4 of four different groups: group 8 q + j sits in lanes 4 j .. 4 j + 3 of
// instruction q.
__device__ __forceinline__ void load_row_stream(float4 (&x)[4], const float* __restrict__ srow, int lane) {
    const float4* s = reinterpret_cast<const float4*>(srow) + lane;
#pragma unroll
    for (int q = 0; q < 4; q++) x[q] = __ldg(s + 32 * q);
}

struct Stream {             // per-warp state (the same in every lane except px)
    const float* base;      // row 0 of the track
    uint32_t* out;          // the unit's slot list in global memory
    int T, lo, hi;          // frames in the track; first / one-past-last row this warp computes
    int row0, row_end;      // rows whose peaks this warp reports
    int unit, block_end;    // the peak unit (256-frame block) being filled; first row of the next one
    int track;
    uint32_t* unit_count;
    int32_t* track_status;
    int n_peaks;
    int jb;                 // rows of the current 25-row block already in (1..25), counted from lo
    float px;               // per lane: max of C5 over those rows
};

// C5 = max(A[g-2 .. g+2]) from the lane's own group maximum
__device__ __forceinline__ float five_group_max(float A) {
    // out-of-range shuffles return the lane's own A, which never changes a maximum that already contains A
    const float am1 = __shfl_up_sync(AID_FULL_MASK, A, 1), am2 = __shfl_up_sync(AID_FULL_MASK, A, 2);
    const float ap1 = __shfl_down_sync(AID_FULL_MASK, A, 1), ap2 = __shfl_down_sync(AID_FULL_MASK, A, 2);
    return fmaxf(fmaxf(fmaxf(A, am1), fmaxf(am2, ap1)), ap2);
}

// Row pass, register part: group maximum A and 5-group maximum C5 of one row loaded by load_row_stream. Pure
// register/shuffle code, so the two rows of a trip can be interleaved by the scheduler. The four partial maxima of a
// lane belong to four groups; a 4 x 4 transpose-reduce inside each quad of lanes (3 shuffles) leaves lane 4 j + p with
// the maximum of group 8 q(p) + j, q(p) = 2 (p & 1) + (p >> 1), and one more shuffle hands group g to lane g.
__device__ __forceinline__ void row_reduce(const float4 (&x)[4], float& A, float& c5, int lane) {
    float a[4];
#pragma unroll
    for (int q = 0; q < 4; q++) a[q] = fmaxf(fmaxf(x[q].x, x[q].y), fmaxf(x[q].z, x[q].w));
    const bool odd = lane & 1, hi = lane & 2;
    const float r1 = __shfl_xor_sync(AID_FULL_MASK, odd ? a[0] : a[2], 1);
    const float r2 = __shfl_xor_sync(AID_FULL_MASK, odd ? a[1] : a[3], 1);
    const float b0 = fmaxf(odd ? a[2] : a[0], r1), b1 = fmaxf(odd ? a[3] : a[1], r2);
    const float r3 = __shfl_xor_sync(AID_FULL_MASK, hi ? b0 : b1, 2);
    const float c = fmaxf(hi ? b1 : b0, r3);
    const int q = lane >> 3;                                 // this lane's group is 8 q + (lane & 7)
    A = __shfl_sync(AID_FULL_MASK, c, 4 * (lane & 7) + ((q >> 1) | ((q & 1) << 1)));
    c5 = five_group_max(A);
}

// Row pass, ring part: van Herk block bookkeeping, A / C5 / candidate flag into the rings.
template <bool PREFETCH = false>
__device__ __forceinline__ void row_commit(WarpSmem& sm, Stream& st, float A, float c5, int r, int lane) {
    if (st.jb == kBlock) {                                   // the previous 25-row block is complete: turn its C5
        float sfx = -1.0f;                                   // entries into suffix maxima, start a new block
#pragma unroll
        for (int k = 1; k <= kBlock; k++) {
            const int sl = (r - k) & (kRing - 1);
            sfx = fmaxf(sfx, sm.C5[sl][lane]);
            sm.C5[sl][lane] = sfx;
        }
        st.jb = 0;
        st.px = -1.0f;
    }
    st.jb++;
    st.px = fmaxf(st.px, c5);
    const int slot = r & (kRing - 1);
    const bool cand = r >= st.row0 && r < st.row_end && A == c5 && A > AID_PEAK_MIN_S;
    if constexpr (PREFETCH) {      // A/B only (measured: 1.59 -> 1.58 ms for 6.5 GB of extra DRAM reads per launch, not used): pull a
        if (cand)                  // candidate group's 16 bins towards the L2 twelve rows before verify_row may read them
            asm volatile("prefetch.global.L2 [%0];" :: "l"(st.base + (int64_t)r * AID_NBINS + 16 * lane));
    }
    sm.A[slot][lane] = cand ? -A : A;
    sm.C5[slot][lane] = c5;
}

// ascending bitonic sort of N (power of two) keys by one warp; `a` is shared or global memory
__device__ __forceinline__ void warp_sort(uint32_t* a, int N, int lane) {
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < N; i += 32) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t x = a[i], y = a[p];
                    if ((x > y) == ((i & k) == 0)) { a[i] = y; a[p] = x; }
                }
            }
            __syncwarp();
        }
}

// The current block is complete: order its peaks by (t, f), publish its list, move on to the next block.
__device__ __forceinline__ void flush_block(WarpSmem& sm, Stream& st, int lane) {
    const int n = min(st.n_peaks, AID_PEAK_BLOCK_CAP);
    int N = 1;
    while (N < n) N <<= 1;
    __syncwarp();
    if (n <= kBuf) {
        for (int i = n + lane; i < N; i += 32) sm.buf[i] = 0xffffffffu;
        __syncwarp();
        warp_sort(sm.buf, N, lane);
        for (int i = lane; i < n; i += 32) st.out[i] = sm.buf[i];
    } else {                                               // degenerate input: finish in the slot list itself
        for (int i = lane; i < kBuf; i += 32) st.out[i] = sm.buf[i];
        for (int i = n + lane; i < N; i += 32) st.out[i] = 0xffffffffu;
        __syncwarp();
        warp_sort(st.out, N, lane);
    }
    if (lane == 0) {
        st.unit_count[st.unit] = (uint32_t)n;
        if (st.n_peaks > AID_PEAK_BLOCK_CAP) atomicOr(st.track_status + st.track, AID_TRACK_PEAK_OVERFLOW);
    }
    __syncwarp();
    st.n_peaks = 0;
    st.unit++;
    st.block_end += AID_PEAK_BLOCK_FRAMES;
    st.out += AID_PEAK_BLOCK_CAP;
}

// Column pass + exact settlement for row c; `last` is the newest row in the rings (c + 12, or hi - 1 at the end
// of the track). The 25-row block structure is counted from row lo.
__device__ __forceinline__ void verify_row(WarpSmem& sm, Stream& st, int c, int last, int lane) {
    if (c >= st.block_end) flush_block(sm, st, lane);       // rows are verified in order: the previous block is done
    // max of C5 over rows [max(c-12, lo), last]: prefix of the current block, plus the suffix maxima of the previous
    // block from the window's first row on (they were written over the C5 ring when that block was completed)
    const int first = max(c - kHalfT, st.lo);
    const int blk0 = last - st.jb + 1;                      // first row of the current block
    float mc = st.px;
    if (first < blk0) mc = fmaxf(mc, sm.C5[first & (kRing - 1)][lane]);
    else if (first > blk0) {                                // only at the end of a track: the window is shorter than
        mc = -1.0f;                                         // the block; its rows are all in the (raw) current block
        for (int rw = first; rw <= last; rw++) mc = fmaxf(mc, sm.C5[rw & (kRing - 1)][lane]);
    }
    const float av = sm.A[c & (kRing - 1)][lane];
    const float v = fabsf(av);
    const bool pass = av < 0.0f && mc <= v;
    if (!__any_sync(AID_FULL_MASK, pass)) return;           // ~3 rows in 4 end here
    // which points of the group equal its maximum? (re-read the 16 bins: L1/L2 hit, rare)
    uint32_t keep = 0;
    if (pass) {
        float4 x[4];
        load_row(x, st.base + (int64_t)c * AID_NBINS, lane);
        const float w[16] = {x[0].x, x[0].y, x[0].z, x[0].w, x[1].x, x[1].y, x[1].z, x[1].w,
                             x[2].x, x[2].y, x[2].z, x[2].w, x[3].x, x[3].y, x[3].z, x[3].w};
#pragma unroll
        for (int i = 0; i < 16; i++) keep |= w[i] == v ? (1u << i) : 0u;
        if (lane == 0) keep &= ~((1u << AID_PEAK_MIN_BIN) - 1u);
    }
    // points at i <= 3 also need group lane-3, points at i >= 12 also need group lane+3, on every row of the window
    if (__any_sync(AID_FULL_MASK, (keep & 0xf00fu) != 0)) {
        float ma = -1.0f;
        for (int rw = first; rw <= last; rw++) ma = fmaxf(ma, fabsf(sm.A[rw & (kRing - 1)][lane]));
        // a missing neighbour group returns the lane's own maximum, which is <= mc <= v: harmless
        const float ml = __shfl_up_sync(AID_FULL_MASK, ma, 3), mr = __shfl_down_sync(AID_FULL_MASK, ma, 3);
        keep &= (ml <= v ? 0x000fu : 0u) | 0x0ff0u | (mr <= v ? 0xf000u : 0u);
    }
    // settle survivors one at a time: one lane per window row (the lanes read each other's A columns: make the ring
    // writes of this trip visible first)
    __syncwarp();
    for (;;) {
        const uint32_t who = __ballot_sync(AID_FULL_MASK, keep != 0);
        if (!who) break;
        const int src = __ffs(who) - 1;
        const uint32_t k = __shfl_sync(AID_FULL_MASK, keep, src);
        const float vv = __shfl_sync(AID_FULL_MASK, v, src);
        const int f = 16 * src + __ffs(k) - 1;
        if (lane == src) keep &= keep - 1;
        const int rw = c - kHalfT + lane;
        bool ok = true;
        if (lane <= 2 * kHalfT && rw >= 0 && rw < st.T) {
            // The edge bins of this row lie inside two aligned groups whose maxima are still in the A ring: if neither
            // exceeds the candidate, the row is settled without touching global memory (the usual case -- the
            // candidate is the maximum of 80 bins x 25 rows). Only rows where an edge group holds something larger
            // are re-read and tested bin by bin. (This removed most of the kernel's DRAM traffic above the algorithmic
            // bytes: by now the rows have left the L2, profiles/r02_peaks.md.)
            const int i = f & 15, g = f >> 4;
            const int gl = (i <= 3 ? g - 3 : g - 2) - 1, gr = i >= 12 ? g + 4 : g + 3;
            const float* arow = sm.A[rw & (kRing - 1)];
            const float el = gl >= 0 ? fabsf(arow[gl]) : -1.0f, er = gr <= 31 ? fabsf(arow[gr]) : -1.0f;
            if (fmaxf(el, er) > vv) ok = edges_le(st.base + (int64_t)rw * AID_NBINS, f, vv);
        }
        if (__all_sync(AID_FULL_MASK, ok)) {
            const uint32_t e = ((uint32_t)c << AID_PEAK_F_BITS) | (uint32_t)f;
            if (lane == 0) {
                if (st.n_peaks < kBuf) sm.buf[st.n_peaks] = e;
                else if (st.n_peaks < AID_PEAK_BLOCK_CAP) st.out[st.n_peaks] = e;
            }
            st.n_peaks++;
        }
    }
}

#ifndef AID_PEAKS_MIN_CTAS
#define AID_PEAKS_MIN_CTAS 4
#endif
// SUMMARY: the group maxima come from the STFT kernel (gmax[row][32], stft.cu GroupMax) instead of the row itself: the warp
// streams 128 B per row, four rows per trip with the next four in flight, and touches the spectrogram only where a
// candidate survives (verify_row). Same rings, same column pass, same settlement: the peaks are bit-identical.
#ifndef AID_PEAKS_SUM_CTAS
#define AID_PEAKS_SUM_CTAS 5
#endif
template <bool SUMMARY>
__global__ void __launch_bounds__(kWarpsPerCta * 32, SUMMARY ? AID_PEAKS_SUM_CTAS : AID_PEAKS_MIN_CTAS)
k_peaks(const float* __restrict__ spec, const float* __restrict__ gmax, const aid_peak_unit* __restrict__ units,
        const aid_peak_run* __restrict__ runs, int n_runs, uint32_t* __restrict__ slots, uint32_t* __restrict__ unit_count,
        int32_t* __restrict__ track_status) {
    __shared__ WarpSmemT<SUMMARY> s_all[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int run_id = blockIdx.x * kWarpsPerCta + warp;
    if (run_id >= n_runs) return;
    WarpSmem& sm = s_all[warp];
    const aid_peak_run run = runs[run_id];
    const aid_peak_unit u = units[run.first_unit];
    Stream st;
    st.T = u.n_frames;
    st.row0 = u.row0;
    st.row_end = min(st.T, u.row0 + run.n_blocks * AID_PEAK_BLOCK_FRAMES);
    st.lo = max(0, st.row0 - kHalfT);
    st.hi = min(st.T, st.row_end + kHalfT);
    st.base = spec + u.spec_row0 * AID_NBINS;
    st.unit = run.first_unit;
    st.block_end = u.row0 + AID_PEAK_BLOCK_FRAMES;
    st.out = slots + (int64_t)run.first_unit * AID_PEAK_BLOCK_CAP;
    st.track = u.track;
    st.unit_count = unit_count;
    st.track_status = track_status;
    st.n_peaks = 0;
    st.jb = 0;
    st.px = -1.0f;

    if constexpr (SUMMARY) {
        // ONE row per trip of ONE loop (commit row r, verify row r - 12; the rows past the end of the track only verify),
        // so the kernel holds a single copy of the column pass and the settlement: with them inlined once per row of a
        // four-row trip the kernel was 43 KB of code and stalled on instruction fetch (9.6 "no instruction" stall cycles
        // per issued instruction, profiles/r02_peaks_summary.md). A row is one 128 B line per warp, so it is latency, not
        // bandwidth, that has to be covered: the next kStage rows are in flight as cp.async copies into a small ring
        // (cp.async.wait_group counts GROUPS, so waiting for the oldest copy does not wait for the newest -- a register
        // ring filled by one LDG in a rolled loop does: all its loads share one counted scoreboard).
        const float* g = gmax + u.spec_row0 * 32 + lane;     // this lane's column of the summary
        float (*stage)[32] = s_all[warp].stage;
        auto fetch = [&](int row) {                          // this lane's value of `row` -> its slot (clamped: keeps the groups uniform)
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&stage[row & (kStage - 1)][lane]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(g + (int64_t)min(row, st.hi - 1) * 32) : "memory");
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        for (int k = 0; k < kStage; k++) fetch(st.lo + k);
        const int r_end = st.row_end + kHalfT;               // the last row verified is row_end - 1
#pragma unroll 1
        for (int r = st.lo; r < r_end; r++) {
            if (r < st.hi) {
                asm volatile("cp.async.wait_group %0;" :: "n"(kStage - 1) : "memory");
                const float a = stage[r & (kStage - 1)][lane];
                fetch(r + kStage);
                row_commit(sm, st, a, five_group_max(a), r, lane);
            }
            const int c = r - kHalfT;
            if (c >= st.row0) verify_row(sm, st, c, min(r, st.hi - 1), lane);
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        flush_block(sm, st, lane);
        return;
    } else {
    // two rows per trip, the next two already in flight
    float4 na[4], nb[4];
    load_row_stream(na, st.base + (int64_t)st.lo * AID_NBINS, lane);
    load_row_stream(nb, st.base + (int64_t)min(st.lo + 1, st.hi - 1) * AID_NBINS, lane);
    for (int r = st.lo; r < st.hi; r += 2) {
        const float4 xa[4] = {na[0], na[1], na[2], na[3]};
        const float4 xb[4] = {nb[0], nb[1], nb[2], nb[3]};
        if (r + 2 < st.hi) {
            load_row_stream(na, st.base + (int64_t)(r + 2) * AID_NBINS, lane);
            load_row_stream(nb, st.base + (int64_t)min(r + 3, st.hi - 1) * AID_NBINS, lane);
        }
        float A0, c0, A1, c1;
        row_reduce(xa, A0, c0, lane);
        row_reduce(xb, A1, c1, lane);
        row_commit(sm, st, A0, c0, r, lane);
        if (r - kHalfT >= st.row0) verify_row(sm, st, r - kHalfT, r, lane);
        if (r + 1 < st.hi) {
            row_commit(sm, st, A1, c1, r + 1, lane);
            if (r + 1 - kHalfT >= st.row0) verify_row(sm, st, r + 1 - kHalfT, r + 1, lane);
        }
    }
    }
    // rows whose window is cut by the end of the track
    for (int c = max(st.row0, st.hi - kHalfT); c < st.row_end; c++) verify_row(sm, st, c, st.hi - 1, lane);
    flush_block(sm, st, lane);
}

// slots (per unit, fixed capacity) -> dense per-batch peak list in (track, t, f) order
__global__ void k_peak_compact(const uint32_t* __restrict__ slots, const uint32_t* __restrict__ unit_count,
                               const uint32_t* __restrict__ unit_pos, const aid_peak_unit* __restrict__ units,
                               uint32_t* __restrict__ peaks, uint32_t* __restrict__ peak_track) {
    const int uidx = blockIdx.x;
    const uint32_t n = unit_count[uidx], pos = unit_pos[uidx];
    const uint32_t track = (uint32_t)units[uidx].track;
    const uint32_t* src = slots + (int64_t)uidx * AID_PEAK_BLOCK_CAP;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        peaks[pos + i] = src[i];
        peak_track[pos + i] = track;
    }
}

}  // namespace

cudaError_t aid_launch_peaks(const float* d_spec, const float* d_gmax, const aid_peak_unit* d_units, const aid_peak_run* d_runs,
                             int n_runs, uint32_t* d_slots, uint32_t* d_unit_count, int32_t* d_track_status,
                             cudaStream_t st) {
    if (n_runs <= 0) return cudaSuccess;
    const int grid = (n_runs + kWarpsPerCta - 1) / kWarpsPerCta;
    if (d_gmax) k_peaks<true><<<grid, kWarpsPerCta * 32, 0, st>>>(d_spec, d_gmax, d_units, d_runs, n_runs, d_slots, d_unit_count, d_track_status);
    else k_peaks<false><<<grid, kWarpsPerCta * 32, 0, st>>>(d_spec, nullptr, d_units, d_runs, n_runs, d_slots, d_unit_count, d_track_status);
    return cudaGetLastError();
}

cudaError_t aid_launch_peak_compact(const uint32_t* d_slots, const uint32_t* d_unit_count,
                                    const uint32_t* d_unit_pos, const aid_peak_unit* d_units, int n_units,
                                    uint32_t* d_peaks, uint32_t* d_peak_track, cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
    k_peak_compact<<<n_units, 128, 0, st>>>(d_slots, d_unit_count, d_unit_pos, d_units, d_peaks, d_peak_track);
    return cudaGetLastError();
}
