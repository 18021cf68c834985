// peaks.cu -- 2-D local-maximum constellation peak picker (103 bins x 25 frames), sm_100a.
//
// Replaces stage a5 of SURVEY.md section 8(a) (event-point picking inside the external `olaf_c`,
// reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the result:
// oracle/aid_oracle.c aid_oracle_peaks(); bit-exact on the same spectrogram.
//
// Design (DESIGN.md "Peak kernel"): one CTA (8 warps) streams the rows of one aligned 256-frame block of one
// track plus a 12-row halo on each side, 16 rows per step (two per warp, both loads in flight together), and
// keeps only two tiny per-row summaries in shared memory -- never the spectrogram itself:
//    A [row][g]  = maximum of the aligned 16-bin group g (32 groups per row, one per lane)
//    C5[row][g]  = max(A[g-2 .. g+2])
//  * Row pass (a warp owns a whole 2 KB row, 16 contiguous bins per lane): one max-reduction per lane gives A,
//    four shuffles give C5. A point can only be the maximum of its 103-bin window if it equals its group
//    maximum and that equals C5 (the five groups lie inside every window of the group). Such "group
//    candidates" (~1.6 % of the points) are queued with one shared-memory atomic per row.
//  * Column pass 1 (one thread per candidate): the whole groups inside the window -- g-2..g+2, plus g-3 or g+3
//    when the bin sits at the edge of its group -- are tested on all 25 rows using C5 and A: 24-50
//    shared-memory loads, rejects ~96 %.
//  * Column pass 2 (one warp per survivor, one lane per row): the at most 30 window bins outside the whole
//    groups are re-read from global memory (streamed moments ago: L2 hits) and tested exactly.
//  * Peaks are appended in no particular order and put in (t, f) order by a bitonic network sized to their count.
// Every comparison is <= against the candidate's own value, so exact ties behave as the specification says.
// HBM traffic: every spectrogram row is read once per block (+24/256 halo rows, L2 hits); peaks out are noise.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kStep = 2 * kWarps;           // rows per step
constexpr int kRing = 64;                   // rows of A / C5 kept (>= 2 * kStep + 24; power of two)
constexpr int kRingStride = kRing + 1;      // [group][row slot], odd stride: conflict-free both ways
constexpr int kQ1 = (kStep + 12) * AID_ROW_CAND_CAP;   // group candidates waiting for column pass 1
constexpr int kQ2 = 256;                    // survivors waiting for column pass 2 (overflow is handled inline)
constexpr int kHalfF = AID_PEAK_HALF_F, kHalfT = AID_PEAK_HALF_T;

struct Smem {
    float A[32 * kRingStride];
    float C5[32 * kRingStride];
    uint32_t q1[2][kQ1];
    uint32_t q2[2][kQ2];
    float q2v[2][kQ2];
    uint32_t peaks[AID_PEAK_BLOCK_CAP];
    int n1[2];
    int n2[2];
    int npeaks;
    int fail;
};

// The window bins of (row, f) that lie outside the whole groups, tested exactly against v.
// i = f & 15, g = f >> 4. Whole groups inside the window: g-2..g+2, plus g-3 if i <= 3, plus g+3 if i >= 12;
// that leaves at most 15 bins on each side. All loads are issued before the first compare (one L2 round trip).
__device__ __forceinline__ bool edges_le(const float* __restrict__ row, int f, float v) {
    const int i = f & 15, g = f >> 4;
    const int left_lo = max(f - kHalfF, 0), left_hi = 16 * (i <= 3 ? g - 3 : g - 2) - 1;
    const int right_lo = 16 * (i >= 12 ? g + 4 : g + 3), right_hi = min(f + kHalfF, AID_NBINS - 1);
    float x[30];
#pragma unroll
    for (int k = 0; k < 15; k++) {
        x[k] = left_lo + k <= left_hi ? __ldg(row + left_lo + k) : -1.0f;
        x[15 + k] = right_lo + k <= right_hi ? __ldg(row + right_lo + k) : -1.0f;
    }
    float m = x[0];
#pragma unroll
    for (int k = 1; k < 30; k++) m = fmaxf(m, x[k]);
    return m <= v;
}

__device__ __forceinline__ void push_peak(Smem& sm, uint32_t e) {
    const int p = atomicAdd(&sm.npeaks, 1);
    if (p < AID_PEAK_BLOCK_CAP) sm.peaks[p] = e;
}

__device__ __forceinline__ void load_row(float4 (&x)[4], const float* __restrict__ srow, int lane) {
    const float4* s = reinterpret_cast<const float4*>(srow) + lane * 4;
#pragma unroll
    for (int q = 0; q < 4; q++) x[q] = __ldg(s + q);
}

// One warp, one row already in registers: A, C5 into the rings; group candidates into q1[qsel].
__device__ __forceinline__ void row_pass(Smem& sm, const float4 (&x)[4], int row, bool store, bool emit, int qsel, int lane) {
    const float v[16] = {x[0].x, x[0].y, x[0].z, x[0].w, x[1].x, x[1].y, x[1].z, x[1].w,
                         x[2].x, x[2].y, x[2].z, x[2].w, x[3].x, x[3].y, x[3].z, x[3].w};
    float A = v[0];
#pragma unroll
    for (int i = 1; i < 16; i++) A = fmaxf(A, v[i]);
    // out-of-range shuffles return the lane's own A, which never changes a maximum that already contains A
    const float am1 = __shfl_up_sync(AID_FULL_MASK, A, 1), am2 = __shfl_up_sync(AID_FULL_MASK, A, 2);
    const float ap1 = __shfl_down_sync(AID_FULL_MASK, A, 1), ap2 = __shfl_down_sync(AID_FULL_MASK, A, 2);
    const float c5 = fmaxf(fmaxf(fmaxf(A, am1), fmaxf(am2, ap1)), ap2);
    if (store) {
        const int slot = row & (kRing - 1);
        sm.A[lane * kRingStride + slot] = A;
        sm.C5[lane * kRingStride + slot] = c5;
    }
    const bool cand = emit && A == c5 && A > AID_PEAK_MIN_S;
    if (!__any_sync(AID_FULL_MASK, cand)) return;
    uint32_t mask = 0;
    if (cand) {
#pragma unroll
        for (int i = 0; i < 16; i++) mask |= v[i] == A ? (1u << i) : 0u;
        if (lane == 0) mask &= ~((1u << AID_PEAK_MIN_BIN) - 1u);
    }
    const int total = __reduce_add_sync(AID_FULL_MASK, __popc(mask));
    if (total > AID_ROW_CAND_CAP) {                          // capacity rule of aid_params.h: the track fails
        if (lane == 0) sm.fail = 1;
        return;
    }
    if (mask) {
        int pos = atomicAdd(&sm.n1[qsel], __popc(mask));
        for (; mask; mask &= mask - 1, pos++)
            sm.q1[qsel][pos] = ((uint32_t)row << AID_PEAK_F_BITS) | (uint32_t)(16 * lane + __ffs(mask) - 1);
    }
}

// column pass 2 for one survivor: one lane per row of the window, exact test of the bins outside the whole groups
__device__ __forceinline__ void settle_survivor(Smem& sm, const float* __restrict__ base, int T, uint32_t e, float v, int lane) {
    const int row = (int)(e >> AID_PEAK_F_BITS), f = (int)(e & (AID_NBINS - 1));
    const int rw = row - kHalfT + lane;
    bool ok = true;
    if (lane <= 2 * kHalfT && rw >= 0 && rw < T) ok = edges_le(base + (int64_t)rw * AID_NBINS, f, v);
    if (__all_sync(AID_FULL_MASK, ok) && lane == 0) push_peak(sm, e);
}

__global__ void __launch_bounds__(kThreads, 3)
k_peaks(const float* __restrict__ spec, const aid_peak_unit* __restrict__ units,
        uint32_t* __restrict__ slots, uint32_t* __restrict__ unit_count, int32_t* __restrict__ track_status) {
    __shared__ Smem sm;
    const aid_peak_unit u = units[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = u.n_frames;
    const int row_end = u.row0 + u.n_rows;
    const int lo = max(0, u.row0 - kHalfT);
    const int hi = min(T, row_end + kHalfT);
    const float* base = spec + u.spec_row0 * AID_NBINS;

    if (tid == 0) { sm.fail = 0; sm.n1[0] = 0; sm.n1[1] = 0; sm.n2[0] = 0; sm.n2[1] = 0; sm.npeaks = 0; }
    __syncthreads();

    // rows step+warp and step+warp+8 belong to this warp; surplus warps redo the last row without side effects
    float4 x0[4], x1[4];
    load_row(x0, base + (int64_t)min(lo + warp, hi - 1) * AID_NBINS, lane);
    load_row(x1, base + (int64_t)min(lo + warp + kWarps, hi - 1) * AID_NBINS, lane);

    int par = 0;
    for (int step = lo; step < hi; step += kStep, par ^= 1) {
        const int r0 = step + warp, r1 = r0 + kWarps;
        row_pass(sm, x0, min(r0, hi - 1), r0 < hi, r0 >= u.row0 && r0 < row_end, par, lane);
        row_pass(sm, x1, min(r1, hi - 1), r1 < hi, r1 >= u.row0 && r1 < row_end, par, lane);
        if (step + kStep < hi) {                             // next step's rows fly during the column passes
            load_row(x0, base + (int64_t)min(r0 + kStep, hi - 1) * AID_NBINS, lane);
            load_row(x1, base + (int64_t)min(r1 + kStep, hi - 1) * AID_NBINS, lane);
        }
        __syncthreads();
        const int done = min(step + kStep, hi);
        const int vhi = done == hi ? row_end : min(row_end, done - kHalfT);   // rows < vhi have their full window

        // ---- column pass 1 (one thread per candidate): whole groups on all rows of the window, from the rings.
        // Rows are clamped into the track, which only repeats rows that are in the window anyway.
        const int nq = sm.n1[par];
        for (int k = tid; k < nq; k += kThreads) {
            const uint32_t e = sm.q1[par][k];
            const int row = (int)(e >> AID_PEAK_F_BITS), f = (int)(e & (AID_NBINS - 1));
            if (row >= vhi) {                                 // window not complete yet: look again next step
                sm.q1[par ^ 1][atomicAdd(&sm.n1[par ^ 1], 1)] = e;
                continue;
            }
            const int g = f >> 4, i = f & 15;
            int gx = i <= 3 ? g - 3 : (i >= 12 ? g + 3 : g);  // the one extra whole group, if any (else g again)
            if (gx < 0 || gx > 31) gx = g;
            const float* c5 = sm.C5 + g * kRingStride;
            const float* ax = sm.A + gx * kRingStride;
            const float v = sm.A[g * kRingStride + (row & (kRing - 1))];
            float m = -1.0f;
#pragma unroll
            for (int d = -kHalfT; d <= kHalfT; d++) {
                const int s = min(max(row + d, 0), T - 1) & (kRing - 1);
                m = fmaxf(m, fmaxf(c5[s], ax[s]));
            }
            if (m > v) continue;
            const int p = atomicAdd(&sm.n2[par], 1);
            if (p < kQ2) { sm.q2[par][p] = e; sm.q2v[par][p] = v; continue; }
            // survivor queue full (degenerate input): settle this one here, serially
            bool ok = true;
            for (int d = max(-kHalfT, -row); d <= min(kHalfT, T - 1 - row) && ok; d++)
                ok = edges_le(base + (int64_t)(row + d) * AID_NBINS, f, v);
            if (ok) push_peak(sm, e);
        }
        // ---- column pass 2 for the survivors of the PREVIOUS step (no barrier needed in between), last warps first
        const int n2 = min(sm.n2[par ^ 1], kQ2);
        for (int k = kWarps - 1 - warp; k < n2; k += kWarps) settle_survivor(sm, base, T, sm.q2[par ^ 1][k], sm.q2v[par ^ 1][k], lane);
        __syncthreads();
        if (tid == 0) { sm.n1[par] = 0; sm.n2[par ^ 1] = 0; }
    }
    __syncthreads();
    {   // survivors of the last step
        const int n2 = min(sm.n2[par ^ 1], kQ2);
        for (int k = warp; k < n2; k += kWarps) settle_survivor(sm, base, T, sm.q2[par ^ 1][k], sm.q2v[par ^ 1][k], lane);
    }
    __syncthreads();

    // peaks were appended in no particular order: sort them (t, f) ascending, then write the unit's slot list
    const int n_found = sm.npeaks;
    const int n = min(n_found, AID_PEAK_BLOCK_CAP);
    int N = 1;
    while (N < n) N <<= 1;
    for (int i = n + tid; i < N; i += kThreads) sm.peaks[i] = 0xffffffffu;
    __syncthreads();
    for (int k = 2; k <= N; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < N; i += kThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const uint32_t a = sm.peaks[i], b = sm.peaks[p];
                    if ((a > b) == ((i & k) == 0)) { sm.peaks[i] = b; sm.peaks[p] = a; }
                }
            }
            __syncthreads();
        }
    uint32_t* out = slots + (int64_t)blockIdx.x * AID_PEAK_BLOCK_CAP;
    for (int i = tid; i < n; i += kThreads) out[i] = sm.peaks[i];
    if (tid == 0) {
        unit_count[blockIdx.x] = (uint32_t)n;
        if (n_found > AID_PEAK_BLOCK_CAP || sm.fail) atomicOr(track_status + u.track, AID_TRACK_PEAK_OVERFLOW);
    }
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

cudaError_t aid_launch_peaks(const float* d_spec, const aid_peak_unit* d_units, int n_units,
                             uint32_t* d_slots, uint32_t* d_unit_count, int32_t* d_track_status,
                             cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
    k_peaks<<<n_units, kThreads, 0, st>>>(d_spec, d_units, d_slots, d_unit_count, d_track_status);
    return cudaGetLastError();
}

cudaError_t aid_launch_peak_compact(const uint32_t* d_slots, const uint32_t* d_unit_count,
                                    const uint32_t* d_unit_pos, const aid_peak_unit* d_units, int n_units,
                                    uint32_t* d_peaks, uint32_t* d_peak_track, cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
    k_peak_compact<<<n_units, 128, 0, st>>>(d_slots, d_unit_count, d_unit_pos, d_units, d_peaks, d_peak_track);
    return cudaGetLastError();
}
