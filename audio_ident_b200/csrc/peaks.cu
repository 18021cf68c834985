// peaks.cu -- 2-D local-maximum constellation peak picker (103 bins x 25 frames), sm_100a.
//
// Replaces stage a5 of SURVEY.md section 8(a) (event-point picking inside the external `olaf_c`,
// reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the result:
// oracle/aid_oracle.c aid_oracle_peaks(); bit-exact on the same spectrogram.
//
// Design (DESIGN.md "Peak kernel"): one CTA streams the rows of one aligned 256-frame block of one
// track (plus a 12-row halo on each side).
//  * Row pass: a warp owns a whole 512-bin row, 16 contiguous bins per lane. The 103-bin sliding
//    maximum M1 is built from per-lane prefix/suffix maxima and 38 warp shuffles (no shared memory),
//    written into a 40-row shared-memory ring, and the row's candidates (S == M1, gates passed) are
//    queued with one shared-memory atomic per row.
//  * Column pass: a candidate is a peak iff no M1 value in the 24 neighbouring rows of its column
//    exceeds it; only candidates (about 1 % of the points) pay for the time direction.
//  * Candidates and peaks travel through small shared-memory queues; the block's peaks (a few dozen) are
//    put in (t, f) order by a bitonic network sized to their count before they are written.
// HBM traffic: every spectrogram row is read once per block (+24/256 halo rows); peaks out are noise.
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;       // rows per step
constexpr int kRing = 40;                   // M1 rows kept in shared memory (>= kWarps + 24)
constexpr int kQueue = 2048;                // candidates waiting for the column pass (<= 28 rows x 64)

// ring rows are stored permuted so that the 16-bins-per-lane register layout writes conflict-free
// 16-byte chunks: bin f = 16*l + 4*q + c  ->  128*q + 4*l + c
__device__ __forceinline__ int phys(int f) { return ((f >> 2) & 3) * 128 + (f >> 4) * 4 + (f & 3); }

struct Smem {
    float ring[kRing][AID_NBINS];
    uint32_t queue[2][kQueue];
    uint32_t peaks[AID_PEAK_BLOCK_CAP];
    int qn[2];
    int npeaks;
    int fail;
};

// One warp: row of S -> M1 row into the ring, row candidates (S == M1, gates passed) into the queue.
// Out-of-range shuffles return the lane's own value, which is never above the lane's own group maximum and
// therefore never changes a window maximum: the clipped window needs no masking.
__device__ __forceinline__ void row_pass(Smem& sm, const float* __restrict__ srow, int row, bool store, bool emit,
                                         int qsel, int lane) {
    float v[16];
    const float4* src = reinterpret_cast<const float4*>(srow) + lane * 4;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const float4 x = __ldg(src + q);
        v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
    }
    float pre[16], suf[16];
    pre[0] = v[0];
#pragma unroll
    for (int i = 1; i < 16; i++) pre[i] = fmaxf(pre[i - 1], v[i]);
    suf[15] = v[15];
#pragma unroll
    for (int i = 14; i >= 0; i--) suf[i] = fmaxf(suf[i + 1], v[i]);
    const float A = pre[15];
    const float am1 = __shfl_up_sync(AID_FULL_MASK, A, 1), am2 = __shfl_up_sync(AID_FULL_MASK, A, 2);
    const float am3 = __shfl_up_sync(AID_FULL_MASK, A, 3);
    const float ap1 = __shfl_down_sync(AID_FULL_MASK, A, 1), ap2 = __shfl_down_sync(AID_FULL_MASK, A, 2);
    const float ap3 = __shfl_down_sync(AID_FULL_MASK, A, 3);
    const float c5 = fmaxf(fmaxf(fmaxf(A, am1), fmaxf(am2, ap1)), ap2);
    const float c5l = fmaxf(c5, am3), c5r = fmaxf(c5, ap3);

    float* dst = sm.ring[row % kRing];
    uint32_t mask = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        float m1[4];
#pragma unroll
        for (int c = 0; c < 4; c++) {
            const int i = 4 * q + c;
            // window of bin 16*lane + i is [16*lane + i - 51, 16*lane + i + 51]
            const float L = i >= 3 ? __shfl_up_sync(AID_FULL_MASK, suf[i - 3], 3) : __shfl_up_sync(AID_FULL_MASK, suf[i + 13], 4);
            const float R = i <= 12 ? __shfl_down_sync(AID_FULL_MASK, pre[i + 3], 3) : __shfl_down_sync(AID_FULL_MASK, pre[i - 13], 4);
            m1[c] = fmaxf(i < 3 ? c5l : (i > 12 ? c5r : c5), fmaxf(L, R));
            mask |= (v[i] == m1[c] && v[i] > AID_PEAK_MIN_S) ? (1u << i) : 0u;
        }
        if (store) *reinterpret_cast<float4*>(dst + 128 * q + 4 * lane) = make_float4(m1[0], m1[1], m1[2], m1[3]);
    }
    if (lane == 0) mask &= ~((1u << AID_PEAK_MIN_BIN) - 1u);
    if (!emit) mask = 0;
    if (!__any_sync(AID_FULL_MASK, mask != 0)) return;
    const int cnt = __popc(mask);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(AID_FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(AID_FULL_MASK, incl, 31);
    const int kept = min(total, AID_ROW_CAND_CAP);
    int base = 0;
    if (lane == 0) {
        base = atomicAdd(&sm.qn[qsel], kept);
        if (total > AID_ROW_CAND_CAP) sm.fail = 1;
    }
    base = __shfl_sync(AID_FULL_MASK, base, 0);
    int pos = incl - cnt;
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        if (pos < kept && base + pos < kQueue) sm.queue[qsel][base + pos] = ((uint32_t)row << AID_PEAK_F_BITS) | (uint32_t)(16 * lane + i);
        pos++;
    }
}

__global__ void __launch_bounds__(kThreads, 2)
k_peaks(const float* __restrict__ spec, const aid_peak_unit* __restrict__ units,
        uint32_t* __restrict__ slots, uint32_t* __restrict__ unit_count, int32_t* __restrict__ track_status) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem& sm = *reinterpret_cast<Smem*>(smem_raw);
    const aid_peak_unit u = units[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int T = u.n_frames;
    const int row_end = u.row0 + u.n_rows;
    const int lo = max(0, u.row0 - AID_PEAK_HALF_T);
    const int hi = min(T, row_end + AID_PEAK_HALF_T);
    const float* base = spec + u.spec_row0 * AID_NBINS;

    if (tid == 0) { sm.fail = 0; sm.qn[0] = 0; sm.qn[1] = 0; sm.npeaks = 0; }
    __syncthreads();

    int par = 0;
    for (int step = lo; step < hi; step += kWarps, par ^= 1) {
        const int r = step + warp;
        const int rr = min(r, hi - 1);                       // every warp runs the pass; surplus warps redo the last row
        row_pass(sm, base + (int64_t)rr * AID_NBINS, rr, r < hi, r >= u.row0 && r < row_end, par, lane);
        __syncthreads();
        const int done = min(step + kWarps, hi);
        const int vhi = done == hi ? row_end : min(row_end, done - AID_PEAK_HALF_T);   // rows < vhi have their full window
        const int nq = min(sm.qn[par], kQueue);
        for (int k = tid; k < nq; k += kThreads) {
            const uint32_t e = sm.queue[par][k];
            const int row = (int)(e >> AID_PEAK_F_BITS);
            if (row >= vhi) {                                 // window not complete yet: look again next step
                const int p = atomicAdd(&sm.qn[par ^ 1], 1);
                if (p < kQueue) sm.queue[par ^ 1][p] = e;
                continue;
            }
            const int pf = phys((int)(e & (AID_NBINS - 1)));
            const int s0 = row % kRing;
            const float v = sm.ring[s0][pf];
            const int dlo = max(-AID_PEAK_HALF_T, -row), dhi = min(AID_PEAK_HALF_T, T - 1 - row);
            int s = s0 - AID_PEAK_HALF_T;
            if (s < 0) s += kRing;
            float m = -1.0f;
#pragma unroll
            for (int d = -AID_PEAK_HALF_T; d <= AID_PEAK_HALF_T; d++) {
                if (d != 0 && d >= dlo && d <= dhi) m = fmaxf(m, sm.ring[s][pf]);
                s = s + 1 == kRing ? 0 : s + 1;
            }
            if (m <= v) {
                const int p = atomicAdd(&sm.npeaks, 1);
                if (p < AID_PEAK_BLOCK_CAP) sm.peaks[p] = e;
            }
        }
        __syncthreads();
        if (tid == 0) sm.qn[par] = 0;                        // next use of this queue is two steps away
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
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(k_peaks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_peaks, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return e;
        configured = true;
    }
    k_peaks<<<n_units, kThreads, sizeof(Smem), st>>>(d_spec, d_units, d_slots, d_unit_count, d_track_status);
    return cudaGetLastError();
}

cudaError_t aid_launch_peak_compact(const uint32_t* d_slots, const uint32_t* d_unit_count,
                                    const uint32_t* d_unit_pos, const aid_peak_unit* d_units, int n_units,
                                    uint32_t* d_peaks, uint32_t* d_peak_track, cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
    k_peak_compact<<<n_units, 128, 0, st>>>(d_slots, d_unit_count, d_unit_pos, d_units, d_peaks, d_peak_track);
    return cudaGetLastError();
}
