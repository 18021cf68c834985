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
//    written into a 48-row shared-memory ring, and the row's candidates (S == M1, gates passed) are
//    recorded in bin order with a warp scan.
//  * Column pass: a candidate is a peak iff no M1 value in the 24 neighbouring rows of its column
//    exceeds it; only candidates (about 1 % of the points) pay for the time direction.
//  * Peaks leave the CTA already ordered by (t, f): block-wide ballot compaction, no sort, no atomics.
// HBM traffic: every spectrogram row is read once per block (+24/256 halo rows); peaks out are noise.
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;       // rows per step
constexpr int kRing = 48;                   // M1 rows kept in shared memory
constexpr int kCandRows = 32;               // candidate lists kept per row (ring)
constexpr float kNeg = -1.0f;               // below every S (S >= 0)

// ring rows are stored permuted so that the 16-bins-per-lane register layout writes conflict-free
// 16-byte chunks: bin f = 16*l + 4*q + c  ->  128*q + 4*l + c
__device__ __forceinline__ int phys(int f) { return ((f >> 2) & 3) * 128 + (f >> 4) * 4 + (f & 3); }

struct Smem {
    float ring[kRing][AID_NBINS];
    uint16_t cand_f[kCandRows][AID_ROW_CAND_CAP];
    int cand_n[kCandRows];
    uint32_t peaks[AID_PEAK_BLOCK_CAP];
    int wcnt[2][kWarps];
    int fail;
};

__device__ __forceinline__ float up(float v, int d, int lane) {
    const float r = __shfl_up_sync(AID_FULL_MASK, v, d);
    return lane >= d ? r : kNeg;
}
__device__ __forceinline__ float down(float v, int d, int lane) {
    const float r = __shfl_down_sync(AID_FULL_MASK, v, d);
    return lane + d < 32 ? r : kNeg;
}

// One warp: row of S -> M1 row into the ring, candidates into the row's list.
__device__ __forceinline__ void row_pass(Smem& sm, const float* __restrict__ srow, int row, bool emit, int lane) {
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
    const float am1 = up(A, 1, lane), am2 = up(A, 2, lane), am3 = up(A, 3, lane);
    const float ap1 = down(A, 1, lane), ap2 = down(A, 2, lane), ap3 = down(A, 3, lane);
    const float c5 = fmaxf(fmaxf(fmaxf(A, am1), fmaxf(am2, ap1)), ap2);

    float m1[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        // window of bin 16*lane + i is [16*lane + i - 51, 16*lane + i + 51]
        float L, R;
        if (i >= 3) L = up(suf[i - 3], 3, lane);
        else        L = fmaxf(up(suf[i + 13], 4, lane), am3);
        if (i <= 12) R = down(pre[i + 3], 3, lane);
        else         R = fmaxf(down(pre[i - 13], 4, lane), ap3);
        m1[i] = fmaxf(c5, fmaxf(L, R));
    }
    float* dst = sm.ring[row % kRing];
#pragma unroll
    for (int q = 0; q < 4; q++)
        *reinterpret_cast<float4*>(dst + 128 * q + 4 * lane) =
            make_float4(m1[4 * q], m1[4 * q + 1], m1[4 * q + 2], m1[4 * q + 3]);

    if (!emit) return;
    uint32_t mask = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const bool c = v[i] == m1[i] && v[i] > AID_PEAK_MIN_S && (16 * lane + i) >= AID_PEAK_MIN_BIN;
        mask |= c ? (1u << i) : 0u;
    }
    int cnt = __popc(mask), incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(AID_FULL_MASK, incl, d);
        if (lane >= d) incl += o;
    }
    const int total = __shfl_sync(AID_FULL_MASK, incl, 31);
    int pos = incl - cnt;
    const int slot = row % kCandRows;
    while (mask) {
        const int i = __ffs(mask) - 1;
        mask &= mask - 1;
        if (pos < AID_ROW_CAND_CAP) sm.cand_f[slot][pos] = (uint16_t)(16 * lane + i);
        pos++;
    }
    if (lane == 0) {
        sm.cand_n[slot] = total < AID_ROW_CAND_CAP ? total : AID_ROW_CAND_CAP;
        if (total > AID_ROW_CAND_CAP) sm.fail = 1;
    }
}

__global__ void __launch_bounds__(kThreads)
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

    if (tid == 0) sm.fail = 0;
    __syncthreads();

    int n_out = 0;          // peaks emitted so far (same value in every thread)
    int verified = u.row0;  // rows < verified have been through the column pass
    int par = 0;
    for (int step = lo; step < hi; step += kWarps) {
        const int r = step + warp;
        if (r < hi) row_pass(sm, base + (int64_t)r * AID_NBINS, r, r >= u.row0 && r < row_end, lane);
        __syncthreads();
        const int done = min(step + kWarps, hi);
        const int vhi = done == hi ? row_end : min(row_end, done - AID_PEAK_HALF_T);
        // column pass over rows [verified, vhi): 8 rows x 64 candidate slots per sweep
        for (int rb = verified; rb < vhi; rb += kThreads / AID_ROW_CAND_CAP) {
            const int row = rb + tid / AID_ROW_CAND_CAP, s = tid % AID_ROW_CAND_CAP;
            bool is_peak = false;
            int f = 0;
            if (row < vhi && s < sm.cand_n[row % kCandRows]) {
                f = sm.cand_f[row % kCandRows][s];
                const int pf = phys(f);
                const float v = sm.ring[row % kRing][pf];
                float m = kNeg;
#pragma unroll
                for (int d = -AID_PEAK_HALF_T; d <= AID_PEAK_HALF_T; d++) {
                    const int rr = row + d;
                    if (d != 0 && rr >= 0 && rr < T) m = fmaxf(m, sm.ring[rr % kRing][pf]);
                }
                is_peak = m <= v;
            }
            const uint32_t bal = __ballot_sync(AID_FULL_MASK, is_peak);
            if (lane == 0) sm.wcnt[par][warp] = __popc(bal);
            __syncthreads();
            int before = 0, all = 0;
#pragma unroll
            for (int w = 0; w < kWarps; w++) {
                const int c = sm.wcnt[par][w];
                all += c;
                before += w < warp ? c : 0;
            }
            if (is_peak) {
                const int pos = n_out + before + __popc(bal & ((1u << lane) - 1));
                if (pos < AID_PEAK_BLOCK_CAP) sm.peaks[pos] = ((uint32_t)row << AID_PEAK_F_BITS) | (uint32_t)f;
            }
            n_out += all;
            par ^= 1;
        }
        verified = max(verified, vhi);
        __syncthreads();     // ring rows and candidate slots may be overwritten by the next step
    }

    const int n = min(n_out, AID_PEAK_BLOCK_CAP);
    uint32_t* out = slots + (int64_t)blockIdx.x * AID_PEAK_BLOCK_CAP;
    for (int i = tid; i < n; i += kThreads) out[i] = sm.peaks[i];
    if (tid == 0) {
        unit_count[blockIdx.x] = (uint32_t)n;
        if (n_out > AID_PEAK_BLOCK_CAP || sm.fail) atomicOr(track_status + u.track, AID_TRACK_PEAK_OVERFLOW);
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
