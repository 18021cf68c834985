// hasher.cu -- warp-cooperative landmark pairing: sorted peaks -> (hash, t_anchor), sm_100a.
//
// Replaces stage a6 of SURVEY.md section 8(a) (fingerprint construction inside the external `olaf_c`,
// reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the result:
// oracle/aid_oracle.c aid_oracle_hashes(); bit-exact, same order.
//
// One warp per anchor: the 32 lanes test the next 32 peaks of the same track in (t, f) order, a ballot
// ranks the qualifying targets, and the first AID_FANOUT are kept. The pass runs twice: once to count
// (so that a device-wide scan can give every anchor its dense output position) and once to write.
// Bytes are negligible next to the spectrogram; this kernel is latency/launch bound.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <bool kWrite>
__global__ void __launch_bounds__(kThreads)
k_hash(const uint32_t* __restrict__ peaks, const uint32_t* __restrict__ peak_track,
       const uint32_t* __restrict__ peak_off, const uint32_t* __restrict__ n_peaks_total,
       uint32_t* __restrict__ cnt, const uint32_t* __restrict__ pos,
       uint32_t* __restrict__ hash, uint32_t* __restrict__ t_anchor, int64_t hash_cap,
       int32_t* __restrict__ overflow) {
    const int lane = threadIdx.x & 31;
    const uint32_t n_total = *n_peaks_total;
    const uint32_t warps = gridDim.x * (kThreads / 32);
    if (!kWrite && blockIdx.x == 0 && threadIdx.x == 0) cnt[n_total] = 0;   // the scan runs over n_total + 1 values
    for (uint32_t a = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); a < n_total; a += warps) {
        const uint32_t ka = peaks[a];
        const uint32_t end = peak_off[peak_track[a] + 1];
        const int t1 = (int)(ka >> AID_PEAK_F_BITS), f1 = (int)(ka & (AID_NBINS - 1));
        const uint32_t out0 = kWrite ? pos[a] : 0;
        int taken = 0;
        for (uint32_t j0 = a + 1; j0 < end && taken < AID_FANOUT; j0 += 32) {
            const uint32_t j = j0 + lane;
            bool ok = false, past = false;
            int f2 = 0, dt = 0;
            if (j < end) {
                const uint32_t kb = peaks[j];
                f2 = (int)(kb & (AID_NBINS - 1));
                dt = (int)(kb >> AID_PEAK_F_BITS) - t1;
                const int df = f2 > f1 ? f2 - f1 : f1 - f2;
                past = dt > AID_DT_MAX;
                ok = dt >= AID_DT_MIN && !past && df >= AID_DF_MIN && df <= AID_DF_MAX;
            }
            const uint32_t bal = __ballot_sync(AID_FULL_MASK, ok);
            const int rank = taken + __popc(bal & ((1u << lane) - 1));
            if (kWrite && ok && rank < AID_FANOUT) {
                const int64_t o = (int64_t)out0 + rank;
                if (o < hash_cap) { hash[o] = AID_HASH(f1, f2, dt); t_anchor[o] = (uint32_t)t1; }
                else *overflow = 1;
            }
            taken += __popc(bal);
            if (__any_sync(AID_FULL_MASK, past)) break;
        }
        if (!kWrite && lane == 0) cnt[a] = (uint32_t)min(taken, AID_FANOUT);
    }
}

__global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                             uint32_t* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

}  // namespace

static int hash_grid(int64_t max_peaks) {
    int64_t need = (max_peaks + (kThreads / 32) - 1) / (kThreads / 32);
    const int64_t cap = 148 * 8;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

cudaError_t aid_launch_hash_count(const uint32_t* d_peaks, const uint32_t* d_peak_track,
                                  const uint32_t* d_peak_off, const uint32_t* d_n_peaks_total,
                                  int64_t max_peaks, uint32_t* d_cnt, cudaStream_t st) {
    k_hash<false><<<hash_grid(max_peaks), kThreads, 0, st>>>(d_peaks, d_peak_track, d_peak_off, d_n_peaks_total,
                                                              d_cnt, nullptr, nullptr, nullptr, 0, nullptr);
    return cudaGetLastError();
}

cudaError_t aid_launch_hash_write(const uint32_t* d_peaks, const uint32_t* d_peak_track,
                                  const uint32_t* d_peak_off, const uint32_t* d_n_peaks_total,
                                  int64_t max_peaks, const uint32_t* d_pos,
                                  uint32_t* d_hash, uint32_t* d_t, int64_t hash_cap, int32_t* d_overflow,
                                  cudaStream_t st) {
    k_hash<true><<<hash_grid(max_peaks), kThreads, 0, st>>>(d_peaks, d_peak_track, d_peak_off, d_n_peaks_total,
                                                             nullptr, d_pos, d_hash, d_t, hash_cap, d_overflow);
    return cudaGetLastError();
}

cudaError_t aid_launch_gather_u32(const uint32_t* d_src, const uint32_t* d_idx, uint32_t* d_dst, int n,
                                  cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    k_gather_u32<<<(n + 255) / 256, 256, 0, st>>>(d_src, d_idx, d_dst, n);
    return cudaGetLastError();
}
