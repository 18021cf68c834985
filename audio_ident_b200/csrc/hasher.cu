// hasher.cu -- landmark pairing: sorted peaks -> (hash, t_anchor), sm_100a.
//
// Replaces stage a6 of SURVEY.md section 8(a) (fingerprint construction inside the external `olaf_c`,
// reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the result:
// oracle/aid_oracle.c aid_oracle_hashes(); bit-exact, same order.
//
// One THREAD per anchor (round 2, second half; round 1 gave a warp to every anchor): the lane walks the peaks that follow
// its anchor in (t, f) order until it holds AID_FANOUT targets or the time window is over -- a handful of steps, since a
// window of AID_DT_MAX frames holds few peaks -- and consecutive lanes read consecutive peaks, so the loads coalesce.
// The warp-per-anchor form spent its time on a chain of four dependent loads per anchor (peak, track, track end, 32
// candidates) for a ballot that usually found fewer than eight: 0.145 ms per pass and 2048 tracks against 0.02 ms now.
// The pass runs twice: once to count (so that a device-wide scan can give every anchor its dense output position) and
// once to write. Bytes are negligible next to the spectrogram.
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
    const uint32_t n_total = *n_peaks_total;
    const uint32_t threads = gridDim.x * kThreads;
    if (!kWrite && blockIdx.x == 0 && threadIdx.x == 0) cnt[n_total] = 0;   // the scan runs over n_total + 1 values
    for (uint32_t a = blockIdx.x * kThreads + threadIdx.x; a < n_total; a += threads) {
        const uint32_t ka = peaks[a];
        const uint32_t end = peak_off[peak_track[a] + 1];
        const int t1 = (int)(ka >> AID_PEAK_F_BITS), f1 = (int)(ka & (AID_NBINS - 1));
        const int64_t out0 = kWrite ? (int64_t)pos[a] : 0;
        int taken = 0;
        for (uint32_t j = a + 1; j < end && taken < AID_FANOUT; j++) {
            const uint32_t kb = peaks[j];
            const int dt = (int)(kb >> AID_PEAK_F_BITS) - t1;
            if (dt > AID_DT_MAX) break;                                     // peaks are in (t, f) order: nothing later qualifies
            const int f2 = (int)(kb & (AID_NBINS - 1));
            const int df = f2 > f1 ? f2 - f1 : f1 - f2;
            if (dt >= AID_DT_MIN && df >= AID_DF_MIN && df <= AID_DF_MAX) {
                if (kWrite) {
                    const int64_t o = out0 + taken;
                    if (o < hash_cap) { hash[o] = AID_HASH(f1, f2, dt); t_anchor[o] = (uint32_t)t1; }
                    else *overflow = 1;
                }
                taken++;
            }
        }
        if (!kWrite) cnt[a] = (uint32_t)taken;
    }
}

__global__ void k_gather_u32(const uint32_t* __restrict__ src, const uint32_t* __restrict__ idx,
                             uint32_t* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

}  // namespace

static int hash_grid(int64_t max_peaks) {
    int64_t need = (max_peaks + kThreads - 1) / kThreads;
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
