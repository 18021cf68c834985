// synth.cu -- deterministic device-side synthetic corpus for bench.py (never used for parity inputs:
// those come from audio_ident_b200/synth.py; the bench copies a sample of these tracks back to the
// host and checks the GPU fingerprints of exactly those samples against the oracle).
//
// Track g is a sum of Gaussian-windowed tone bursts over a white-noise floor, like the host generator
// (SURVEY.md section 8(d)), but laid out so that any sample can be produced independently: time is cut
// into cells of kCell samples and every cell owns kPerCell bursts whose centres lie inside it, all
// derived from a counter-based hash of (seed, track, cell, burst).
#include "common.cuh"

namespace {

constexpr int kCell = 1600;        // 0.1 s
constexpr int kPerCell = 4;        // 40 bursts per second
constexpr int kReach = 3;          // a burst is cut at 3 sigma <= 0.225 s < 3 cells
constexpr int kBursts = (2 * kReach + 1) * kPerCell;

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ float u01(uint64_t h) { return (float)(h >> 40) * (1.0f / 16777216.0f); }

struct Burst { float centre, inv_sigma, omega, amp, phase, reach; };

__global__ void __launch_bounds__(256)
k_synth(float* __restrict__ pcm, int64_t first_track, int64_t track_stride, int64_t samples_per_track, uint64_t seed) {
    __shared__ Burst s_b[kBursts];
    const int64_t track = first_track + track_stride * blockIdx.y;
    const int cell = blockIdx.x;
    const uint64_t tkey = mix64(seed ^ mix64((uint64_t)track));
    if (threadIdx.x < kBursts) {
        const int c = cell - kReach + (int)threadIdx.x / kPerCell, k = threadIdx.x % kPerCell;
        Burst b;
        b.amp = 0.0f; b.centre = 0.0f; b.inv_sigma = 1.0f; b.omega = 0.0f; b.phase = 0.0f; b.reach = 0.0f;
        if (c >= 0) {
            const uint64_t h = mix64(tkey ^ ((uint64_t)c * 8 + k) * 0xD1B54A32D192ED03ull);
            const float dur = (0.030f + 0.270f * u01(mix64(h + 1))) * AID_SAMPLE_RATE;
            const float sigma = dur * 0.25f;
            b.centre = ((float)(c - cell) + u01(mix64(h + 2))) * kCell;       // relative to this cell's start
            b.inv_sigma = 1.0f / sigma;
            b.reach = 3.0f * sigma;
            const float f = 150.0f * __expf(u01(mix64(h + 3)) * 3.912023f);      // log-uniform 150 .. 7500 Hz
            b.omega = 6.283185307f * f / AID_SAMPLE_RATE;
            b.amp = 0.05f + 0.45f * u01(mix64(h + 4));
            b.phase = 6.283185307f * u01(mix64(h + 5));
        }
        s_b[threadIdx.x] = b;
    }
    __syncthreads();
    float* out = pcm + (int64_t)blockIdx.y * samples_per_track;
    for (int i = threadIdx.x; i < kCell; i += blockDim.x) {
        const int64_t n = (int64_t)cell * kCell + i;
        if (n >= samples_per_track) break;
        const float t = (float)i;
        float acc = 0.0f;
#pragma unroll 4
        for (int j = 0; j < kBursts; j++) {
            const Burst b = s_b[j];
            const float d = t - b.centre;
            if (fabsf(d) <= b.reach) {
                const float z = d * b.inv_sigma;
                acc += b.amp * __expf(-0.5f * z * z) * __sinf(b.omega * d + b.phase);
            }
        }
        // noise floor at about -50 dBFS: sum of four uniforms, variance 1/3 -> scale
        const uint64_t h = mix64(tkey ^ (uint64_t)n * 0xA24BAED4963EE407ull);
        const float g = (u01(h) + u01(mix64(h)) + u01(mix64(h + 7)) + u01(mix64(h + 13)) - 2.0f) * 1.7320508f;
        acc = 0.45f * acc + 0.0028f * g;
        out[n] = fminf(1.0f, fmaxf(-1.0f, acc));
    }
}

}  // namespace

cudaError_t aid_launch_synth(float* d_pcm, int64_t first_track, int n_tracks, int64_t samples_per_track,
                             uint64_t seed, cudaStream_t st, int64_t track_stride) {
    if (n_tracks <= 0 || samples_per_track <= 0) return cudaSuccess;
    const int cells = (int)((samples_per_track + kCell - 1) / kCell);
    for (int t0 = 0; t0 < n_tracks; t0 += 32768) {
        const int nt = n_tracks - t0 < 32768 ? n_tracks - t0 : 32768;
        k_synth<<<dim3(cells, nt), 256, 0, st>>>(d_pcm + (int64_t)t0 * samples_per_track, first_track + track_stride * t0,
                                                 track_stride, samples_per_track, seed);
    }
    return cudaGetLastError();
}
