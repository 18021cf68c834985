// stft.cu -- batched framing + Hamming window + 1024-point real FFT + log(1 + power), sm_100a.
//
// Replaces stage a4 of SURVEY.md section 8(a): the STFT that the reference leaves to the external
// `olaf_c` process (reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the
// result: oracle/aid_oracle.c frame_spectrum(); tolerance AID_SPEC_TOL.
//
// Design (see DESIGN.md "STFT kernel"):
//  * One warp owns a run of consecutive frames of one track. Two frames a, b = a+1 are packed into one
//    1024-point COMPLEX transform z = a + i*b, factored 32 x 32: lane n1 holds z[n1 + 32*n2] for
//    n2 = 0..31 in registers, does a 32-point FFT over n2, multiplies by W_1024^(n1*k1), the warp
//    transposes through its private shared-memory tile, lane k1 does the second 32-point FFT over n1
//    and ends up holding Z[k1 + 32*k2]. Z[N-k] lives in lane (32-k1)&31, so the two real spectra are
//    separated with one shuffle per value.
//  * Because the hop is 128 = 4*32 samples, lane n1 needs x[32*m + n1] for a window of m that slides by 4
//    per frame: PCM goes global -> registers with fully coalesced 128 B warp loads and every sample is
//    read from HBM once per unit (8x frame overlap is served from the register ring, never re-read).
//  * Rows of the spectrogram (512 floats = 2 KB) are written with 128 B coalesced warp stores.
// The kernel is FP32-issue-bound, not HBM-bound (about 1.3 k FP32 instructions per lane per frame pair).
#include "common.cuh"

namespace {

constexpr int kWarpsPerCta = 4;
constexpr int kTileStride = 33;                        // 32 x 33 floats: conflict-free transpose
constexpr int kTileFloats = 32 * kTileStride;

// W_32^k = cos(2 pi k/32) - i sin(2 pi k/32), k = 0..15
__device__ constexpr float kC32[16] = {
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254524f,
    0.70710678118654757f, 0.55557023301960229f, 0.38268343236508984f, 0.19509032201612833f,
    0.0f, -0.19509032201612819f, -0.38268343236508973f, -0.55557023301960196f,
    -0.70710678118654746f, -0.83146961230254535f, -0.92387953251128674f, -0.98078528040323043f};
__device__ constexpr float kS32[16] = {
    0.0f, 0.19509032201612825f, 0.38268343236508978f, 0.55557023301960218f,
    0.70710678118654746f, 0.83146961230254524f, 0.92387953251128674f, 0.98078528040323043f,
    1.0f, 0.98078528040323043f, 0.92387953251128674f, 0.83146961230254546f,
    0.70710678118654757f, 0.55557023301960218f, 0.38268343236508989f, 0.19509032201612861f};

__host__ __device__ constexpr int bitrev5(int i) {
    return ((i & 1) << 4) | ((i & 2) << 2) | (i & 4) | ((i & 8) >> 2) | ((i & 16) >> 4);
}

// In-register 32-point complex FFT, radix-2 decimation in frequency, fully unrolled so every
// twiddle is an immediate. Output X[k] is left in element bitrev5(k).
__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
#pragma unroll
    for (int s = 0; s < 5; s++) {
        const int half = 16 >> s;
#pragma unroll
        for (int g = 0; g < (1 << s); g++) {
#pragma unroll
            for (int k = 0; k < half; k++) {
                const int a = g * 2 * half + k, b = a + half;
                const int tw = k << s;                       // W_32^tw
                const float ar = re[a], ai = im[a], br = re[b], bi = im[b];
                re[a] = ar + br; im[a] = ai + bi;
                const float dr = ar - br, di = ai - bi;
                if (tw == 0) { re[b] = dr; im[b] = di; }
                else if (tw == 8) { re[b] = di; im[b] = -dr; }            // * (-i)
                else {
                    const float c = kC32[tw], sn = kS32[tw];              // (dr + i di)(c - i sn)
                    re[b] = fmaf(dr, c, di * sn);
                    im[b] = fmaf(di, c, -dr * sn);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_stft(const float* __restrict__ window, const float2* __restrict__ twiddle,
       const float* __restrict__ pcm, const aid_stft_unit* __restrict__ units, int n_units,
       float* __restrict__ spec) {
    __shared__ float2 s_tw[32 * 32];
    __shared__ float s_tile[kWarpsPerCta][2][kTileFloats];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) s_tw[i] = twiddle[i];
    __syncthreads();

    const int unit_id = blockIdx.x * kWarpsPerCta + warp;
    if (unit_id >= n_units) return;
    const aid_stft_unit u = units[unit_id];

    float w[32];
#pragma unroll
    for (int j = 0; j < 32; j++) w[j] = __ldg(window + lane + 32 * j);

    // lane's samples: x[first + 32*m + lane], first = frame0 * 128
    const float* x = pcm + u.pcm_begin;
    const int64_t first = (int64_t)u.frame0 * AID_HOP + lane;
    const int64_t n = u.n_samples;
    auto sample = [&](int m) -> float {
        const int64_t idx = first + 32 * (int64_t)m;
        return idx < n ? __ldg(x + idx) : 0.0f;
    };

    float ring[36];
#pragma unroll
    for (int j = 0; j < 36; j++) ring[j] = sample(j);

    float* tile_re = s_tile[warp][0];
    float* tile_im = s_tile[warp][1];
    const int partner = (32 - lane) & 31;

    for (int p = 0; p < u.n_frames; p += 2) {
        float nxt[8];
#pragma unroll
        for (int j = 0; j < 8; j++) nxt[j] = sample(4 * p + 36 + j);

        float re[32], im[32];
#pragma unroll
        for (int j = 0; j < 32; j++) { re[j] = w[j] * ring[j]; im[j] = w[j] * ring[j + 4]; }

        fft32(re, im);                                   // over n2; Y[k1] in element bitrev5(k1)

        // twiddle W_1024^(lane * k1), then scatter Y[k1] to tile[k1][lane]
#pragma unroll
        for (int k1 = 0; k1 < 32; k1++) {
            const int e = bitrev5(k1);
            float yr = re[e], yi = im[e];
            if (k1 != 0) {
                const float2 t = s_tw[k1 * 32 + lane];
                const float r2 = yr * t.x - yi * t.y;
                yi = fmaf(yr, t.y, yi * t.x);
                yr = r2;
            }
            tile_re[k1 * kTileStride + lane] = yr;
            tile_im[k1 * kTileStride + lane] = yi;
        }
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {                // lane = k1 gathers all n1
            re[n1] = tile_re[lane * kTileStride + n1];
            im[n1] = tile_im[lane * kTileStride + n1];
        }
        __syncwarp();

        fft32(re, im);                                   // over n1; Z[lane + 32*k2] in element bitrev5(k2)

        float* row_a = spec + (u.spec_row + p) * AID_NBINS;
        const bool has_b = p + 1 < u.n_frames;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const int e = bitrev5(k2);
            const float zr = re[e], zi = im[e];
            // mirror Z[1024 - k]: lane (32-k1)&31, k2' = 31 - k2 (k1 != 0) or (32 - k2)&31 (k1 == 0)
            float mr = __shfl_sync(AID_FULL_MASK, re[bitrev5(31 - k2)], partner);
            float mi = __shfl_sync(AID_FULL_MASK, im[bitrev5(31 - k2)], partner);
            if (lane == 0) { mr = re[bitrev5((32 - k2) & 31)]; mi = im[bitrev5((32 - k2) & 31)]; }
            const float ar = zr + mr, ai = zi - mi;      // 2 * X_a[k]
            const float br = zr - mr, bi = zi + mi;      // 2i * X_b[k]
            const float pa = 0.25f * fmaf(ar, ar, ai * ai);
            const float pb = 0.25f * fmaf(br, br, bi * bi);
            const int k = lane + 32 * k2;
            row_a[k] = __logf(1.0f + pa);
            if (has_b) row_a[AID_NBINS + k] = __logf(1.0f + pb);
        }

#pragma unroll
        for (int j = 0; j < 28; j++) ring[j] = ring[j + 8];
#pragma unroll
        for (int j = 0; j < 8; j++) ring[28 + j] = nxt[j];
    }
}

}  // namespace

cudaError_t aid_launch_stft(const aid_tables& tb, const float* d_pcm, const aid_stft_unit* d_units,
                            int n_units, float* d_spec, cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
    const int grid = (n_units + kWarpsPerCta - 1) / kWarpsPerCta;
    k_stft<<<grid, kWarpsPerCta * 32, 0, st>>>(tb.window, tb.twiddle, d_pcm, d_units, n_units, d_spec);
    return cudaGetLastError();
}
