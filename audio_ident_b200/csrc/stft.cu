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

// In-register 32-point complex FFT, radix-2 decimation in time, fully unrolled so every twiddle is an
// immediate. Input x[n] sits in element bitrev5(n); output X[k] is left in element k. A general butterfly
// costs 6 FMA: X = a + w*b (4 FMA), Y = 2a - X (2 FMA).
template <int A, int B, int TW>
__device__ __forceinline__ void butterfly(float (&re)[32], float (&im)[32]) {
    const float ar = re[A], ai = im[A], br = re[B], bi = im[B];
    if constexpr (TW == 0) {
        re[A] = ar + br; im[A] = ai + bi; re[B] = ar - br; im[B] = ai - bi;
    } else if constexpr (TW == 8) {                          // w = -i
        re[A] = ar + bi; im[A] = ai - br; re[B] = ar - bi; im[B] = ai + br;
    } else {
        constexpr float c = kC32[TW], sn = kS32[TW];         // w = c - i sn
        const float xr = fmaf(bi, sn, fmaf(br, c, ar));
        const float xi = fmaf(bi, c, fmaf(-br, sn, ai));
        re[A] = xr; im[A] = xi;
        re[B] = fmaf(2.0f, ar, -xr); im[B] = fmaf(2.0f, ai, -xi);
    }
}

template <int S, int I>
__device__ __forceinline__ void stage(float (&re)[32], float (&im)[32]) {
    if constexpr (I < 16) {
        constexpr int half = 1 << S;
        constexpr int g = I / half, k = I % half;
        butterfly<g * 2 * half + k, g * 2 * half + k + half, k * (16 >> S)>(re, im);
        stage<S, I + 1>(re, im);
    }
}

__device__ __forceinline__ void fft32(float (&re)[32], float (&im)[32]) {
    stage<0, 0>(re, im); stage<1, 0>(re, im); stage<2, 0>(re, im); stage<3, 0>(re, im); stage<4, 0>(re, im);
}

// log(1 + s4/4) for s4 >= 0: one FFMA, one MUFU.LG2 (the argument is >= 1, so no denormal handling), one FMUL
__device__ __forceinline__ float log1p_quarter(float s4) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(0.25f, s4, 1.0f)));
    return y * 0.69314718055994531f;
}

#ifndef AID_STFT_WINDOW_IN_SMEM
#define AID_STFT_WINDOW_IN_SMEM 1
#endif
#ifndef AID_STFT_MIN_CTAS
#define AID_STFT_MIN_CTAS 4
#endif
__global__ void __launch_bounds__(kWarpsPerCta * 32, AID_STFT_MIN_CTAS)
k_stft(const float* __restrict__ window, const float2* __restrict__ twiddle,
       const float* __restrict__ pcm, const aid_stft_unit* __restrict__ units, int n_units,
       float* __restrict__ spec) {
    __shared__ float2 s_tw[32 * 32];
    __shared__ float2 s_tile[kWarpsPerCta][kTileFloats];
#if AID_STFT_WINDOW_IN_SMEM
    __shared__ float s_win[AID_NFFT];
    for (int i = threadIdx.x; i < AID_NFFT; i += blockDim.x) s_win[i] = window[i];
#endif

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) s_tw[i] = twiddle[i];
    __syncthreads();

    const int unit_id = blockIdx.x * kWarpsPerCta + warp;
    if (unit_id >= n_units) return;
    const aid_stft_unit u = units[unit_id];

#if !AID_STFT_WINDOW_IN_SMEM
    float w[32];
#pragma unroll
    for (int j = 0; j < 32; j++) w[j] = __ldg(window + lane + 32 * j);
#endif

    // lane's samples: xp[32*m], m = 0.. ; rem = samples left from xp (32-bit: a track has < 2^31 samples)
    const int64_t first = (int64_t)u.frame0 * AID_HOP + lane;
    const float* xp = pcm + u.pcm_begin + first;
    int rem = (int)(u.n_samples - first);

    float ring[36];
#pragma unroll
    for (int j = 0; j < 36; j++) ring[j] = 32 * j < rem ? __ldg(xp + 32 * j) : 0.0f;

    float2* tile = s_tile[warp];
    const int partner = (32 - lane) & 31;
    float* row_a = spec + u.spec_row * AID_NBINS + lane;

    for (int p = 0; p < u.n_frames; p += 2) {
        float nxt[8];
#pragma unroll
        for (int j = 0; j < 8; j++) nxt[j] = 32 * (36 + j) < rem ? __ldg(xp + 32 * (36 + j)) : 0.0f;
        xp += 2 * AID_HOP;
        rem -= 2 * AID_HOP;

        float re[32], im[32];
#pragma unroll
        for (int j = 0; j < 32; j++) {
#if AID_STFT_WINDOW_IN_SMEM
            const float wj = s_win[lane + 32 * j];
#else
            const float wj = w[j];
#endif
            re[bitrev5(j)] = wj * ring[j]; im[bitrev5(j)] = wj * ring[j + 4];
        }

        fft32(re, im);                                   // over n2: Y[k1] in element k1

        // twiddle W_1024^(lane * k1), then scatter Y[k1] to tile[k1][lane]
        tile[lane] = make_float2(re[0], im[0]);
#pragma unroll
        for (int k1 = 1; k1 < 32; k1++) {
            const float2 t = s_tw[k1 * 32 + lane];
            tile[k1 * kTileStride + lane] = make_float2(fmaf(re[k1], t.x, -im[k1] * t.y), fmaf(re[k1], t.y, im[k1] * t.x));
        }
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < 32; n1++) {                // lane = k1 gathers all n1
            const float2 z = tile[lane * kTileStride + n1];
            re[bitrev5(n1)] = z.x; im[bitrev5(n1)] = z.y;
        }
        __syncwarp();

        fft32(re, im);                                   // over n1: Z[lane + 32*k2] in element k2

        const bool has_b = p + 1 < u.n_frames;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const float zr = re[k2], zi = im[k2];
            // mirror Z[1024 - k]: lane (32-k1)&31, k2' = 31 - k2 (k1 != 0) or (32 - k2)&31 (k1 == 0)
            const float sr = __shfl_sync(AID_FULL_MASK, re[31 - k2], partner);
            const float si = __shfl_sync(AID_FULL_MASK, im[31 - k2], partner);
            const float mr = lane == 0 ? re[(32 - k2) & 31] : sr;
            const float mi = lane == 0 ? im[(32 - k2) & 31] : si;
            const float ar = zr + mr, ai = zi - mi;      // 2 * X_a[k]
            const float br = zr - mr, bi = zi + mi;      // 2i * X_b[k]
            row_a[32 * k2] = log1p_quarter(fmaf(ar, ar, ai * ai));
            if (has_b) row_a[AID_NBINS + 32 * k2] = log1p_quarter(fmaf(br, br, bi * bi));
        }
        row_a += 2 * AID_NBINS;

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
