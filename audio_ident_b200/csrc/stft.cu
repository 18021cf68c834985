// stft.cu -- batched framing + Hamming window + 1024-point real FFT + log(1 + power), sm_100a.
//
// Replaces stage a4 of SURVEY.md section 8(a): the STFT that the reference leaves to the external
// `olaf_c` process (reference audio-ident-service/app/audio/fingerprint.py:117-125). Definition of the
// result: oracle/aid_oracle.c frame_spectrum(); tolerance AID_SPEC_TOL.
//
// Design (see DESIGN.md "STFT kernel"):
//  * One warp owns a run of consecutive frames of one track. Two frames a, b = a+1 are packed into one
//    1024-point COMPLEX transform z = a + i*b, factored 32 x 32: lane n1 holds z[n1 + 32*n2] for
//    n2 = 0..31 in registers, does a 32-point FFT over n2 (its first stage fused with the window
//    multiply), the warp transposes through its private shared-memory tile, lane k1 does the second
//    32-point transform over n1 with the twiddles W_1024^(n1*k1) folded into its butterflies (AID_STFT_UNFOLD = 1:
//    31 complex multiplies while gathering the row, then constant twiddles as immediates) and ends up
//    holding Z[k1 + 32*k2]. Z[N-k] lives in lane (32-k1)&31, so the two real spectra are separated with
//    one shuffle per value.
//  * Because the hop is 128 = 4*32 samples, lane n1 needs x[32*m + n1] for a window of m that slides by 4
//    per frame: PCM goes global -> registers with fully coalesced 128 B warp loads and every sample is
//    read from HBM once per unit (8x frame overlap is served from the register ring, never re-read).
//  * Rows of the spectrogram (512 floats = 2 KB) are written with 128 B coalesced warp stores.
// The kernel is FP32-issue-bound, not HBM-bound (about 1.1 k FP32 instructions per lane per frame pair);
// packed FFMA2/FADD2 were measured to issue at half rate on sm_100a (tools/microbench/fp32_issue.cu), so the
// butterflies stay scalar.
#include "common.cuh"

namespace {

#ifndef AID_STFT_WARPS
#define AID_STFT_WARPS 4
#endif
constexpr int kWarpsPerCta = AID_STFT_WARPS;
constexpr int kTileStride = 34;                        // float2 per tile row: 16 B aligned rows, conflict-free STS.64 / LDS.128
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

// All five stages with constant twiddles (second transform of the AID_STFT_UNFOLD variant).
__device__ __forceinline__ void fft32_const(float (&re)[32], float (&im)[32]) {
    stage<0, 0>(re, im); stage<1, 0>(re, im); stage<2, 0>(re, im); stage<3, 0>(re, im); stage<4, 0>(re, im);
}

// Stages 1..4 of the DIT transform (stage 0 is fused with the window multiply, see k_stft).
__device__ __forceinline__ void fft32_after_stage0(float (&re)[32], float (&im)[32]) {
    stage<1, 0>(re, im); stage<2, 0>(re, im); stage<3, 0>(re, im); stage<4, 0>(re, im);
}

// ---- second transform: the inter-stage twiddle W_1024^(n1*k1) is folded into the butterflies ----
// Lane k1 needs Z[k2] = sum_n1 y[n1] g^n1 W_32^(n1*k2) with g = W_1024^k1. Splitting n1 into even and odd
// gives the usual DIT recursion with the twist squared at every level, so stage S (half = 2^S) uses the
// twiddles g^(16>>S) * W_(2*half)^k, k < half: lane-dependent values from a table (AID_TWIST_*), 16 complex
// per lane because W^(k + half/2) = -i W^k reuses the same pair with the roles of c and s exchanged.
// All 80 butterflies are general (6 FMA), which is still fewer instructions than 31 complex multiplies
// followed by a transform with constant twiddles, and the table costs 8 LDS.128 instead of 31 LDS.64.
template <int A, int B, bool ROT>
__device__ __forceinline__ void twisted_butterfly(float (&re)[32], float (&im)[32], float c, float s) {
    const float ar = re[A], ai = im[A], br = re[B], bi = im[B];
    float xr, xi;
    if constexpr (!ROT) {                                   // w = c - i s
        xr = fmaf(bi, s, fmaf(br, c, ar));
        xi = fmaf(-br, s, fmaf(bi, c, ai));
    } else {                                                // w = -i (c - i s) = -s - i c
        xr = fmaf(bi, c, fmaf(-br, s, ar));
        xi = fmaf(-br, c, fmaf(-bi, s, ai));
    }
    re[A] = xr; im[A] = xi;
    re[B] = fmaf(2.0f, ar, -xr); im[B] = fmaf(2.0f, ai, -xi);
}

// butterflies k' and k' + half/2 of every group of stage S, for the two twiddles held in one float4
template <int S, int KP, int G>
__device__ __forceinline__ void twisted_groups(float (&re)[32], float (&im)[32], float c, float s) {
    constexpr int half = 1 << S;
    if constexpr (G < 16 / half) {
        constexpr int a0 = G * 2 * half + KP;
        twisted_butterfly<a0, a0 + half, false>(re, im, c, s);
        if constexpr (half >= 2) twisted_butterfly<a0 + half / 2, a0 + half / 2 + half, true>(re, im, c, s);
        twisted_groups<S, KP, G + 1>(re, im, c, s);
    }
}

__device__ __forceinline__ void fft32_twisted(float (&re)[32], float (&im)[32], const float4* tw) {
    const float4 t01 = tw[0];                               // stage 0: g^16 ; stage 1: g^8
    twisted_groups<0, 0, 0>(re, im, t01.x, t01.y);
    twisted_groups<1, 0, 0>(re, im, t01.z, t01.w);
    const float4 t2 = tw[1];                                // stage 2: g^4 W_8^{0,1}
    twisted_groups<2, 0, 0>(re, im, t2.x, t2.y);
    twisted_groups<2, 1, 0>(re, im, t2.z, t2.w);
    const float4 t3a = tw[2], t3b = tw[3];                  // stage 3: g^2 W_16^{0..3}
    twisted_groups<3, 0, 0>(re, im, t3a.x, t3a.y);
    twisted_groups<3, 1, 0>(re, im, t3a.z, t3a.w);
    twisted_groups<3, 2, 0>(re, im, t3b.x, t3b.y);
    twisted_groups<3, 3, 0>(re, im, t3b.z, t3b.w);
#pragma unroll
    for (int q = 0; q < 4; q++) {                           // stage 4: g W_32^{0..7}
        const float4 t = tw[4 + q];
        if (q == 0) { twisted_groups<4, 0, 0>(re, im, t.x, t.y); twisted_groups<4, 1, 0>(re, im, t.z, t.w); }
        if (q == 1) { twisted_groups<4, 2, 0>(re, im, t.x, t.y); twisted_groups<4, 3, 0>(re, im, t.z, t.w); }
        if (q == 2) { twisted_groups<4, 4, 0>(re, im, t.x, t.y); twisted_groups<4, 5, 0>(re, im, t.z, t.w); }
        if (q == 3) { twisted_groups<4, 6, 0>(re, im, t.x, t.y); twisted_groups<4, 7, 0>(re, im, t.z, t.w); }
    }
}

// log(1 + re^2 + im^2): two FFMA (the "+1" rides on the first), one MUFU.LG2 (the argument is >= 1, so no denormal
// handling), one FMUL
__device__ __forceinline__ float log1p_power(float re, float im) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(fmaf(re, re, fmaf(im, im, 1.0f))));
    return y * 0.69314718055994531f;
}

#ifndef AID_STFT_MIN_CTAS
#define AID_STFT_MIN_CTAS 3
#endif
// Timing-only ablation switches for tools/microbench/stft_bench.cu (profiles/r01_stft_v3.md); 0 in the product build.
// 1: one store per value pair instead of two  2: no logarithm  4: no mirror shuffles  8: no transpose through shared
// memory  16: second transform skipped  32: first transform (stages 1-4) skipped. Results are wrong with any bit set.
#ifndef AID_STFT_ABLATE
#define AID_STFT_ABLATE 0
#endif
constexpr int kAblate = AID_STFT_ABLATE;
// 1: the inter-transform twiddle W_1024^(n1*k1) is applied as 31 complex multiplies while the transposed row is
//    gathered, and the second transform uses constant twiddles (FFMA with immediates: two register operands) instead
//    of folding the twiddle into 80 general butterflies whose FFMAs read three registers.
// 2: window products as FMUL + FADD/FADD instead of FMUL + two three-register FFMAs.
// Measured alone (profiles/r01_stft_v3.md): 0 -> 53.4 %, 1 -> 54.6 %, 2 -> 52.6 %, 3 -> 53.8 % of the HBM roofline. Inside
// the ingest step (power-capped) variant 1 gains 0.7 % and its slightly larger rounding error (1.3e-6 vs 7e-7 scaled)
// flips more near-tie peaks against the double-precision oracle (15 vs 9 of 1024 tracks), so 0 stays the default.
#ifndef AID_STFT_UNFOLD
#define AID_STFT_UNFOLD 0
#endif
constexpr int kUnfold = AID_STFT_UNFOLD;
constexpr int kTwistStride = (kUnfold & 1) ? 68 : 36;  // floats per lane row of the twiddle table (16 B aligned, conflict-free LDS.128)
constexpr int kTabStride = 36;                         // floats per lane row of the two lane-major tables (16 B aligned, conflict-free LDS.128)

// 3 CTAs x 4 warps per SM: 142 registers, no spills. Measured alternatives (tools/microbench/stft_bench.cu, 2048 x 30 s
// tracks): 4 CTAs (128 registers, 20 B of spills) 50.9 %, 7 x 2 warps at 144 registers 52.8 %, this shape 52.8 % of the
// HBM roofline; unrolling the pair loop to rename the sample ring away costs more in instruction fetch than the 40 MOVs
// it removes (44 % / 36 % at x2 / x3); twiddles of the first transform in registers instead of immediates 51.8 %;
// window table (partly) in registers: no change. Ablations are in profiles/r01_stft_v3.md.
__global__ void __launch_bounds__(kWarpsPerCta * 32, AID_STFT_MIN_CTAS)
k_stft(const float* __restrict__ window, const float* __restrict__ twist,
       const float* __restrict__ pcm, const aid_stft_unit* __restrict__ units, int n_units,
       float* __restrict__ spec) {
    __shared__ __align__(16) float s_win[32 * kTabStride];       // [lane][j]  = window[lane + 32 j]
    __shared__ __align__(16) float s_twist[32 * kTwistStride];   // [lane][..] = one of the two AID_TWIST layouts (common.cuh)
    __shared__ __align__(16) float2 s_tile[kWarpsPerCta][kTileFloats];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
        s_win[(i & 31) * kTabStride + (i >> 5)] = 0.5f * window[i];      // exact halving: Z = X_a + i X_b after the mirror sums below
        if constexpr (!(kUnfold & 1)) s_twist[(i >> 5) * kTwistStride + (i & 31)] = twist[i];
    }
    if constexpr (kUnfold & 1)
        for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) s_twist[(i >> 6) * kTwistStride + (i & 63)] = twist[AID_TWIST_FOLDED + i];
    __syncthreads();

    const int unit_id = blockIdx.x * kWarpsPerCta + warp;
    if (unit_id >= n_units) return;
    const aid_stft_unit u = units[unit_id];

    // lane's samples: xp[32*m], m = 0.. ; rem = samples left from xp (32-bit: a track has < 2^31 samples)
    const int64_t first = (int64_t)u.frame0 * AID_HOP + lane;
    const float* xp = pcm + u.pcm_begin + first;
    int rem = (int)(u.n_samples - first);

    float ring[36];
#pragma unroll
    for (int j = 0; j < 36; j++) ring[j] = 32 * j < rem ? __ldg(xp + 32 * j) : 0.0f;

    float2* tile = s_tile[warp];
    const float4* tile_row = reinterpret_cast<const float4*>(tile + lane * kTileStride);
    const float4* win4 = reinterpret_cast<const float4*>(s_win + lane * kTabStride);
    const float4* twist4 = reinterpret_cast<const float4*>(s_twist + lane * kTwistStride);
    const int partner = (32 - lane) & 31;
    float* row_a = spec + u.spec_row * AID_NBINS + lane;

    for (int p = 0; p < u.n_frames; p += 2) {
        float nxt[8];
#pragma unroll
        for (int j = 0; j < 8; j++) nxt[j] = 32 * (36 + j) < rem ? __ldg(xp + 32 * (36 + j)) : 0.0f;
        xp += 2 * AID_HOP;
        rem -= 2 * AID_HOP;

        // window multiply fused with DIT stage 0 (pairs samples j and j + 16): frame a in re, frame b = a + 1 in im
        float re[32], im[32];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const float4 wa4 = win4[q], wb4 = win4[4 + q];
            const float wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int ja = 4 * q + r, jb = ja + 16, e = bitrev5(ja);
                const float tr = wb[r] * ring[jb], ti = wb[r] * ring[jb + 4];
                if constexpr (kUnfold & 2) {
                    const float ur = __fmul_rn(wa[r], ring[ja]), ui = __fmul_rn(wa[r], ring[ja + 4]);   // no contraction
                    re[e] = __fadd_rn(ur, tr); re[e + 1] = __fsub_rn(ur, tr);
                    im[e] = __fadd_rn(ui, ti); im[e + 1] = __fsub_rn(ui, ti);
                } else {
                    re[e] = fmaf(wa[r], ring[ja], tr);     re[e + 1] = fmaf(wa[r], ring[ja], -tr);
                    im[e] = fmaf(wa[r], ring[ja + 4], ti); im[e + 1] = fmaf(wa[r], ring[ja + 4], -ti);
                }
            }
        }
        if constexpr (!(kAblate & 32)) fft32_after_stage0(re, im);   // over n2: Y[k1] in element k1

        // transpose: lane n1 scatters Y[k1] to tile[k1][n1]; lane k1 gathers its row two values at a time
        if constexpr (!(kAblate & 8)) {
#pragma unroll
            for (int k1 = 0; k1 < 32; k1++) tile[k1 * kTileStride + lane] = make_float2(re[k1], im[k1]);
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float4 z = tile_row[q];
                if constexpr (kUnfold & 1) {                 // y[n1] * W_1024^(n1*k1), w = c - i s
                    const float4 w = twist4[q];
                    if (q == 0) { re[0] = z.x; im[0] = z.y; }
                    else {
                        re[bitrev5(2 * q)] = fmaf(z.x, w.x, z.y * w.y);
                        im[bitrev5(2 * q)] = fmaf(z.y, w.x, -(z.x * w.y));
                    }
                    re[bitrev5(2 * q + 1)] = fmaf(z.z, w.z, z.w * w.w);
                    im[bitrev5(2 * q + 1)] = fmaf(z.w, w.z, -(z.z * w.w));
                } else {
                    re[bitrev5(2 * q)] = z.x;     im[bitrev5(2 * q)] = z.y;
                    re[bitrev5(2 * q + 1)] = z.z; im[bitrev5(2 * q + 1)] = z.w;
                }
            }
            __syncwarp();
        }

        if constexpr (!(kAblate & 16)) {                     // over n1: Z[lane + 32*k2] in element k2
            if constexpr (kUnfold & 1) fft32_const(re, im); else fft32_twisted(re, im, twist4);
        }

        const bool has_b = p + 1 < u.n_frames;
#pragma unroll
        for (int k2 = 0; k2 < 16; k2++) {
            const float zr = re[k2], zi = im[k2];
            // mirror Z[1024 - k]: lane (32-k1)&31, k2' = 31 - k2 (k1 != 0) or (32 - k2)&31 (k1 == 0)
            const float sr = (kAblate & 4) ? re[31 - k2] : __shfl_sync(AID_FULL_MASK, re[31 - k2], partner);
            const float si = (kAblate & 4) ? im[31 - k2] : __shfl_sync(AID_FULL_MASK, im[31 - k2], partner);
            const float mr = lane == 0 ? re[(32 - k2) & 31] : sr;
            const float mi = lane == 0 ? im[(32 - k2) & 31] : si;
            const float ar = zr + mr, ai = zi - mi;      // X_a[k]   (the window table is pre-scaled by 1/2)
            const float br = zr - mr, bi = zi + mi;      // i * X_b[k]
            if constexpr (kAblate & 2) {
                row_a[32 * k2] = fmaf(ar, ar, ai * ai);
                if (has_b) row_a[AID_NBINS + 32 * k2] = fmaf(br, br, bi * bi);
            } else if constexpr (kAblate & 1) {
                row_a[32 * k2] = log1p_power(ar, ai) + log1p_power(br, bi);
            } else {
                row_a[32 * k2] = log1p_power(ar, ai);
                if (has_b) row_a[AID_NBINS + 32 * k2] = log1p_power(br, bi);
            }
        }
        row_a += 2 * AID_NBINS;

#pragma unroll
        for (int j = 0; j < 28; j++) ring[j] = ring[j + 8];
#pragma unroll
        for (int j = 0; j < 8; j++) ring[28 + j] = nxt[j];
    }
}

#ifdef AID_STFT_TC
#include "../../tools/microbench/stft_tc.cuh"      // experimental tensor-core second transform (micro-benchmarks only)
#endif

}  // namespace

cudaError_t aid_launch_stft(const aid_tables& tb, const float* d_pcm, const aid_stft_unit* d_units,
                            int n_units, float* d_spec, cudaStream_t st) {
    if (n_units <= 0) return cudaSuccess;
#ifdef AID_STFT_TC
    static bool configured = false;
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(k_stft_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StftTcSmem) + 1024);
        if (ce != cudaSuccess) return ce;
        configured = true;
    }
    k_stft_tc<<<(n_units + 3) / 4, 128, sizeof(StftTcSmem) + 1024, st>>>(tb.window, tb.twist, d_pcm, d_units, n_units, d_spec);
    return cudaGetLastError();
#else
    const int grid = (n_units + kWarpsPerCta - 1) / kWarpsPerCta;
    k_stft<<<grid, kWarpsPerCta * 32, 0, st>>>(tb.window, tb.twist, d_pcm, d_units, n_units, d_spec);
    return cudaGetLastError();
#endif
}
