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
// Two kernels with bit-identical output:
//  * k_stft        (round 1) scalar FP32 butterflies; issue bound (1,326 instructions per frame pair, 1,060 of them FP32);
//                  kept as the reference point of the A/B leg in bench.py and of tests/test_gpu_fingerprint.py.
//  * k_stft_packed (round 2, the product kernel) the same operations as two-lane f32x2 instructions (FFMA2 / FADD2 / FMUL2):
//                  897 instructions per frame pair, bound by the FP32 pipe itself; optionally also writes the maxima of
//                  every row's 16-bin groups for the peak kernel. Measurements: profiles/r02_stft_packed.md.
#include "common.cuh"
#include <stdlib.h>

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

// =====================================================================================================
// Packed variant (round 2): the same transform, the same operations on every element in the same order --
// the output is bit-identical to k_stft above -- but issued as f32x2 instructions (FFMA2 / FADD2 / FMUL2,
// PTX fma.rn.f32x2 etc., sm_100+) wherever two elements go through the same butterfly.
//
// Why: k_stft is ISSUE bound, not FMA-pipe bound (ncu: 79 % issue slots busy, FMA pipe 63 %): 1,060 of its
// 1,326 instructions per frame pair are FP32. A packed instruction does two lanes' worth of work in ONE issue
// slot (the FMA pipe is busy two cycles, the slot is free for LDS / STS / SHFL / MUFU / integer work), so the
// FP32 part needs 624 slots instead of 1,060 and the kernel moves from the issue limit to the FMA-pipe limit.
//
// Pairing. A 64-bit register pair holds two elements of the 32-element array whose indices differ in ONE bit P.
// A radix-2 stage S pairs elements differing in bit S, so for S != P both halves of a register pair take part in
// two butterflies of the same shape (packed); for S == P the butterfly runs between the halves (scalar, the halves
// are ordinary 32-bit registers). The scalar stage is chosen where a layout has to change anyway:
//  * first transform: P = 4. Elements (e, e + 16) come from ring samples (2i, 2i + 1) (and the frame b = a + 1
//    uses samples + 4, which keeps the parity), the window table delivers (w[2i], w[2i+1]) as one 64-bit half of an
//    LDS.128, and the twiddles of stages 0-3 are the same for both halves: FFMA2 with a broadcast immediate.
//    Stage 4 (k, k + 16) is the scalar one; its results go straight into the (re, im) pairs of the STS.64.
//  * second transform: P = 0. Stage 0 is the scalar one: it reads the (re, im, re, im) quads of the LDS.128 as they
//    come and writes X and Y = 2a - X into the two halves of a pair. Stages 1-4 then use a DIFFERENT twiddle in
//    each half (k and k + 1): the shared table holds them as (c_k, c_k+1, s_k, s_k+1), one LDS.128 per butterfly
//    shape and no duplicated entries; k + half/2 reuses the same quad with the roles of c and s exchanged.
//  * the two real spectra separate pairwise as well: outputs (k2, k2 + 1) are one pair, their mirrors arrive by
//    two shuffles into the two halves of one pair.
// Timing-only ablations of k_stft_packed for tools/microbench (results are wrong with any bit set; 0 in the product build):
// 1: every table read uses quad 0 (one LDS.128 per table and trip)  2: no transpose through shared memory
// 4: no global stores  8: no mirror shuffles
#ifndef AID_STFT_PABLATE
#define AID_STFT_PABLATE 0
#endif
constexpr int kPA = AID_STFT_PABLATE;
typedef unsigned long long f2;
__device__ __forceinline__ f2 pk(float l, float h) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(l), "f"(h)); return r; }
__device__ __forceinline__ float lo(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a; }
__device__ __forceinline__ float hi(f2 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return b; }
__device__ __forceinline__ f2 bc(float c) { return pk(c, c); }
__device__ __forceinline__ f2 neg2(f2 v) { return pk(-lo(v), -hi(v)); }          // folds into the operand's sign modifier
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }

// first transform, stages 1-3 on pairs (e, e + 16): element indices A, B < 16, constant twiddle for both halves
template <int A, int B, int TW>
__device__ __forceinline__ void pbutterfly(f2 (&re)[16], f2 (&im)[16]) {
    const f2 ar = re[A], ai = im[A], br = re[B], bi = im[B];
    if constexpr (TW == 0) {
        re[A] = add2(ar, br); im[A] = add2(ai, bi); re[B] = sub2(ar, br); im[B] = sub2(ai, bi);
    } else if constexpr (TW == 8) {
        re[A] = add2(ar, bi); im[A] = sub2(ai, br); re[B] = sub2(ar, bi); im[B] = add2(ai, br);
    } else {
        constexpr float c = kC32[TW], sn = kS32[TW];
        const f2 xr = fma2(bi, bc(sn), fma2(br, bc(c), ar));
        const f2 xi = fma2(bi, bc(c), fma2(neg2(br), bc(sn), ai));
        re[A] = xr; im[A] = xi;
        re[B] = fma2(bc(2.0f), ar, neg2(xr)); im[B] = fma2(bc(2.0f), ai, neg2(xi));
    }
}
template <int S, int I>
__device__ __forceinline__ void pstage(f2 (&re)[16], f2 (&im)[16]) {
    if constexpr (I < 8) {
        constexpr int half = 1 << S;
        constexpr int g = I / half, k = I % half;
        pbutterfly<g * 2 * half + k, g * 2 * half + k + half, k * (16 >> S)>(re, im);
        pstage<S, I + 1>(re, im);
    }
}

// first transform, stage 4: between the halves of pair K; X -> element K, Y -> element K + 16
template <int K>
__device__ __forceinline__ void last_stage_store(const f2 (&re)[16], const f2 (&im)[16], float2* tile_col) {
    const float ar = lo(re[K]), ai = lo(im[K]), br = hi(re[K]), bi = hi(im[K]);
    float xr, xi, yr, yi;
    if constexpr (K == 0) { xr = ar + br; xi = ai + bi; yr = ar - br; yi = ai - bi; }
    else if constexpr (K == 8) { xr = ar + bi; xi = ai - br; yr = ar - bi; yi = ai + br; }
    else {
        constexpr float c = kC32[K], sn = kS32[K];
        xr = fmaf(bi, sn, fmaf(br, c, ar));
        xi = fmaf(bi, c, fmaf(-br, sn, ai));
        yr = fmaf(2.0f, ar, -xr); yi = fmaf(2.0f, ai, -xi);
    }
    if constexpr (kPA & 2) { const_cast<f2&>(re[K]) = pk(xr, yr); const_cast<f2&>(im[K]) = pk(xi, yi); }
    else {
        tile_col[K * kTileStride] = make_float2(xr, xi);
        tile_col[(K + 16) * kTileStride] = make_float2(yr, yi);
    }
    if constexpr (K < 15) last_stage_store<K + 1>(re, im, tile_col);
}

// second transform, stages 1-4 on pairs (e, e + 1): pair indices a, b = a + hp; the halves use twiddles k and k + 1
template <int A, int B, bool ROT>
__device__ __forceinline__ void tpbutterfly(f2 (&re)[16], f2 (&im)[16], f2 c, f2 s) {
    const f2 ar = re[A], ai = im[A], br = re[B], bi = im[B];
    f2 xr, xi;
    if constexpr (!ROT) {
        xr = fma2(bi, s, fma2(br, c, ar));
        xi = fma2(neg2(br), s, fma2(bi, c, ai));
    } else {
        xr = fma2(bi, c, fma2(neg2(br), s, ar));
        xi = fma2(neg2(br), c, fma2(neg2(bi), s, ai));
    }
    re[A] = xr; im[A] = xi;
    re[B] = fma2(bc(2.0f), ar, neg2(xr)); im[B] = fma2(bc(2.0f), ai, neg2(xi));
}
// all butterflies of stage S (hp = 2^(S-1) pair positions per half group) that use table entry KP: positions KP and KP + hp/2
template <int S, int KP, int G>
__device__ __forceinline__ void tpgroups(f2 (&re)[16], f2 (&im)[16], f2 c, f2 s) {
    constexpr int hp = 1 << (S - 1);
    if constexpr (G < 8 / hp) {
        constexpr int a0 = G * 2 * hp + KP;
        tpbutterfly<a0, a0 + hp, false>(re, im, c, s);
        if constexpr (hp >= 2) tpbutterfly<a0 + hp / 2, a0 + hp / 2 + hp, true>(re, im, c, s);
        tpgroups<S, KP, G + 1>(re, im, c, s);
    }
}

constexpr int kPTwistStride = 36;      // floats per lane row of the packed twiddle table: 8 quads + stage 0's (c, s) + pad
constexpr int kPackedSmem = kWarpsPerCta * kTileFloats * 8 + 32 * kTabStride * 4 + 32 * kPTwistStride * 4;
static_assert(32 * kTabStride + 32 * kPTwistStride == AID_TWIST_IMAGE_FLOATS, "shared-memory table image: common.cuh");

// One group (two bins, k2 = 2J and 2J + 1, of both frames) of the separation of the two real spectra + log(1 + power) + store.
template <int J>
__device__ __forceinline__ void separate_group(const f2 (&re)[16], const f2 (&im)[16], int lane, int partner,
                                               float* row_a, bool st_a, bool st_b, f2& out_a, f2& out_b) {
    const f2 zr = re[J], zi = im[J];
    // mirrors: partner's elements 31 - 2J (hi of pair 15 - J) and 30 - 2J (lo of pair 15 - J);
    // lane 0: its own elements (32 - 2J) & 31 (lo of pair (16 - J) & 15) and 31 - 2J (hi of pair 15 - J)
    const float s0r = (kPA & 8) ? hi(re[15 - J]) : __shfl_sync(AID_FULL_MASK, hi(re[15 - J]), partner);
    const float s1r = (kPA & 8) ? lo(re[15 - J]) : __shfl_sync(AID_FULL_MASK, lo(re[15 - J]), partner);
    const float s0i = (kPA & 8) ? hi(im[15 - J]) : __shfl_sync(AID_FULL_MASK, hi(im[15 - J]), partner);
    const float s1i = (kPA & 8) ? lo(im[15 - J]) : __shfl_sync(AID_FULL_MASK, lo(im[15 - J]), partner);
    const f2 mr = pk(lane == 0 ? lo(re[(16 - J) & 15]) : s0r, lane == 0 ? hi(re[15 - J]) : s1r);
    const f2 mi = pk(lane == 0 ? lo(im[(16 - J) & 15]) : s0i, lane == 0 ? hi(im[15 - J]) : s1i);
    const f2 ar = add2(zr, mr), ai = sub2(zi, mi);       // X_a[k]   (the window table is pre-scaled by 1/2)
    const f2 br = sub2(zr, mr), bi = add2(zi, mi);       // i * X_b[k]
    const f2 pa = fma2(ar, ar, fma2(ai, ai, bc(1.0f))), pb = fma2(br, br, fma2(bi, bi, bc(1.0f)));
    float la0, la1, lb0, lb1;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la0) : "f"(lo(pa)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(la1) : "f"(hi(pa)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb0) : "f"(lo(pb)));
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lb1) : "f"(hi(pb)));
    const f2 sa = mul2(pk(la0, la1), bc(0.69314718055994531f)), sb = mul2(pk(lb0, lb1), bc(0.69314718055994531f));
    if constexpr (kPA & 4) { st_a = st_a && la0 == 12345.678f; st_b = st_b && lb0 == 12345.678f; }
    if (st_a) { row_a[64 * J] = lo(sa); row_a[64 * J + 32] = hi(sa); }
    if (st_b) { row_a[AID_NBINS + 64 * J] = lo(sb); row_a[AID_NBINS + 64 * J + 32] = hi(sb); }
    out_a = sa; out_b = sb;
}

// ---- group maxima of the finished rows (the peak kernel's first reduction, done while the row is in registers) ----
// The peak kernel (peaks.cu) works on the maxima of the 32 aligned 16-bin groups of a row. Group g = 2 k2 + h is the
// maximum over half-warp h (lanes 16 h .. 16 h + 15) of register k2, so a 16 x 16 transpose-reduce inside each half-warp
// leaves lane L with the maximum of group 2 (L & 15) + (L >> 4): level 1 pairs k2 with k2 + 8 across lanes L ^ 8 (the
// lane whose bit 3 is clear keeps k2 and sends k2 + 8), level 2 pairs k2 with k2 + 4 across L ^ 4, ... 15 shuffles per
// row instead of 64 for sixteen separate reductions; maxima are exact, so the summary is bit-identical to what the
// peak kernel computes from the stored row. (redux.sync.max on half-warp masks was measured too: 18 cycles per
// instruction and SM on this part, 2.4x slower for the whole kernel; profiles/r02_stft_packed.md.)
__device__ __forceinline__ float selb(int bit, float x, float y) {        // bit ? x : y as ONE select (the ?: form on halves of a
    float d;                                                               // register pair compiles to two predicated moves)
    asm("{ .reg .pred p; setp.ne.s32 p, %3, 0; selp.f32 %0, %1, %2, p; }" : "=f"(d) : "f"(x), "f"(y), "r"(bit));
    return d;
}
__device__ __forceinline__ float xmax(float a, float b, int bit, int d) {  // a: index k (kept where bit is clear), b: index k + half
    const float r = __shfl_xor_sync(AID_FULL_MASK, selb(bit, a, b), d);
    return fmaxf(selb(bit, b, a), r);
}
struct GroupMax {            // partial maxima of the two rows of a pair, indexed by k2 (mod 16, 8, 4, 2 from level to level)
    float va[16], vb[16];    // level 0: rows a, b as they leave separate_group J (k2 = 2J, 2J + 1)
    float a1[8], b1[8];      // after level 1 (lanes ^ 8)
    float a2[4], b2[4];      // after level 2 (lanes ^ 4)
};
// separation groups are visited in the order 0, 4, 1, 5, 2, 6, 3, 7 so that each level can run as soon as its operands exist
__host__ __device__ constexpr int sep_order(int i) { return (i >> 1) + 4 * (i & 1); }
template <int I>
__device__ __forceinline__ void group_max_step(GroupMax& g, int lane, float* gmax_row, bool st_a, bool st_b) {
    if constexpr (I & 1) {                                   // groups J = I >> 1 and J + 4 are in: k2 = 2J, 2J + 1 and + 8
        constexpr int k = 2 * (I >> 1);
        const int bit = lane & 8;
        g.a1[k] = xmax(g.va[k], g.va[k + 8], bit, 8);         g.b1[k] = xmax(g.vb[k], g.vb[k + 8], bit, 8);
        g.a1[k + 1] = xmax(g.va[k + 1], g.va[k + 9], bit, 8); g.b1[k + 1] = xmax(g.vb[k + 1], g.vb[k + 9], bit, 8);
    }
    if constexpr (I == 5 || I == 7) {                        // level-1 results k and k + 4
        constexpr int k = I - 5;
        const int bit = lane & 4;
        g.a2[k] = xmax(g.a1[k], g.a1[k + 4], bit, 4);         g.b2[k] = xmax(g.b1[k], g.b1[k + 4], bit, 4);
        g.a2[k + 1] = xmax(g.a1[k + 1], g.a1[k + 5], bit, 4); g.b2[k + 1] = xmax(g.b1[k + 1], g.b1[k + 5], bit, 4);
    }
    if constexpr (I == 7) {
        const int b2 = lane & 2, b1 = lane & 1;
        const float a30 = xmax(g.a2[0], g.a2[2], b2, 2), a31 = xmax(g.a2[1], g.a2[3], b2, 2);
        const float b30 = xmax(g.b2[0], g.b2[2], b2, 2), b31 = xmax(g.b2[1], g.b2[3], b2, 2);
        const float ma = xmax(a30, a31, b1, 1), mb = xmax(b30, b31, b1, 1);
        const int grp = 2 * (lane & 15) + (lane >> 4);
        if (st_a) gmax_row[grp] = ma;
        if (st_b) gmax_row[32 + grp] = mb;
    }
}
template <bool SUM, int I>
__device__ __forceinline__ void separate_step(const f2 (&re)[16], const f2 (&im)[16], int lane, int partner,
                                              float* row_a, float* gmax_row, bool st_a, bool st_b, GroupMax& g) {
    constexpr int J = sep_order(I);
    f2 sa, sb;
    separate_group<J>(re, im, lane, partner, row_a, st_a, st_b, sa, sb);
    if constexpr (SUM) {
        g.va[2 * J] = lo(sa); g.va[2 * J + 1] = hi(sa); g.vb[2 * J] = lo(sb); g.vb[2 * J + 1] = hi(sb);
        group_max_step<I>(g, lane, gmax_row, st_a, st_b);
    }
}

// One group (samples 2I, 2I + 1 and + 16 of both frames) of the window multiply fused with DIT stage 0.
template <int I>
__device__ __forceinline__ void window_group(f2 (&re)[16], f2 (&im)[16], const f2 (&ring)[18], const ulonglong2* win2) {
    constexpr int e = bitrev5(2 * I);                     // pair I = samples (2I, 2I + 1) -> elements (e, e + 16)
    const ulonglong2 wa = win2[(kPA & 1) ? 0 : I >> 1], wb = win2[(kPA & 1) ? 0 : 4 + (I >> 1)];
    const f2 wa2 = (I & 1) ? wa.y : wa.x, wb2 = (I & 1) ? wb.y : wb.x;
    const f2 tr = mul2(wb2, ring[I + 8]), ti = mul2(wb2, ring[I + 10]);
    re[e] = fma2(wa2, ring[I], tr);     re[e + 1] = fma2(wa2, ring[I], neg2(tr));
    im[e] = fma2(wa2, ring[I + 2], ti); im[e + 1] = fma2(wa2, ring[I + 2], neg2(ti));
}

template <bool PIPE, bool SUM, int I>
__device__ __forceinline__ void window_and_separate(f2 (&re)[16], f2 (&im)[16], const f2 (&ring)[18], const ulonglong2* win2,
                                                    const f2 (&zre)[16], const f2 (&zim)[16], int lane, int partner,
                                                    float* row_prev, float* gmax_prev, bool st_a, bool st_b, GroupMax& g) {
    if constexpr (I < 8) {
        window_group<I>(re, im, ring, win2);
        if constexpr (PIPE) separate_step<SUM, I>(zre, zim, lane, partner, row_prev, gmax_prev, st_a, st_b, g);
        window_and_separate<PIPE, SUM, I + 1>(re, im, ring, win2, zre, zim, lane, partner, row_prev, gmax_prev, st_a, st_b, g);
    }
}
template <bool SUM, int I>
__device__ __forceinline__ void separate_all(const f2 (&re)[16], const f2 (&im)[16], int lane, int partner, float* row_a,
                                             float* gmax_row, bool st_a, bool st_b, GroupMax& g) {
    if constexpr (I < 8) {
        separate_step<SUM, I>(re, im, lane, partner, row_a, gmax_row, st_a, st_b, g);
        separate_all<SUM, I + 1>(re, im, lane, partner, row_a, gmax_row, st_a, st_b, g);
    }
}

// PF: lanes 0..8 prefetch into L1 the lines the NEXT trip's sample loads will read. The loads keep their place but hit on
// chip: ptxas has six scoreboards and the transpose's STS read-barrier shares one with the LDGs, so without the prefetch the
// first register reuse after the STS waited for the loads of the same trip to come back from DRAM (one instruction held
// 15 % of all stall samples, profiles/r02_stft_packed.md).
// PIPE: software pipelining across trips. The separation of pair p (32 SHFL, 32 FSEL, 32 MUFU, 32 STG, 80 packed FP32:
// LSU / XU work with little arithmetic) is issued group by group BETWEEN the groups of the window stage of pair p + 1
// (48 packed FP32, 8 LDS.128), so that every warp offers the FMA pipe work all the time instead of in phases; each group
// of the separation frees the registers the next window group fills.
// SUM: the kernel also writes gmax[row][32], the maxima of the 32 aligned 16-bin groups of every row (GroupMax above), so
// that the peak kernel never has to read the spectrogram back except around its few surviving candidates.
template <bool PF, bool PIPE, bool SUM>
__global__ void __launch_bounds__(kWarpsPerCta * 32, AID_STFT_MIN_CTAS)
k_stft_packed(const float* __restrict__ twist, const float* __restrict__ pcm, const aid_stft_unit* __restrict__ units, int n_units,
              float* __restrict__ spec, float* __restrict__ gmax) {
    // dynamic shared memory:
    //   tile  [warps][32 * 34] float2  the warp's transpose tile
    //   win   [32][36] float           [lane][j] = window[lane + 32 j] / 2
    //   twist [32][36] float           [lane][4 e + i]: quad e = 0: stage 1 (c, -s, s, c); e = 1: stage 2; e = 2..3: stage 3;
    //                                  e = 4..7: stage 4, each (c_k, c_k+1, s_k, s_k+1); [lane][32..33] = stage 0 (c, s)
    //   win and twist are one image, built on the host (aid_fill_stft_tables) and copied with 128-bit loads
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 (*s_tile)[kTileFloats] = reinterpret_cast<float2 (*)[kTileFloats]>(smem_raw);
    float* s_win = reinterpret_cast<float*>(s_tile + kWarpsPerCta);
    float* s_twist = s_win + 32 * kTabStride;

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {
        const float4* src = reinterpret_cast<const float4*>(twist + AID_TWIST_IMAGE);
        float4* dst = reinterpret_cast<float4*>(s_win);
        for (int i = threadIdx.x; i < AID_TWIST_IMAGE_FLOATS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    float2* tile = s_tile[warp];
    const float4* tile_row = reinterpret_cast<const float4*>(tile + lane * kTileStride);
    const ulonglong2* win2 = reinterpret_cast<const ulonglong2*>(s_win + lane * kTabStride);
    const ulonglong2* tw2 = reinterpret_cast<const ulonglong2*>(s_twist + lane * kPTwistStride);
    const int partner = (32 - lane) & 31;

    // the grid may be smaller than the unit list (aid_launch_stft_variant): a warp then walks the list with the grid's stride
    for (int unit_id = blockIdx.x * kWarpsPerCta + warp; unit_id < n_units; unit_id += gridDim.x * kWarpsPerCta) {
        const aid_stft_unit u = units[unit_id];

        // lane's samples: xp[32 m], m = 0.. ; rem = samples left from xp (32-bit: a track has < 2^31 samples)
        const int64_t first = (int64_t)u.frame0 * AID_HOP + lane;
        const float* xp = pcm + u.pcm_begin + first;
        int rem = (int)(u.n_samples - first);

        f2 ring[18];                                              // ring[i] = samples (x[32 (2i)], x[32 (2i + 1)]) of this lane
#pragma unroll
        for (int j = 0; j < 18; j++)
            ring[j] = pk(64 * j < rem ? __ldg(xp + 64 * j) : 0.0f, 64 * j + 32 < rem ? __ldg(xp + 64 * j + 32) : 0.0f);
        float* row_a = spec + u.spec_row * AID_NBINS + lane;
        float* gmax_row = SUM ? gmax + u.spec_row * 32 : nullptr;
        GroupMax gm;

        f2 re[16], im[16];                                        // PIPE: Z of the previous pair on entry to a trip
        if constexpr (PIPE) {
#pragma unroll
            for (int j = 0; j < 16; j++) { re[j] = 0; im[j] = 0; }
        }

        // nxt = the eight samples pair p + 1 adds, loaded one trip ahead. The loads are issued AFTER the ring has taken the
        // previous eight (end of the trip): issued before, they share a scoreboard with the loads the ring moves wait for,
        // and the moves then wait for the new loads too (7 % of all stall samples)
        f2 nxt[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            nxt[j] = pk(64 * (18 + j) < rem ? __ldg(xp + 64 * (18 + j)) : 0.0f,
                        64 * (18 + j) + 32 < rem ? __ldg(xp + 64 * (18 + j) + 32) : 0.0f);

        for (int p = 0; p < u.n_frames; p += 2) {
            xp += 2 * AID_HOP;
            rem -= 2 * AID_HOP;
            if constexpr (PF) {      // the loads at the end of this trip read xp[32 (36 + j)], j < 8, of all lanes: 1 KB = 8 or 9 lines
                if (lane < 9 && 32 * (36 + lane) - lane < rem)
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(xp - lane + 32 * (36 + lane)));
            }

            // window multiply fused with DIT stage 0 (samples j and j + 16, two j per instruction), frame a in re, b = a + 1 in im;
            // PIPE: interleaved with the separation of the previous pair (both of its frames exist: only a unit's last pair can be odd)
            f2 nre[16], nim[16];
            window_and_separate<PIPE, SUM, 0>(nre, nim, ring, win2, re, im, lane, partner, row_a - 2 * AID_NBINS, gmax_row - 64,
                                              p > 0, p > 0, gm);
            pstage<1, 0>(nre, nim); pstage<2, 0>(nre, nim); pstage<3, 0>(nre, nim);
            last_stage_store<0>(nre, nim, tile + lane);              // Y[k1] -> tile[k1][n1 = lane]
            __syncwarp();

            // second transform: stage 0 (scalar) straight from the gathered row, n1 = 2q + r pairs with n1 + 16
            {
                const float c0 = s_twist[lane * kPTwistStride + 32], s0 = s_twist[lane * kPTwistStride + 33];
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    float4 za, zb;
                    if constexpr (kPA & 2) {
                        za = make_float4(lo(nre[q]), lo(nim[q]), hi(nre[q]), hi(nim[q]));
                        zb = make_float4(lo(nre[q + 8]), lo(nim[q + 8]), hi(nre[q + 8]), hi(nim[q + 8]));
                    } else { za = tile_row[q]; zb = tile_row[q + 8]; }
#pragma unroll
                    for (int r = 0; r < 2; r++) {
                        const float ar = r ? za.z : za.x, ai = r ? za.w : za.y, br = r ? zb.z : zb.x, bi = r ? zb.w : zb.y;
                        const float xr = fmaf(bi, s0, fmaf(br, c0, ar));
                        const float xi = fmaf(-br, s0, fmaf(bi, c0, ai));
                        const int e = bitrev5(2 * q + r);             // even: elements (e, e + 1) = pair e / 2
                        re[e >> 1] = pk(xr, fmaf(2.0f, ar, -xr));
                        im[e >> 1] = pk(xi, fmaf(2.0f, ai, -xi));
                    }
                }
            }
            __syncwarp();
            {
                const ulonglong2 t1 = tw2[(kPA & 1) ? 0 : 0];
                tpgroups<1, 0, 0>(re, im, t1.x, t1.y);
                const ulonglong2 t2 = tw2[(kPA & 1) ? 0 : 1];
                tpgroups<2, 0, 0>(re, im, t2.x, t2.y);
                const ulonglong2 t3a = tw2[2], t3b = tw2[(kPA & 1) ? 0 : 3];
                tpgroups<3, 0, 0>(re, im, t3a.x, t3a.y);
                tpgroups<3, 1, 0>(re, im, t3b.x, t3b.y);
                const ulonglong2 t4a = tw2[4], t4b = tw2[(kPA & 1) ? 0 : 5];
                tpgroups<4, 0, 0>(re, im, t4a.x, t4a.y);
                tpgroups<4, 1, 0>(re, im, t4b.x, t4b.y);
                const ulonglong2 t4c = tw2[6], t4d = tw2[(kPA & 1) ? 0 : 7];
                tpgroups<4, 2, 0>(re, im, t4c.x, t4c.y);
                tpgroups<4, 3, 0>(re, im, t4d.x, t4d.y);
            }

            if constexpr (!PIPE) separate_all<SUM, 0>(re, im, lane, partner, row_a, gmax_row, true, p + 1 < u.n_frames, gm);
            row_a += 2 * AID_NBINS;
            if constexpr (SUM) gmax_row += 64;

#pragma unroll
            for (int j = 0; j < 14; j++) ring[j] = ring[j + 4];
#pragma unroll
            for (int j = 0; j < 4; j++) ring[14 + j] = nxt[j];
#pragma unroll
            for (int j = 0; j < 4; j++)
                nxt[j] = pk(64 * (18 + j) < rem ? __ldg(xp + 64 * (18 + j)) : 0.0f,
                            64 * (18 + j) + 32 < rem ? __ldg(xp + 64 * (18 + j) + 32) : 0.0f);
        }
        if constexpr (PIPE)
            separate_all<SUM, 0>(re, im, lane, partner, row_a - 2 * AID_NBINS, gmax_row - 64, u.n_frames > 0,
                                 (u.n_frames & 1) == 0 && u.n_frames > 0, gm);
    }
}

#ifdef AID_STFT_TC
#include "../../tools/microbench/stft_tc.cuh"      // experimental tensor-core second transform (micro-benchmarks only)
#endif

}  // namespace

// =====================================================================================================
// Warp-specialised form of the packed kernel (variant bit 4): a PRODUCER warp runs the first transform of a unit (sample
// ring, window stage, stages 1-4, STS of the transposed tile) and a CONSUMER warp the second (gather + stage 0, stages
// 1-4, separation, log, stores, group maxima); the two hand tiles over through two shared-memory buffers guarded by
// mbarriers (full / empty, 32 arrivals each: every lane releases its own stores / loads). Neither role holds the other's
// state, so the kernel fits 128 registers and an SM holds 2 CTAs x 8 warps = 16 warps instead of 12: the question it
// answers is whether the packed kernel's missing overlap of its shared-memory phases with arithmetic (profiles/
// r02_stft_packed.md section 4) is a matter of resident warps. Same operations on every element: bit-identical output.
constexpr int kWsPairs = 4;                                   // (producer, consumer) pairs = units per CTA
constexpr int kWsSmem = kWsPairs * 2 * kTileFloats * 8 + 32 * kTabStride * 4 + 32 * kPTwistStride * 4 + kWsPairs * 4 * 8;

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{ .reg .pred p;\n"
                 "W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                 "@!p bra W;\n}" :: "r"(bar), "r"(parity) : "memory");
}

template <bool SUM>
__global__ void __launch_bounds__(kWsPairs * 64, 2)
k_stft_ws(const float* __restrict__ twist, const float* __restrict__ pcm, const aid_stft_unit* __restrict__ units, int n_units,
          float* __restrict__ spec, float* __restrict__ gmax) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float2 (*s_tile)[kTileFloats] = reinterpret_cast<float2 (*)[kTileFloats]>(smem_raw);      // [pair * 2 + buffer]
    float* s_win = reinterpret_cast<float*>(s_tile + kWsPairs * 2);
    float* s_twist = s_win + 32 * kTabStride;
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_twist + 32 * kPTwistStride);   // [pair][full0, full1, empty0, empty1]

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pair = warp >> 1;
    const bool producer = (warp & 1) == 0;
    {
        const float4* src = reinterpret_cast<const float4*>(twist + AID_TWIST_IMAGE);
        float4* dst = reinterpret_cast<float4*>(s_win);
        for (int i = threadIdx.x; i < AID_TWIST_IMAGE_FLOATS / 4; i += blockDim.x) dst[i] = __ldg(src + i);
        if (threadIdx.x < kWsPairs * 4) mbar_init((uint32_t)__cvta_generic_to_shared(s_bar + threadIdx.x), 32);
    }
    __syncthreads();

    const int unit_id = blockIdx.x * kWsPairs + pair;
    if (unit_id >= n_units) return;
    const aid_stft_unit u = units[unit_id];
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(s_bar + pair * 4);               // full[b] = bar0 + 8 b, empty[b] = bar0 + 16 + 8 b
    float2* tile0 = s_tile[pair * 2];

    if (producer) {
        const ulonglong2* win2 = reinterpret_cast<const ulonglong2*>(s_win + lane * kTabStride);
        const int64_t first = (int64_t)u.frame0 * AID_HOP + lane;
        const float* xp = pcm + u.pcm_begin + first;
        int rem = (int)(u.n_samples - first);
        f2 ring[18];
#pragma unroll
        for (int j = 0; j < 18; j++)
            ring[j] = pk(64 * j < rem ? __ldg(xp + 64 * j) : 0.0f, 64 * j + 32 < rem ? __ldg(xp + 64 * j + 32) : 0.0f);
        int trip = 0;
        for (int p = 0; p < u.n_frames; p += 2, trip++) {
            xp += 2 * AID_HOP;
            rem -= 2 * AID_HOP;
            if (lane < 9 && 32 * (36 + lane) - lane + 2 * AID_HOP < rem + 2 * AID_HOP)      // lines the loads of the NEXT trip's end will read
                asm volatile("prefetch.global.L1 [%0];" :: "l"(xp - lane + 2 * AID_HOP + 32 * (36 + lane) - 2 * AID_HOP));
            f2 nre[16], nim[16], zre[16], zim[16];
            GroupMax gm;
            window_and_separate<false, false, 0>(nre, nim, ring, win2, zre, zim, lane, 0, nullptr, nullptr, false, false, gm);
            pstage<1, 0>(nre, nim); pstage<2, 0>(nre, nim); pstage<3, 0>(nre, nim);
            const int b = trip & 1, use = trip >> 1;
            if (use > 0) mbar_wait(bar0 + 16 + 8 * b, (uint32_t)((use - 1) & 1));             // the consumer has read this buffer's previous tile
            last_stage_store<0>(nre, nim, tile0 + b * kTileFloats + lane);
            mbar_arrive(bar0 + 8 * b);
#pragma unroll
            for (int j = 0; j < 14; j++) ring[j] = ring[j + 4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                ring[14 + j] = pk(64 * (14 + j) < rem ? __ldg(xp + 64 * (14 + j)) : 0.0f,
                                  64 * (14 + j) + 32 < rem ? __ldg(xp + 64 * (14 + j) + 32) : 0.0f);
        }
    } else {
        const ulonglong2* tw2 = reinterpret_cast<const ulonglong2*>(s_twist + lane * kPTwistStride);
        const int partner = (32 - lane) & 31;
        float* row_a = spec + u.spec_row * AID_NBINS + lane;
        float* gmax_row = SUM ? gmax + u.spec_row * 32 : nullptr;
        const float c0 = s_twist[lane * kPTwistStride + 32], s0 = s_twist[lane * kPTwistStride + 33];
        int trip = 0;
        for (int p = 0; p < u.n_frames; p += 2, trip++) {
            const int b = trip & 1, use = trip >> 1;
            const float4* tile_row = reinterpret_cast<const float4*>(tile0 + b * kTileFloats + lane * kTileStride);
            f2 re[16], im[16];
            mbar_wait(bar0 + 8 * b, (uint32_t)(use & 1));
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const float4 za = tile_row[q], zb = tile_row[q + 8];
#pragma unroll
                for (int r = 0; r < 2; r++) {
                    const float ar = r ? za.z : za.x, ai = r ? za.w : za.y, br = r ? zb.z : zb.x, bi = r ? zb.w : zb.y;
                    const float xr = fmaf(bi, s0, fmaf(br, c0, ar));
                    const float xi = fmaf(-br, s0, fmaf(bi, c0, ai));
                    const int e = bitrev5(2 * q + r);
                    re[e >> 1] = pk(xr, fmaf(2.0f, ar, -xr));
                    im[e >> 1] = pk(xi, fmaf(2.0f, ai, -xi));
                }
            }
            mbar_arrive(bar0 + 16 + 8 * b);
            {
                const ulonglong2 t1 = tw2[0];
                tpgroups<1, 0, 0>(re, im, t1.x, t1.y);
                const ulonglong2 t2 = tw2[1];
                tpgroups<2, 0, 0>(re, im, t2.x, t2.y);
                const ulonglong2 t3a = tw2[2], t3b = tw2[3];
                tpgroups<3, 0, 0>(re, im, t3a.x, t3a.y);
                tpgroups<3, 1, 0>(re, im, t3b.x, t3b.y);
                const ulonglong2 t4a = tw2[4], t4b = tw2[5];
                tpgroups<4, 0, 0>(re, im, t4a.x, t4a.y);
                tpgroups<4, 1, 0>(re, im, t4b.x, t4b.y);
                const ulonglong2 t4c = tw2[6], t4d = tw2[7];
                tpgroups<4, 2, 0>(re, im, t4c.x, t4c.y);
                tpgroups<4, 3, 0>(re, im, t4d.x, t4d.y);
            }
            GroupMax gm;
            separate_all<SUM, 0>(re, im, lane, partner, row_a, gmax_row, true, p + 1 < u.n_frames, gm);
            row_a += 2 * AID_NBINS;
            if constexpr (SUM) gmax_row += 64;
        }
    }
}

template <bool PF, bool PIPE, bool SUM>
static cudaError_t launch_packed(int grid, const aid_tables& tb, const float* d_pcm, const aid_stft_unit* d_units, int n_units,
                                 float* d_spec, float* d_gmax, cudaStream_t st) {
    static bool configured = false;         // 43.5 KB of dynamic shared memory: below the 48 KB default, set for robustness
    if (!configured) {
        cudaError_t ce = cudaFuncSetAttribute(k_stft_packed<PF, PIPE, SUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kPackedSmem);
        if (ce != cudaSuccess) return ce;
        configured = true;
    }
    k_stft_packed<PF, PIPE, SUM><<<grid, kWarpsPerCta * 32, kPackedSmem, st>>>(tb.twist, d_pcm, d_units, n_units, d_spec, d_gmax);
    return cudaGetLastError();
}

// variant 0: k_stft (scalar FP32, round 1); otherwise k_stft_packed (f32x2) with bit 1: separation software-pipelined into
// the next trip's window stage, bit 2: L1 prefetch of the next trip's samples, bit 3: persistent grid (1, 3, 5, 7, + 8).
// The spectrogram is bit-identical for every variant (tools/microbench/stft_bench.cu, tests/test_gpu_fingerprint.py).
// d_gmax (optional, packed variants only): [rows][32] maxima of the 16-bin groups of every row, for the peak kernel.
cudaError_t aid_launch_stft_variant(int variant, const aid_tables& tb, const float* d_pcm, const aid_stft_unit* d_units,
                                    int n_units, float* d_spec, float* d_gmax, cudaStream_t st) {
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
    int grid = (n_units + kWarpsPerCta - 1) / kWarpsPerCta;
    if (variant == 0) {
        if (d_gmax) return cudaErrorInvalidValue;                 // the scalar kernel has no summary output
        k_stft<<<grid, kWarpsPerCta * 32, 0, st>>>(tb.window, tb.twist, d_pcm, d_units, n_units, d_spec);
        return cudaGetLastError();
    }
    if (variant & 8) {                      // persistent grid: SMs x AID_STFT_MIN_CTAS CTAs walk the unit list (measured slower: the warps of
        static int resident = 0;            // an SM then run in lock step through the trip's phases; profiles/r02_stft_packed.md)
        if (!resident) {
            int dev = 0, sms = 0;
            cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            resident = sms * AID_STFT_MIN_CTAS;
        }
        if (grid > resident) grid = resident;
    }
    if (variant & 16) {                     // warp-specialised form (A/B): 4 units per CTA of 8 warps
        static bool configured = false;
        if (!configured) {
            cudaError_t ce = cudaFuncSetAttribute(k_stft_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem);
            if (ce == cudaSuccess) ce = cudaFuncSetAttribute(k_stft_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWsSmem);
            if (ce != cudaSuccess) return ce;
            configured = true;
        }
        const int g = (n_units + kWsPairs - 1) / kWsPairs;
        if (d_gmax) k_stft_ws<true><<<g, kWsPairs * 64, kWsSmem, st>>>(tb.twist, d_pcm, d_units, n_units, d_spec, d_gmax);
        else k_stft_ws<false><<<g, kWsPairs * 64, kWsSmem, st>>>(tb.twist, d_pcm, d_units, n_units, d_spec, nullptr);
        return cudaGetLastError();
    }
    const bool pipe = variant & 2, pf = variant & 4;
#define AID_P(PF, PIPE) (d_gmax ? launch_packed<PF, PIPE, true>(grid, tb, d_pcm, d_units, n_units, d_spec, d_gmax, st) \
                                : launch_packed<PF, PIPE, false>(grid, tb, d_pcm, d_units, n_units, d_spec, nullptr, st))
    return pf ? (pipe ? AID_P(true, true) : AID_P(true, false)) : (pipe ? AID_P(false, true) : AID_P(false, false));
#undef AID_P
#endif
}

int aid_stft_default_variant() {
    static const int variant = [] { const char* v = getenv("AID_STFT_VARIANT"); return v ? atoi(v) : 5; }();
    return variant;
}
