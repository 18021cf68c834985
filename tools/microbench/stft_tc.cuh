// stft_tc.cuh -- EXPERIMENT, included by audio_ident_b200/csrc/stft.cu only when a micro-benchmark defines AID_STFT_TC
// (tools/microbench/build.sh builds bin/stft_tc); the product library never sees it. It is textually part of stft.cu's
// anonymous namespace so that it can reuse the first transform, the tables and the epilogue helpers.
}  // namespace
#include <cuda_fp16.h>
namespace {
// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENT (tools/microbench only; the product build never defines AID_STFT_TC): the second 32-point transform on
// tcgen05. DESIGN.md section 7 item 1. Same units, same first transform and same epilogue as k_stft; between them lane
// n1 multiplies Y[n1][k1] by W_1024^(n1 k1), splits re / im into FP16 hi + lo and stores them as row (k1, warp) of the
// data operand A[128 x 64] (K index = 2 n1 + {re, im}); thread 0 issues 3 x 4 MMAs (hi*hi, hi*lo, lo*hi) against the
// constant operand B[64 x 64] = W_32 as a real matrix (column 2 k2 + {re, im}); warp w reads TMEM lanes 32 w .. 32 w + 31
// back and finds Z[k1 + 32 k2] in lane k1, element k2 -- the layout the epilogue of k_stft starts from.
// Descriptors, layout and TMEM read-back are those verified by tools/microbench/umma_probe.cu.
// Status: compiles for sm_100a; not yet run (the round's GPU budget was spent). No overlap of MMA and CUDA-core work yet.
constexpr uint32_t kTcSbo = 128;
constexpr uint32_t kTcLboA = (128 / 8) * 128 + 16;       // + 16 B: the 8 K-chunks of a row land in different banks
constexpr uint32_t kTcLboB = (64 / 8) * 128;
constexpr uint32_t kTcABytes = 8 * kTcLboA, kTcBBytes = 8 * kTcLboB;
constexpr uint32_t kTcCols = 64;
constexpr uint32_t kTcInstr = (1u << 4) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // F32 acc, F16 x F16, K-major, N = 64, M = 128

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t tc_desc(uint32_t addr, uint32_t lbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (uint64_t)((lbo >> 4) & 0x3fffu) << 16 | (uint64_t)((kTcSbo >> 4) & 0x3fffu) << 32 |
           (uint64_t)1 << 46;
}
// x = hi + lo with hi the 11 leading bits of x (exact in FP16 for normal values), both packed for (re, im)
__device__ __forceinline__ void tc_split(float xr, float xi, uint32_t& hi, uint32_t& lo) {
    const float hr = __uint_as_float(__float_as_uint(xr) & 0xffffe000u), hi_ = __uint_as_float(__float_as_uint(xi) & 0xffffe000u);
    const __half2 h = __floats2half2_rn(hr, hi_), l = __floats2half2_rn(xr - hr, xi - hi_);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

struct StftTcSmem {
    alignas(1024) unsigned char a_hi[kTcABytes];
    alignas(128) unsigned char a_lo[kTcABytes];
    alignas(128) unsigned char b_hi[kTcBBytes];
    alignas(128) unsigned char b_lo[kTcBBytes];
    alignas(16) float win[32 * kTabStride];
    alignas(16) float tw[32 * 68];
    alignas(8) uint64_t mbar;
    uint32_t tmem_slot;
    int max_pairs;
};

__global__ void __launch_bounds__(128, 3)
k_stft_tc(const float* __restrict__ window, const float* __restrict__ twist,
          const float* __restrict__ pcm, const aid_stft_unit* __restrict__ units, int n_units,
          float* __restrict__ spec) {
    extern __shared__ __align__(1024) unsigned char tc_raw[];
    StftTcSmem& sm = *reinterpret_cast<StftTcSmem*>(tc_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (int i = tid; i < 32 * 32; i += 128) sm.win[(i & 31) * kTabStride + (i >> 5)] = 0.5f * window[i];
    for (int i = tid; i < 32 * 64; i += 128) sm.tw[(i >> 6) * 68 + (i & 63)] = twist[AID_TWIST_FOLDED + i];
    // B[col = 2 k2 + c_out][k = 2 n1 + c_in]:  Zr = sum wr yr - wi yi,  Zi = sum wi yr + wr yi,  W = wr + i wi = e^(-2 pi i k2 n1 / 32)
    for (int i = tid; i < 64 * 64; i += 128) {
        const int col = i >> 6, k = i & 63, k2 = col >> 1, co = col & 1, n1 = k >> 1, ci = k & 1;
        float sn, cs;
        sincospif((float)((k2 * n1) & 31) * (1.0f / 16.0f), &sn, &cs);
        const float wr = cs, wi = -sn;
        const float v = co == 0 ? (ci == 0 ? wr : -wi) : (ci == 0 ? wi : wr);
        const __half h = __float2half_rn(v), l = __float2half_rn(v - __half2float(h));
        const uint32_t off = (k / 8) * kTcLboB + (col / 8) * kTcSbo + (col % 8) * 16 + (k % 8) * 2;
        *reinterpret_cast<__half*>(sm.b_hi + off) = h;
        *reinterpret_cast<__half*>(sm.b_lo + off) = l;
    }
    if (tid == 0) sm.max_pairs = 0;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(tc_smem(&sm.tmem_slot)), "n"(kTcCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(tc_smem(&sm.mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = sm.tmem_slot;

    const int unit_id = blockIdx.x * 4 + warp;
    const bool live = unit_id < n_units;
    aid_stft_unit u{};
    if (live) u = units[unit_id];
    const int my_pairs = live ? (u.n_frames + 1) / 2 : 0;
    if (lane == 0) atomicMax(&sm.max_pairs, my_pairs);
    __syncthreads();
    const int max_pairs = sm.max_pairs;

    const int64_t first = (int64_t)u.frame0 * AID_HOP + lane;
    const float* xp = pcm + u.pcm_begin + first;
    int rem = live ? (int)(u.n_samples - first) : 0;
    float ring[36];
#pragma unroll
    for (int j = 0; j < 36; j++) ring[j] = 32 * j < rem ? __ldg(xp + 32 * j) : 0.0f;
    const float4* win4 = reinterpret_cast<const float4*>(sm.win + lane * kTabStride);
    const float4* tw4 = reinterpret_cast<const float4*>(sm.tw + lane * 68);
    const int partner = (32 - lane) & 31;
    float* row_a = spec + u.spec_row * AID_NBINS + lane;
    // row m = 32 warp + k1 of A, K chunk lane / 4, bytes 4 (lane % 4) .. + 3 inside the 16-byte core-matrix row
    const uint32_t a_lane = (lane >> 2) * kTcLboA + (lane & 3) * 4;

    for (int t = 0; t < max_pairs; t++) {
        const bool mine = t < my_pairs;
        float re[32], im[32];
        if (mine) {
            float nxt[8];
#pragma unroll
            for (int j = 0; j < 8; j++) nxt[j] = 32 * (36 + j) < rem ? __ldg(xp + 32 * (36 + j)) : 0.0f;
            xp += 2 * AID_HOP;
            rem -= 2 * AID_HOP;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const float4 wa4 = win4[q], wb4 = win4[4 + q];
                const float wa[4] = {wa4.x, wa4.y, wa4.z, wa4.w}, wb[4] = {wb4.x, wb4.y, wb4.z, wb4.w};
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int ja = 4 * q + r, jb = ja + 16, e = bitrev5(ja);
                    const float tr = wb[r] * ring[jb], ti = wb[r] * ring[jb + 4];
                    re[e] = fmaf(wa[r], ring[ja], tr);     re[e + 1] = fmaf(wa[r], ring[ja], -tr);
                    im[e] = fmaf(wa[r], ring[ja + 4], ti); im[e + 1] = fmaf(wa[r], ring[ja + 4], -ti);
                }
            }
            fft32_after_stage0(re, im);                      // Y[n1 = lane][k1] in element k1
#pragma unroll
            for (int j = 0; j < 28; j++) ring[j] = ring[j + 8];
#pragma unroll
            for (int j = 0; j < 8; j++) ring[28 + j] = nxt[j];
            // twiddle W_1024^(lane k1) = c - i s, split, store row (k1, warp)
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float4 w = tw4[q];
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int k1 = 2 * q + h;
                    const float c = h ? w.z : w.x, sn = h ? w.w : w.y;
                    const float yr = k1 == 0 ? re[0] : fmaf(re[k1], c, im[k1] * sn);
                    const float yi = k1 == 0 ? im[0] : fmaf(im[k1], c, -(re[k1] * sn));
                    uint32_t hi, lo;
                    tc_split(yr, yi, hi, lo);
                    const uint32_t m = 32 * warp + k1, off = a_lane + (m >> 3) * kTcSbo + (m & 7) * 16;
                    *reinterpret_cast<uint32_t*>(sm.a_hi + off) = hi;
                    *reinterpret_cast<uint32_t*>(sm.a_lo + off) = lo;
                }
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // the previous trip's TMEM loads are done
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t ah = tc_smem(sm.a_hi), al = tc_smem(sm.a_lo), bh = tc_smem(sm.b_hi), bl = tc_smem(sm.b_lo);
            const uint32_t pa[3] = {ah, ah, al}, pb[3] = {bh, bl, bh};
#pragma unroll
            for (int p = 0; p < 3; p++)
#pragma unroll
                for (int ks = 0; ks < 4; ks++) {
                    const uint64_t da = tc_desc(pa[p] + ks * 2 * kTcLboA, kTcLboA), db = tc_desc(pb[p] + ks * 2 * kTcLboB, kTcLboB);
                    const uint32_t acc = (p | ks) ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 :: "r"(tmem), "l"(da), "l"(db), "r"(kTcInstr), "r"(acc) : "memory");
                }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(tc_smem(&sm.mbar)) : "memory");
        }
        {
            const uint32_t bar = tc_smem(&sm.mbar), parity = (uint32_t)(t & 1);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (mine) {
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {
                uint32_t v[16];
                const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                             : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                               "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                             : "r"(taddr) : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 8; j++) { re[c0 / 2 + j] = __uint_as_float(v[2 * j]); im[c0 / 2 + j] = __uint_as_float(v[2 * j + 1]); }
            }
            const bool has_b = 2 * t + 1 < u.n_frames;
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                const float zr = re[k2], zi = im[k2];
                const float sr = __shfl_sync(AID_FULL_MASK, re[31 - k2], partner);
                const float si = __shfl_sync(AID_FULL_MASK, im[31 - k2], partner);
                const float mr = lane == 0 ? re[(32 - k2) & 31] : sr;
                const float mi = lane == 0 ? im[(32 - k2) & 31] : si;
                const float ar = zr + mr, ai = zi - mi, br = zr - mr, bi = zi + mi;
                row_a[32 * k2] = log1p_power(ar, ai);
                if (has_b) row_a[AID_NBINS + 32 * k2] = log1p_power(br, bi);
            }
            row_a += 2 * AID_NBINS;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(kTcCols) : "memory");
}
