// umma_probe.cu -- smallest possible tcgen05 program: D[128 x 64] = A[128 x 64] * B[64 x 64]^T with FP16 operands in the
// no-swizzle K-major canonical shared-memory layout and an FP32 accumulator in TMEM, checked against a host product.
//
// Why it exists: DESIGN.md section 7 plans the STFT's second transform as FP16-split products on tcgen05. Before any of
// that kernel is written the shared-memory descriptor (start / leading-dimension / stride byte offsets, version bit),
// the instruction descriptor (formats, majors, M, N) and the TMEM read-back (lane = row, column = n) have to be right,
// and each can only be verified on hardware. This file encodes them as understood from the CUTLASS headers vendored in
// the image (cute/arch/mma_sm100_desc.hpp, cute/atom/mma_traits_sm100.hpp) and prints PASS / the first mismatches.
// With `x3` it runs the split product the STFT needs (hi*hi + hi*lo + lo*hi into one accumulator) on FP32 inputs and
// reports the error against a double-precision product.
//
// With `hankel` (not yet run) it asks whether a descriptor may describe OVERLAPPING rows: A[m][k] = s[8 m + k] read straight
// from a contiguous FP16 array (leading-dimension offset 16 B = the next 8 halves, core-matrix rows 16 B apart), which is
// what a tensor-core STFT fed with raw samples instead of per-frame operand tiles would need.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o bin/umma_probe umma_probe.cu
// Run:   ./bin/umma_probe [x3 | hankel]
// Status: run on a B200 at the end of round 1 (gpurun): "PASS: 0 of 8192 entries off, worst absolute error 0 (exact
// small integers)" and "PASS: 0 of 8192 entries off, worst absolute error 0.00015 (FP16 hi+lo split, 3 products ...)" --
// i.e. the descriptors, the layout and the TMEM read-back below are right, and three MMAs into one accumulator give
// FP32-grade products (1.5e-4 absolute on sums of magnitude ~1e3).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int M = 128, N = 64, K = 64, UMMA_K = 16;
// canonical K-major, no swizzle: a core matrix is 8 rows x 16 bytes (8 halves of K), rows 16 B apart (128 B per core
// matrix); core matrices along M (or N) are SBO apart, the two 16-byte K chunks of one MMA are LBO apart.
constexpr uint32_t SBO = 128;
constexpr uint32_t LBO_A = (M / 8) * 128, LBO_B = (N / 8) * 128;
constexpr uint32_t A_BYTES = (K / 8) * LBO_A, B_BYTES = (K / 8) * LBO_B;      // 16 KB, 8 KB
constexpr uint32_t TMEM_COLS = 64;

__host__ __device__ inline uint32_t a_off(int m, int k) { return (k / 8) * LBO_A + (m / 8) * SBO + (m % 8) * 16 + (k % 8) * 2; }
__host__ __device__ inline uint32_t b_off(int n, int k) { return (k / 8) * LBO_B + (n / 8) * SBO + (n % 8) * 16 + (k % 8) * 2; }

__device__ inline uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// cute::UMMA::SmemDescriptor: start address [0,14) >> 4, leading byte offset [16,30) >> 4, stride byte offset [32,46) >> 4,
// version [46,48) = 1 on Blackwell, base offset 0, lbo mode 0, layout type [61,64) = 0 (no swizzle)
__device__ inline uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fffu) | (uint64_t)((lbo >> 4) & 0x3fffu) << 16 | (uint64_t)((sbo >> 4) & 0x3fffu) << 32 |
           (uint64_t)1 << 46;
}

// cute::UMMA::InstrDescriptor for kind::f16: D format [4,6) = 1 (F32), A format [7,10) = 0 (F16), B format [10,13) = 0,
// A major [15] = 0 (K), B major [16] = 0 (K), N >> 3 at [17,23), M >> 4 at [24,29)
constexpr uint32_t kInstrDesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);

__global__ void __launch_bounds__(128, 1)
k_probe(const __half* __restrict__ A, const __half* __restrict__ B, int n_terms, int hankel, float* __restrict__ D) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                           // n_terms x A_BYTES
    unsigned char* sB = smem + 3 * A_BYTES;             // n_terms x B_BYTES
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + 3 * A_BYTES + 3 * B_BYTES);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int t = 0; t < n_terms; t++) {
        if (hankel) {                                   // the 1-D sequence itself; rows will overlap
            for (int i = tid; i < 8 * (M - 1) + K; i += blockDim.x) reinterpret_cast<__half*>(sA)[i] = A[i];
        } else
        for (int i = tid; i < M * K; i += blockDim.x)
            *reinterpret_cast<__half*>(sA + t * A_BYTES + a_off(i / K, i % K)) = A[(size_t)t * M * K + i];
        for (int i = tid; i < N * K; i += blockDim.x)
            *reinterpret_cast<__half*>(sB + t * B_BYTES + b_off(i / K, i % K)) = B[(size_t)t * N * K + i];
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "n"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(mbar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        // products accumulate into one TMEM tile: term pairs (a, b) = (0, 0) and, for the split, (0, 1) and (1, 0)
        const int pa[3] = {0, 0, 1}, pb[3] = {0, 1, 0};
        const int n_prod = n_terms == 1 ? 1 : 3;
        bool first = true;
        for (int p = 0; p < n_prod; p++)
            for (int ks = 0; ks < K / UMMA_K; ks++) {
                const uint64_t da = hankel ? smem_desc(smem_u32(sA) + ks * 32, 16, SBO)      // K step of 16 halves = 32 B
                                           : smem_desc(smem_u32(sA + pa[p] * A_BYTES) + ks * 2 * LBO_A, LBO_A, SBO);
                const uint64_t db = smem_desc(smem_u32(sB + pb[p] * B_BYTES) + ks * 2 * LBO_B, LBO_B, SBO);
                const uint32_t acc = first ? 0u : 1u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                             "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             :: "r"(tmem), "l"(da), "l"(db), "r"(kInstrDesc), "r"(acc) : "memory");
                first = false;
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smem_u32(mbar)) : "memory");
    }
    // everybody waits for the MMAs (phase 0 of the barrier)
    {
        const uint32_t bar = smem_u32(mbar);
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // warp w owns TMEM lanes 32 w .. 32 w + 31 = rows of D; 64 columns as 4 x 16
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                       "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 16; j++) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(TMEM_COLS) : "memory");
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)

int main(int argc, char** argv) {
    const bool split = argc > 1 && argv[1][0] == 'x';
    const bool hankel = argc > 1 && argv[1][0] == 'h';
    const int n_terms = split ? 2 : 1;
    std::vector<double> a(M * K), b(N * K);
    srand(7);
    for (auto& v : a) v = split ? (rand() / (double)RAND_MAX * 2.0 - 1.0) : (double)(rand() % 9 - 4);      // exact small integers, or reals in [-1, 1]
    std::vector<double> seq(8 * (M - 1) + K);
    if (hankel) {                                           // a[m][k] = seq[8 m + k]
        for (auto& v : seq) v = (double)(rand() % 9 - 4);
        for (int m = 0; m < M; m++) for (int k = 0; k < K; k++) a[m * K + k] = seq[8 * m + k];
    }
    for (auto& v : b) v = split ? (rand() / (double)RAND_MAX * 60.0 - 30.0) : (double)(rand() % 9 - 4);
    std::vector<__half> ha((size_t)n_terms * M * K), hb((size_t)n_terms * N * K);
    for (int i = 0; i < M * K; i++) {
        const __half hi = __float2half((float)a[i]);
        ha[i] = hi;
        if (split) ha[(size_t)M * K + i] = __float2half((float)a[i] - __half2float(hi));
    }
    for (int i = 0; i < N * K; i++) {
        const __half hi = __float2half((float)b[i]);
        hb[i] = hi;
        if (split) hb[(size_t)N * K + i] = __float2half((float)b[i] - __half2float(hi));
    }
    __half *dA, *dB; float* dD;
    CK(cudaMalloc(&dA, ha.size() * 2)); CK(cudaMalloc(&dB, hb.size() * 2)); CK(cudaMalloc(&dD, M * N * 4));
    CK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0xff, M * N * 4));
    const size_t smem = 3 * A_BYTES + 3 * B_BYTES + 64;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (hankel) {                                           // the device gets the sequence, not the matrix
        for (size_t j = 0; j < seq.size(); j++) ha[j] = __float2half((float)seq[j]);
        CK(cudaMemcpy(dA, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
    }
    k_probe<<<1, 128, smem>>>(dA, dB, n_terms, hankel ? 1 : 0, dD);
    CK(cudaDeviceSynchronize());
    std::vector<float> d(M * N);
    CK(cudaMemcpy(d.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    double worst = 0; int bad = 0;
    for (int m = 0; m < M; m++)
        for (int n = 0; n < N; n++) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += a[m * K + k] * b[n * K + k];
            const double err = fabs((double)d[m * N + n] - ref);
            worst = fmax(worst, err);
            if (err > (split ? 1e-3 : 0.0)) { if (bad < 8) printf("  D[%d][%d] = %g, expected %g\n", m, n, d[m * N + n], ref); bad++; }
        }
    printf("%s: %d of %d entries off, worst absolute error %.3g (%s)\n", bad ? "FAIL" : "PASS", bad, M * N, worst,
           split ? "FP16 hi+lo split, 3 products, |A| <= 1, |B| <= 30, K = 64" : hankel ? "overlapping rows A[m][k] = s[8 m + k]" : "exact small integers");
    return bad ? 2 : 0;
}
