// fp32_issue.cu -- issue-rate microbenchmark for scalar vs packed (f32x2) FP32 instructions on sm_100a.
// Decides whether the STFT butterflies should use FFMA2/FADD2. Build: nvcc -gencode arch=compute_100a,code=sm_100a
// Prints warp-instructions per clock per SM sub-partition (4 per SM) from in-kernel clock64() deltas.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;

#define REP16(x) x x x x x x x x x x x x x x x x

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, float seed, int iters) {
    float a[8], b[8];
    u64 A[8], B[8];
    for (int i = 0; i < 8; i++) {
        a[i] = seed * (threadIdx.x + i); b[i] = seed + i;
        float2 t = make_float2(a[i], b[i]);
        A[i] = *(u64*)&t; B[i] = *(u64*)&t;
    }
    float c = seed * 1.0001f, d = seed * 0.5f;
    float2 cd2 = make_float2(c, d);
    u64 C = *(u64*)&cd2;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {          // FFMA reg,reg,reg : 8 independent chains
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c));
        } else if (MODE == 1) {   // FFMA with immediate multiplier
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %1, 0f3F7FF000, %0;" : "+f"(a[i]) : "f"(b[i]));
        } else if (MODE == 2) {   // FFMA2 reg
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
        } else if (MODE == 3) {   // FADD reg
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
        } else if (MODE == 4) {   // FADD2
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(B[i]));
        } else if (MODE == 5) {   // FMUL
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(c));
        } else if (MODE == 6) {   // FMUL2
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(A[i]) : "l"(C));
        } else if (MODE == 7) {   // FFMA + FADD interleaved 1:1 (two pipes?)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c));
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i + 1]) : "f"(b[i + 1]));
                }
        } else if (MODE == 8) {   // FFMA2 + SHFL 3:1
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i += 4) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i + 1]) : "l"(B[i + 1]), "l"(C));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i + 2]) : "l"(B[i + 2]), "l"(C));
                    b[i] = __shfl_xor_sync(0xffffffffu, b[i], 1);
                }
        } else if (MODE == 9) {   // FFMA(reg) + SHFL 3:1
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i += 4) {
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i + 1]), "f"(c));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i + 1]) : "f"(b[i + 1]), "f"(c));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i + 2]) : "f"(b[i + 2]), "f"(c));
                    b[i] = __shfl_xor_sync(0xffffffffu, b[i], 1);
                }
        } else if (MODE == 10) {  // FFMA2 + FFMA(imm) 1:1
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
                    asm volatile("fma.rn.f32 %0, %1, 0f3F7FF000, %0;" : "+f"(a[i + 1]) : "f"(b[i + 1]));
                }
        } else if (MODE == 11) {  // FFMA2 + FADD2 1:1
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
                    asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(A[i + 1]) : "l"(B[i + 1]));
                }
        } else if (MODE == 13) {  // blocks: 16 FFMA2 then 16 scalar FFMA (the packed STFT's shape: packed stages, one scalar stage)
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
#pragma unroll
            for (int r = 0; r < 2; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i]) : "f"(b[i]), "f"(c));
        } else if (MODE == 14) {  // FFMA2 with a broadcast immediate multiplier
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) asm volatile("{.reg .b64 t; mov.b64 t, {0f3F7FF000, 0f3F7FF000}; fma.rn.f32x2 %0, %1, t, %0;}" : "+l"(A[i]) : "l"(B[i]));
        } else if (MODE == 15) {  // FFMA2 + scalar FFMA 2:2 (scalar instructions in adjacent pairs)
#pragma unroll
            for (int r = 0; r < 4; r++)
#pragma unroll
                for (int i = 0; i < 8; i += 4) {
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i + 1]) : "l"(B[i + 1]), "l"(C));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i + 2]) : "f"(b[i + 2]), "f"(c));
                    asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(a[i + 3]) : "f"(b[i + 3]), "f"(c));
                }
        } else if (MODE == 12) {  // FFMA2 + MUFU.LG2 7:1
#pragma unroll
            for (int r = 0; r < 4; r++) {
#pragma unroll
                for (int i = 0; i < 7; i++) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(A[i]) : "l"(B[i]), "l"(C));
                asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[7]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0; u64 S = 0;
    for (int i = 0; i < 8; i++) { s += a[i] + b[i]; S ^= A[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)(S & 0xff);
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
    const int iters = 2048, grid = 148;
    float* out; long long* cyc;
    cudaMalloc(&out, grid * 1024 * 4); cudaMalloc(&cyc, grid * 32 * 8);
    k<MODE><<<grid, threads>>>(out, cyc, 1.0f, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(out, cyc, 1.0f, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[148 * 32];
    cudaMemcpy(h, cyc, grid * (threads / 32) * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < grid * (threads / 32); i++) mx = h[i] > mx ? h[i] : mx;
    const double per_smsp = (double)(threads / 32) / 4.0 * iters * 32.0;   // warp-instructions per sub-partition
    printf("%-28s warps/SMSP %d  %.3f warp-instr/clk/SMSP  (%.3f ms, %lld cyc, err %d)\n", name, threads / 128,
           per_smsp / (double)mx, ms, mx, (int)cudaGetLastError());
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512}) {
        run<0>("FFMA reg", threads); run<1>("FFMA imm", threads); run<2>("FFMA2 reg", threads);
        run<3>("FADD", threads); run<4>("FADD2", threads); run<5>("FMUL", threads); run<6>("FMUL2", threads);
        run<7>("FFMA+FADD 1:1", threads); run<8>("FFMA2+SHFL 3:1", threads); run<9>("FFMA+SHFL 3:1", threads);
        run<10>("FFMA2+FFMAimm 1:1", threads); run<11>("FFMA2+FADD2 1:1", threads); run<12>("FFMA2+LG2 7:1", threads);
        run<13>("16 FFMA2 | 16 FFMA blocks", threads); run<14>("FFMA2 imm", threads); run<15>("FFMA2+FFMA 2:2", threads);
    }
    return 0;
}
