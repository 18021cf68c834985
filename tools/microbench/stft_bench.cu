// stft_bench.cu -- A/B timing and a double-precision spot check of k_stft, outside the engine.
// Build (tools/microbench/build.sh): compiles audio_ident_b200/csrc/stft.cu with -DAID_STFT_BENCH_VARIANT flags
// Usage: stft_bench [tracks=2048] [seconds=30] [reps=5] [variant=1]   (variant 0: scalar k_stft, 1: packed f32x2 k_stft_packed;
//        a variant other than 0 is also compared bit for bit with variant 0 over the whole spectrogram)
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <algorithm>
#include "../../audio_ident_b200/csrc/common.cuh"

__global__ void k_fill(float* p, int64_t n, uint32_t seed) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t h = (uint32_t)i * 2654435761u ^ seed; h ^= h >> 15; h *= 2246822519u; h ^= h >> 13; h *= 3266489917u; h ^= h >> 16;
        // a few tones + noise, so that the spectrum has both loud bins and near-nulls
        const float t = (float)(i % 480000) * (1.0f / 16000.0f);
        p[i] = 0.4f * __sinf(6.2831853f * 440.0f * t) + 0.2f * __sinf(6.2831853f * 3110.0f * t + 1.0f) +
               1e-3f * ((float)(h >> 8) * (1.0f / 8388608.0f) - 1.0f);
    }
}

// every entry of gmax must be the exact maximum of its 16-bin group of the stored row
__global__ void k_check_gmax(const float* spec, const float* gmax, int64_t rows, unsigned long long* cnt) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows * 32; i += (int64_t)gridDim.x * blockDim.x) {
        const float* g = spec + (i >> 5) * AID_NBINS + (i & 31) * 16;
        float m = g[0];
        for (int k = 1; k < 16; k++) m = fmaxf(m, g[k]);
        c += __float_as_uint(m) != __float_as_uint(gmax[i]);
    }
    if (c) atomicAdd(cnt, c);
}

__global__ void k_diff(const uint32_t* a, const uint32_t* b, int64_t n, unsigned long long* cnt) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c += a[i] != b[i];
    if (c) atomicAdd(cnt, c);
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

int main(int argc, char** argv) {
    const int tracks = argc > 1 ? atoi(argv[1]) : 2048;
    const double seconds = argc > 2 ? atof(argv[2]) : 30.0;
    const int reps = argc > 3 ? atoi(argv[3]) : 5;
    const int variant = argc > 4 ? atoi(argv[4]) : 7;
    const int with_gmax = argc > 5 ? atoi(argv[5]) : 0;      // 1: the kernel also writes the group maxima (checked against the rows)
    const int64_t ns = (int64_t)(seconds * 16000.0);
    const int64_t T = (ns - AID_NFFT) / AID_HOP + 1;
    std::vector<aid_stft_unit> units;
    for (int i = 0; i < tracks; i++)
        for (int64_t f0 = 0; f0 < T; f0 += AID_STFT_UNIT_FRAMES) {
            aid_stft_unit u; u.pcm_begin = i * ns; u.n_samples = ns; u.spec_row = i * T + f0; u.frame0 = (int32_t)f0;
            u.n_frames = (int32_t)std::min<int64_t>(AID_STFT_UNIT_FRAMES, T - f0); units.push_back(u);
        }
    float *d_pcm, *d_spec, *d_win, *d_tw; aid_stft_unit* d_units;
    CK(cudaMalloc(&d_pcm, tracks * ns * 4)); CK(cudaMalloc(&d_spec, tracks * T * AID_NBINS * 4));
    CK(cudaMalloc(&d_units, units.size() * sizeof(aid_stft_unit)));
    CK(cudaMemcpy(d_units, units.data(), units.size() * sizeof(aid_stft_unit), cudaMemcpyHostToDevice));
    std::vector<float> win(AID_NFFT), tw(AID_TWIST_FLOATS);
    aid_fill_stft_tables(win.data(), tw.data());
    CK(cudaMalloc(&d_win, 4096)); CK(cudaMalloc(&d_tw, tw.size() * 4));
    CK(cudaMemcpy(d_win, win.data(), 4096, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_tw, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice));
    float* d_gmax = nullptr;
    if (with_gmax && variant != 0) CK(cudaMalloc(&d_gmax, (size_t)tracks * T * 32 * 4));
    k_fill<<<148 * 8, 256>>>(d_pcm, tracks * ns, 12345u);
    CK(cudaDeviceSynchronize());
    aid_tables tb{d_win, d_tw};
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int w = 0; w < 2; w++) CK(aid_launch_stft_variant(variant, tb, d_pcm, d_units, (int)units.size(), d_spec, d_gmax, 0));
    CK(cudaDeviceSynchronize());
    float best = 1e9f, sum = 0;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0);
        CK(aid_launch_stft_variant(variant, tb, d_pcm, d_units, (int)units.size(), d_spec, d_gmax, 0));
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms); sum += ms;
    }
    const double bytes = (double)tracks * ns * 4 + (double)tracks * T * AID_NBINS * 4;
    printf("variant %d: ", variant);
    printf("k_stft %d tracks x %.0f s: best %.3f ms avg %.3f ms  -> %.1f GB/s algorithmic = %.1f %% of 6558.4\n", tracks, seconds,
           best, sum / reps, bytes / (sum / reps) * 1e-6, bytes / (sum / reps) * 1e-6 / 6558.4 * 100.0);

    if (variant != 0) {          // bit-for-bit against the scalar kernel
        float* d_ref; unsigned long long* d_cnt; unsigned long long h_cnt = 0;
        const int64_t n = (int64_t)tracks * T * AID_NBINS;
        CK(cudaMalloc(&d_ref, n * 4)); CK(cudaMalloc(&d_cnt, 8)); CK(cudaMemset(d_cnt, 0, 8));
        CK(aid_launch_stft_variant(0, tb, d_pcm, d_units, (int)units.size(), d_ref, nullptr, 0));
        k_diff<<<148 * 8, 256>>>((const uint32_t*)d_spec, (const uint32_t*)d_ref, n, d_cnt);
        CK(cudaMemcpy(&h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost));
        printf("bitwise vs variant 0: %llu of %lld values differ\n", h_cnt, (long long)n);
        cudaFree(d_ref); cudaFree(d_cnt);
    }
    if (d_gmax) {
        unsigned long long* d_cnt; unsigned long long h_cnt = 0;
        CK(cudaMalloc(&d_cnt, 8)); CK(cudaMemset(d_cnt, 0, 8));
        k_check_gmax<<<148 * 8, 256>>>(d_spec, d_gmax, (int64_t)tracks * T, d_cnt);
        CK(cudaMemcpy(&h_cnt, d_cnt, 8, cudaMemcpyDeviceToHost));
        printf("group maxima: %llu of %lld entries differ from the stored rows\n", h_cnt, (long long)tracks * T * 32);
        cudaFree(d_cnt);
    }
    // spot check against a double-precision DFT: tolerance |dS| <= AID_SPEC_TOL * max(|S|, 1)
    std::vector<float> pcm(ns), row(AID_NBINS);
    double worst = 0; int bad = 0;
    const int checks[6][2] = {{0, 0}, {0, 1}, {1, 777}, {tracks - 1, (int)T - 1}, {tracks / 2, (int)T - 2}, {3 % tracks, 63}};
    for (auto& c : checks) {
        const int tr = c[0]; const int64_t f = std::min<int64_t>(c[1], T - 1);
        CK(cudaMemcpy(pcm.data(), d_pcm + tr * ns, ns * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(row.data(), d_spec + (tr * T + f) * AID_NBINS, AID_NBINS * 4, cudaMemcpyDeviceToHost));
        for (int k = 0; k < AID_NBINS; k++) {
            double sr = 0, si = 0;
            for (int n = 0; n < AID_NFFT; n++) {
                const double v = (double)((float)((double)win[n])) * (double)pcm[f * AID_HOP + n];
                const double a = -2.0 * M_PI * (double)((int64_t)n * k % AID_NFFT) / AID_NFFT;
                sr += v * cos(a); si += v * sin(a);
            }
            const double S = log(1.0 + sr * sr + si * si);
            const double err = fabs((double)row[k] - S) / std::max(fabs(S), 1.0);
            worst = std::max(worst, err); if (err > AID_SPEC_TOL) bad++;
        }
    }
    printf("spot check: worst scaled error %.3g (tolerance %.3g), %d bins out of tolerance\n", worst, (double)AID_SPEC_TOL, bad);
    return bad ? 2 : 0;
}
