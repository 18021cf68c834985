// scan.cu -- device-wide exclusive prefix sum over uint32 (reduce / scan-of-sums / scan-and-add).
// Used three times on the path: peak-slot compaction, hash output offsets, and the index's
// bucket offsets (2^24 counters -> first posting of every hash).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kItems = 16;                       // per thread, as 4 rounds of one uint4
constexpr int kTile = kThreads * kItems;         // 4096 values per CTA

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(AID_FULL_MASK, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// the scan covers min(n, *d_n + 1) values when a device-side length is given
__device__ __forceinline__ int64_t eff_len(int64_t n, const uint32_t* d_n) {
    if (!d_n) return n;
    const int64_t m = (int64_t)(*d_n) + 1;
    return m < n ? m : n;
}

__device__ __forceinline__ uint4 load4(const uint32_t* in, int64_t i, int64_t n) {
    if (i + 3 < n && (reinterpret_cast<uintptr_t>(in + i) & 15) == 0) return *reinterpret_cast<const uint4*>(in + i);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (i < n) v.x = in[i];
    if (i + 1 < n) v.y = in[i + 1];
    if (i + 2 < n) v.z = in[i + 2];
    if (i + 3 < n) v.w = in[i + 3];
    return v;
}

__global__ void __launch_bounds__(kThreads)
k_scan_reduce(const uint32_t* __restrict__ in, int64_t n, const uint32_t* __restrict__ d_n,
              uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_w[kThreads / 32];
    n = eff_len(n, d_n);
    const int64_t base = (int64_t)blockIdx.x * kTile;
    if (base >= n) return;
    uint32_t acc = 0;
#pragma unroll
    for (int r = 0; r < kItems / 4; r++) {
        const uint4 v = load4(in, base + (int64_t)r * kThreads * 4 + threadIdx.x * 4, n);
        acc += v.x + v.y + v.z + v.w;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(AID_FULL_MASK, acc, d);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kThreads / 32; w++) t += s_w[w];
        sums[blockIdx.x] = t;
    }
}

// one CTA: exclusive scan of the tile sums in place; total -> *total
__global__ void __launch_bounds__(1024)
k_scan_sums(uint32_t* __restrict__ sums, int64_t n, const uint32_t* __restrict__ d_n,
            uint32_t* __restrict__ total) {
    __shared__ uint32_t s_w[32];
    __shared__ uint32_t s_carry, s_chunk;
    const int64_t nb = (eff_len(n, d_n) + kTile - 1) / kTile;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < nb; b += 1024) {
        const int64_t i = b + threadIdx.x;
        const uint32_t v = i < nb ? sums[i] : 0;
        const uint32_t incl = warp_incl_scan(v, lane);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = s_w[lane];
            const uint32_t wi = warp_incl_scan(w, lane);
            s_w[lane] = wi - w;
            if (lane == 31) s_chunk = wi;
        }
        __syncthreads();
        const uint32_t carry = s_carry;
        if (i < nb) sums[i] = carry + s_w[warp] + incl - v;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_chunk;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total) *total = s_carry;
}

__global__ void __launch_bounds__(kThreads)
k_scan_final(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
             const uint32_t* __restrict__ d_n, const uint32_t* __restrict__ sums) {
    __shared__ uint32_t s_w[kThreads / 32];
    __shared__ uint32_t s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    n = eff_len(n, d_n);
    const int64_t base = (int64_t)blockIdx.x * kTile;
    if (base >= n) return;
    if (threadIdx.x == 0) s_carry = sums[blockIdx.x];
    __syncthreads();
#pragma unroll 1
    for (int r = 0; r < kItems / 4; r++) {
        const int64_t i = base + (int64_t)r * kThreads * 4 + threadIdx.x * 4;
        const uint4 v = load4(in, i, n);
        const uint32_t mine = v.x + v.y + v.z + v.w;
        const uint32_t incl = warp_incl_scan(mine, lane);
        if (lane == 31) s_w[warp] = incl;
        __syncthreads();
        uint32_t wbefore = 0, all = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; w++) {
            const uint32_t c = s_w[w];
            all += c;
            wbefore += w < warp ? c : 0;
        }
        const uint32_t carry = s_carry;
        uint32_t e = carry + wbefore + incl - mine;
        if (i < n) out[i] = e;
        e += v.x; if (i + 1 < n) out[i + 1] = e;
        e += v.y; if (i + 2 < n) out[i + 2] = e;
        e += v.z; if (i + 3 < n) out[i + 3] = e;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + all;
        __syncthreads();
    }
}

}  // namespace

size_t aid_scan_tmp_elems(int64_t n) { return (size_t)((n + kTile - 1) / kTile) + 1; }

cudaError_t aid_launch_scan_u32(const uint32_t* d_in, uint32_t* d_out, int64_t n, uint32_t* d_tmp,
                                uint32_t* d_total, const uint32_t* d_n, cudaStream_t st) {
    if (n <= 0) {
        if (d_total) return cudaMemsetAsync(d_total, 0, sizeof(uint32_t), st);
        return cudaSuccess;
    }
    const int64_t nb = (n + kTile - 1) / kTile;
    k_scan_reduce<<<(unsigned)nb, kThreads, 0, st>>>(d_in, n, d_n, d_tmp);
    k_scan_sums<<<1, 1024, 0, st>>>(d_tmp, n, d_n, d_total);
    k_scan_final<<<(unsigned)nb, kThreads, 0, st>>>(d_in, d_out, n, d_n, d_tmp);
    return cudaGetLastError();
}
