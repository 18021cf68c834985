// dedup.cu -- resident Chromaprint store and the content-duplicate scan, sm_100a.
//
// Replaces the O(#tracks) Python loop of the reference's ingest step 5:
//   audio-ident-service/app/audio/dedup.py:127-167  _fingerprint_similarity  (32-bit Hamming similarity with a
//                                                    length penalty, compared over the overlapping prefix)
//   audio-ident-service/app/audio/dedup.py:170-222  check_content_duplicate  (rows whose duration lies within
//                                                    [0.9 d, 1.1 d]; strictly-greater running best; >= threshold)
// Definition of the result: oracle/aid_oracle.c aid_oracle_dedup_scan(), pinned against the reference's own
// functions by tests/golden/dedup_contract.json.
//
// Design: the fingerprints of all stored tracks are one uint32 array in HBM, every row padded to 16 B, with a row
// table (`start`, `len`, float64 `duration`). A scan is one streaming pass: a warp takes 32 consecutive rows, reads
// their table entries with one coalesced load each, drops the rows outside the query's duration window (the SQL
// WHERE of the reference) with a ballot, and for every remaining row XORs the overlapping prefix against the query
// with 512 B warp loads (uint4 per lane, up to four in flight) and popcounts. The query sits in shared memory.
// The similarity is evaluated with the reference's operations in the reference's order in IEEE double
// (matching/total, min/max, one multiply), so it is bit-identical to Python's floats; a float estimate first skips
// rows that cannot reach the running best. Each CTA keeps its best (similarity, lowest row); a second kernel folds
// the per-CTA candidates. Integer and byte work, bound by HBM: algorithmic bytes = 4 * sum over candidate rows of
// min(len_row, len_query) + 20 B of row table per row.
#include "common.cuh"
#include <algorithm>
#include <cstring>
#include <new>
#include <string>
#include <vector>

namespace {

constexpr int kDedupWarps = 8;                 // warps per CTA; a warp owns 32 consecutive rows
constexpr int kDedupRowsPerCta = kDedupWarps * 32;
constexpr int kDedupMaxQueryWords = 12288;     // query words staged in shared memory (48 KB); longer queries read global memory

struct DedupBest {                             // running best of the reference's loop: similarity > best, first row wins ties
    double sim;
    int64_t row;
};

__device__ __forceinline__ bool better(double s, int64_t r, const DedupBest& b) {
    return s > b.sim || (s == b.sim && s > 0.0 && r < b.row);
}

__device__ __forceinline__ uint32_t popc4(uint4 a, uint4 b) {
    return __popc(a.x ^ b.x) + __popc(a.y ^ b.y) + __popc(a.z ^ b.z) + __popc(a.w ^ b.w);
}

// differing bits of words [idx, idx+4) of a row against the query, counting only words below min_len
__device__ __forceinline__ uint32_t diff4(uint4 a, uint4 b, int idx, int min_len) {
    if (idx + 4 <= min_len) return popc4(a, b);
    uint32_t d = __popc(a.x ^ b.x);
    if (idx + 1 < min_len) d += __popc(a.y ^ b.y);
    if (idx + 2 < min_len) d += __popc(a.z ^ b.z);
    return d;                                   // idx + 3 >= min_len here
}

// CTA (q, blk): query q = blockIdx.x % nq against rows [blk * 256, blk * 256 + 256). The query index is the fast one so
// that the CTAs that stream the same rows for different queries run together and share them in L2.
// Rows start at multiples of 4 words (16 B) so that a lane reads one uint4; `start` and `len` replace the caller's
// ragged offsets inside the store.
__global__ void __launch_bounds__(kDedupWarps * 32, 4)
k_dedup_scan(const uint32_t* __restrict__ words, const int64_t* __restrict__ start, const int32_t* __restrict__ len,
             const double* __restrict__ dur, int64_t n_rows, const uint32_t* __restrict__ q_words,
             const int64_t* __restrict__ q_off, const double* __restrict__ q_lo, const double* __restrict__ q_hi, int nq,
             DedupBest* __restrict__ partial) {
    extern __shared__ __align__(16) uint32_t s_q[];
    __shared__ DedupBest s_best[kDedupWarps];
    const int q = blockIdx.x % nq, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t blk = blockIdx.x / nq;
    const uint32_t* qw = q_words + q_off[q];
    const int64_t q_len64 = q_off[q + 1] - q_off[q];
    const int q_len = (int)q_len64;                              // < 2^26 words, checked by aid_dedup_scan
    const int q_staged = min(kDedupMaxQueryWords, q_len);
    const int q_padded = (q_staged + 3) & ~3;
    for (int i = threadIdx.x; i < q_padded; i += blockDim.x) s_q[i] = i < q_staged ? qw[i] : 0u;
    __syncthreads();
    const double lo = q_lo[q], hi = q_hi[q];

    // the warp's 32 rows: one coalesced load of the row table, then the candidates one after the other
    const int64_t my_row = blk * kDedupRowsPerCta + warp * 32 + lane;
    int64_t my_start = 0; int my_len = 0; bool cand = false;
    if (my_row < n_rows) {
        const double d = dur[my_row];
        my_start = start[my_row]; my_len = len[my_row];
        cand = d >= lo && d <= hi && my_len > 0 && q_len > 0;    // dedup.py:192-197; an empty side gives 0.0 (:145-151)
    }
    uint32_t todo = __ballot_sync(AID_FULL_MASK, cand);
    DedupBest best{0.0, -1};
    float best_f = 0.0f;
    const int64_t row_base = blk * kDedupRowsPerCta + warp * 32;

    // exact evaluation of one row, dedup.py:153-167; `differing` is the warp-wide count
    auto evaluate = [&](int r, int r_len, uint32_t differing) {
        const int min_len = min(r_len, q_len), max_len = max(r_len, q_len);
        const int64_t total_bits = (int64_t)min_len * 32;
        const int64_t matching_bits = total_bits - (int64_t)differing;
        // a float estimate (relative error below 1e-6) rules out rows that can neither beat nor tie the running best;
        // only the others pay for the exact double arithmetic
        const float est = __fdividef((float)matching_bits, 32.0f * (float)max_len);
        if (est >= best_f * 0.99999f) {
            // same operations in the same order in IEEE double (the counts are exact)
            const double length_penalty = __ddiv_rn((double)min_len, (double)max_len);
            const double sim = __dmul_rn(__ddiv_rn((double)matching_bits, (double)total_bits), length_penalty);
            if (better(sim, row_base + r, best)) { best.sim = sim; best.row = row_base + r; best_f = (float)sim; }
        }
    };

    while (todo) {                                               // two candidate rows per trip: 2 KB in flight per warp
        const int r0 = __ffs(todo) - 1;
        todo &= todo - 1;
        const bool two = todo != 0;
        const int r1 = two ? __ffs(todo) - 1 : r0;
        todo &= todo - 1;                                        // no-op when todo is already 0
        const int64_t s0 = __shfl_sync(AID_FULL_MASK, my_start, r0), s1 = __shfl_sync(AID_FULL_MASK, my_start, r1);
        const int l0 = __shfl_sync(AID_FULL_MASK, my_len, r0), l1 = __shfl_sync(AID_FULL_MASK, my_len, r1);
        const int m0 = min(l0, q_len), m1 = two ? min(l1, q_len) : 0;
        const int g0 = min(q_staged, m0), g1 = min(q_staged, m1);        // words compared against the shared-memory copy
        const uint4* w0 = reinterpret_cast<const uint4*>(words + s0);
        const uint4* w1 = reinterpret_cast<const uint4*>(words + s1);
        uint32_t d0 = 0, d1 = 0;
        const int g = max(g0, g1);
        for (int base = 0; base < g; base += 256) {
            const int i0 = base + 4 * lane, i1 = i0 + 128;
            uint4 a00, a01, a10, a11;
            if (i0 < g0) a00 = __ldg(w0 + (i0 >> 2));
            if (i1 < g0) a01 = __ldg(w0 + (i1 >> 2));
            if (i0 < g1) a10 = __ldg(w1 + (i0 >> 2));
            if (i1 < g1) a11 = __ldg(w1 + (i1 >> 2));
            if (i0 < g) {
                const uint4 b0 = *reinterpret_cast<const uint4*>(s_q + i0);
                if (i0 < g0) d0 += diff4(a00, b0, i0, g0);
                if (i0 < g1) d1 += diff4(a10, b0, i0, g1);
            }
            if (i1 < g) {
                const uint4 b1 = *reinterpret_cast<const uint4*>(s_q + i1);
                if (i1 < g0) d0 += diff4(a01, b1, i1, g0);
                if (i1 < g1) d1 += diff4(a11, b1, i1, g1);
            }
        }
        for (int j = g0 + lane; j < m0; j += 32) d0 += __popc(__ldg(words + s0 + j) ^ __ldg(qw + j));   // query longer than the stage
        for (int j = g1 + lane; j < m1; j += 32) d1 += __popc(__ldg(words + s1 + j) ^ __ldg(qw + j));
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            d0 += __shfl_xor_sync(AID_FULL_MASK, d0, o);
            d1 += __shfl_xor_sync(AID_FULL_MASK, d1, o);
        }
        evaluate(r0, l0, d0);
        if (two) evaluate(r1, l1, d1);
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < kDedupWarps; w++)
            if (better(s_best[w].sim, s_best[w].row, best)) best = s_best[w];
        partial[(int64_t)q * (gridDim.x / nq) + blk] = best;
    }
}

__global__ void k_dedup_fold(const DedupBest* __restrict__ partial, int n_partial, int64_t* __restrict__ best_row,
                             double* __restrict__ best_sim) {
    __shared__ DedupBest s_best[8];
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    DedupBest best{0.0, -1};
    for (int i = threadIdx.x; i < n_partial; i += blockDim.x) {
        const DedupBest c = partial[(int64_t)q * n_partial + i];
        if (better(c.sim, c.row, best)) best = c;
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        DedupBest c;
        c.sim = __shfl_xor_sync(AID_FULL_MASK, best.sim, o);
        c.row = __shfl_xor_sync(AID_FULL_MASK, best.row, o);
        if (better(c.sim, c.row, best)) best = c;
    }
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++)
            if (better(s_best[w].sim, s_best[w].row, best)) best = s_best[w];
        best_row[q] = best.row; best_sim[q] = best.sim;
    }
}

template <typename T>
struct DBuf {                                   // grow-only device array that keeps its contents
    T* p = nullptr;
    int64_t cap = 0;
    cudaError_t reserve(int64_t n, int64_t keep, cudaStream_t st) {
        if (n <= cap) return cudaSuccess;
        int64_t want = cap ? cap : 1024;
        while (want < n) want += want / 2 + 1024;
        T* np = nullptr;
        cudaError_t ce = cudaMalloc(&np, (size_t)want * sizeof(T));
        if (ce != cudaSuccess) return ce;
        if (keep > 0) {
            ce = cudaMemcpyAsync(np, p, (size_t)keep * sizeof(T), cudaMemcpyDeviceToDevice, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
            if (ce != cudaSuccess) { cudaFree(np); return ce; }
        }
        if (p) cudaFree(p);
        p = np; cap = want;
        return cudaSuccess;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct aid_dedup {
    int device = 0;
    std::string err;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DBuf<uint32_t> words, q_words;           // words: every row padded to a multiple of 4 words
    DBuf<int64_t> start, q_off, best_row;    // start[row]: first word of the row in `words`
    DBuf<int32_t> len;                       // len[row]: words in the row
    DBuf<double> dur, q_lo, q_hi, best_sim;
    DBuf<DedupBest> partial;
    int64_t n_rows = 0, n_words = 0;
    int64_t launches = 0;
    int sm_count = 148;
    float last_scan_ms = 0.f;
    int64_t last_scan_bytes = 0;
};

static int dd_fail(aid_dedup* d, cudaError_t ce, const char* what) {
    d->err = std::string(what) + ": " + cudaGetErrorString(ce);
    cudaGetLastError();
    return AID_E_CUDA;
}
#define DD_CUDA(d, call) do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) return dd_fail(d, ce_, #call); } while (0)

extern "C" int aid_dedup_create(int device, aid_dedup** out) {
    if (!out) return AID_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return AID_E_CUDA; }
    aid_dedup* d = new (std::nothrow) aid_dedup();
    if (!d) return AID_E_ARG;
    d->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&d->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&d->ev0) != cudaSuccess || cudaEventCreate(&d->ev1) != cudaSuccess) {
        cudaGetLastError(); delete d; return AID_E_CUDA;
    }
    cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, device);
    cudaFuncSetAttribute(k_dedup_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, kDedupMaxQueryWords * 4);
    *out = d;
    return AID_OK;
}

extern "C" void aid_dedup_destroy(aid_dedup* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    cudaStreamSynchronize(d->st);
    d->words.release(); d->q_words.release(); d->start.release(); d->len.release(); d->q_off.release(); d->best_row.release();
    d->dur.release(); d->q_lo.release(); d->q_hi.release(); d->best_sim.release(); d->partial.release();
    cudaEventDestroy(d->ev0); cudaEventDestroy(d->ev1);
    cudaStreamDestroy(d->st);
    delete d;
}

extern "C" const char* aid_dedup_last_error(const aid_dedup* d) { return d ? d->err.c_str() : "null store"; }
extern "C" int64_t aid_dedup_size(const aid_dedup* d) { return d ? d->n_rows : 0; }
extern "C" int64_t aid_dedup_launch_count(const aid_dedup* d) { return d ? d->launches : 0; }

extern "C" int aid_dedup_add(aid_dedup* d, const uint32_t* words, const int64_t* word_off, const double* duration,
                             int n, int64_t* first_row) {
    if (!d || n < 0 || (n > 0 && (!word_off || !duration))) return AID_E_ARG;
    if (first_row) *first_row = d->n_rows;
    if (n == 0) return AID_OK;
    if (word_off[0] != 0) return AID_E_ARG;
    for (int i = 0; i < n; i++) if (word_off[i + 1] < word_off[i] || word_off[i + 1] - word_off[i] >= ((int64_t)1 << 26)) return AID_E_ARG;
    if (word_off[n] > 0 && !words) return AID_E_ARG;
    // repack: every row starts at a multiple of 4 words (the scan reads uint4), padding is zero
    std::vector<int64_t> h_start((size_t)n);
    std::vector<int32_t> h_len((size_t)n);
    int64_t padded = 0;
    for (int i = 0; i < n; i++) {
        const int64_t l = word_off[i + 1] - word_off[i];
        h_start[i] = d->n_words + padded; h_len[i] = (int32_t)l;
        padded += (l + 3) & ~(int64_t)3;
    }
    std::vector<uint32_t> h_words((size_t)padded, 0u);
    for (int i = 0; i < n; i++)
        if (h_len[i]) memcpy(h_words.data() + (h_start[i] - d->n_words), words + word_off[i], (size_t)h_len[i] * 4);
    DD_CUDA(d, cudaSetDevice(d->device));
    DD_CUDA(d, d->words.reserve(d->n_words + padded + 4, d->n_words, d->st));
    DD_CUDA(d, d->start.reserve(d->n_rows + n, d->n_rows, d->st));
    DD_CUDA(d, d->len.reserve(d->n_rows + n, d->n_rows, d->st));
    DD_CUDA(d, d->dur.reserve(d->n_rows + n, d->n_rows, d->st));
    if (padded) DD_CUDA(d, cudaMemcpyAsync(d->words.p + d->n_words, h_words.data(), (size_t)padded * 4, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaMemcpyAsync(d->start.p + d->n_rows, h_start.data(), (size_t)n * 8, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaMemcpyAsync(d->len.p + d->n_rows, h_len.data(), (size_t)n * 4, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaMemcpyAsync(d->dur.p + d->n_rows, duration, (size_t)n * 8, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaStreamSynchronize(d->st));
    d->n_rows += n; d->n_words += padded;
    return AID_OK;
}

// Device-resident variant of the scan (queries already in HBM, results left in HBM); the host entry point wraps it.
static int dedup_scan_dev(aid_dedup* d, const uint32_t* dq_words, const int64_t* dq_off, const double* dq_lo,
                          const double* dq_hi, int nq, int64_t max_q_len, int64_t* d_best_row, double* d_best_sim) {
    const int64_t n_blk = std::max<int64_t>(1, (d->n_rows + kDedupRowsPerCta - 1) / kDedupRowsPerCta);
    if (n_blk * nq >= ((int64_t)1 << 31)) return AID_E_ARG;
    DD_CUDA(d, d->partial.reserve(n_blk * nq, 0, d->st));
    const size_t smem = (size_t)((std::min<int64_t>(kDedupMaxQueryWords, std::max<int64_t>(max_q_len, 1)) + 3) & ~3) * 4;
    k_dedup_scan<<<(unsigned)(n_blk * nq), kDedupWarps * 32, smem, d->st>>>(d->words.p, d->start.p, d->len.p, d->dur.p, d->n_rows,
                                                                          dq_words, dq_off, dq_lo, dq_hi, nq, d->partial.p);
    DD_CUDA(d, cudaGetLastError());
    k_dedup_fold<<<nq, 256, 0, d->st>>>(d->partial.p, (int)n_blk, d_best_row, d_best_sim);
    DD_CUDA(d, cudaGetLastError());
    d->launches += 2;
    return AID_OK;
}

extern "C" int aid_dedup_scan(aid_dedup* d, const uint32_t* q_words, const int64_t* q_off, const double* q_lo,
                              const double* q_hi, int nq, int64_t* best_row, double* best_sim) {
    if (!d || nq < 0 || (nq > 0 && (!q_off || !q_lo || !q_hi || !best_row || !best_sim))) return AID_E_ARG;
    if (nq == 0) return AID_OK;
    if (nq > 65535 || q_off[0] != 0) return AID_E_ARG;
    int64_t max_q_len = 0;
    for (int i = 0; i < nq; i++) {
        if (q_off[i + 1] < q_off[i] || q_off[i + 1] - q_off[i] >= ((int64_t)1 << 26)) return AID_E_ARG;
        max_q_len = std::max(max_q_len, q_off[i + 1] - q_off[i]);
    }
    const int64_t nw = q_off[nq];
    if (nw > 0 && !q_words) return AID_E_ARG;
    DD_CUDA(d, cudaSetDevice(d->device));
    DD_CUDA(d, d->q_words.reserve(std::max<int64_t>(nw, 1), 0, d->st));
    DD_CUDA(d, d->q_off.reserve(nq + 1, 0, d->st));
    DD_CUDA(d, d->q_lo.reserve(nq, 0, d->st));
    DD_CUDA(d, d->q_hi.reserve(nq, 0, d->st));
    DD_CUDA(d, d->best_row.reserve(nq, 0, d->st));
    DD_CUDA(d, d->best_sim.reserve(nq, 0, d->st));
    if (nw) DD_CUDA(d, cudaMemcpyAsync(d->q_words.p, q_words, (size_t)nw * 4, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaMemcpyAsync(d->q_off.p, q_off, (size_t)(nq + 1) * 8, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaMemcpyAsync(d->q_lo.p, q_lo, (size_t)nq * 8, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaMemcpyAsync(d->q_hi.p, q_hi, (size_t)nq * 8, cudaMemcpyHostToDevice, d->st));
    DD_CUDA(d, cudaEventRecord(d->ev0, d->st));
    const int rc = dedup_scan_dev(d, d->q_words.p, d->q_off.p, d->q_lo.p, d->q_hi.p, nq, max_q_len, d->best_row.p, d->best_sim.p);
    if (rc != AID_OK) return rc;
    DD_CUDA(d, cudaEventRecord(d->ev1, d->st));
    DD_CUDA(d, cudaMemcpyAsync(best_row, d->best_row.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, d->st));
    DD_CUDA(d, cudaMemcpyAsync(best_sim, d->best_sim.p, (size_t)nq * 8, cudaMemcpyDeviceToHost, d->st));
    DD_CUDA(d, cudaStreamSynchronize(d->st));
    DD_CUDA(d, cudaEventElapsedTime(&d->last_scan_ms, d->ev0, d->ev1));
    return AID_OK;
}

extern "C" double aid_dedup_last_scan_ms(const aid_dedup* d) { return d ? (double)d->last_scan_ms : 0.0; }
