// exchange.cu -- sharded identification: row exchange over peer memory fused into the ranking kernel, sm_100a.
//
// the index is sharded over the GPUs of one box, every rank probes its shard for every query
// window and the per-rank rows (count, track, offset, q_first, q_last) have to meet before they can be ranked --
// the reference has one LMDB, so one `olaf_c query` sees everything (audio-ident-service/app/audio/fingerprint.py:185-193)
// and its rows arrive "sorted by match_count descending" (fingerprint.py:299-301). Definition of the merged result:
// audio_ident_b200/sharded.py merge_rows() (the union of the ranks' rows ordered by (count desc, global track asc,
// offset asc), first AID_MAX_ROWS kept), itself tested equal to one unsharded index.
//
// Instead of "k_rank -> NCCL all-gather of fixed 50-row blocks -> sort" this file gives every rank a receive window
// in its own HBM (layout in index.h, RowSink), maps the windows of all ranks into each other (CUDA IPC between the
// processes torchrun starts, plain pointers inside one process) and lets k_rank (match.cu) store a window's rows
// straight into all of them while it is ranking the next windows: only the rows that exist cross NVLink (a handful
// per window instead of 1000 B), there is no collective call and no host synchronisation. k_merge_blocks then waits,
// on the device, for the epoch flags of all ranks and merges the blocks, one warp per window.
//
// aid_identify_exchange_dev runs the whole step that way: rank r fingerprints its slice of the window batch,
// k_publish_hashes stores those fingerprints into every rank's window (instead of an all-gather padded to the largest
// rank, whose size needed a host synchronisation), k_wait_hashes holds the stream until all slices have arrived,
// then probe/vote, the fused ranking + row stores, and the merge.
#include <algorithm>
#include <climits>
#include <cstdint>
#include <cstring>
#include "engine.h"
#include "index.h"

struct aid_exchange {
    aid_engine* e = nullptr;
    int rank = 0, world = 0, max_q = 0;
    XchgLayout lay;
    unsigned char* window = nullptr;                 // this rank's receive window
    unsigned char* peer[AID_MAX_RANKS] = {};
    bool ipc_open[AID_MAX_RANKS] = {};
    bool connected = false;
    uint32_t epoch = 0;
    uint32_t* d_done = nullptr;                      // [0] k_rank counter, [1] timeout seen, [2] hash capacity exceeded, [3] publish counter
    int64_t timeout_ms = 20000;
    int clock_khz = 2000000;                         // SM clock for the timeout (queried once: the attribute is slow)
};

namespace {

constexpr int kMergeWarps = 4;
constexpr int kMergeCap = AID_MAX_RANKS * AID_MAX_ROWS;

// spins until *flag has reached `epoch` (flags only grow; a peer may already be one epoch ahead)
__device__ __forceinline__ bool wait_flag(const uint32_t* flag, uint32_t epoch, long long timeout_cycles) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if ((int32_t)(v - epoch) >= 0) return true;
        if (clock64() - t0 > timeout_cycles) return false;
        __nanosleep(100);
    }
}

// ---- query fingerprints of this rank's slice of the batch -> every rank's window
struct HashSink {
    XchgLayout lay;
    int rank = 0;
    uint32_t epoch = 0;
    unsigned char* window[AID_MAX_RANKS] = {};
    uint32_t* done = nullptr;                  // CTA completion counter (local)
};

__global__ void __launch_bounds__(256)
k_publish_hashes(const uint32_t* __restrict__ hash, const uint32_t* __restrict__ t, const uint32_t* __restrict__ hash_off,
                 const int32_t* __restrict__ status, int count, int first_window, const HashSink sink) {
    const int parity = (int)(sink.epoch & 1u), P = sink.lay.world;
    const uint32_t total = count > 0 ? hash_off[count] : 0u;
    const bool fits = total <= sink.lay.hcap;
    const size_t slot = (size_t)sink.rank * sink.lay.hcap;
    const uint32_t stride = gridDim.x * blockDim.x, i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (fits)
        for (uint32_t i = i0; i < total; i += stride) {
            const uint32_t h = hash[i], tt = t[i];
            for (int p = 0; p < P; p++) {
                sink.lay.hash(sink.window[p], parity)[slot + i] = h;
                sink.lay.t(sink.window[p], parity)[slot + i] = tt;
            }
        }
    for (uint32_t j = i0; j < (uint32_t)count; j += stride) {
        const uint32_t b = hash_off[j], len = fits && (status[j] & 3) == 0 ? hash_off[j + 1] - b : 0u;
        for (int p = 0; p < P; p++) {
            sink.lay.begins(sink.window[p], parity)[first_window + j] = (uint32_t)slot + b;
            sink.lay.lens(sink.window[p], parity)[first_window + j] = len;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t prev = atomicAdd(sink.done, 1u);
        if (prev == gridDim.x - 1) {
            *sink.done = 0;
            __threadfence_system();
            for (int p = 0; p < P; p++) {
                XchgLayout::hash_status(sink.window[p])[sink.rank] = fits ? 0u : 1u;     // 1: slice larger than hcap
                asm volatile("st.release.sys.global.u32 [%0], %1;" :: "l"(XchgLayout::hash_flag(sink.window[p]) + sink.rank),
                             "r"(sink.epoch) : "memory");
            }
        }
    }
}

// one warp holds the stream until every rank's fingerprints of this epoch are in this rank's window
// (err[0] = timeout seen, err[1] = a slice did not fit: both sticky until aid_exchange_status reports and clears them;
// while err[0] is set k_match probes nothing -- the hash window is only partly written -- and the merge reports -1 rows)
__global__ void k_wait_hashes(unsigned char* window, int world, uint32_t epoch, uint32_t* err, long long timeout_cycles) {
    const int lane = threadIdx.x;
    if (lane >= world) return;
    if (!wait_flag(XchgLayout::hash_flag(window) + lane, epoch, timeout_cycles)) { atomicExch(err, 1u); return; }
    if (XchgLayout::hash_status(window)[lane] != 0) atomicExch(err + 1, 1u);
}

__global__ void __launch_bounds__(kMergeWarps * 32)
k_merge_blocks(unsigned char* __restrict__ window, const XchgLayout lay, uint32_t epoch, int n_q, int max_rows,
               aid_match_row* __restrict__ rows, int32_t* __restrict__ n_rows, uint32_t* __restrict__ err,
               long long timeout_cycles) {
    const int world = lay.world;
    __shared__ uint64_t s_key[kMergeWarps][kMergeCap];      // (0xffffffff - count) << 32 | global track
    __shared__ int32_t s_off[kMergeWarps][kMergeCap];
    __shared__ uint16_t s_src[kMergeWarps][kMergeCap];      // source rank << 6 | row

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kMergeWarps + warp;
    if (q >= n_q) return;

    // every rank's block of this epoch has to be complete: lane p watches the flag rank p publishes
    bool ok = *reinterpret_cast<volatile uint32_t*>(err) == 0;     // an earlier wait of this exchange already gave up
    if (ok && lane < world) {
        ok = wait_flag(XchgLayout::row_flag(window) + lane, epoch, timeout_cycles);
    }
    if (!__all_sync(AID_FULL_MASK, ok)) {                    // a peer never delivered: report, do not invent rows
        if (lane == 0) { atomicExch(err, 1u); n_rows[q] = -1; }
        return;
    }

    const int parity = (int)(epoch & 1u);
    const int cnt = lane < world ? min(max(lay.counts(window, parity, lane)[q], 0), AID_MAX_ROWS) : 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(AID_FULL_MASK, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(AID_FULL_MASK, incl, 31);
    for (int p = 0; p < world; p++) {
        const int c = __shfl_sync(AID_FULL_MASK, cnt, p), b = __shfl_sync(AID_FULL_MASK, incl - cnt, p);
        const aid_match_row* src = lay.rows(window, parity, p) + (int64_t)q * AID_MAX_ROWS;
        for (int i = lane; i < c; i += 32) {
            const aid_match_row r = src[i];
            s_key[warp][b + i] = (uint64_t)(0xffffffffu - (uint32_t)r.count) << 32 | r.track;
            s_off[warp][b + i] = r.offset;
            s_src[warp][b + i] = (uint16_t)(p << 6 | i);
        }
    }
    __syncwarp();
    // (track, offset) is unique in the union (a track lives on one rank), so counting smaller keys is a permutation
    for (int i = lane; i < total; i += 32) {
        const uint64_t k = s_key[warp][i];
        const int32_t o = s_off[warp][i];
        int pos = 0;
        for (int j = 0; j < total; j++) {
            const uint64_t kj = s_key[warp][j];
            pos += (kj < k) || (kj == k && s_off[warp][j] < o);
        }
        if (pos < max_rows) {
            const int s = s_src[warp][i];
            rows[(int64_t)q * max_rows + pos] =
                (lay.rows(window, parity, s >> 6) + (int64_t)q * AID_MAX_ROWS)[s & 63];
        }
    }
    if (lane == 0) n_rows[q] = min(total, max_rows);
}

int fail(aid_exchange* x, cudaError_t ce, const char* what) { return aid_fail_cuda(x->e, ce, what); }
#define X_CUDA(x, call) do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) return fail((x), ce_, #call); } while (0)

}  // namespace

extern "C" int aid_exchange_create(aid_engine* e, int rank, int world, int max_queries, int64_t max_hashes_per_rank,
                                   aid_exchange** out) {
    if (!e || !out || world < 1 || world > AID_MAX_RANKS || rank < 0 || rank >= world || max_queries < 1 ||
        max_hashes_per_rank < 0 || max_hashes_per_rank >= ((int64_t)1 << 28)) return AID_E_ARG;
    if (max_hashes_per_rank == 0)        // default: 1024 per window of a rank's slice (a 3.5 s window has ~300)
        max_hashes_per_rank = std::min<int64_t>(((int64_t)max_queries + world - 1) / world * 1024, ((int64_t)1 << 28) - 1);
    AID_CUDA(e, cudaSetDevice(e->device));
    aid_exchange* x = new aid_exchange();
    x->e = e; x->rank = rank; x->world = world; x->max_q = max_queries;
    x->lay = XchgLayout::make(world, max_queries, (uint32_t)max_hashes_per_rank);
    const size_t bytes = x->lay.bytes;
    cudaError_t ce = cudaMalloc(&x->window, bytes);          // a plain cudaMalloc: IPC cannot export pool memory
    if (ce == cudaSuccess) ce = cudaMemset(x->window, 0, bytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&x->d_done, 16);      // [0] k_rank counter, [1] timeout, [2] capacity, [3] publish counter
    if (ce == cudaSuccess) ce = cudaMemset(x->d_done, 0, 16);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
        if (x->window) cudaFree(x->window);
        if (x->d_done) cudaFree(x->d_done);
        delete x;
        return aid_fail_cuda(e, ce, "aid_exchange_create");
    }
    int khz = 0;
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, e->device) == cudaSuccess && khz > 0) x->clock_khz = khz;
    else cudaGetLastError();
    x->peer[rank] = x->window;
    x->connected = world == 1;
    *out = x;
    return AID_OK;
}

extern "C" void aid_exchange_destroy(aid_exchange* x) {
    if (!x) return;
    cudaSetDevice(x->e->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < x->world; p++)
        if (x->ipc_open[p]) cudaIpcCloseMemHandle(x->peer[p]);
    if (x->window) cudaFree(x->window);
    if (x->d_done) cudaFree(x->d_done);
    cudaGetLastError();
    delete x;
}

extern "C" int aid_exchange_handle(aid_exchange* x, uint8_t* handle) {
    if (!x || !handle) return AID_E_ARG;
    static_assert(sizeof(cudaIpcMemHandle_t) == AID_IPC_HANDLE_BYTES, "handle size");
    X_CUDA(x, cudaSetDevice(x->e->device));
    cudaIpcMemHandle_t h;
    X_CUDA(x, cudaIpcGetMemHandle(&h, x->window));
    std::memcpy(handle, &h, sizeof h);
    return AID_OK;
}

extern "C" int aid_exchange_connect(aid_exchange* x, const uint8_t* handles) {
    if (!x || !handles) return AID_E_ARG;
    X_CUDA(x, cudaSetDevice(x->e->device));
    for (int p = 0; p < x->world; p++) {
        if (p == x->rank || x->ipc_open[p]) continue;
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + (size_t)p * AID_IPC_HANDLE_BYTES, sizeof h);
        void* ptr = nullptr;
        X_CUDA(x, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        x->peer[p] = static_cast<unsigned char*>(ptr);
        x->ipc_open[p] = true;
    }
    x->connected = true;
    return AID_OK;
}

extern "C" int aid_exchange_connect_local(aid_exchange* x, aid_exchange* const* peers) {
    if (!x || !peers) return AID_E_ARG;
    X_CUDA(x, cudaSetDevice(x->e->device));
    for (int p = 0; p < x->world; p++) {
        if (p == x->rank) continue;
        const aid_exchange* y = peers[p];
        if (!y || y->world != x->world || y->max_q != x->max_q || y->rank != p) return AID_E_ARG;
        if (y->e->device != x->e->device) {
            const cudaError_t ce = cudaDeviceEnablePeerAccess(y->e->device, 0);
            if (ce == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else if (ce != cudaSuccess) return fail(x, ce, "cudaDeviceEnablePeerAccess");
        }
        x->peer[p] = y->window;
    }
    x->connected = true;
    return AID_OK;
}

extern "C" int aid_exchange_set_timeout_ms(aid_exchange* x, int64_t ms) {
    if (!x || ms < 1) return AID_E_ARG;
    x->timeout_ms = ms;
    return AID_OK;
}

extern "C" int aid_exchange_status(aid_exchange* x) {
    if (!x) return AID_E_ARG;
    X_CUDA(x, cudaSetDevice(x->e->device));
    uint32_t err[2] = {0, 0};
    X_CUDA(x, cudaDeviceSynchronize());            // the engine's streams are non-blocking: a plain cudaMemcpy would not wait for them
    X_CUDA(x, cudaMemcpy(err, x->d_done + 1, 8, cudaMemcpyDeviceToHost));
    if (err[0] || err[1]) X_CUDA(x, cudaMemset(x->d_done + 1, 0, 8));      // reported once: a transient failure must not fail every later step
    if (err[0]) {
        x->e->err = "a peer rank did not deliver its block within the exchange timeout";
        return AID_E_TIMEOUT;
    }
    if (err[1]) {
        x->e->err = "a rank's query fingerprints exceed max_hashes_per_rank of the exchange";
        return AID_E_CAPACITY;
    }
    return AID_OK;
}

// probe/vote + ranking with the row stores into every window (match.cu), then the merge of all ranks' blocks
static int match_and_merge(aid_engine* e, aid_exchange* x, uint32_t epoch, const uint32_t* d_hash, const uint32_t* d_t,
                           const uint32_t* d_hash_off, const uint32_t* d_hash_len, const int32_t* d_status, int n_q,
                           const uint32_t* d_track_map, int64_t n_map, aid_match_row* d_rows, int max_rows,
                           int32_t* d_n_rows, cudaStream_t st) {
    RowSink sink;
    sink.lay = x->lay; sink.rank = x->rank; sink.epoch = epoch;
    for (int p = 0; p < x->world; p++) sink.window[p] = x->peer[p];
    sink.done = x->d_done;
    sink.track_map = d_track_map; sink.n_map = (uint32_t)n_map;
    int rc = aid_match_device_out(e, d_hash, d_t, d_hash_off, d_hash_len, d_status, n_q, nullptr, max_rows, nullptr, sink, st,
                                  x->d_done + 1);
    if (rc) return rc;
    const long long timeout_cycles = (long long)x->timeout_ms * x->clock_khz;
    { StageTimer tm(e, st, 7);
    k_merge_blocks<<<(n_q + kMergeWarps - 1) / kMergeWarps, kMergeWarps * 32, 0, st>>>(
        x->window, x->lay, epoch, n_q, max_rows, d_rows, d_n_rows, x->d_done + 1, timeout_cycles); }
    AID_CUDA(e, cudaGetLastError());
    e->launches += 1;
    return AID_OK;
}

extern "C" int aid_match_exchange_dev(aid_engine* e, aid_exchange* x, const uint32_t* d_hash, const uint32_t* d_t_anchor,
                                      const uint32_t* d_hash_off, const uint32_t* d_hash_len, const int32_t* d_status,
                                      int n_queries, const uint32_t* d_track_map, int64_t n_map, aid_match_row* d_rows,
                                      int max_rows, int32_t* d_n_rows, void* stream) {
    if (!e || !x || x->e != e || !x->connected || n_queries < 0 || n_queries > x->max_q || max_rows < 1 ||
        max_rows > AID_MAX_ROWS || n_map < 0 || n_map >= ((int64_t)1 << 32)) return AID_E_ARG;
    if (n_queries > 0 && (!d_hash_off || !d_rows || !d_n_rows)) return AID_E_ARG;
    if (n_queries == 0) return AID_OK;               // nothing to publish: every rank passes the same batch
    AID_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->slot[0].st;
    return match_and_merge(e, x, ++x->epoch, d_hash, d_t_anchor, d_hash_off, d_hash_len, d_status, n_queries, d_track_map,
                           n_map, d_rows, max_rows, d_n_rows, st);
}

int aid_fingerprint_windows_core(aid_engine* e, const float* d_pcm, int64_t base, const int64_t* win_begin,
                                 const int64_t* win_end, int n, aid_fp_device_result* out, void* stream);   // engine.cu

// The whole sharded step. d_pcm points at sample `base` of the caller's numbering; window i is samples
// [win_begin[i], win_end[i]) of it -- windows may overlap (the three sub-windows of a 5 s clip, exact.py:48-52, or the
// sliding windows of a long recording share their samples instead of being copied out three / two times).
static int identify_core(aid_engine* e, aid_exchange* x, const float* d_pcm, int64_t base, const int64_t* win_begin,
                         const int64_t* win_end, int n_windows, const uint32_t* d_track_map, int64_t n_map,
                         aid_match_row* d_rows, int max_rows, int32_t* d_n_rows, cudaStream_t st) {
    // A vote window is at most AID_QUERY_MAX_FRAMES frames (k_match packs t_query into 15 bits of the vote key and 16 bits
    // of q_first / q_last). Checked for ALL windows and before the epoch moves: every rank passes the same batch, so every
    // rank returns the same status and the ranks' epochs stay in step.
    for (int i = 0; i < n_windows; i++) {
        if (win_end[i] < win_begin[i]) return AID_E_ARG;
        if (aid_num_frames(win_end[i] - win_begin[i]) > AID_QUERY_MAX_FRAMES) return AID_E_TOO_LONG;
    }
    const int P = x->world, r = x->rank;
    const int lo = (int)((int64_t)r * n_windows / P), hi = (int)((int64_t)(r + 1) * n_windows / P);
    aid_fp_device_result fp{};
    int rc = aid_fingerprint_windows_core(e, d_pcm, base, win_begin + lo, win_end + lo, hi - lo, &fp, st);     // this rank's slice only
    if (rc) return rc;
    const uint32_t epoch = ++x->epoch;
    const long long timeout_cycles = (long long)x->timeout_ms * x->clock_khz;
    HashSink hs;
    hs.lay = x->lay; hs.rank = r; hs.epoch = epoch; hs.done = x->d_done + 3;
    for (int p = 0; p < P; p++) hs.window[p] = x->peer[p];
    { StageTimer tm(e, st, 7);
    int sms = 148;
    k_publish_hashes<<<2 * sms, 256, 0, st>>>(fp.d_hash, fp.d_t_anchor, fp.d_hash_off, fp.d_status, hi - lo, lo, hs);
    k_wait_hashes<<<1, 32, 0, st>>>(x->window, P, epoch, x->d_done + 1, timeout_cycles); }
    AID_CUDA(e, cudaGetLastError());
    e->launches += 2;
    const int parity = (int)(epoch & 1u);
    const uint32_t* wh = x->lay.hash(x->window, parity);
    return match_and_merge(e, x, epoch, wh, x->lay.t(x->window, parity), x->lay.begins(x->window, parity),
                           x->lay.lens(x->window, parity), nullptr, n_windows, d_track_map, n_map, d_rows, max_rows,
                           d_n_rows, st);
}

static int identify_args_ok(aid_engine* e, aid_exchange* x, int n_windows, int max_rows, int64_t n_map) {
    return e && x && x->e == e && x->connected && n_windows >= 0 && n_windows <= x->max_q && max_rows >= 1 &&
           max_rows <= AID_MAX_ROWS && n_map >= 0 && n_map < ((int64_t)1 << 32);
}

extern "C" int aid_identify_exchange_dev(aid_engine* e, aid_exchange* x, const float* d_pcm, const int64_t* sample_off,
                                         int n_windows, const uint32_t* d_track_map, int64_t n_map,
                                         aid_match_row* d_rows, int max_rows, int32_t* d_n_rows, void* stream) {
    if (!identify_args_ok(e, x, n_windows, max_rows, n_map) || !sample_off) return AID_E_ARG;
    if (n_windows > 0 && (!d_rows || !d_n_rows)) return AID_E_ARG;
    if (n_windows == 0) return AID_OK;
    AID_CUDA(e, cudaSetDevice(e->device));
    return identify_core(e, x, d_pcm, 0, sample_off, sample_off + 1, n_windows, d_track_map, n_map, d_rows, max_rows, d_n_rows,
                         stream ? (cudaStream_t)stream : e->slot[0].st);
}

extern "C" int aid_identify_exchange_windows_dev(aid_engine* e, aid_exchange* x, const float* d_pcm, const int64_t* win_begin,
                                                 const int64_t* win_end, int n_windows, const uint32_t* d_track_map,
                                                 int64_t n_map, aid_match_row* d_rows, int max_rows, int32_t* d_n_rows,
                                                 void* stream) {
    if (!identify_args_ok(e, x, n_windows, max_rows, n_map)) return AID_E_ARG;
    if (n_windows > 0 && (!d_rows || !d_n_rows || !win_begin || !win_end)) return AID_E_ARG;
    if (n_windows == 0) return AID_OK;
    for (int i = 0; i < n_windows; i++) if (win_begin[i] < 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    return identify_core(e, x, d_pcm, 0, win_begin, win_end, n_windows, d_track_map, n_map, d_rows, max_rows, d_n_rows,
                         stream ? (cudaStream_t)stream : e->slot[0].st);
}

// Host buffers in, host buffers out: what a service process hands over (the windows' PCM in pinned or pageable host
// memory) and what it gets back (rows of its own slice of the batch, or of all windows). Rank r only needs the samples
// its windows [r*n/P, (r+1)*n/P) cover, so that span is all that crosses PCIe -- once, however much the windows
// overlap; the merged rows of every window end up on every rank and `rows_first / rows_count` says which of them this
// caller wants copied back.
static int identify_host(aid_engine* e, aid_exchange* x, const float* pcm, const int64_t* win_begin, const int64_t* win_end,
                         int n_windows, const uint32_t* d_track_map, int64_t n_map, int rows_first, int rows_count,
                         aid_match_row* rows, int max_rows, int32_t* n_rows) {
    if (!identify_args_ok(e, x, n_windows, max_rows, n_map) || rows_first < 0 || rows_count < 0 ||
        rows_first + (int64_t)rows_count > n_windows) return AID_E_ARG;
    if (rows_count > 0 && (!rows || !n_rows)) return AID_E_ARG;
    if (n_windows == 0) return AID_OK;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    Index* ix = e->index;
    const int P = x->world, r = x->rank;
    // The batch is cut into up to kParts (8) consecutive parts, each a complete step (fingerprint, exchange, vote, merge) on the
    // engine's stream, and ALL uploads are queued first on a second stream: while part k is computed, part k + 1 crosses
    // PCIe. A step's compute (7-10 ms for 4,096 queries) then hides behind the copy (24-51 ms) instead of following it.
    constexpr int kParts = 8;
    // Only on a single rank: with P ranks the contract is that rank r is handed, and uploads, exactly the samples of ITS slice
    // [r n / P, (r + 1) n / P) of the batch (a caller may pass a pointer that is valid for that slice only), and a part of
    // the batch would give it other windows.
    const int parts = P == 1 ? std::max(1, std::min(kParts, n_windows / 64)) : 1;      // at least 64 windows per part
    int64_t span0[kParts], span1[kParts];
    int64_t all0 = INT64_MAX, all1 = INT64_MIN;
    for (int k = 0; k < parts; k++) {
        const int w0 = (int)((int64_t)k * n_windows / parts), w1 = (int)((int64_t)(k + 1) * n_windows / parts), nk = w1 - w0;
        const int lo = w0 + (int)((int64_t)r * nk / P), hi = w0 + (int)((int64_t)(r + 1) * nk / P);
        int64_t a = INT64_MAX, b = INT64_MIN;
        for (int i = lo; i < hi; i++) {
            if (win_end[i] < win_begin[i]) return AID_E_ARG;
            a = std::min(a, win_begin[i]); b = std::max(b, win_end[i]);
        }
        if (hi <= lo) { a = 0; b = 0; }
        span0[k] = a; span1[k] = b;
        if (b > a) { all0 = std::min(all0, a); all1 = std::max(all1, b); }
    }
    const int64_t samples = all1 > all0 ? all1 - all0 : 0;
    if (samples == 0) all0 = 0;
    if (samples > 0 && !pcm) return AID_E_ARG;
    AID_CUDA(e, s.pcm.ensure((size_t)std::max<int64_t>(samples, 1) * sizeof(float)));
    AID_CUDA(e, ix->rows.ensure((size_t)n_windows * max_rows * sizeof(aid_match_row)));
    AID_CUDA(e, ix->rows_n.ensure((size_t)n_windows * 4));
    cudaStream_t copy_st = parts > 1 ? e->slot[1].st : s.st;
    cudaEvent_t landed[kParts] = {};
    AID_CUDA(e, cudaStreamSynchronize(s.st));          // the PCM buffer may still be read by an earlier call on this engine
    int rc = AID_OK;
    // issue order: upload k, then step k (whose small descriptor upload queues behind upload k on the copy engine and in
    // front of upload k + 1), so step k runs while part k + 1 is on the bus
    for (int k = 0; k < parts && rc == AID_OK; k++) {
        const int w0 = (int)((int64_t)k * n_windows / parts), w1 = (int)((int64_t)(k + 1) * n_windows / parts);
        if (span1[k] > span0[k]) {
            cudaError_t ce = cudaMemcpyAsync(s.pcm.as<float>() + (span0[k] - all0), pcm + span0[k],
                                             (size_t)(span1[k] - span0[k]) * sizeof(float), cudaMemcpyHostToDevice, copy_st);
            if (ce != cudaSuccess) { rc = aid_fail_cuda(e, ce, "cudaMemcpyAsync(query pcm)"); break; }
        }
        if (parts > 1) {
            cudaError_t ce = cudaEventCreateWithFlags(&landed[k], cudaEventDisableTiming);
            if (ce == cudaSuccess) ce = cudaEventRecord(landed[k], copy_st);
            if (ce == cudaSuccess) ce = cudaStreamWaitEvent(s.st, landed[k], 0);
            if (ce != cudaSuccess) { rc = aid_fail_cuda(e, ce, "cudaEventRecord(query pcm)"); break; }
        }
        rc = identify_core(e, x, s.pcm.as<float>(), all0, win_begin + w0, win_end + w0, w1 - w0, d_track_map, n_map,
                           ix->rows.as<aid_match_row>() + (int64_t)w0 * max_rows, max_rows, ix->rows_n.as<int32_t>() + w0, s.st);
    }
    if (parts > 1) {
        cudaStreamSynchronize(copy_st);
        for (int k = 0; k < parts; k++) if (landed[k]) cudaEventDestroy(landed[k]);
    }
    if (rc) { cudaStreamSynchronize(s.st); return rc; }
    if (rows_count > 0) {
        AID_CUDA(e, cudaMemcpyAsync(rows, ix->rows.as<aid_match_row>() + (int64_t)rows_first * max_rows,
                                    (size_t)rows_count * max_rows * sizeof(aid_match_row), cudaMemcpyDeviceToHost, s.st));
        AID_CUDA(e, cudaMemcpyAsync(n_rows, ix->rows_n.as<int32_t>() + rows_first, (size_t)rows_count * 4, cudaMemcpyDeviceToHost, s.st));
    }
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    for (int i = 0; i < rows_count; i++) if (n_rows[i] < 0) return AID_E_TIMEOUT;
    return AID_OK;
}

extern "C" int aid_identify_exchange_host(aid_engine* e, aid_exchange* x, const float* pcm, const int64_t* sample_off,
                                          int n_windows, const uint32_t* d_track_map, int64_t n_map,
                                          int rows_first, int rows_count, aid_match_row* rows, int max_rows,
                                          int32_t* n_rows) {
    if (!sample_off) return AID_E_ARG;
    return identify_host(e, x, pcm, sample_off, sample_off + 1, n_windows, d_track_map, n_map, rows_first, rows_count, rows,
                         max_rows, n_rows);
}

extern "C" int aid_identify_exchange_windows_host(aid_engine* e, aid_exchange* x, const float* pcm, const int64_t* win_begin,
                                                  const int64_t* win_end, int n_windows, const uint32_t* d_track_map,
                                                  int64_t n_map, int rows_first, int rows_count, aid_match_row* rows,
                                                  int max_rows, int32_t* n_rows) {
    if (n_windows > 0 && (!win_begin || !win_end)) return AID_E_ARG;
    for (int i = 0; i < n_windows; i++) if (win_begin[i] < 0) return AID_E_ARG;
    return identify_host(e, x, pcm, win_begin, win_end, n_windows, d_track_map, n_map, rows_first, rows_count, rows, max_rows,
                         n_rows);
}
