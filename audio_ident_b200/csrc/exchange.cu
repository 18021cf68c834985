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
#include <cstring>
#include "engine.h"
#include "index.h"

struct aid_exchange {
    aid_engine* e = nullptr;
    int rank = 0, world = 0, max_q = 0;
    unsigned char* window = nullptr;                 // this rank's receive window
    unsigned char* peer[AID_MAX_RANKS] = {};
    bool ipc_open[AID_MAX_RANKS] = {};
    bool connected = false;
    uint32_t epoch = 0;
    uint32_t* d_done = nullptr;                      // [0] = k_rank completion counter, [1] = error word
    int64_t timeout_ms = 20000;
};

namespace {

constexpr int kMergeWarps = 4;
constexpr int kMergeCap = AID_MAX_RANKS * AID_MAX_ROWS;

__global__ void __launch_bounds__(kMergeWarps * 32)
k_merge_blocks(unsigned char* __restrict__ window, int world, int max_q, uint32_t epoch, int n_q, int max_rows,
               aid_match_row* __restrict__ rows, int32_t* __restrict__ n_rows, uint32_t* __restrict__ err,
               long long timeout_cycles) {
    __shared__ uint64_t s_key[kMergeWarps][kMergeCap];      // (0xffffffff - count) << 32 | global track
    __shared__ int32_t s_off[kMergeWarps][kMergeCap];
    __shared__ uint16_t s_src[kMergeWarps][kMergeCap];      // source rank << 6 | row

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = blockIdx.x * kMergeWarps + warp;
    if (q >= n_q) return;

    // every rank's block of this epoch has to be complete: lane p watches the flag rank p publishes
    bool ok = true;
    if (lane < world) {
        const uint32_t* flag = reinterpret_cast<const uint32_t*>(window) + lane;
        const long long t0 = clock64();
        for (;;) {
            uint32_t v;
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
            if ((int32_t)(v - epoch) >= 0) break;
            if (clock64() - t0 > timeout_cycles) { ok = false; break; }
            __nanosleep(100);
        }
    }
    if (!__all_sync(AID_FULL_MASK, ok)) {                    // a peer never delivered: report, do not invent rows
        if (lane == 0) { atomicExch(err, 1u); n_rows[q] = -1; }
        return;
    }

    const int parity = (int)(epoch & 1u);
    const int cnt = lane < world ? min(max(xchg_counts(window, world, max_q, parity, lane)[q], 0), AID_MAX_ROWS) : 0;
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(AID_FULL_MASK, incl, d);
        if (lane >= d) incl += v;
    }
    const int total = __shfl_sync(AID_FULL_MASK, incl, 31);
    for (int p = 0; p < world; p++) {
        const int c = __shfl_sync(AID_FULL_MASK, cnt, p), b = __shfl_sync(AID_FULL_MASK, incl - cnt, p);
        const aid_match_row* src = xchg_rows(window, world, max_q, parity, p) + (int64_t)q * AID_MAX_ROWS;
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
                (xchg_rows(window, world, max_q, parity, s >> 6) + (int64_t)q * AID_MAX_ROWS)[s & 63];
        }
    }
    if (lane == 0) n_rows[q] = min(total, max_rows);
}

int fail(aid_exchange* x, cudaError_t ce, const char* what) { return aid_fail_cuda(x->e, ce, what); }
#define X_CUDA(x, call) do { cudaError_t ce_ = (call); if (ce_ != cudaSuccess) return fail((x), ce_, #call); } while (0)

}  // namespace

extern "C" int aid_exchange_create(aid_engine* e, int rank, int world, int max_queries, aid_exchange** out) {
    if (!e || !out || world < 1 || world > AID_MAX_RANKS || rank < 0 || rank >= world || max_queries < 1) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    aid_exchange* x = new aid_exchange();
    x->e = e; x->rank = rank; x->world = world; x->max_q = max_queries;
    const size_t bytes = xchg_window_bytes(world, max_queries);
    cudaError_t ce = cudaMalloc(&x->window, bytes);          // a plain cudaMalloc: IPC cannot export pool memory
    if (ce == cudaSuccess) ce = cudaMemset(x->window, 0, bytes);
    if (ce == cudaSuccess) ce = cudaMalloc(&x->d_done, 8);
    if (ce == cudaSuccess) ce = cudaMemset(x->d_done, 0, 8);
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
    if (ce != cudaSuccess) {
        if (x->window) cudaFree(x->window);
        if (x->d_done) cudaFree(x->d_done);
        delete x;
        return aid_fail_cuda(e, ce, "aid_exchange_create");
    }
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
    uint32_t err = 0;
    X_CUDA(x, cudaMemcpy(&err, x->d_done + 1, 4, cudaMemcpyDeviceToHost));
    if (err) {
        x->e->err = "a peer rank did not deliver its rows within the exchange timeout";
        return AID_E_TIMEOUT;
    }
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
    RowSink sink;
    sink.world = x->world; sink.rank = x->rank; sink.max_q = x->max_q;
    sink.epoch = ++x->epoch;
    for (int p = 0; p < x->world; p++) sink.window[p] = x->peer[p];
    sink.done = x->d_done;
    sink.track_map = d_track_map; sink.n_map = (uint32_t)n_map;
    int rc = aid_match_device_out(e, d_hash, d_t_anchor, d_hash_off, d_hash_len, d_status, n_queries, nullptr, max_rows,
                                  nullptr, sink, st);
    if (rc) return rc;
    int clock_khz = 0;
    AID_CUDA(e, cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, e->device));
    const long long timeout_cycles = (long long)x->timeout_ms * (clock_khz > 0 ? clock_khz : 2000000);
    { StageTimer tm(e, st, 7);
    k_merge_blocks<<<(n_queries + kMergeWarps - 1) / kMergeWarps, kMergeWarps * 32, 0, st>>>(
        x->window, x->world, x->max_q, sink.epoch, n_queries, max_rows, d_rows, d_n_rows, x->d_done + 1, timeout_cycles); }
    AID_CUDA(e, cudaGetLastError());
    e->launches += 1;
    return AID_OK;
}
