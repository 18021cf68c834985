// engine.cu -- host side of the C ABI: engine lifetime, ragged-batch planning, workspace, and the
// fingerprint entry points (include/audio_ident_b200.h). Index and query entry points: index.cu, match.cu.
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include "engine.h"
#include <cstdlib>

// ------------------------------------------------------------------------------------------ library
extern "C" int aid_abi_version(void) { return AID_ABI_VERSION; }

extern "C" const char* aid_strerror(int status) {
    switch (status) {
        case AID_OK: return "ok";
        case AID_E_CUDA: return "CUDA error";
        case AID_E_ARG: return "bad argument";
        case AID_E_CAPACITY: return "output buffer too small";
        case AID_E_TOO_LONG: return "track or query too long";
        case AID_E_NOT_FOUND: return "unknown track";
        case AID_E_IO: return "index directory I/O error";
        case AID_E_FORMAT: return "index files have the wrong format";
        case AID_E_FULL: return "index is full";
        case AID_E_TIMEOUT: return "a peer rank did not deliver its rows in time";
        default: return "unknown status";
    }
}

extern "C" void aid_get_params(int32_t* out) {
    out[0] = AID_SAMPLE_RATE; out[1] = AID_NFFT; out[2] = AID_HOP; out[3] = AID_NBINS;
    out[4] = AID_PEAK_HALF_F; out[5] = AID_PEAK_HALF_T; out[6] = AID_PEAK_MIN_BIN;
    out[7] = AID_DT_MIN; out[8] = AID_DT_MAX; out[9] = AID_DF_MIN; out[10] = AID_DF_MAX;
    out[11] = AID_FANOUT; out[12] = AID_MIN_VOTES; out[13] = AID_MAX_ROWS;
    out[14] = AID_QUERY_MAX_FRAMES; out[15] = AID_SEG_TRACKS;
}

extern "C" int aid_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return AID_E_CUDA; }
    return n;
}

extern "C" int64_t aid_num_frames(int64_t n_samples) {
    return n_samples < AID_NFFT ? 0 : (n_samples - AID_NFFT) / AID_HOP + 1;
}

int aid_fail_cuda(aid_engine* e, cudaError_t ce, const char* what) {
    if (e) {
        char buf[512];
        snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(ce));
        e->err = buf;
    }
    cudaGetLastError();
    return AID_E_CUDA;
}

// ------------------------------------------------------------------------------------------- engine
void Slot::release() {
    DevBuf* bufs[] = {&pcm, &desc, &spec, &gmax, &slots, &unit_pos, &peaks, &peak_track, &peak_off, &pos, &hash, &t,
                      &hash_off, &status, &scan_tmp, &misc};
    for (DevBuf* b : bufs) b->release();
    h_desc.release(); h_small.release();
    if (done) cudaEventDestroy(done);
    if (st) cudaStreamDestroy(st);
    done = nullptr; st = nullptr;
}

extern "C" int aid_engine_create(int device, aid_engine** out) {
    if (!out) return AID_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || device < 0 || device >= n) { cudaGetLastError(); return AID_E_CUDA; }
    aid_engine* e = new aid_engine();
    e->stft_variant = aid_stft_default_variant();
    if (const char* v = getenv("AID_PEAK_SUMMARY")) e->peak_summary = atoi(v) != 0;     // measurement runs
    e->device = device;
    auto fail = [&](cudaError_t ce, const char* what) { int r = aid_fail_cuda(e, ce, what); fprintf(stderr, "audio_ident_b200: %s\n", e->err.c_str()); aid_engine_destroy(e); return r; };
    cudaError_t ce;
    if ((ce = cudaSetDevice(device)) != cudaSuccess) return fail(ce, "cudaSetDevice");
    for (int i = 0; i < 2; i++) {
        if ((ce = cudaStreamCreateWithFlags(&e->slot[i].st, cudaStreamNonBlocking)) != cudaSuccess) return fail(ce, "cudaStreamCreate");
        if ((ce = cudaEventCreateWithFlags(&e->slot[i].done, cudaEventDisableTiming)) != cudaSuccess) return fail(ce, "cudaEventCreate");
    }
    // constant tables (definition: aid_fill_stft_tables in common.cuh)
    std::vector<float> win(AID_NFFT), tw(AID_TWIST_FLOATS);
    aid_fill_stft_tables(win.data(), tw.data());
    if ((ce = e->d_window.ensure(win.size() * sizeof(float))) != cudaSuccess) return fail(ce, "cudaMalloc(window)");
    if ((ce = e->d_twiddle.ensure(tw.size() * sizeof(float))) != cudaSuccess) return fail(ce, "cudaMalloc(twist)");
    if ((ce = cudaMemcpy(e->d_window.p, win.data(), win.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(ce, "cudaMemcpy(window)");
    if ((ce = cudaMemcpy(e->d_twiddle.p, tw.data(), tw.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess) return fail(ce, "cudaMemcpy(twist)");
    e->tables.window = e->d_window.as<float>();
    e->tables.twist = e->d_twiddle.as<float>();
    e->index = aid_index_new();
    *out = e;
    return AID_OK;
}

extern "C" void aid_engine_destroy(aid_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaDeviceSynchronize();
    if (e->index) aid_index_free(e, e->index);
    for (int i = 0; i < 2; i++) e->slot[i].release();
    for (auto& r : e->stage_recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (cudaEvent_t ev : e->event_pool) cudaEventDestroy(ev);
    e->d_window.release(); e->d_twiddle.release();
    delete e;
}

extern "C" const char* aid_last_error(const aid_engine* e) { return e ? e->err.c_str() : ""; }
extern "C" int64_t aid_launch_count(const aid_engine* e) { return e ? e->launches : 0; }

extern "C" int aid_engine_sync(aid_engine* e) {
    if (!e) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    for (int i = 0; i < 2; i++) AID_CUDA(e, cudaStreamSynchronize(e->slot[i].st));
    return AID_OK;
}

extern "C" int aid_engine_set_max_batch_frames(aid_engine* e, int64_t frames) {
    if (!e || frames < 1) return AID_E_ARG;
    e->max_batch_frames = frames;
    return AID_OK;
}

// --------------------------------------------------------------------------------------- planning
// Tracks [first, first+count) of a ragged batch. Tracks longer than frame_limit frames are planned
// with zero frames and flagged AID_TRACK_TOO_LONG.
int aid_build_plan(const int64_t* sample_off, int first, int count, int64_t frame_limit, Plan& plan) {
    return aid_build_plan_windows(sample_off + first, sample_off + first + 1, count, sample_off[first], frame_limit, plan);
}

// General form: track / window i is pcm[begin[i] .. end[i]) -- the ranges may overlap (the three sub-windows of a 5 s
// query clip, the sliding windows of a long recording) or leave gaps; pcm_begin of a unit is relative to `base`, the
// sample the caller's device pointer stands on.
int aid_build_plan_windows(const int64_t* begin_of, const int64_t* end_of, int count, int64_t base, int64_t frame_limit,
                           Plan& plan) {
    plan = Plan();
    plan.n_tracks = count;
    plan.frames.resize(count);
    plan.frame_off.assign(count + 1, 0);
    plan.host_status.assign(count, AID_TRACK_OK);
    plan.first_punit.assign(count + 1, 0);
    // blocks streamed back to back by one warp of the peak kernel: longer runs re-read fewer halo rows, but the
    // launch needs several waves of warps (148 SMs x 16 resident warps) to balance
    int64_t all_blocks = 0;
    for (int i = 0; i < count; i++) {
        const int64_t T = aid_num_frames(end_of[i] - begin_of[i]);
        if (T <= frame_limit) all_blocks += (T + AID_PEAK_BLOCK_FRAMES - 1) / AID_PEAK_BLOCK_FRAMES;
    }
    const int64_t run_blocks = std::max<int64_t>(1, std::min<int64_t>(AID_PEAK_RUN_BLOCKS, all_blocks / (8 * 148 * 16)));
    for (int i = 0; i < count; i++) {
        const int64_t begin = begin_of[i], end = end_of[i];
        if (end < begin) return AID_E_ARG;
        int64_t T = aid_num_frames(end - begin);
        if (T == 0) plan.host_status[i] |= AID_TRACK_EMPTY;
        if (T > frame_limit) { T = 0; plan.host_status[i] |= AID_TRACK_TOO_LONG; }
        plan.frames[i] = T;
        plan.frame_off[i + 1] = plan.frame_off[i] + T;
        for (int64_t f0 = 0; f0 < T; f0 += AID_STFT_UNIT_FRAMES) {
            aid_stft_unit u;
            u.pcm_begin = begin - base;
            u.n_samples = end - begin;
            u.spec_row = plan.frame_off[i] + f0;
            u.frame0 = (int32_t)f0;
            u.n_frames = (int32_t)std::min<int64_t>(AID_STFT_UNIT_FRAMES, T - f0);
            plan.sunits.push_back(u);
        }
        plan.first_punit[i] = (uint32_t)plan.punits.size();
        for (int64_t r0 = 0; r0 < T; r0 += AID_PEAK_BLOCK_FRAMES * run_blocks) {
            aid_peak_run run;
            const int64_t blocks_left = (T - r0 + AID_PEAK_BLOCK_FRAMES - 1) / AID_PEAK_BLOCK_FRAMES;
            run.n_blocks = (int32_t)std::min<int64_t>(run_blocks, blocks_left);
            run.first_unit = (int32_t)(plan.first_punit[i] + r0 / AID_PEAK_BLOCK_FRAMES);
            plan.pruns.push_back(run);
        }
        for (int64_t r0 = 0; r0 < T; r0 += AID_PEAK_BLOCK_FRAMES) {
            aid_peak_unit u;
            u.spec_row0 = plan.frame_off[i];
            u.track = i;
            u.n_frames = (int32_t)T;
            u.row0 = (int32_t)r0;
            u.n_rows = (int32_t)std::min<int64_t>(AID_PEAK_BLOCK_FRAMES, T - r0);
            plan.punits.push_back(u);
        }
    }
    plan.first_punit[count] = (uint32_t)plan.punits.size();
    // full-length runs first, the shorter tails of the tracks after them: a launch is a wave or two of warps (148 SMs x
    // 16-20 resident warps), so the last warps to start should be the short ones. The order of the run list decides only
    // which warp streams which run -- every run writes its own blocks' slot lists.
    {
        const int64_t full_rows = run_blocks * AID_PEAK_BLOCK_FRAMES;
        std::stable_partition(plan.pruns.begin(), plan.pruns.end(), [&](const aid_peak_run& r) {
            const aid_peak_unit& u = plan.punits[r.first_unit];
            return (int64_t)u.n_frames - u.row0 >= full_rows;
        });
    }
    plan.total_frames = plan.frame_off[count];
    plan.peak_cap = (int64_t)plan.punits.size() * AID_PEAK_BLOCK_CAP;
    plan.hash_cap = plan.peak_cap * AID_FANOUT;
    if (plan.peak_cap + 1 >= ((int64_t)1 << 32) || plan.hash_cap + 1 >= ((int64_t)1 << 32)) return AID_E_TOO_LONG;
    return AID_OK;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

int aid_slot_prepare(aid_engine* e, Slot& s, const Plan& plan, bool need_pcm, int64_t pcm_samples) {
    const size_t n = (size_t)plan.n_tracks, nu = plan.punits.size();
    if (need_pcm) AID_CUDA(e, s.pcm.ensure((size_t)std::max<int64_t>(pcm_samples, 1) * sizeof(float)));
    AID_CUDA(e, s.spec.ensure((size_t)std::max<int64_t>(plan.total_frames, 1) * AID_NBINS * sizeof(float)));
    AID_CUDA(e, s.gmax.ensure((size_t)std::max<int64_t>(plan.total_frames, 1) * 32 * sizeof(float)));
    AID_CUDA(e, s.slots.ensure((size_t)std::max<int64_t>(plan.peak_cap, 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.unit_pos.ensure((nu + 1) * 2 * sizeof(uint32_t)));           // counts, then positions
    AID_CUDA(e, s.peaks.ensure((size_t)(plan.peak_cap + 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.peak_track.ensure((size_t)(plan.peak_cap + 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.peak_off.ensure((n + 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.pos.ensure((size_t)(plan.peak_cap + 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.hash.ensure((size_t)std::max<int64_t>(plan.hash_cap, 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.t.ensure((size_t)std::max<int64_t>(plan.hash_cap, 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.hash_off.ensure((n + 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.status.ensure((n + 1) * sizeof(int32_t)));
    AID_CUDA(e, s.scan_tmp.ensure(aid_scan_tmp_elems(plan.peak_cap + 1) * sizeof(uint32_t)));
    AID_CUDA(e, s.misc.ensure(256));
    const size_t b0 = align256(plan.sunits.size() * sizeof(aid_stft_unit));
    const size_t b1 = align256(nu * sizeof(aid_peak_unit));
    const size_t b2 = align256((n + 1) * sizeof(uint32_t));
    const size_t b3 = align256(plan.pruns.size() * sizeof(aid_peak_run));
    AID_CUDA(e, s.desc.ensure(b0 + b1 + b2 + b3 + 256));
    AID_CUDA(e, s.h_desc.ensure(b0 + b1 + b2 + b3 + 256));
    AID_CUDA(e, s.h_small.ensure((n + 1) * (sizeof(uint32_t) + sizeof(int32_t)) + 1024));
    s.d_sunits = reinterpret_cast<aid_stft_unit*>(s.desc.as<char>());
    s.d_punits = reinterpret_cast<aid_peak_unit*>(s.desc.as<char>() + b0);
    s.d_first_punit = reinterpret_cast<uint32_t*>(s.desc.as<char>() + b0 + b1);
    s.d_pruns = reinterpret_cast<aid_peak_run*>(s.desc.as<char>() + b0 + b1 + b2);
    return AID_OK;
}

extern "C" int aid_engine_set_kernels(aid_engine* e, int stft_variant, int peak_summary) {
    if (!e || stft_variant < 0 || stft_variant > 31) return AID_E_ARG;
    e->stft_variant = stft_variant;
    e->peak_summary = peak_summary != 0;
    return AID_OK;
}

extern "C" int aid_engine_set_stage_timing(aid_engine* e, int on) {
    if (!e) return AID_E_ARG;
    e->timing = on != 0;
    return AID_OK;
}

extern "C" int aid_engine_stage_times(aid_engine* e, double* ms, int64_t* launches) {
    if (!e || !ms || !launches) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    for (auto& r : e->stage_recs) {
        float t = 0.0f;
        AID_CUDA(e, cudaEventSynchronize(r.b));
        AID_CUDA(e, cudaEventElapsedTime(&t, r.a, r.b));
        if (r.stage >= 0 && r.stage < 8) { ms[r.stage] += t; launches[r.stage] += 1; }
        e->event_pool.push_back(r.a); e->event_pool.push_back(r.b);
    }
    e->stage_recs.clear();
    return AID_OK;
}

// misc words: [0] n_peaks_total, [1] hash overflow flag, [2] n_hash_total
int aid_run_fingerprint(aid_engine* e, Slot& s, const Plan& plan, const float* d_pcm, cudaStream_t st) {
    const int n = plan.n_tracks;
    const int nsu = (int)plan.sunits.size(), npu = (int)plan.punits.size();
    // descriptors: one pinned staging copy, one H2D
    char* h = s.h_desc.as<char>();
    const size_t b0 = (size_t)((char*)s.d_punits - (char*)s.d_sunits);
    const size_t b1 = (size_t)((char*)s.d_first_punit - (char*)s.d_punits);
    const size_t b2 = (size_t)((char*)s.d_pruns - (char*)s.d_first_punit);
    const size_t b3 = plan.pruns.size() * sizeof(aid_peak_run);
    AID_CUDA(e, cudaEventSynchronize(s.done));      // the previous upload from this staging buffer has landed
    if (nsu) memcpy(h, plan.sunits.data(), (size_t)nsu * sizeof(aid_stft_unit));
    if (npu) memcpy(h + b0, plan.punits.data(), (size_t)npu * sizeof(aid_peak_unit));
    memcpy(h + b0 + b1, plan.first_punit.data(), (size_t)(n + 1) * sizeof(uint32_t));
    if (b3) memcpy(h + b0 + b1 + b2, plan.pruns.data(), b3);
    AID_CUDA(e, cudaMemcpyAsync(s.desc.p, h, b0 + b1 + b2 + b3, cudaMemcpyHostToDevice, st));
    AID_CUDA(e, cudaEventRecord(s.done, st));
    AID_CUDA(e, cudaMemsetAsync(s.status.p, 0, (size_t)(n + 1) * sizeof(int32_t), st));
    AID_CUDA(e, cudaMemsetAsync(s.misc.p, 0, 256, st));

    uint32_t* unit_cnt = s.unit_pos.as<uint32_t>();
    uint32_t* unit_pos = unit_cnt + (npu + 1);
    uint32_t* misc = s.misc.as<uint32_t>();
    float* gmax = e->peak_summary && e->stft_variant != 0 ? s.gmax.as<float>() : nullptr;
    { StageTimer tm(e, st, 0);
      AID_CUDA(e, aid_launch_stft_variant(e->stft_variant, e->tables, d_pcm, s.d_sunits, nsu, s.spec.as<float>(), gmax, st)); }
    { StageTimer tm(e, st, 1);
      AID_CUDA(e, aid_launch_peaks(s.spec.as<float>(), gmax, s.d_punits, s.d_pruns, (int)plan.pruns.size(),
                                   s.slots.as<uint32_t>(), unit_cnt, s.status.as<int32_t>(), st)); }
    // unit counts -> dense positions (unit_pos[npu] = total peaks)
    { StageTimer tm2(e, st, 2);
    AID_CUDA(e, cudaMemsetAsync(unit_cnt + npu, 0, sizeof(uint32_t), st));
    AID_CUDA(e, aid_launch_scan_u32(unit_cnt, unit_pos, npu + 1, s.scan_tmp.as<uint32_t>(), misc + 0, nullptr, st));
    AID_CUDA(e, aid_launch_peak_compact(s.slots.as<uint32_t>(), unit_cnt, unit_pos, s.d_punits, npu,
                                        s.peaks.as<uint32_t>(), s.peak_track.as<uint32_t>(), st));
    AID_CUDA(e, aid_launch_gather_u32(unit_pos, s.d_first_punit, s.peak_off.as<uint32_t>(), n + 1, st));
    }
    // per-anchor hash counts -> dense positions (pos[n_peaks] = total hashes)
    StageTimer tm3(e, st, 3);
    AID_CUDA(e, aid_launch_hash_count(s.peaks.as<uint32_t>(), s.peak_track.as<uint32_t>(), s.peak_off.as<uint32_t>(),
                                      misc + 0, plan.peak_cap, s.pos.as<uint32_t>(), st));
    AID_CUDA(e, aid_launch_scan_u32(s.pos.as<uint32_t>(), s.pos.as<uint32_t>(), plan.peak_cap + 1,
                                    s.scan_tmp.as<uint32_t>(), misc + 2, misc + 0, st));
    AID_CUDA(e, aid_launch_hash_write(s.peaks.as<uint32_t>(), s.peak_track.as<uint32_t>(), s.peak_off.as<uint32_t>(),
                                      misc + 0, plan.peak_cap, s.pos.as<uint32_t>(), s.hash.as<uint32_t>(),
                                      s.t.as<uint32_t>(), plan.hash_cap, reinterpret_cast<int32_t*>(misc + 1), st));
    AID_CUDA(e, aid_launch_gather_u32(s.pos.as<uint32_t>(), s.peak_off.as<uint32_t>(), s.hash_off.as<uint32_t>(),
                                      n + 1, st));
    e->launches += (nsu > 0) + (npu > 0) * 2 + 3 + 3 + 2 + 2;
    return AID_OK;
}

// ------------------------------------------------------------------------- fingerprint entry points
// d_pcm points at sample `base`; window i is samples [win_begin[i], win_end[i]) of that numbering (overlaps allowed)
static int fingerprint_windows_dev(aid_engine* e, const float* d_pcm, int64_t base, const int64_t* win_begin,
                                   const int64_t* win_end, int n_tracks, aid_fp_device_result* out, void* stream);
int aid_fingerprint_windows_core(aid_engine* e, const float* d_pcm, int64_t base, const int64_t* win_begin,
                                 const int64_t* win_end, int n, aid_fp_device_result* out, void* stream) {
    return fingerprint_windows_dev(e, d_pcm, base, win_begin, win_end, n, out, stream);
}
static int fingerprint_windows_dev(aid_engine* e, const float* d_pcm, int64_t base, const int64_t* win_begin,
                                   const int64_t* win_end, int n_tracks, aid_fp_device_result* out, void* stream) {
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    cudaStream_t st = stream ? (cudaStream_t)stream : s.st;
    Plan plan;
    int rc = aid_build_plan_windows(win_begin, win_end, n_tracks, base, AID_MAX_FRAMES, plan);
    if (rc != AID_OK) return rc;
    if ((rc = aid_slot_prepare(e, s, plan, false, 0)) != AID_OK) return rc;
    if ((rc = aid_run_fingerprint(e, s, plan, d_pcm, st)) != AID_OK) return rc;
    // host-known status bits (empty / too long) are OR-ed in on the device so d_status is complete
    bool any = false;
    for (int32_t v : plan.host_status) any |= v != 0;
    if (any) {
        int32_t* hs = s.h_small.as<int32_t>();
        AID_CUDA(e, cudaStreamSynchronize(st));
        AID_CUDA(e, cudaMemcpy(hs, s.status.p, (size_t)n_tracks * sizeof(int32_t), cudaMemcpyDeviceToHost));
        for (int i = 0; i < n_tracks; i++) hs[i] |= plan.host_status[i];
        AID_CUDA(e, cudaMemcpyAsync(s.status.p, hs, (size_t)n_tracks * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    }
    out->d_hash = s.hash.as<uint32_t>();
    out->d_t_anchor = s.t.as<uint32_t>();
    out->d_hash_off = s.hash_off.as<uint32_t>();
    out->d_peaks = s.peaks.as<uint32_t>();
    out->d_peak_off = s.peak_off.as<uint32_t>();
    out->d_status = s.status.as<int32_t>();
    out->d_spec = s.spec.as<float>();
    out->total_frames = plan.total_frames;
    return AID_OK;
}

extern "C" int aid_fingerprint_dev(aid_engine* e, const float* d_pcm, const int64_t* sample_off, int n_tracks,
                                   aid_fp_device_result* out, void* stream) {
    if (!e || !sample_off || !out || n_tracks < 0 || (!d_pcm && n_tracks > 0 && sample_off[n_tracks] > sample_off[0]))
        return AID_E_ARG;
    return fingerprint_windows_dev(e, d_pcm + sample_off[0], sample_off[0], sample_off, sample_off + 1, n_tracks, out, stream);
}

extern "C" int aid_fingerprint_windows_dev(aid_engine* e, const float* d_pcm, const int64_t* win_begin, const int64_t* win_end,
                                           int n_windows, aid_fp_device_result* out, void* stream) {
    if (!e || !out || n_windows < 0 || (n_windows > 0 && (!win_begin || !win_end || !d_pcm))) return AID_E_ARG;
    for (int i = 0; i < n_windows; i++) if (win_begin[i] < 0 || win_end[i] < win_begin[i]) return AID_E_ARG;
    return fingerprint_windows_dev(e, d_pcm, 0, win_begin, win_end, n_windows, out, stream);
}

namespace {
struct Pending {          // a sub-batch whose kernels are queued and whose results are not yet copied out
    bool active = false;
    int first = 0, count = 0;
    Plan plan;
};
}

// Copies one finished sub-batch out of slot s into the caller's dense arrays. hash_base is the
// caller-array position of the sub-batch's first hash; returns the new position via *hash_next.
static int collect_fingerprints(aid_engine* e, Slot& s, const Pending& pb, uint32_t* hash, uint32_t* t_anchor,
                                int64_t hash_cap, int64_t* hash_off, int32_t* status, int64_t hash_base,
                                int64_t* hash_next) {
    const int n = pb.count;
    uint32_t* h_off = s.h_small.as<uint32_t>();
    int32_t* h_st = reinterpret_cast<int32_t*>(h_off + (n + 1));
    AID_CUDA(e, cudaMemcpyAsync(h_off, s.hash_off.p, (size_t)(n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaMemcpyAsync(h_st, s.status.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    bool any_failed = false;
    for (int i = 0; i < n; i++) {
        h_st[i] |= pb.plan.host_status[i];
        any_failed |= (h_st[i] & (AID_TRACK_PEAK_OVERFLOW | AID_TRACK_TOO_LONG)) != 0;
    }
    int64_t pos = hash_base;
    if (!any_failed) {
        const int64_t total = (int64_t)h_off[n] - h_off[0];
        if (pos + total > hash_cap) return AID_E_CAPACITY;
        if (total > 0) {
            AID_CUDA(e, cudaMemcpyAsync(hash + pos, s.hash.as<uint32_t>() + h_off[0], (size_t)total * 4, cudaMemcpyDeviceToHost, s.st));
            AID_CUDA(e, cudaMemcpyAsync(t_anchor + pos, s.t.as<uint32_t>() + h_off[0], (size_t)total * 4, cudaMemcpyDeviceToHost, s.st));
        }
        for (int i = 0; i < n; i++) {
            hash_off[pb.first + i] = pos + ((int64_t)h_off[i] - h_off[0]);
            status[pb.first + i] = h_st[i];
        }
        pos += total;
    } else {
        for (int i = 0; i < n; i++) {
            hash_off[pb.first + i] = pos;
            status[pb.first + i] = h_st[i];
            if (h_st[i] & (AID_TRACK_PEAK_OVERFLOW | AID_TRACK_TOO_LONG)) continue;
            const int64_t cnt = (int64_t)h_off[i + 1] - h_off[i];
            if (pos + cnt > hash_cap) return AID_E_CAPACITY;
            if (cnt > 0) {
                AID_CUDA(e, cudaMemcpyAsync(hash + pos, s.hash.as<uint32_t>() + h_off[i], (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.st));
                AID_CUDA(e, cudaMemcpyAsync(t_anchor + pos, s.t.as<uint32_t>() + h_off[i], (size_t)cnt * 4, cudaMemcpyDeviceToHost, s.st));
            }
            pos += cnt;
        }
    }
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    *hash_next = pos;
    return AID_OK;
}

// number of tracks from `first` whose frames fit one sub-batch (at least one)
static int take_tracks(const int64_t* sample_off, int first, int n_tracks, int64_t max_frames) {
    int64_t frames = 0;
    int i = first;
    while (i < n_tracks) {
        const int64_t T = aid_num_frames(sample_off[i + 1] - sample_off[i]);
        if (i > first && frames + T > max_frames) break;
        frames += T;
        i++;
        if (i - first >= (1 << 20)) break;
    }
    return i - first;
}

extern "C" int aid_fingerprint_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks,
                                    uint32_t* hash, uint32_t* t_anchor, int64_t hash_cap,
                                    int64_t* hash_off, int32_t* status) {
    if (!e || !sample_off || !hash_off || !status || n_tracks < 0 || hash_cap < 0 || (hash_cap > 0 && (!hash || !t_anchor)))
        return AID_E_ARG;
    if (n_tracks > 0 && !pcm && sample_off[n_tracks] > sample_off[0]) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Pending pend[2];
    int64_t hash_pos = 0;
    int which = 0, rc = AID_OK;
    for (int first = 0; first < n_tracks && rc == AID_OK;) {
        const int count = take_tracks(sample_off, first, n_tracks, e->max_batch_frames);
        Slot& s = e->slot[which];
        Pending& pb = pend[which];
        pb.first = first; pb.count = count;
        if ((rc = aid_build_plan(sample_off, first, count, AID_MAX_FRAMES, pb.plan)) != AID_OK) break;
        const int64_t samples = sample_off[first + count] - sample_off[first];
        if ((rc = aid_slot_prepare(e, s, pb.plan, true, samples)) != AID_OK) break;
        if (samples > 0)
            AID_CUDA(e, cudaMemcpyAsync(s.pcm.p, pcm + sample_off[first], (size_t)samples * sizeof(float), cudaMemcpyHostToDevice, s.st));
        if ((rc = aid_run_fingerprint(e, s, pb.plan, s.pcm.as<float>(), s.st)) != AID_OK) break;
        pb.active = true;
        // while this sub-batch runs, drain the previous one from the other slot
        Pending& prev = pend[which ^ 1];
        if (prev.active) {
            rc = collect_fingerprints(e, e->slot[which ^ 1], prev, hash, t_anchor, hash_cap, hash_off, status, hash_pos, &hash_pos);
            prev.active = false;
        }
        first += count;
        which ^= 1;
    }
    // drain in submission order: the slot used before the last one first
    for (int k = 0; k < 2; k++) {
        Pending& p = pend[which ^ k];
        if (!p.active) continue;
        if (rc == AID_OK)
            rc = collect_fingerprints(e, e->slot[which ^ k], p, hash, t_anchor, hash_cap, hash_off, status, hash_pos, &hash_pos);
        else
            cudaStreamSynchronize(e->slot[which ^ k].st);
        p.active = false;
    }
    if (rc == AID_OK) hash_off[n_tracks] = hash_pos;
    return rc;
}

// ------------------------------------------------------------------------------ single-stage entries
extern "C" int aid_stft_host(aid_engine* e, const float* pcm, const int64_t* sample_off, int n_tracks, float* spec) {
    if (!e || !sample_off || n_tracks < 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    Plan plan;
    int rc = aid_build_plan(sample_off, 0, n_tracks, AID_MAX_FRAMES, plan);
    if (rc != AID_OK) return rc;
    const int64_t samples = n_tracks ? sample_off[n_tracks] - sample_off[0] : 0;
    if ((rc = aid_slot_prepare(e, s, plan, true, samples)) != AID_OK) return rc;
    if (plan.total_frames == 0) return AID_OK;
    if (!pcm || !spec) return AID_E_ARG;
    AID_CUDA(e, cudaMemcpyAsync(s.pcm.p, pcm + sample_off[0], (size_t)samples * sizeof(float), cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.desc.p, plan.sunits.data(), plan.sunits.size() * sizeof(aid_stft_unit), cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, aid_launch_stft_variant(e->stft_variant, e->tables, s.pcm.as<float>(), s.d_sunits, (int)plan.sunits.size(), s.spec.as<float>(), nullptr, s.st));
    e->launches += 1;
    AID_CUDA(e, cudaMemcpyAsync(spec, s.spec.p, (size_t)plan.total_frames * AID_NBINS * sizeof(float), cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    return AID_OK;
}

// builds a Plan from frame counts only (no PCM): used by the peaks-only entry
static int plan_from_frames(const int64_t* frame_off, int n_tracks, Plan& plan) {
    std::vector<int64_t> fake(n_tracks + 1, 0);
    for (int i = 0; i < n_tracks; i++) {
        const int64_t T = frame_off[i + 1] - frame_off[i];
        if (T < 0) return AID_E_ARG;
        fake[i + 1] = fake[i] + (T == 0 ? 0 : (T - 1) * AID_HOP + AID_NFFT);
    }
    return aid_build_plan(fake.data(), 0, n_tracks, AID_MAX_FRAMES, plan);
}

extern "C" int aid_peaks_host(aid_engine* e, const float* spec, const int64_t* frame_off, int n_tracks,
                              uint32_t* peaks, int64_t peak_cap, int64_t* peak_off, int32_t* status) {
    if (!e || !frame_off || !peak_off || !status || n_tracks < 0 || frame_off[0] != 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    Plan plan;
    int rc = plan_from_frames(frame_off, n_tracks, plan);
    if (rc != AID_OK) return rc;
    for (int i = 0; i < n_tracks; i++) if (plan.host_status[i] & AID_TRACK_TOO_LONG) return AID_E_TOO_LONG;
    if ((rc = aid_slot_prepare(e, s, plan, false, 0)) != AID_OK) return rc;
    const int npu = (int)plan.punits.size(), n = n_tracks;
    for (int i = 0; i <= n; i++) peak_off[i] = 0;
    for (int i = 0; i < n; i++) status[i] = plan.host_status[i];
    if (plan.total_frames == 0) return AID_OK;
    if (!spec) return AID_E_ARG;
    AID_CUDA(e, cudaMemcpyAsync(s.spec.p, spec, (size_t)plan.total_frames * AID_NBINS * sizeof(float), cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.d_punits, plan.punits.data(), (size_t)npu * sizeof(aid_peak_unit), cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.d_first_punit, plan.first_punit.data(), (size_t)(n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.d_pruns, plan.pruns.data(), plan.pruns.size() * sizeof(aid_peak_run), cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemsetAsync(s.status.p, 0, (size_t)(n + 1) * sizeof(int32_t), s.st));
    uint32_t* unit_cnt = s.unit_pos.as<uint32_t>();
    uint32_t* unit_pos = unit_cnt + (npu + 1);
    AID_CUDA(e, aid_launch_peaks(s.spec.as<float>(), nullptr, s.d_punits, s.d_pruns, (int)plan.pruns.size(), s.slots.as<uint32_t>(), unit_cnt, s.status.as<int32_t>(), s.st));
    AID_CUDA(e, cudaMemsetAsync(unit_cnt + npu, 0, sizeof(uint32_t), s.st));
    AID_CUDA(e, aid_launch_scan_u32(unit_cnt, unit_pos, npu + 1, s.scan_tmp.as<uint32_t>(), nullptr, nullptr, s.st));
    AID_CUDA(e, aid_launch_peak_compact(s.slots.as<uint32_t>(), unit_cnt, unit_pos, s.d_punits, npu, s.peaks.as<uint32_t>(), s.peak_track.as<uint32_t>(), s.st));
    AID_CUDA(e, aid_launch_gather_u32(unit_pos, s.d_first_punit, s.peak_off.as<uint32_t>(), n + 1, s.st));
    e->launches += 7;
    std::vector<uint32_t> h_off(n + 1);
    std::vector<int32_t> h_st(n);
    AID_CUDA(e, cudaMemcpyAsync(h_off.data(), s.peak_off.p, (size_t)(n + 1) * sizeof(uint32_t), cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaMemcpyAsync(h_st.data(), s.status.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    if ((int64_t)h_off[n] > peak_cap) return AID_E_CAPACITY;
    if (h_off[n] > 0) {
        if (!peaks) return AID_E_ARG;
        AID_CUDA(e, cudaMemcpy(peaks, s.peaks.p, (size_t)h_off[n] * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i <= n; i++) peak_off[i] = h_off[i];
    for (int i = 0; i < n; i++) status[i] |= h_st[i];
    return AID_OK;
}

extern "C" int aid_hashes_host(aid_engine* e, const uint32_t* peaks, const int64_t* peak_off, int n_tracks,
                               uint32_t* hash, uint32_t* t_anchor, int64_t hash_cap, int64_t* hash_off) {
    if (!e || !peak_off || !hash_off || n_tracks < 0 || peak_off[0] != 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[0];
    const int n = n_tracks;
    const int64_t np = peak_off[n];
    for (int i = 0; i <= n; i++) hash_off[i] = 0;
    if (np == 0) return AID_OK;
    if (!peaks || np + 1 >= ((int64_t)1 << 32) / AID_FANOUT) return AID_E_ARG;
    std::vector<uint32_t> off32(n + 1), track(np);
    for (int i = 0; i <= n; i++) off32[i] = (uint32_t)peak_off[i];
    for (int i = 0; i < n; i++) {
        if (peak_off[i + 1] < peak_off[i]) return AID_E_ARG;
        for (int64_t a = peak_off[i]; a < peak_off[i + 1]; a++) track[a] = (uint32_t)i;
    }
    const int64_t hcap = np * AID_FANOUT;
    AID_CUDA(e, s.peaks.ensure((size_t)(np + 1) * 4));
    AID_CUDA(e, s.peak_track.ensure((size_t)(np + 1) * 4));
    AID_CUDA(e, s.peak_off.ensure((size_t)(n + 1) * 4));
    AID_CUDA(e, s.pos.ensure((size_t)(np + 1) * 4));
    AID_CUDA(e, s.hash.ensure((size_t)hcap * 4));
    AID_CUDA(e, s.t.ensure((size_t)hcap * 4));
    AID_CUDA(e, s.hash_off.ensure((size_t)(n + 1) * 4));
    AID_CUDA(e, s.scan_tmp.ensure(aid_scan_tmp_elems(np + 1) * 4));
    AID_CUDA(e, s.misc.ensure(256));
    uint32_t* misc = s.misc.as<uint32_t>();
    const uint32_t np32 = (uint32_t)np;
    AID_CUDA(e, cudaMemsetAsync(s.misc.p, 0, 256, s.st));
    AID_CUDA(e, cudaMemcpyAsync(misc, &np32, 4, cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.peaks.p, peaks, (size_t)np * 4, cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.peak_track.p, track.data(), (size_t)np * 4, cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, cudaMemcpyAsync(s.peak_off.p, off32.data(), (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s.st));
    AID_CUDA(e, aid_launch_hash_count(s.peaks.as<uint32_t>(), s.peak_track.as<uint32_t>(), s.peak_off.as<uint32_t>(), misc, np, s.pos.as<uint32_t>(), s.st));
    AID_CUDA(e, aid_launch_scan_u32(s.pos.as<uint32_t>(), s.pos.as<uint32_t>(), np + 1, s.scan_tmp.as<uint32_t>(), misc + 2, misc, s.st));
    AID_CUDA(e, aid_launch_hash_write(s.peaks.as<uint32_t>(), s.peak_track.as<uint32_t>(), s.peak_off.as<uint32_t>(), misc, np, s.pos.as<uint32_t>(), s.hash.as<uint32_t>(), s.t.as<uint32_t>(), hcap, reinterpret_cast<int32_t*>(misc + 1), s.st));
    AID_CUDA(e, aid_launch_gather_u32(s.pos.as<uint32_t>(), s.peak_off.as<uint32_t>(), s.hash_off.as<uint32_t>(), n + 1, s.st));
    e->launches += 6;
    std::vector<uint32_t> h_off(n + 1);
    AID_CUDA(e, cudaMemcpyAsync(h_off.data(), s.hash_off.p, (size_t)(n + 1) * 4, cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    if ((int64_t)h_off[n] > hash_cap) return AID_E_CAPACITY;
    if (h_off[n] > 0) {
        if (!hash || !t_anchor) return AID_E_ARG;
        AID_CUDA(e, cudaMemcpy(hash, s.hash.p, (size_t)h_off[n] * 4, cudaMemcpyDeviceToHost));
        AID_CUDA(e, cudaMemcpy(t_anchor, s.t.p, (size_t)h_off[n] * 4, cudaMemcpyDeviceToHost));
    }
    for (int i = 0; i <= n; i++) hash_off[i] = h_off[i];
    return AID_OK;
}

// ------------------------------------------------------------------------------------------ helpers
extern "C" int aid_device_alloc(aid_engine* e, int64_t bytes, void** d_ptr) {
    if (!e || !d_ptr || bytes < 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    AID_CUDA(e, cudaMalloc(d_ptr, (size_t)std::max<int64_t>(bytes, 1)));
    return AID_OK;
}
extern "C" int aid_device_free(aid_engine* e, void* d_ptr) {
    if (!e) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    AID_CUDA(e, cudaFree(d_ptr));
    return AID_OK;
}
extern "C" int aid_copy_to_device(aid_engine* e, void* d_dst, const void* h_src, int64_t bytes) {
    if (!e || bytes < 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    AID_CUDA(e, cudaMemcpy(d_dst, h_src, (size_t)bytes, cudaMemcpyHostToDevice));
    return AID_OK;
}
extern "C" int aid_copy_to_host(aid_engine* e, void* h_dst, const void* d_src, int64_t bytes) {
    if (!e || bytes < 0) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    for (int i = 0; i < 2; i++) AID_CUDA(e, cudaStreamSynchronize(e->slot[i].st));
    AID_CUDA(e, cudaMemcpy(h_dst, d_src, (size_t)bytes, cudaMemcpyDeviceToHost));
    return AID_OK;
}
extern "C" int aid_synth_tracks_strided_dev(aid_engine* e, float* d_pcm, int64_t first_track, int64_t track_stride,
                                            int n_tracks, int64_t samples_per_track, uint64_t seed, void* stream) {
    if (!e || (!d_pcm && n_tracks > 0) || n_tracks < 0 || samples_per_track < 0 || track_stride < 1) return AID_E_ARG;
    AID_CUDA(e, cudaSetDevice(e->device));
    cudaStream_t st = stream ? (cudaStream_t)stream : e->slot[0].st;
    AID_CUDA(e, aid_launch_synth(d_pcm, first_track, n_tracks, samples_per_track, seed, st, track_stride));
    e->launches += (n_tracks + 32767) / 32768;
    return AID_OK;
}
extern "C" int aid_synth_tracks_dev(aid_engine* e, float* d_pcm, int64_t first_track, int n_tracks,
                                    int64_t samples_per_track, uint64_t seed, void* stream) {
    return aid_synth_tracks_strided_dev(e, d_pcm, first_track, 1, n_tracks, samples_per_track, seed, stream);
}
