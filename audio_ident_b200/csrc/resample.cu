// resample.cu -- 48 kHz -> 16 kHz polyphase decimator (factor 3), sm_100a.
//
// SURVEY.md section 8(f)-2, the decode feed: the reference's decode_dual_rate starts TWO ffmpeg children per file, one
// per output rate (reference audio-ident-service/app/audio/decode.py:74-87, :37-58), and once fingerprinting is fast
// those two processes are the ingest wall time. With this kernel one 48 kHz decode is enough: the 16 kHz stream the
// fingerprint path needs is derived on the GPU, where the PCM is going anyway.
//
// Definition (oracle: scipy.signal.resample_poly(x, 1, 3) -- its default design, restated in oracle/np_oracle.py):
//   h = 61-tap linear-phase low-pass, firwin(61, 1/3, window = ("kaiser", 5.0)), unity gain at DC, computed in double
//       and rounded to float32 once (aid_resample_taps);
//   y[j] = sum_{m=-30..30} h[m + 30] * x[3 j + m],  x = 0 outside [0, n),  j < ceil(n / 3)
// i.e. zero phase: output sample j sits on input sample 3 j. PARITY UNPINNED against the reference: ffmpeg's own
// resampler (libswresample) uses a different filter, and there is no ffmpeg in this image to record vectors from.
//
// Bound: HBM (12 B read + 4 B written per output sample; 61 FMA per output is far below the FP32 ridge). A CTA stages
// 3 * 256 + 60 inputs in shared memory with coalesced loads; every thread then reads its 61-sample window at stride 3
// (conflict-free: 3 is coprime to the 32 banks).
#include "engine.h"

namespace {

constexpr int kTaps = 61, kHalf = 30, kOutPerCta = 256, kIn = 3 * kOutPerCta + 2 * kHalf;
__constant__ float c_taps[kTaps];

double bessel_i0(double x) {
    double sum = 1.0, term = 1.0;
    for (int k = 1; k < 64; k++) { term *= (x / (2.0 * k)) * (x / (2.0 * k)); sum += term; if (term < 1e-18 * sum) break; }
    return sum;
}

__global__ void __launch_bounds__(kOutPerCta)
k_resample3(const float* __restrict__ x, int64_t n_in, float* __restrict__ y, int64_t n_out) {
    __shared__ float s[kIn];
    const int64_t j0 = (int64_t)blockIdx.x * kOutPerCta;
    const int64_t i0 = 3 * j0 - kHalf;
    for (int i = threadIdx.x; i < kIn; i += kOutPerCta) {
        const int64_t g = i0 + i;
        s[i] = g >= 0 && g < n_in ? x[g] : 0.0f;
    }
    __syncthreads();
    const int64_t j = j0 + threadIdx.x;
    if (j >= n_out) return;
    const float* w = s + 3 * threadIdx.x;
    float acc = 0.0f;
#pragma unroll
    for (int m = 0; m < kTaps; m++) acc = fmaf(c_taps[m], w[m], acc);
    y[j] = acc;
}

}  // namespace

extern "C" int64_t aid_resample_out_len(int64_t n_in) { return n_in <= 0 ? 0 : (n_in + 2) / 3; }

extern "C" void aid_resample_taps(float* taps) {
    // scipy.signal.firwin(61, 1/3, window=("kaiser", 5.0)): cutoff relative to Nyquist, scaled to unity gain at DC
    const double cutoff = 1.0 / 3.0, beta = 5.0, alpha = 0.5 * (kTaps - 1);
    double h[kTaps], sum = 0.0;
    for (int n = 0; n < kTaps; n++) {
        const double m = n - alpha, a = M_PI * cutoff * m;
        const double sinc = m == 0.0 ? 1.0 : sin(a) / a;
        const double r = m / alpha;
        h[n] = cutoff * sinc * bessel_i0(beta * sqrt(1.0 - r * r)) / bessel_i0(beta);
        sum += h[n];
    }
    for (int n = 0; n < kTaps; n++) taps[n] = (float)(h[n] / sum);
}

static int upload_taps(aid_engine* e) {
    static int uploaded_for = -1;                     // per device (constant memory is per context)
    if (uploaded_for == e->device) return AID_OK;
    float taps[kTaps];
    aid_resample_taps(taps);
    AID_CUDA(e, cudaMemcpyToSymbol(c_taps, taps, sizeof taps));
    uploaded_for = e->device;
    return AID_OK;
}

extern "C" int aid_resample_48k_to_16k_dev(aid_engine* e, const float* d_in, int64_t n_in, float* d_out, void* stream) {
    if (!e || n_in < 0 || (n_in > 0 && (!d_in || !d_out))) return AID_E_ARG;
    if (n_in == 0) return AID_OK;
    AID_CUDA(e, cudaSetDevice(e->device));
    int rc = upload_taps(e);
    if (rc) return rc;
    const int64_t n_out = aid_resample_out_len(n_in);
    cudaStream_t st = stream ? (cudaStream_t)stream : e->slot[0].st;
    k_resample3<<<(unsigned)((n_out + kOutPerCta - 1) / kOutPerCta), kOutPerCta, 0, st>>>(d_in, n_in, d_out, n_out);
    AID_CUDA(e, cudaGetLastError());
    e->launches += 1;
    return AID_OK;
}

extern "C" int aid_resample_48k_to_16k_host(aid_engine* e, const float* pcm48, int64_t n_in, float* pcm16) {
    if (!e || n_in < 0 || (n_in > 0 && (!pcm48 || !pcm16))) return AID_E_ARG;
    if (n_in == 0) return AID_OK;
    AID_CUDA(e, cudaSetDevice(e->device));
    Slot& s = e->slot[1];                              // the second slot: an ingest on slot 0 is not disturbed
    const int64_t n_out = aid_resample_out_len(n_in);
    AID_CUDA(e, s.pcm.ensure((size_t)n_in * sizeof(float)));
    AID_CUDA(e, s.spec.ensure((size_t)n_out * sizeof(float)));
    AID_CUDA(e, cudaMemcpyAsync(s.pcm.p, pcm48, (size_t)n_in * sizeof(float), cudaMemcpyHostToDevice, s.st));
    int rc = aid_resample_48k_to_16k_dev(e, s.pcm.as<float>(), n_in, s.spec.as<float>(), s.st);
    if (rc) return rc;
    AID_CUDA(e, cudaMemcpyAsync(pcm16, s.spec.p, (size_t)n_out * sizeof(float), cudaMemcpyDeviceToHost, s.st));
    AID_CUDA(e, cudaStreamSynchronize(s.st));
    return AID_OK;
}
