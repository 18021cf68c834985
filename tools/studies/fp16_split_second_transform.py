#!/usr/bin/env python
"""Numerical feasibility of running the SECOND 32-point transform of k_stft on tensor cores.

The 1024-point transform of a frame pair z = a + i b is factored 32 x 32 (csrc/stft.cu): a first transform over n2 in
FP32, then Z[k] = sum_{n1 < 32} W_1024^(n1 k) Y[n1][k mod 32]. For a fixed k1 = k mod 32 that is a 32 x 32 complex
matrix times a vector, and over many frame pairs a GEMM. Tensor cores take FP16 / BF16 / TF32 operands; this script
emulates, in numpy, FP16 operands split in two (x = hi + lo, lo scaled by 2^11 so that it stays normal), three products
per real product (hi*hi + hi*lo + lo*hi, every product exact in FP32, FP32 accumulation), and compares the stored value
S = log(1 + |X|^2) with the double-precision oracle under the repo's tolerance |dS| <= 1e-4 * max(|S|, 1)
(include/aid_params.h AID_SPEC_TOL). It also runs the unsplit variants to show why they are not enough.

Usage: python tools/studies/fp16_split_second_transform.py [n_tracks] [seconds]
CPU only; not part of the product or of the tests.
"""
from __future__ import annotations

import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from audio_ident_b200 import synth  # noqa: E402

N, HOP, TOL = 1024, 128, 1e-4


def window() -> np.ndarray:
    n = np.arange(N, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * np.pi * n / (N - 1))).astype(np.float32)


def split16(x: np.ndarray):
    hi = x.astype(np.float16)
    lo = ((x - hi.astype(np.float32)) * np.float32(2048.0)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def round_to(x: np.ndarray, kind: str) -> np.ndarray:
    if kind == "fp16":
        return x.astype(np.float16).astype(np.float32)
    if kind in ("bf16", "tf32"):
        keep = 16 if kind == "bf16" else 19                  # bits kept of the 32 (sign + 8 exp + 7 / 10 mantissa)
        u = x.astype(np.float32).view(np.uint32).astype(np.uint64)
        drop = 32 - keep
        u = ((u + (1 << (drop - 1))) >> drop) << drop          # round to nearest (ties up: good enough here)
        return u.astype(np.uint32).view(np.float32)
    raise ValueError(kind)


def real_gemm(a: np.ndarray, b: np.ndarray, mode: str) -> np.ndarray:
    """a [M, K] constants, b [K, P] data, both float32; returns float32 [M, P] as the tensor core would."""
    if mode == "fp32":
        return (a.astype(np.float32) @ b.astype(np.float32)).astype(np.float32)
    if mode == "fp16x3":
        ah, al = split16(a); bh, bl = split16(b)
        s = np.float32(1.0 / 2048.0)
        return (ah @ bh + (ah @ bl) * s + (al @ bh) * s).astype(np.float32)
    if mode == "fp16x3u":            # the same without scaling lo: small lo parts fall into FP16 subnormals (absolute step 6e-8),
        ah = a.astype(np.float16).astype(np.float32); al = (a - ah).astype(np.float16).astype(np.float32)   # which is what three
        bh = b.astype(np.float16).astype(np.float32); bl = (b - bh).astype(np.float16).astype(np.float32)   # MMAs into one accumulator see
        return (ah @ bh + ah @ bl + al @ bh).astype(np.float32)
    return (round_to(a, mode) @ round_to(b, mode)).astype(np.float32)


def second_transform(Y: np.ndarray, mode: str) -> np.ndarray:
    """Y [32 n1, 32 k1, P] complex64 -> Z [1024, P] complex64 with Z[k1 + 32 k2] = sum_n1 W^(n1 k) Y[n1, k1]."""
    P = Y.shape[2]
    Z = np.empty((N, P), np.complex64)
    n1 = np.arange(32)
    for k1 in range(32):
        k = k1 + 32 * np.arange(32)
        F = np.exp(-2j * np.pi * np.outer(k, n1) / N)          # [32 k2, 32 n1], double
        Fr, Fi = F.real.astype(np.float32), F.imag.astype(np.float32)
        yr, yi = np.ascontiguousarray(Y[:, k1, :].real), np.ascontiguousarray(Y[:, k1, :].imag)
        zr = real_gemm(Fr, yr, mode) - real_gemm(Fi, yi, mode)
        zi = real_gemm(Fr, yi, mode) + real_gemm(Fi, yr, mode)
        Z[k] = zr + 1j * zi
    return Z


def both_transforms(zz: np.ndarray, mode: str) -> np.ndarray:
    """zz [P, n2, n1] complex64 -> Z [1024, P]: BOTH 32-point transforms as products with the one constant matrix W_32
    (64 x 64 as a real matrix), the inter-transform twiddle W_1024^(n1 k1) applied elementwise in FP32 in between."""
    P = zz.shape[0]
    W = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    Wr, Wi = W.real.astype(np.float32), W.imag.astype(np.float32)
    # first: Y[k1, (n1, p)] = sum_n2 W[k1, n2] z[n2, (n1, p)]
    d = np.ascontiguousarray(np.transpose(zz, (1, 2, 0))).reshape(32, 32 * P)         # [n2, n1 * P]
    dr, di = np.ascontiguousarray(d.real), np.ascontiguousarray(d.imag)
    yr = real_gemm(Wr, dr, mode) - real_gemm(Wi, di, mode)
    yi = real_gemm(Wr, di, mode) + real_gemm(Wi, dr, mode)
    Y = (yr + 1j * yi).astype(np.complex64).reshape(32, 32, P)                        # [k1, n1, p]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / N).astype(np.complex64)   # [k1, n1]
    Y = (Y * tw[:, :, None]).astype(np.complex64)
    # second: Z[k1 + 32 k2, p] = sum_n1 W[k2, n1] Y[k1, n1, p]
    e = np.ascontiguousarray(np.transpose(Y, (1, 0, 2))).reshape(32, 32 * P)           # [n1, k1 * P]
    er, ei = np.ascontiguousarray(e.real), np.ascontiguousarray(e.imag)
    zr = real_gemm(Wr, er, mode) - real_gemm(Wi, ei, mode)
    zi = real_gemm(Wr, ei, mode) + real_gemm(Wi, er, mode)
    Zk = (zr + 1j * zi).astype(np.complex64).reshape(32, 32, P)                       # [k2, k1, p]
    return Zk.reshape(N, P)                                                           # k = k1 + 32 k2


def raw_sample_transforms(x: np.ndarray, T: int, w: np.ndarray, mode: str) -> np.ndarray:
    """The sliding-operand variant: the window is folded into 32 first-stage matrices
    G_n1[k1][n2] = w[n1 + 32 n2] W_32^(n2 k1) and the data operand is the RAW sample sequence (split into FP16 hi + lo once
    per sample, not per frame). Returns Z [1024, T / 2] for the packed frame pairs, second stage as in both_transforms."""
    n2 = np.arange(32)
    Y = np.empty((T, 32, 32), np.complex64)                                   # [t, k1, n1]
    for n1 in range(32):
        G = (w[n1 + 32 * n2].astype(np.float64)[None, :] * np.exp(-2j * np.pi * np.outer(np.arange(32), n2) / 32))   # [k1, n2]
        Gr, Gi = G.real.astype(np.float32), G.imag.astype(np.float32)
        seq = x[n1::32]                                                       # x[n1 + 32 j]
        idx = n2[:, None] + 4 * np.arange(T)[None, :]                         # j = n2 + 4 t
        d = np.ascontiguousarray(seq[idx]).astype(np.float32)                 # [n2, t]  (the Hankel operand)
        Y[:, :, n1] = (real_gemm(Gr, d, mode) + 1j * real_gemm(Gi, d, mode)).T.astype(np.complex64)
    z = (0.5 * (Y[0::2] + 1j * Y[1::2])).astype(np.complex64)                 # pack frame pairs: [P, k1, n1]
    P = z.shape[0]
    tw = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / N).astype(np.complex64)
    z = (z * tw[None, :, :]).astype(np.complex64)
    W = np.exp(-2j * np.pi * np.outer(np.arange(32), np.arange(32)) / 32)
    Wr, Wi = W.real.astype(np.float32), W.imag.astype(np.float32)
    e = np.ascontiguousarray(np.transpose(z, (2, 1, 0))).reshape(32, 32 * P)   # [n1, k1 * P]
    er, ei = np.ascontiguousarray(e.real), np.ascontiguousarray(e.imag)
    zr = real_gemm(Wr, er, mode) - real_gemm(Wi, ei, mode)
    zi = real_gemm(Wr, ei, mode) + real_gemm(Wi, er, mode)
    return (zr + 1j * zi).astype(np.complex64).reshape(32, 32, P).reshape(N, P)


def study(n_tracks: int, seconds: float) -> None:
    w = window()
    modes = ["fp32", "fp16x3", "tf32", "fp16", "bf16", "both:fp16x3", "both:fp16x3u", "both:tf32", "raw:fp16x3u", "raw:fp32"]
    worst = {m: 0.0 for m in modes}
    bad = {m: 0 for m in modes}
    total = 0
    for tr in range(n_tracks):
        x = synth.make_track(1000 + tr, seconds)
        T = (len(x) - N) // HOP + 1
        T -= T % 2
        idx = np.arange(N)[None, :] + HOP * np.arange(T)[:, None]
        fr = x[idx] * w[None, :]                                  # [T, 1024] float32
        ref = np.fft.fft(fr.astype(np.float64), axis=1)[:, :512]
        S_ref = np.log1p(ref.real ** 2 + ref.imag ** 2)
        z = (0.5 * (fr[0::2] + 1j * fr[1::2])).astype(np.complex64)   # pairs, pre-scaled as the kernel does; [P, 1024]
        P = z.shape[0]
        zz = z.reshape(P, 32, 32)                                 # [P, n2, n1]  (n = n1 + 32 n2)
        # first transform over n2 in FP32 (numpy computes it in double; rounded to complex64 once, like the kernel's registers)
        Y = np.fft.fft(zz.astype(np.complex128), axis=1).astype(np.complex64)   # [P, k1, n1]
        Y = np.ascontiguousarray(np.transpose(Y, (2, 1, 0)))      # [n1, k1, P]
        for m in modes:
            if m.startswith("raw:"):
                Z = raw_sample_transforms(x, T, w, m[4:]).T
            else:
                Z = (both_transforms(zz, m[5:]) if m.startswith("both:") else second_transform(Y, m)).T   # [P, 1024]
            Zm = np.conj(np.roll(Z[:, ::-1], 1, axis=1))          # conj(Z[N - k])
            Xa = (Z + Zm)[:, :512]
            Xb = (-1j * (Z - Zm))[:, :512]
            S = np.empty((T, 512), np.float64)
            S[0::2] = np.log1p(Xa.real.astype(np.float64) ** 2 + Xa.imag.astype(np.float64) ** 2)
            S[1::2] = np.log1p(Xb.real.astype(np.float64) ** 2 + Xb.imag.astype(np.float64) ** 2)
            err = np.abs(S - S_ref) / np.maximum(np.abs(S_ref), 1.0)
            worst[m] = max(worst[m], float(err.max()))
            bad[m] += int((err > TOL).sum())
        total += T * 512
    print(f"{n_tracks} tracks x {seconds:g} s = {total} stored values; tolerance {TOL:g} * max(|S|, 1)")
    for m in modes:
        what = ("both transforms in " + m[5:] if m.startswith("both:") else
                "raw samples x window-folded matrices, " + m[4:] if m.startswith("raw:") else "second transform in " + m)
        print(f"  {what:52s}: worst scaled error {worst[m]:.3g}, values out of tolerance {bad[m]}")


if __name__ == "__main__":
    study(int(sys.argv[1]) if len(sys.argv) > 1 else 4, float(sys.argv[2]) if len(sys.argv) > 2 else 10.0)
