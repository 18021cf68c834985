#!/usr/bin/env python
"""Writes tests/golden/clip10_fingerprint.npz: this repo's own golden vectors for BASELINE.json configs[0]
(one synthetic 10 s clip, seed 42). The reference has none for these stages (parity
unpinned), so they pin the oracle against accidental change, not against the reference.

Produced by the independent numpy/scipy restatement (oracle/np_oracle.py), NOT by oracle/aid_oracle.c."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from audio_ident_b200 import synth  # noqa: E402
from oracle import np_oracle as npo  # noqa: E402

clip = synth.make_track(0, 10.0)
S = npo.stft(clip)
pk = npo.peaks(S)
h, t = npo.hashes(pk)
rows = np.arange(0, S.shape[0], 97)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "clip10_fingerprint.npz"),
                    pcm_head=clip[:64], pcm_sum=np.float64(clip.astype(np.float64).sum()), n_samples=len(clip),
                    spec_rows=rows, spec_values=S[rows], spec_sum=np.float64(S.astype(np.float64).sum()),
                    peaks=pk, hash=h, t_anchor=t)
print("frames", S.shape[0], "peaks", len(pk), "hashes", len(h))
