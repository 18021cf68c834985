"""Host logic of the STFT kernel's constant tables (csrc/common.cuh aid_fill_stft_tables): the shared-memory image the packed
kernel copies with 128-bit loads must hold exactly the values of the base tables -- window / 2 lane-major, and the folded
twiddles regrouped two butterflies per quad (stage 1 as (c, -s, s, c)) -- and the base tables must be the double-precision
formulas rounded to float once. Compiled with g++ from the product header; no GPU involved."""
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CUDA_INC = "/usr/local/cuda/include"

SRC = textwrap.dedent(r'''
    #include <cstdio>
    #include <cmath>
    #include <vector>
    #include "audio_ident_b200/csrc/common.cuh"
    int main() {
        std::vector<float> win(AID_NFFT), tw(AID_TWIST_FLOATS);
        aid_fill_stft_tables(win.data(), tw.data());
        const double two_pi = 6.283185307179586476925286766559;
        int bad = 0;
        for (int i = 0; i < AID_NFFT; i++)
            bad += win[i] != (float)(AID_WIN_A0 - AID_WIN_A1 * cos(two_pi * i / (double)(AID_NFFT - 1)));
        const float* iw = tw.data() + AID_TWIST_IMAGE;
        const float* it = iw + 32 * 36;
        for (int i = 0; i < AID_NFFT; i++) bad += iw[(i & 31) * 36 + (i >> 5)] != 0.5f * win[i];
        for (int lane = 0; lane < 32; lane++) for (int j = 32; j < 36; j++) bad += iw[lane * 36 + j] != 0.0f;
        for (int k1 = 0; k1 < 32; k1++) {
            // folded twiddle i of lane k1: stage S (half = 2^S), k < max(half / 2, 1): w = g^(16 >> S) W_(2 half)^k, g = W_1024^k1
            float c[16], s[16]; int i = 0;
            for (int S = 0; S < 5; S++) {
                const int half = 1 << S, nk = half >= 2 ? half / 2 : 1;
                for (int k = 0; k < nk; k++, i++) {
                    const double a = two_pi * ((double)(k1 * (16 >> S)) / 1024.0 + (double)k / (double)(2 * half));
                    c[i] = (float)cos(a); s[i] = (float)sin(a);
                    bad += tw[k1 * 32 + 2 * i] != c[i] || tw[k1 * 32 + 2 * i + 1] != s[i];
                }
            }
            const float* o = it + k1 * 36;
            bad += o[0] != c[1] || o[1] != -s[1] || o[2] != s[1] || o[3] != c[1];          // stage 1: halves k and k + 1 = -i w
            for (int e = 1; e < 8; e++)                                                    // stages 2-4: (c_k, c_k+1, s_k, s_k+1)
                bad += o[4 * e] != c[2 * e] || o[4 * e + 1] != c[2 * e + 1] || o[4 * e + 2] != s[2 * e] || o[4 * e + 3] != s[2 * e + 1];
            bad += o[32] != c[0] || o[33] != s[0] || o[34] != 0.0f || o[35] != 0.0f;       // stage 0
            for (int n1 = 0; n1 < 32; n1++) {
                const double a = two_pi * (double)(k1 * n1) / 1024.0;
                bad += tw[AID_TWIST_FOLDED + k1 * 64 + 2 * n1] != (float)cos(a) || tw[AID_TWIST_FOLDED + k1 * 64 + 2 * n1 + 1] != (float)sin(a);
            }
        }
        printf("%d\n", bad);
        return bad != 0;
    }
''')


@pytest.mark.skipif(not os.path.isdir(CUDA_INC), reason="CUDA headers not installed")
def test_stft_table_image_matches_the_base_tables(tmp_path):
    src = tmp_path / "tables.cpp"
    src.write_text(SRC)
    exe = tmp_path / "tables"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I", ROOT, "-I", CUDA_INC, "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and out.stdout.strip() == "0", out.stdout + out.stderr
