"""The index arithmetic of the general STFT kernels (csrc/stft_core.cuh: load/window, Stockham radix-4/2 passes, real-FFT
untangle, adjoint packing) compiled for the HOST and checked against numpy for every transform size the library accepts.
The CUDA kernels in csrc/stft.cu call exactly these functions (one thread block per frame)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("stft") / "libstft_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tools", "exp", "stft_host_test.cpp")], check=True)
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("N,win", [(2048, 2048), (4096, 4096), (2048, 1024), (2048, 256), (1024, 1024), (512, 512), (256, 256), (4096, 512)])
def test_forward_and_adjoint_match_numpy(host_lib, N, win):
    rng = np.random.default_rng(N + win)
    L = 9000
    wav = (rng.random(L) - 0.5).astype(np.float32)
    for s0 in (-N // 2, 1000, L - N + 300):                     # left reflection, interior, right reflection
        X = np.zeros((N // 2 + 1) * 2, np.float32)
        host_lib.stft_frame_fwd(_p(wav), L, s0, N, win, 256, _p(X))
        idx = np.abs(s0 + np.arange(N))
        idx = np.where(idx >= L, 2 * (L - 1) - idx, idx)
        j = np.arange(N) - (N - win) // 2
        w = np.where((j >= 0) & (j < win), 0.5 - 0.5 * np.cos(2 * np.pi * j / win), 0.0)     # torch.hann_window(win), centred
        ref = np.fft.rfft(wav[idx].astype(np.float64) * w)
        got = X[0::2] + 1j * X[1::2]
        assert np.abs(got - ref).max() / np.abs(ref).max() < 2e-6
        G = rng.standard_normal((N // 2 + 1) * 2).astype(np.float32)
        S = np.zeros(N, np.float32)
        host_lib.stft_frame_adj(_p(G), N, 256, _p(S))
        Gc = G[0::2] + 1j * G[1::2]
        k, n = np.arange(N // 2 + 1)[:, None], np.arange(N)[None, :]
        Sref = (Gc[:, None] * np.exp(2j * np.pi * k * n / N)).real.sum(0)       # d/dx[n] of <G, rfft(x)> (real inner product)
        assert np.abs(S - Sref).max() / np.abs(Sref).max() < 2e-6
