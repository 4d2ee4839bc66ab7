"""Oracle (test infrastructure): STFT-magnitude / mel / log-mel features.

Restates /root/reference/src/easevoice/module/mel_processing.py:
  spectrogram_torch      :40-74   reflect-pad (n_fft-hop)/2, periodic Hann, rFFT,
                                   sqrt(re^2 + im^2 + 1e-6)
  spec_to_mel_torch      :77-90   mel_basis @ spec, log(clamp(., 1e-5))
  mel_spectrogram_torch  :93-142  the two above fused
and the third-party filterbank the reference calls at :82-84,106-108:
  librosa.filters.mel (librosa 0.9.2, pinned in the reference's uv.lock),
  Slaney mel scale + Slaney area normalisation, computed in float64 and cast to
  float32 (librosa absent from this image; the published algorithm is restated).

Two implementations are given on purpose: a torch one that mirrors the
reference call-for-call (torch.stft), and an independent numpy float64 one
(explicit framing + np.fft.rfft) used to judge which of two fp32 answers is
closer to the truth.
"""
import math
import numpy as np
import torch


# ----------------------------------------------------------------------------
# librosa.filters.mel restated (Slaney scale, norm="slaney", htk=False)
# ----------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-30) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr, n_fft, n_mels, fmin=0.0, fmax=None):
    """[n_mels, n_fft//2+1] float32, equals librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=)."""
    if fmax is None:
        fmax = sr / 2.0
    n_bins = n_fft // 2 + 1
    fft_f = np.linspace(0.0, sr / 2.0, n_bins)
    mel_pts = np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2)
    hz = _mel_to_hz(mel_pts)
    fdiff = np.diff(hz)
    ramps = hz[:, None] - fft_f[None, :]
    w = np.zeros((n_mels, n_bins), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        w[i] = np.maximum(0.0, np.minimum(lower, upper))
    enorm = 2.0 / (hz[2:n_mels + 2] - hz[:n_mels])
    w *= enorm[:, None]
    return w.astype(np.float32)


# ----------------------------------------------------------------------------
# torch restatement, call-for-call (fp32)
# ----------------------------------------------------------------------------
def spectrogram(y, n_fft, hop, win):
    """y [B, L] -> |X| [B, n_fft//2+1, T]; mel_processing.py:40-74."""
    pad = int((n_fft - hop) / 2)
    yp = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    window = torch.hann_window(win, dtype=y.dtype, device=y.device)
    X = torch.stft(yp, n_fft, hop_length=hop, win_length=win, window=window, center=False,
                   pad_mode="reflect", normalized=False, onesided=True, return_complex=True)
    X = torch.view_as_real(X)
    return torch.sqrt(X.pow(2).sum(-1) + 1e-6)


def spec_to_mel(spec, n_fft, n_mels, sr, fmin, fmax):
    """|X| [B, F, T] -> log-mel [B, n_mels, T]; mel_processing.py:77-90 (+:8-14)."""
    basis = torch.from_numpy(mel_filterbank(sr, n_fft, n_mels, fmin, fmax)).to(device=spec.device, dtype=spec.dtype)
    return torch.log(torch.clamp(torch.matmul(basis, spec), min=1e-5))


def mel_spectrogram(y, n_fft, n_mels, sr, hop, win, fmin, fmax):
    """y [B, L] -> log-mel [B, n_mels, T]; mel_processing.py:93-142."""
    return spec_to_mel(spectrogram(y, n_fft, hop, win), n_fft, n_mels, sr, fmin, fmax)


# ----------------------------------------------------------------------------
# independent float64 numpy restatement (truth estimate)
# ----------------------------------------------------------------------------
def spectrogram_f64(y, n_fft, hop, win):
    y = np.asarray(y, dtype=np.float64)
    pad = int((n_fft - hop) / 2)
    yp = np.pad(y, ((0, 0), (pad, pad)), mode="reflect")
    n = np.arange(win, dtype=np.float64)
    window = 0.5 - 0.5 * np.cos(2.0 * np.pi * n / win)  # periodic Hann == torch.hann_window
    T = (yp.shape[1] - n_fft) // hop + 1
    idx = np.arange(n_fft)[None, :] + hop * np.arange(T)[:, None]
    frames = yp[:, idx] * window[None, None, :]
    X = np.fft.rfft(frames, n=n_fft, axis=-1)           # [B, T, F]
    mag = np.sqrt(X.real ** 2 + X.imag ** 2 + 1e-6)
    return mag.transpose(0, 2, 1)                        # [B, F, T]


def mel_spectrogram_f64(y, n_fft, n_mels, sr, hop, win, fmin, fmax):
    spec = spectrogram_f64(y, n_fft, hop, win)
    basis = mel_filterbank(sr, n_fft, n_mels, fmin, fmax).astype(np.float64)
    return np.log(np.maximum(basis @ spec, 1e-5))


def kat_sines(sr=22050, seconds=1.0, n=8):
    """BASELINE.json config 1 input: y[b,k] = 0.5 sin(2 pi 220 (b+1) k / sr), f64 -> f32."""
    k = np.arange(int(sr * seconds), dtype=np.float64)
    y = np.stack([0.5 * np.sin(2 * np.pi * 220.0 * (b + 1) * k / sr) for b in range(n)])
    return torch.from_numpy(y.astype(np.float32))
