"""Drop-in for ddsp/loss.py (`train.py:48`, `solver.py:112`): single- and random-scale spectral losses in stock
torch ops (torch.stft; no torchaudio needed).  Not on the synthesizer hot path -- provided so that
`solver.train` runs end to end with the ddsp_b200 modules.

SSSLoss (loss.py:8-25): magnitude spectrogram with a periodic Hann window of n_fft samples, hop n_fft*(1-overlap),
no centring, normalised by the window's L2 norm (`torchaudio.transforms.Spectrogram(power=1, normalized=True,
center=False)`); loss = mean_b ||S_t - S_p||_F / ||S_t + S_p||_F  +  alpha * mean |log S_t - log S_p|.
RSSLoss (loss.py:28-43): the mean of n_scale SSSLosses with n_fft drawn uniformly from [fft_min, fft_max).
"""
import torch
import torch.nn.functional as F
from torch import nn


class SSSLoss(nn.Module):
    def __init__(self, n_fft=111, alpha=1.0, overlap=0, eps=1e-7):
        super().__init__()
        self.n_fft, self.alpha, self.eps = int(n_fft), alpha, eps
        self.hop = int(n_fft * (1 - overlap))
        self.register_buffer('window', torch.hann_window(self.n_fft), persistent=False)

    def spec(self, x):
        w = self.window.to(x.dtype)
        s = torch.stft(x, self.n_fft, hop_length=self.hop, win_length=self.n_fft, window=w, center=False,
                       normalized=False, onesided=True, return_complex=True)
        return s.abs() / w.pow(2).sum().sqrt()

    def forward(self, x_true, x_pred):
        s_true = self.spec(x_true) + self.eps
        s_pred = self.spec(x_pred) + self.eps
        converge = torch.mean(torch.linalg.norm(s_true - s_pred, dim=(1, 2)) / torch.linalg.norm(s_true + s_pred, dim=(1, 2)))
        return converge + self.alpha * F.l1_loss(s_true.log(), s_pred.log())


class RSSLoss(nn.Module):
    def __init__(self, fft_min, fft_max, n_scale, alpha=1.0, overlap=0, eps=1e-7, device='cuda'):
        super().__init__()
        self.fft_min, self.fft_max, self.n_scale = fft_min, fft_max, n_scale
        self.alpha, self.overlap, self.eps, self.device = alpha, overlap, eps, device
        self.lossdict = {}                       # built lazily: the reference constructs all fft_max - fft_min modules up front

    def _loss(self, n_fft):
        if n_fft not in self.lossdict:
            self.lossdict[n_fft] = SSSLoss(n_fft, self.alpha, self.overlap, self.eps).to(self.device)
        return self.lossdict[n_fft]

    def to(self, *args, **kwargs):
        dev = args[0] if args and not isinstance(args[0], torch.dtype) else kwargs.get('device')
        if dev is not None:
            self.device = dev
            for m in self.lossdict.values():
                m.to(dev)
        return super().to(*args, **kwargs)

    def forward(self, x_pred, x_true):
        value = 0.
        n_ffts = torch.randint(self.fft_min, self.fft_max, (self.n_scale,))     # same draw as the reference (CPU generator)
        for n_fft in n_ffts:
            value = value + self._loss(int(n_fft))(x_true, x_pred)
        return value / self.n_scale
