"""A minimal conv VAE with the interface the solvers use (reference models.py:301-355: cdim, zdim,
encoder.image_size, encode/decode/sample/forward, encoder returns mu/logvar as chunk views).  Test helper only."""
import torch
import torch.nn as nn


class _Encoder(nn.Module):
    def __init__(self, cdim, zdim, image_size):
        super().__init__()
        self.image_size = image_size
        self.conv = nn.Sequential(nn.Conv2d(cdim, 16, 5, 2, 2), nn.LeakyReLU(0.2), nn.Conv2d(16, 32, 3, 2, 1), nn.LeakyReLU(0.2))
        self.fc = nn.Linear(32 * (image_size // 4) ** 2, 2 * zdim)

    def forward(self, x):
        y = self.fc(self.conv(x).flatten(1))
        return y.chunk(2, dim=1)                       # pitch-2*zdim views, like models.py:242-244


class _Decoder(nn.Module):
    def __init__(self, cdim, zdim, image_size):
        super().__init__()
        self.s = image_size // 4
        self.fc = nn.Linear(zdim, 32 * self.s * self.s)
        self.deconv = nn.Sequential(nn.ReLU(), nn.ConvTranspose2d(32, 16, 4, 2, 1), nn.ReLU(), nn.ConvTranspose2d(16, cdim, 4, 2, 1))

    def forward(self, z):
        return self.deconv(self.fc(z).view(z.size(0), 32, self.s, self.s))


class TinyVAE(nn.Module):
    def __init__(self, cdim=3, zdim=32, image_size=16, reparameterize=None):
        super().__init__()
        self.cdim, self.zdim = cdim, zdim
        self.encoder = _Encoder(cdim, zdim, image_size)
        self.decoder = _Decoder(cdim, zdim, image_size)
        self._reparameterize = reparameterize

    def encode(self, x):
        return self.encoder(x)

    def decode(self, z):
        return self.decoder(z)

    def sample(self, z):
        return self.decode(z)

    def forward(self, x, deterministic=False):
        mu, logvar = self.encode(x)
        z = mu if deterministic else self._reparameterize(mu, logvar)
        return mu, logvar, z, self.decode(z)
