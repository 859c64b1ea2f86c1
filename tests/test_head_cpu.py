"""intro_tc_vae_b200.head: dense mu / logvar from the encoder's fc parameters == fc(y).chunk(2, dim=1) (models.py:240-244),
values and gradients; runs on the CPU (plain torch GEMMs)."""
import torch

from intro_tc_vae_b200 import head


def test_split_head_equals_linear_chunk():
    torch.manual_seed(0)
    fc = torch.nn.Linear(24, 2 * 8)
    y0 = torch.randn(5, 24)
    w1, w2 = torch.randn(5, 8), torch.randn(5, 8)
    y = y0.clone().requires_grad_(True)
    mu_r, lv_r = fc(y).chunk(2, dim=1)
    ((mu_r * w1).sum() + (lv_r * w2).sum()).backward()
    ref = (mu_r.detach(), lv_r.detach(), y.grad.clone(), fc.weight.grad.clone(), fc.bias.grad.clone())
    fc.zero_grad()
    y = y0.clone().requires_grad_(True)
    mu, lv = head.split_head(y, fc)
    assert mu.is_contiguous() and lv.is_contiguous() and mu.shape == (5, 8)
    ((mu * w1).sum() + (lv * w2).sum()).backward()
    for got, want in zip((mu.detach(), lv.detach(), y.grad, fc.weight.grad, fc.bias.grad), ref):
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-6)
    buf = torch.empty(5, 8)
    mu2, _ = head.split_head(y0, fc, mu_out=buf)
    assert mu2.data_ptr() == buf.data_ptr()
    torch.testing.assert_close(buf, ref[0], rtol=1e-5, atol=1e-6)


def test_attach_replaces_the_encoder_forward_only():
    class Enc(torch.nn.Module):                      # the interface of the reference's models.Encoder: main + fc
        def __init__(self):
            super().__init__()
            self.main = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, 1, 1), torch.nn.ReLU())
            self.fc = torch.nn.Linear(4 * 8 * 8, 2 * 6)

        def forward(self, x):
            y = self.fc(self.main(x).view(x.size(0), -1))
            return y.chunk(2, dim=1)

    torch.manual_seed(1)
    enc = Enc()
    x = torch.randn(3, 3, 8, 8)
    mu_r, lv_r = enc(x)
    keys = list(enc.state_dict().keys())
    head.attach(enc)
    mu, lv = enc(x)
    assert list(enc.state_dict().keys()) == keys
    assert mu.is_contiguous() and lv.is_contiguous() and not mu_r.is_contiguous()
    torch.testing.assert_close(mu, mu_r, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lv, lv_r, rtol=1e-5, atol=1e-6)
