"""Drop-in check against the REAL reference (staged unmodified under oracle/_ref by oracle/ref_loader.py):

* the reference's own ``IntroTCSovler.train_step`` / ``TCSovler.train_step`` (solvers/intro.py:56-196, solvers/vae.py:89-136)
  on the reference's own ``SoftIntroVAE`` run on top of the kernels after ``intro_tc_vae_b200.install()``;
* with identical seeds the installed step returns the same losses as the untouched reference run in eager torch on the
  same GPU;
* ``solver.compute_kl_loss`` of the installed reference equals the untouched reference's on the same encoder outputs.
"""
import math

import pytest
import torch

from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="oracle/_ref not staged")]

N_DATA = 16704


@pytest.fixture(autouse=True)
def _restore_package_writer():
    """install() aliases the package's SingletonWriter to the reference's; undo that so later tests see the package's own."""
    from intro_tc_vae_b200 import utils as fast_utils
    saved = fast_utils.SingletonWriter
    yield
    fast_utils.SingletonWriter = saved


class _Dataset:
    def __len__(self):
        return N_DATA


def _build(solver_name, dev, installed, seed=0, B=48, zdim=32, image=16):
    """Model + solver from the reference's modules (fresh import inside the current ref_loader.on_path() context)."""
    import models
    if installed:
        import intro_tc_vae_b200
        intro_tc_vae_b200.install()
    from solvers.intro_tc import IntroTCSovler
    from solvers.tc import TCSovler
    from utils import SingletonWriter
    SingletonWriter().writer = None                  # what train.py:100-103 does without TensorBoard
    SingletonWriter().cur_iter = 0
    torch.manual_seed(seed)
    model = models.SoftIntroVAE(arch="conv", cdim=3, zdim=zdim, channels=(16, 32), image_size=image).to(dev)
    opt_e = torch.optim.Adam(model.encoder.parameters(), lr=2e-4)
    opt_d = torch.optim.Adam(model.decoder.parameters(), lr=2e-4)
    common = dict(dataset=_Dataset(), model=model, batch_size=B, optimizer_e=opt_e, optimizer_d=opt_d, recon_loss_type="mse",
                  beta_kl=0.5, beta_rec=0.75, device=dev, use_amp=False, grad_scaler=None, writer=None, test_iter=1000, clip=100.0)
    if solver_name == "tc":
        return model, TCSovler(**common)
    return model, IntroTCSovler(**common, beta_neg=512.0, gamma_r=1e-8)


@pytest.mark.parametrize("solver_name", ["tc", "intro-tc"])
def test_reference_train_step_runs_on_the_kernels_and_matches_the_untouched_reference(solver_name):
    dev = torch.device("cuda:0")
    B = 48
    batch = torch.rand(B, 3, 16, 16, generator=torch.Generator().manual_seed(5))
    outs = {}
    for installed in (False, True):
        with ref_loader.on_path():
            from intro_tc_vae_b200 import _lib
            lib = _lib.load()
            model, solver = _build(solver_name, dev, installed)
            torch.manual_seed(11)                        # CPU generator: the fake-batch noise; CUDA generator: the reparameterize eps
            torch.cuda.manual_seed(11)
            c0 = lib.tcelbo_launch_count()
            out = solver.train_step(batch, 0)
            launched = lib.tcelbo_launch_count() - c0
            assert all(math.isfinite(v) for v in out.values() if v is not None), out
            # the installed reference must actually run the library's kernels; the untouched one must not
            assert (launched > 0) == installed, (installed, launched)
            outs[installed] = out
    ref, got = outs[False], outs[True]
    # loss_dec of the Soft-Intro step is evaluated after the encoder's Adam update, which amplifies last-bit differences of the
    # gradients (the update is lr * g / sqrt(v)); the other three are pure forward values of identical parameters
    for key, tol in (("loss_enc", 1e-4), ("loss_kl", 1e-4), ("loss_rec", 1e-4), ("loss_dec", 2e-3)):
        assert abs(got[key] - ref[key]) <= tol * max(abs(ref[key]), 1e-3), (key, got[key], ref[key])


def test_installed_compute_kl_loss_equals_reference_on_identical_latents():
    """Identical (z, mu, logvar): the installed reference solver's compute_kl_loss (mean / none / explicit beta) vs the untouched
    reference's own ops on the CPU, at the north-star tolerances, gradients included."""
    dev = torch.device("cuda:0")
    B, D = 96, 32
    g = torch.Generator().manual_seed(17)
    mu_c, lv_c, eps_c = torch.randn(B, D, generator=g), -2.0 + torch.randn(B, D, generator=g), torch.randn(B, D, generator=g)
    ref = {}
    with ref_loader.on_path():
        _, solver = _build("intro-tc", torch.device("cpu"), installed=False, B=B, zdim=D)
        mu, lv = mu_c.clone().requires_grad_(True), lv_c.clone().requires_grad_(True)
        z = mu + eps_c * torch.exp(0.5 * lv)
        loss = solver.compute_kl_loss(z, mu, lv)
        rows = solver.compute_kl_loss(z, mu, lv, reduce="none", beta=512.0)
        (loss + 1e-3 * rows.mean()).backward()
        ref = dict(loss=loss.item(), rows=rows.detach(), gmu=mu.grad, glv=lv.grad)
    with ref_loader.on_path():
        _, solver = _build("intro-tc", dev, installed=True, B=B, zdim=D)
        mu, lv = mu_c.to(dev).requires_grad_(True), lv_c.to(dev).requires_grad_(True)
        z = mu + eps_c.to(dev) * torch.exp(0.5 * lv)
        loss = solver.compute_kl_loss(z, mu, lv)
        rows = solver.compute_kl_loss(z, mu, lv, reduce="none", beta=512.0)
        (loss + 1e-3 * rows.mean()).backward()
        rel = lambda a, b: ((a.detach().cpu().double() - b.double()).abs().max() / b.double().abs().max()).item()   # noqa: E731
        assert abs(loss.item() - ref["loss"]) <= 1e-5 * abs(ref["loss"])
        assert rel(rows, ref["rows"]) <= 1e-5
        assert rel(mu.grad, ref["gmu"]) <= 1e-4
        assert rel(lv.grad, ref["glv"]) <= 1e-4


@pytest.mark.parametrize("capture", [False, True])
def test_graphed_soft_intro_step_matches_the_reference_train_step(capture):
    """intro_tc_vae_b200.train_step.SoftIntroTCStep (no host synchronisation, optionally two CUDA graphs) performs the update of
    the reference's IntroTCSovler.train_step (solvers/intro.py:56-196): same seeds -> same losses on the first step and the same
    parameters afterwards (up to Adam's amplification of last-bit gradient differences), on the reference's own SoftIntroVAE."""
    dev = torch.device("cuda:0")
    B = 48
    batch = torch.rand(B, 3, 16, 16, generator=torch.Generator().manual_seed(5))
    with ref_loader.on_path():
        model, solver = _build("intro-tc", dev, installed=False)
        torch.manual_seed(11)
        torch.cuda.manual_seed(11)
        noise = torch.randn(B, model.zdim)                      # what solvers/intro.py:61 draws first from the CPU generator
        torch.manual_seed(11)
        init_params = [p.detach().clone() for p in model.parameters()]
        ref_out = solver.train_step(batch, 0)
        ref_params = [p.detach().clone() for p in model.parameters()]
    with ref_loader.on_path():
        import models
        import intro_tc_vae_b200
        from intro_tc_vae_b200.train_step import SoftIntroTCStep
        intro_tc_vae_b200.install()
        torch.manual_seed(0)
        model = models.SoftIntroVAE(arch="conv", cdim=3, zdim=32, channels=(16, 32), image_size=16).to(dev)
        opt_e = torch.optim.Adam(model.encoder.parameters(), lr=2e-4, capturable=True)
        opt_d = torch.optim.Adam(model.decoder.parameters(), lr=2e-4, capturable=True)
        step = SoftIntroTCStep(model, opt_e, opt_d, N_DATA, batch.shape, recon_loss_type="mse", beta_kl=0.5, beta_rec=0.75,
                               beta_neg=512.0, gamma_r=1e-8, clip=100.0, capture=capture)
        torch.cuda.manual_seed(11)
        out = step(batch.to(dev), noise.to(dev))
        step.check_finite()
        got = {k: v.item() for k, v in out.items()}
        for key, tol in (("loss_enc", 1e-4), ("loss_kl", 1e-4), ("loss_rec", 1e-4), ("loss_dec", 2e-3)):
            assert abs(got[key] - ref_out[key]) <= tol * max(abs(ref_out[key]), 1e-3), (capture, key, got[key], ref_out[key])
        assert abs(max(got["norm_e"], got["norm_d"]) - ref_out["L2"]) <= 1e-3 * ref_out["L2"]
        # Adam's first update is lr * g / |g| per weight, so weights whose gradient is ~0 may step the other way on last-bit
        # differences; the two update vectors must still point the same way overall
        u_ref = torch.cat([(q - p0).flatten() for q, p0 in zip(ref_params, init_params)])
        u_got = torch.cat([(p.detach() - p0).flatten() for p, p0 in zip(model.parameters(), init_params)])
        cos = torch.dot(u_ref, u_got) / (u_ref.norm() * u_got.norm())
        assert cos.item() > 0.99, (capture, cos.item())


def test_train_step_throughput_reference_vs_dropin_vs_graphed_step():
    """BASELINE configs[1] shape (64x64x3 images, conv arch with channels 64-128-256-512, z_dim 128, batch 64) on the reference's own
    SoftIntroVAE: images/s of (a) the untouched reference's IntroTCSovler.train_step in eager torch on this GPU, (b) the same
    train_step after install() (loss terms on the kernels), (c) SoftIntroTCStep (two CUDA graphs, no host synchronisation).
    Recorded in gpurun_out/r2_train_dropin.json; the kernels must not make the reference's own step slower."""
    import json
    import os
    import time
    dev = torch.device("cuda:0")
    B, steps = 64, 8
    batch = torch.rand(B, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    out = {}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        return B * steps / (time.perf_counter() - t0)

    for installed in (False, True):
        with ref_loader.on_path():
            import models
            if installed:
                import intro_tc_vae_b200
                intro_tc_vae_b200.install()
            from solvers.intro_tc import IntroTCSovler
            from utils import SingletonWriter
            SingletonWriter().writer, SingletonWriter().cur_iter = None, 0
            torch.manual_seed(0)
            model = models.SoftIntroVAE(arch="conv", cdim=3, zdim=128, channels=(64, 128, 256, 512), image_size=64).to(dev)
            opt_e = torch.optim.Adam(model.encoder.parameters(), lr=2e-4)
            opt_d = torch.optim.Adam(model.decoder.parameters(), lr=2e-4)
            solver = IntroTCSovler(dataset=_Dataset(), model=model, batch_size=B, optimizer_e=opt_e, optimizer_d=opt_d, recon_loss_type="mse",
                                   beta_kl=0.5, beta_rec=0.75, beta_neg=512.0, gamma_r=1e-8, device=dev, use_amp=False, grad_scaler=None,
                                   writer=None, test_iter=1000, clip=100.0)
            out["installed_reference_train_step" if installed else "untouched_reference_train_step"] = timed(lambda: solver.train_step(batch, 0))
            if installed:
                from intro_tc_vae_b200.train_step import SoftIntroTCStep
                torch.manual_seed(0)
                model2 = models.SoftIntroVAE(arch="conv", cdim=3, zdim=128, channels=(64, 128, 256, 512), image_size=64).to(dev)
                oe = torch.optim.Adam(model2.encoder.parameters(), lr=2e-4, capturable=True)
                od = torch.optim.Adam(model2.decoder.parameters(), lr=2e-4, capturable=True)
                step = SoftIntroTCStep(model2, oe, od, N_DATA, batch.shape, recon_loss_type="mse", beta_kl=0.5, beta_rec=0.75, beta_neg=512.0,
                                       gamma_r=1e-8, clip=100.0)
                real, noise = batch.to(dev), torch.randn(B, 128, device=dev)
                out["graphed_soft_intro_step"] = timed(lambda: step(real, noise))
                step.check_finite()
    out = {k: round(v, 1) for k, v in out.items()}
    out["unit"] = "images/s, one B200, 64x64x3, conv arch, z_dim 128, batch 64, fp32"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "r2_train_dropin.json"), "w") as f:
        json.dump(out, f)
    assert out["installed_reference_train_step"] > 0.9 * out["untouched_reference_train_step"], out
    assert out["graphed_soft_intro_step"] > out["installed_reference_train_step"], out
