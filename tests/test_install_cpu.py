"""install(): the staged reference checkout (oracle/_ref, see oracle/ref_loader.py) is routed through this package.
CPU-side check of the patching only -- no kernels run here."""
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.stage(), reason="reference not staged (oracle/_ref) and /root/reference absent")


def test_install_patches_every_binding_of_the_loss_helpers():
    with ref_loader.on_path():
        import intro_tc_vae_b200
        from intro_tc_vae_b200 import losses as fast_losses
        from intro_tc_vae_b200 import ops as fast_ops
        from intro_tc_vae_b200 import utils as fast_utils
        saved_writer = fast_utils.SingletonWriter
        try:
            patched = intro_tc_vae_b200.install()
            import models as ref_models
            import ops as ref_ops
            import solvers.intro as ref_intro
            import solvers.tc as ref_tc
            import solvers.vae as ref_vae
            import utils as ref_utils
            from solvers.intro_tc import IntroTCSovler
            assert ref_ops.total_correlation is fast_ops.total_correlation
            assert ref_tc.total_correlation is fast_ops.total_correlation      # solvers/tc.py:5-11 binds by name
            assert ref_tc.kl_divergence is fast_ops.kl_divergence
            assert ref_vae.kl_divergence is fast_ops.kl_divergence             # solvers/vae.py:22
            assert ref_vae.reconstruction_loss is fast_losses.reconstruction_loss
            assert ref_models.reparameterize is fast_ops.reparameterize        # models.py:5
            assert ref_intro.reparameterize is fast_ops.reparameterize         # solvers/intro.py:14
            assert ref_ops.reparameterize is fast_ops.reparameterize
            assert ref_tc.TCSovler._compute_kl_loss_simple.__module__ == "intro_tc_vae_b200.solvers.tc"
            assert ref_tc.TCSovler._compute_kl_loss_full.__module__ == "intro_tc_vae_b200.solvers.tc"
            assert ref_tc.TCSovler.compute_kl_loss.__module__ == "solvers.tc"       # dispatcher untouched
            assert IntroTCSovler.compute_kl_loss.__module__ == "solvers.intro_tc"   # forwarder untouched, resolves TCSovler at call time
            # the patched loss methods log through the reference's singleton (train.py:100-103,212 update that one)
            assert fast_utils.SingletonWriter is ref_utils.SingletonWriter
            ref_utils.SingletonWriter().cur_iter = 41
            from intro_tc_vae_b200.solvers.tc import _singleton_writer
            assert _singleton_writer().cur_iter == 41
            assert "reparameterize" in patched["models"] and "reconstruction_loss" in patched["solvers.vae"]
            # the package's solver names are the reference's own classes
            from intro_tc_vae_b200 import solvers as fast_solvers
            assert fast_solvers.IntroTCSovler is IntroTCSovler and fast_solvers.TCSovler is ref_tc.TCSovler
        finally:
            fast_utils.SingletonWriter = saved_writer
