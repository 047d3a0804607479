"""GAN losses with the reference's class names and method signatures (huangzh13/StyleGAN.pytorch models/Losses.py).

``LogisticGAN`` (the default, config.py:21) is the hot path: softplus means and the R1 penalty's sum of squares run
on sgb200 kernels, and the penalty's double backward flows through the differentiable backward Functions of the
discriminator.  The other losses operate on the [B,1] logits with a handful of torch ops."""
import os

import torch
import torch.nn as nn
from torch.nn import BCEWithLogitsLoss

from . import _ops


class GANLoss:
    """Base class: holds the discriminator the loss is evaluated with."""

    def __init__(self, dis):
        self.dis = dis

    def dis_loss(self, real_samps, fake_samps, height, alpha):
        raise NotImplementedError("dis_loss method has not been implemented")

    def gen_loss(self, real_samps, fake_samps, height, alpha):
        raise NotImplementedError("gen_loss method has not been implemented")


class ConditionalGANLoss:
    """BCE loss for the conditional discriminator (reference Losses.py:54-89)."""

    def __init__(self, dis):
        self.criterion = BCEWithLogitsLoss()
        self.dis = dis

    def dis_loss(self, real_samps, fake_samps, labels, height, alpha):
        assert real_samps.device == fake_samps.device, "Real and Fake samples are not on the same device"
        dev = fake_samps.device
        r = self.dis(real_samps, height, alpha, labels_in=labels)
        f = self.dis(fake_samps, height, alpha, labels_in=labels)
        real_loss = self.criterion(torch.squeeze(r), torch.ones(real_samps.shape[0], device=dev))
        fake_loss = self.criterion(torch.squeeze(f), torch.zeros(fake_samps.shape[0], device=dev))
        return (real_loss + fake_loss) / 2

    def gen_loss(self, _, fake_samps, labels, height, alpha):
        preds = self.dis(fake_samps, height, alpha, labels_in=labels)
        return self.criterion(torch.squeeze(preds), torch.ones(fake_samps.shape[0], device=fake_samps.device))


class StandardGAN(GANLoss):
    """Non-saturating BCE GAN loss (reference Losses.py:96-133)."""

    def __init__(self, dis):
        super().__init__(dis)
        self.criterion = BCEWithLogitsLoss()

    def dis_loss(self, real_samps, fake_samps, height, alpha):
        assert real_samps.device == fake_samps.device, "Real and Fake samples are not on the same device"
        dev = fake_samps.device
        r = self.dis(real_samps, height, alpha)
        f = self.dis(fake_samps, height, alpha)
        real_loss = self.criterion(torch.squeeze(r), torch.ones(real_samps.shape[0], device=dev))
        fake_loss = self.criterion(torch.squeeze(f), torch.zeros(fake_samps.shape[0], device=dev))
        return (real_loss + fake_loss) / 2

    def gen_loss(self, _, fake_samps, height, alpha):
        # the reference unpacks a 3-tuple here (Losses.py:131) although the discriminator returns one tensor
        preds = self.dis(fake_samps, height, alpha)
        return self.criterion(torch.squeeze(preds), torch.ones(fake_samps.shape[0], device=fake_samps.device))


class HingeGAN(GANLoss):
    def __init__(self, dis):
        super().__init__(dis)

    def dis_loss(self, real_samps, fake_samps, height, alpha):
        r = self.dis(real_samps, height, alpha)
        f = self.dis(fake_samps, height, alpha)
        return torch.mean(torch.relu(1 - r)) + torch.mean(torch.relu(1 + f))

    def gen_loss(self, _, fake_samps, height, alpha):
        return -torch.mean(self.dis(fake_samps, height, alpha))


class RelativisticAverageHingeGAN(GANLoss):
    def __init__(self, dis):
        super().__init__(dis)

    def _diffs(self, real_samps, fake_samps, height, alpha):
        r = self.dis(real_samps, height, alpha)
        f = self.dis(fake_samps, height, alpha)
        return r - torch.mean(f), f - torch.mean(r)

    def dis_loss(self, real_samps, fake_samps, height, alpha):
        rf, fr = self._diffs(real_samps, fake_samps, height, alpha)
        return torch.mean(torch.relu(1 - rf)) + torch.mean(torch.relu(1 + fr))

    def gen_loss(self, real_samps, fake_samps, height, alpha):
        rf, fr = self._diffs(real_samps, fake_samps, height, alpha)
        return torch.mean(torch.relu(1 + rf)) + torch.mean(torch.relu(1 - fr))


class LogisticGAN(GANLoss):
    """Logistic loss with the R1 gradient penalty on reals (reference Losses.py:192-229)."""

    # The reference evaluates D(real) twice per discriminator step: once for the logistic term (Losses.py:214) and once,
    # with requires_grad on the images, inside R1Penalty (Losses.py:203).  With `shared_real_pass` the two share one
    # forward graph (same function value, same gradients up to summation order): one discriminator forward and one
    # backward through it fewer per step, 3 of the step's 18 network passes.  On by default; `models.set_shared_r1(False)`
    # or SG_SHARED_R1=0 restores the two-pass form (tools/check_shared_r1.py and tests/test_gpu_coverage.py compare them).
    shared_real_pass = os.environ.get("SG_SHARED_R1", "1") == "1"

    def __init__(self, dis):
        super().__init__(dis)

    @staticmethod
    def _penalty_from(real_img, real_logit):
        """sum over the batch of |d D(real) / d real|^2, with the graph kept for the second backward."""
        with _ops.skip_param_grads():        # only d/d(image) is wanted from this backward
            real_grads = torch.autograd.grad(outputs=real_logit, inputs=real_img,
                                             grad_outputs=torch.ones_like(real_logit), create_graph=True,
                                             retain_graph=True)[0]
        return _ops.SqSum.apply(real_grads, 1.0).reshape(())

    def R1Penalty(self, real_img, height, alpha):
        real_img = real_img.detach().requires_grad_(True)
        real_logit = self.dis(real_img, height, alpha)
        return self._penalty_from(real_img, real_logit)

    def dis_loss(self, real_samps, fake_samps, height, alpha, r1_gamma=10.0):
        shared = self.shared_real_pass and r1_gamma != 0.0 and torch.is_grad_enabled()
        if shared:
            real_samps = real_samps.detach().requires_grad_(True)
        r_preds = self.dis(real_samps, height, alpha)
        f_preds = self.dis(fake_samps, height, alpha)
        loss = _ops.SoftplusMean.apply(f_preds, 1.0) + _ops.SoftplusMean.apply(r_preds, -1.0)
        loss = loss.reshape(())
        if shared:
            loss = loss + self._penalty_from(real_samps, r_preds) * (r1_gamma * 0.5)
        elif r1_gamma != 0.0:
            loss = loss + self.R1Penalty(real_samps.detach(), height, alpha) * (r1_gamma * 0.5)
        return loss

    def gen_loss(self, _, fake_samps, height, alpha):
        f_preds = self.dis(fake_samps, height, alpha)
        return _ops.SoftplusMean.apply(f_preds, -1.0).reshape(())
