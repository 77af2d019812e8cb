"""Fused losses as autograd functions (forward and gradient in one kernel launch)."""
from __future__ import annotations

import torch

from . import ops


class _N2NLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out, sub2, den1, den2, lam):
        loss3, grad = ops.n2n_loss_fwdbwd(out, sub2, den1, den2, lam, 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    def backward(ctx, g, _g3):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None, None, None


def n2n_loss(noisy_output, noisy_target, sub1_denoised, sub2_denoised, Lambda):
    """training_script.md:146-153: returns (loss_all, [loss_all, loss1, loss2]).  Only
    ``noisy_output`` receives a gradient (the other three are no-grad in the reference loop)."""
    return _N2NLoss.apply(noisy_output, noisy_target.detach(), sub1_denoised.detach(), sub2_denoised.detach(),
                          float(Lambda))


class _L1GradLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, lambda_grad):
        loss3, grad = ops.l1grad_loss_fwdbwd(pred, target, lambda_grad, 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    def backward(ctx, g, _g3):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def l1_grad_loss(pred, clean, lambda_grad):
    """finetune.py:283-285: L1(pred, clean) + lambda_grad * gradient_loss(pred, clean) ->
    (loss, [loss, loss_l1, loss_grad])."""
    return _L1GradLoss.apply(pred, clean.detach(), float(lambda_grad))


class _StructureLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, pred2, target, alpha, beta, gamma):
        loss4, g1, g2 = ops.structure_loss_fwdbwd(pred, pred2, target, alpha, beta, gamma, 1.0, want_grad=True)
        ctx.save_for_backward(g1, g2)
        ctx.mark_non_differentiable(loss4)
        return loss4[0].clone(), loss4

    @staticmethod
    def backward(ctx, g, _g4):
        g1, g2 = ctx.saved_tensors
        return g1 * g, g2 * g, None, None, None, None


class Structure_loss(torch.nn.Module):
    """Drop-in ``util.Structure_loss`` (util.py:41-70), the criterion of the fork's live training loop
    (train.py:322, :361-363): ``criterion(network(noisy), network(clean), clean)`` =
    alpha*L1(pred, target) + beta*TV(pred2) + gamma*L1(pred2, target), forward and both gradients in ONE kernel.
    ``last_terms`` holds the device tensor [loss, pixel, TV, consistency] of the most recent call (pixel is the
    ``F.l1_loss(noisy_output, clean)`` the reference logs separately, train.py:365)."""

    def __init__(self, alpha: float = 1.0, beta: float = .5, gamma: float = .5, reduction: str = 'mean'):
        super().__init__()
        if reduction != 'mean':
            raise NotImplementedError("Structure_loss: only reduction='mean' (the reference's default and only use) is implemented")
        self.alpha, self.beta, self.gamma, self.reduction = alpha, beta, gamma, reduction
        self.last_terms = None

    def forward(self, pred, pred2, target):
        loss, self.last_terms = _StructureLoss.apply(pred, pred2, target.detach(), float(self.alpha), float(self.beta),
                                                      float(self.gamma))
        return loss


class _IqslLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target, t1, t2, tau, margin, ce_factor, eps):
        loss3, grad = ops.iqsl_loss_fwdbwd(pred, target, t1, t2, tau, margin, ce_factor, eps, 1.0, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.mark_non_differentiable(loss3)
        return loss3[0].clone(), loss3

    @staticmethod
    def backward(ctx, g, _g3):
        (grad,) = ctx.saved_tensors
        return (grad * g,) + (None,) * 7


def iqsl_loss(pred, target, t1, t2, tau=0.1, margin=0.0, ce_factor=0.5, eps=1e-6):
    """finetune_iqsl.py:291-383 (same signature): Intensity-Quantized Structural Loss — 3-class surrogate segmentation
    (dark / mid / bright by the thresholds t1, t2), multi-class Dice over the batch + ce_factor * soft cross-entropy.
    Forward and gradient w.r.t. ``pred`` in two fused kernels (the Dice couples all pixels through nine global sums)."""
    if pred.dim() == 3:
        pred = pred.unsqueeze(1)
    if target.dim() == 3:
        target = target.unsqueeze(1)
    loss, _terms = _IqslLoss.apply(pred, target.detach(), float(t1), float(t2), float(tau), float(margin), float(ce_factor), float(eps))
    return loss
