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
