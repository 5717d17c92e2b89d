"""torch.autograd bridge for the QAOA gradient path (the reference's equivalent is the TensorFlow graph built
by tensorflowbk.tensormul, tensorflowbk.py:134-152, driven from examples/qaoa_maxcut.py:37-87).

Gate parameters are torch tensors; gate operators are tiny host tensors built from them by the formulas in
stdgates.py (autograd-tracked). The state stays in HBM; each `tensormul` is one node:

  forward   out = U psi                         qfb_apply_dense / qfb_apply_diag
  backward  grad_psi = U^H g                    same kernels
            grad_U[r][c] = sum_groups g[r] conj(psi[c])      qfb_gate_grad (two-pass deterministic reduction)

following torch's convention for complex leaves (grad = conjugate Wirtinger derivative; checked against
torch.matmul's own backward in tests/test_autograd.py). The read-out node is sum_i d_i |psi_i|^2
(qfb_expect_diag); its backward and that of `probabilities` are single elementwise products on the device.
"""
from typing import Sequence

import numpy as np
import torch

from . import engine


def needs_grad(tensor) -> bool:
    return isinstance(tensor, torch.Tensor) and tensor.requires_grad and torch.is_grad_enabled()


class _ApplyOperator(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mat: torch.Tensor, psi: torch.Tensor, bits: Sequence[int]):
        bits = tuple(int(b) for b in bits)
        mat_np = mat.detach().cpu().numpy()
        out = engine.apply_operator(psi.detach(), mat_np, bits)
        ctx.save_for_backward(mat, psi)
        ctx.bits = bits
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        mat, psi = ctx.saved_tensors
        bits = ctx.bits
        g = grad_out.contiguous()
        grad_mat = grad_psi = None
        if ctx.needs_input_grad[0]:
            if len(bits) > 3:
                raise NotImplementedError('operator gradients are implemented for up to 3-qubit gates')
            grad_mat = engine.gate_grad(g, psi.detach(), bits).to(mat.device).reshape(mat.shape)
        if ctx.needs_input_grad[1]:
            mat_h = np.ascontiguousarray(mat.detach().cpu().numpy().conj().T)
            grad_psi = engine.apply_operator(g, mat_h, bits)
        return grad_mat, grad_psi, None


def tensormul_autograd(tensor0: torch.Tensor, tensor1: torch.Tensor, indices: Sequence[int]) -> torch.Tensor:
    """Differentiable tensormul(op, amplitudes, indices)."""
    n = tensor1.dim()
    k = len(indices)
    bits = [n - 1 - int(i) for i in indices]
    mat = tensor0.reshape(1 << k, 1 << k)
    if mat.is_cuda:
        mat = mat.cpu()
    return _ApplyOperator.apply(mat, tensor1.contiguous(), bits)


class _ExpectDiag(torch.autograd.Function):
    @staticmethod
    def forward(ctx, psi: torch.Tensor, diag: torch.Tensor):
        ctx.save_for_backward(psi, diag)
        return engine.expectation_diag(psi.detach(), diag.reshape(-1))

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        psi, diag = ctx.saved_tensors
        grad = (2.0 * grad_out) * diag.reshape(psi.shape) * psi.detach()
        return grad, None


def expectation_autograd(psi: torch.Tensor, diag_hermitian) -> torch.Tensor:
    if isinstance(diag_hermitian, torch.Tensor):
        diag = diag_hermitian.real if diag_hermitian.is_complex() else diag_hermitian
        diag = diag.to(device=psi.device, dtype=torch.float64)
    else:
        diag = torch.from_numpy(np.ascontiguousarray(np.real(np.asarray(diag_hermitian)),
                                                     dtype=np.float64)).to(psi.device)
    return _ExpectDiag.apply(psi.contiguous(), diag.contiguous())


class _Probabilities(torch.autograd.Function):
    @staticmethod
    def forward(ctx, psi: torch.Tensor):
        ctx.save_for_backward(psi)
        return engine.probabilities(psi.detach())

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        (psi,) = ctx.saved_tensors
        return 2.0 * grad_out * psi.detach()


def probabilities_autograd(psi: torch.Tensor) -> torch.Tensor:
    return _Probabilities.apply(psi.contiguous())


def inner_autograd(t0: torch.Tensor, t1: torch.Tensor) -> torch.Tensor:
    """<t0|t1> with gradient support (fidelity-style losses); elementwise product + sum on the device."""
    return torch.sum(torch.conj(t0.reshape(-1)) * t1.reshape(-1))
