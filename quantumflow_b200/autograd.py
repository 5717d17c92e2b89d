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

Small states (<= 12 qubits, the QAOA workload): a whole run of 1- and 2-qubit unitary gates is ONE node and ONE launch
each way (`run_small_circuit`, csrc/qfb_small.cu): the state stays in the shared memory of a CTA for all gates, and the
backward pass is the reverse sweep of the adjoint method -- it recomputes every intermediate state with U^H instead
of saving it, so the node keeps the final state and the matrices only.
"""
from typing import Optional, Sequence

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


# ---------------------------------------------------------------------------------------------------------
# whole small circuits as one node (csrc/qfb_small.cu)
# ---------------------------------------------------------------------------------------------------------

SMALL_MAX_BITS = 12        # state + adjoint vector in the shared memory of one CTA


def small_circuit_shape_ok(gates, nbits: int, min_gates: int = 2) -> bool:
    """Structural half of the eligibility test: 1- and 2-qubit gates on <= 12 qubits."""
    return nbits <= SMALL_MAX_BITS and len(gates) >= min_gates and all(g.qubit_nb in (1, 2) for g in gates)


def _all_unitary(flat: np.ndarray, ks: Sequence[int]) -> bool:
    """Are all matrices packed in `flat` (2^k x 2^k row-major each, in order) unitary? Vectorised per arity."""
    sizes = np.asarray([1 << (2 * k) for k in ks])
    starts = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    for k in (1, 2):
        sel = [i for i, kk in enumerate(ks) if kk == k]
        if not sel:
            continue
        dim = 1 << k
        idx = (starts[sel][:, None] + np.arange(dim * dim)[None, :]).reshape(-1)
        m = flat[idx].reshape(len(sel), dim, dim)
        if not np.allclose(m @ m.conj().transpose(0, 2, 1), np.eye(dim)[None], atol=1e-12):
            return False
    return True


class _RunSmallCircuit(torch.autograd.Function):
    """(psi, gate records on the device, all matrices packed into one HOST tensor) -> final state. The packing is a
    torch.cat outside this node, so autograd routes the packed gradient back to every gate's own tensor."""

    @staticmethod
    def forward(ctx, psi: torch.Tensor, desc: torch.Tensor, flat: torch.Tensor):
        from . import _lib
        lib = _lib.load()
        nbits = psi.dim() if psi.dim() > 1 else int(psi.numel()).bit_length() - 1
        flat_dev = flat.detach().to(psi.device, non_blocking=True)
        src = psi.detach().contiguous()
        out = torch.empty_like(src)
        _lib.check(lib.qfb_small_circuit_run(out.data_ptr(), src.data_ptr(), nbits, 1, desc.shape[0], desc.data_ptr(),
                                             flat_dev.data_ptr(), flat_dev.numel(), engine._stream()))
        ctx.save_for_backward(out, desc, flat_dev)
        ctx.nbits = nbits
        ctx.flat_device = flat.device
        return out

    @staticmethod
    def backward(ctx, grad_out: torch.Tensor):
        from . import _lib
        lib = _lib.load()
        out, desc, flat_dev = ctx.saved_tensors
        ngates = desc.shape[0]
        g = grad_out.contiguous()
        grad_flat = torch.empty_like(flat_dev)
        grad_in = torch.empty_like(out)
        scratch = torch.empty(int(lib.qfb_small_circuit_scratch_doubles(1, ngates)), dtype=torch.float64,
                              device=out.device)
        _lib.check(lib.qfb_small_circuit_adjoint(out.data_ptr(), g.data_ptr(), ctx.nbits, 1, ngates, desc.data_ptr(),
                                                 flat_dev.data_ptr(), flat_dev.numel(), grad_flat.data_ptr(),
                                                 grad_in.data_ptr(), scratch.data_ptr(), engine._stream()))
        return (grad_in if ctx.needs_input_grad[0] else None, None,
                grad_flat.to(ctx.flat_device) if ctx.needs_input_grad[2] else None)


def run_small_circuit(psi: torch.Tensor, gates, bits_of) -> Optional[torch.Tensor]:
    """psi -> U_G ... U_1 psi for a run of 1- and 2-qubit gates (small_circuit_shape_ok) as one autograd node and one
    launch each way; None when a gate is not unitary (the reverse sweep un-applies gates with U^H), so that the
    caller takes the gate-by-gate path. `bits_of(gate)` = index bit of every gate qubit (gate qubit 0 first)."""
    rows, ks, at = [], [], 0
    for g in gates:
        bits = bits_of(g)
        k = len(bits)
        rows.append((k, bits[0], bits[1] if k > 1 else 0, at))
        ks.append(k)
        at += 1 << (2 * k)
    flat = torch.cat([g.tensor.reshape(-1) if isinstance(g.tensor, torch.Tensor)
                      else torch.from_numpy(g.matrix().reshape(-1)) for g in gates])
    if flat.dtype != torch.complex128:
        flat = flat.to(torch.complex128)
    if flat.is_cuda:
        flat = flat.cpu()
    if not _all_unitary(flat.detach().numpy(), ks):
        return None
    desc = torch.tensor(rows, dtype=torch.int32).to(psi.device, non_blocking=True)
    return _RunSmallCircuit.apply(psi, desc, flat)
