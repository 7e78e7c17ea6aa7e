"""Loss value + cotangent on the device, differentiable w.r.t. the prediction.

    mse(yhat, y)                          = mean(abs2, yhat - y)                        (VMH.md:105-109; Flux.Losses.mse [DEP])
    logitcrossentropy(yhat, y, mask=None) = mean(-sum(y .* logsoftmax(yhat[:, mask]); dims=1))   (graph_node.md:100-106)

Arrays are Julia-shaped `(features, items)` stored column-major, like everything else at the layer boundary; the kernels
(`ngpde_mse_loss`, `ngpde_logit_cross_entropy`) write the loss as a device scalar and the cotangent of `yhat` in the same
pass, with fixed-order two-stage reductions.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib, ops
from .graph import from_rowmajor, rowmajor

Tensor = torch.Tensor


def _ws(dev) -> Tensor:
    return torch.empty(int(_lib.load().ngpde_loss_workspace_bytes()), dtype=torch.uint8, device=dev)


class _Mse(torch.autograd.Function):
    @staticmethod
    def forward(ctx, yhat: Tensor, y: Tensor):
        if not yhat.is_cuda:
            raise _lib.NgpdeError("mse: CUDA tensors only (no CPU fallback)")
        a, b = ops._f32c(yhat, "yhat"), ops._f32c(y, "y")
        if a.shape != b.shape:
            raise ValueError(f"DimensionMismatch: {tuple(a.shape)} vs {tuple(b.shape)}")
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        d = torch.empty_like(a)
        ws = _ws(a.device)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().ngpde_mse_loss(a.data_ptr(), b.data_ptr(), a.numel(), loss.data_ptr(), d.data_ptr(),
                                                  ws.data_ptr(), ws.numel(), ops._stream(a.device)))
        ops.LAUNCHES["count"] += 2
        ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (d,) = ctx.saved_tensors
        return d * g, None


def mse(yhat: Tensor, y: Tensor) -> Tensor:
    # element order is irrelevant to a mean over all entries, so the Julia-shaped views are used as stored
    return _Mse.apply(yhat.T if yhat.dim() == 2 else yhat, y.T if y.dim() == 2 else y)


class _LogitCE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, yhat_rm: Tensor, y_rm: Tensor, mask: Optional[Tensor]):
        if not yhat_rm.is_cuda:
            raise _lib.NgpdeError("logitcrossentropy: CUDA tensors only (no CPU fallback)")
        a, b = ops._f32c(yhat_rm, "yhat"), ops._f32c(y_rm, "y")
        n, c = a.shape
        idx = None
        if mask is not None:
            idx = mask.to(device=a.device)
            if idx.dtype == torch.bool:
                idx = torch.nonzero(idx, as_tuple=False).reshape(-1)
            idx = idx.to(torch.int32).contiguous()
        nm = n if idx is None else idx.numel()
        if tuple(b.shape) != (nm, c):
            raise ValueError(f"DimensionMismatch: y has shape {tuple(b.shape)[::-1]}, expected ({c}, {nm})")
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        d = torch.empty_like(a)
        ws = _ws(a.device)
        with torch.cuda.device(a.device):
            _lib.check(_lib.load().ngpde_logit_cross_entropy(a.data_ptr(), n, c, b.data_ptr(),
                                                             None if idx is None else idx.data_ptr(), nm, loss.data_ptr(),
                                                             d.data_ptr(), ws.data_ptr(), ws.numel(), ops._stream(a.device)))
        ops.LAUNCHES["count"] += 2
        ctx.save_for_backward(d)
        return loss

    @staticmethod
    def backward(ctx, g: Tensor):
        (d,) = ctx.saved_tensors
        return d * g, None, None


def logitcrossentropy(yhat: Tensor, y: Tensor, mask: Optional[Tensor] = None) -> Tensor:
    """yhat (C, N); y (C, n_masked) one-hot / soft targets of the masked columns; mask: index vector (0-based) or bool (N,)."""
    return _LogitCE.apply(rowmajor(yhat), rowmajor(y), mask)
