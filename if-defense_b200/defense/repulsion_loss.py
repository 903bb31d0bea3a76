"""RepulsionLoss with the reference's interface (ConvONet/defense/repulsion_loss.py:6-71): a module whose
call takes pred [B,N,3] and returns the per-cloud loss [B], differentiable w.r.t. pred.  kNN, gather, pair
terms and the scatter-add backward are one fused kernel; the [B,N,N] distance matrix is never materialised,
so the reference's OOM back-off (knn_batch halving, :25-39) has nothing left to do -- the argument is kept
for signature compatibility."""
import torch
import torch.nn as nn

from .. import capi


class _RepulsionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, k, radius, h, eps):
        capi.require_gpu()
        x = pred.detach().float().contiguous()
        B, K, _ = x.shape
        L = capi.lib()
        ws_bytes = L.ifd_knn_repulsion_workspace_bytes(B, K)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        loss = torch.empty(B, dtype=torch.float32, device=x.device)
        grad = torch.empty_like(x)
        capi.check(L.ifd_knn_repulsion(capi.ptr(x), B, K, k, radius, h, eps, None, None, capi.ptr(loss), capi.ptr(grad),
                                       capi.ptr(ws), ws_bytes, capi.stream()), "ifd_knn_repulsion")
        ctx.save_for_backward(grad)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        (grad,) = ctx.saved_tensors
        return grad * grad_loss.view(-1, 1, 1), None, None, None, None


class RepulsionLoss(nn.Module):
    def __init__(self, nn_size=5, radius=0.07, h=0.03, eps=1e-12, knn_batch=None):
        super().__init__()
        self.nn_size, self.radius, self.h, self.eps, self.knn_batch = nn_size, radius, h, eps, knn_batch

    def get_repulsion_loss(self, pred):
        return _RepulsionFn.apply(pred, self.nn_size, self.radius, self.h, self.eps)

    def forward(self, pred):
        return self.get_repulsion_loss(pred)


repulsion_loss = RepulsionLoss()
