"""SORDefense with the reference's interface (ConvONet/defense/SOR.py:5-74): call with [B,K,3], get a list
of B tensors [K_i,3].  The float64 kNN / mean / std / threshold run in one kernel per batch; only the ragged
selection (boolean indexing, as in the reference :48) stays in torch."""
import torch
import torch.nn as nn

from .. import capi


class SORDefense(nn.Module):
    def __init__(self, k=2, alpha=1.1, sor_batch=None):
        super().__init__()
        self.k, self.alpha, self.sor_batch = k, alpha, sor_batch   # sor_batch kept for signature parity

    def outlier_removal(self, x):
        capi.require_gpu()
        pc = x.detach().float().contiguous()
        B, K, _ = pc.shape
        keep = torch.empty((B, K), dtype=torch.uint8, device=pc.device)
        capi.check(capi.lib().ifd_sor(capi.ptr(pc), B, K, self.k, float(self.alpha), capi.ptr(keep), None, capi.stream()),
                   "ifd_sor")
        mask = keep.bool()
        return [x[i][mask[i]] for i in range(B)]

    def forward(self, x):
        with torch.no_grad():
            return self.outlier_removal(x)
