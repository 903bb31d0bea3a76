"""TEST INFRASTRUCTURE ONLY. Pure-torch restatement of the two torch_scatter==2.0.5 entry points the
reference encoder calls (ConvONet/src/encoder/pointnet.py:5,77,112).  torch_scatter is an un-vendored
third-party dependency (requirements.txt:59), so this shim *is* the definition of the boundary:
parity unpinned by the reference's own tests (there are none).

  scatter_mean(src, index, dim=-1, out=None, dim_size=None): sum into bins, divide by clamp(count, 1)
  scatter_max (src, index, dim=-1, out=None, dim_size=None) -> (values, argmax); empty bins give 0 and
                                                                argmax == src.size(dim)
"""
import torch


def _expand(index, src):
    return index.expand_as(src) if index.shape != src.shape else index


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    assert dim in (-1, src.dim() - 1)
    index = _expand(index, src)
    if out is None:
        size = list(src.shape)
        size[-1] = int(dim_size) if dim_size is not None else int(index.max()) + 1
        out = src.new_zeros(size)
    out.scatter_add_(-1, index, src)
    cnt = torch.zeros_like(out).scatter_add_(-1, index, torch.ones_like(src))
    out.div_(cnt.clamp_(min=1))
    return out


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    assert dim in (-1, src.dim() - 1) and out is None
    index = _expand(index, src)
    size = list(src.shape)
    size[-1] = int(dim_size) if dim_size is not None else int(index.max()) + 1
    val = src.new_full(size, float("-inf"))
    val.scatter_reduce_(-1, index, src, reduce="amax", include_self=True)
    # argmax: first position attaining the max in each bin
    n = src.shape[-1]
    pos = torch.arange(n, device=src.device).expand_as(src)
    hit = src == val.gather(-1, index)
    cand = torch.where(hit, pos, torch.full_like(pos, n))
    arg = torch.full(size, n, dtype=torch.long, device=src.device)
    arg.scatter_reduce_(-1, index, cand, reduce="amin", include_self=True)
    val = torch.where(torch.isinf(val), torch.zeros_like(val), val)
    return val, arg
