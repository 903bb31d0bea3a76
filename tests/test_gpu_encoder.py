"""-m gpu: the encoder operators (ifd_plane_bins, ifd_scatter_max_gather, ifd_scatter_mean_cl) against the oracle's
restatement of coordinate2index / torch_scatter semantics (bit-exact: max is order-free, the mean sums each bin in
ascending point order like a sequential scatter_add), and the product encoder on the GPU against the oracle encoder."""
import numpy as np
import pytest
import torch

from ifdefense_b200 import capi, convonet, models, synth
from oracle import torch_port as tp
from tests.gpu_util import dev

pytestmark = pytest.mark.gpu
PL = ("xz", "xy", "yz")


def bins_of(p, R=64, padding=0.1):
    x = dev(p)
    B, T, _ = x.shape
    out = torch.empty((3, B, T), dtype=torch.int32, device="cuda")
    capi.check(capi.lib().ifd_plane_bins(capi.ptr(x), B, T, R, padding, capi.ptr(out), capi.stream()))
    return out


def cloud(B, T, seed, spread=0.5):
    r = np.random.default_rng(seed)
    p = r.uniform(-spread, spread, size=(B, T, 3)).astype(np.float32)
    p[:, :5] = r.uniform(-0.7, 0.7, size=(B, 5, 3))         # a few points beyond the padded cube: both clamps
    p[:, 5] = p[:, 6]                                       # a duplicate point
    return p


@pytest.mark.parametrize("B,T,R,spread", [(3, 600, 64, 0.5), (2, 300, 64, 0.05), (1, 37, 16, 0.5), (2, 2048, 32, 0.5)])
def test_bins_and_scatter_ops_bit_exact(B, T, R, spread):
    p = cloud(B, T, seed=T, spread=spread)
    want_idx = {pl: tp.coordinate2index(tp.normalize_coordinate(torch.from_numpy(p).clone(), plane=pl), R) for pl in PL}
    bins = bins_of(p, R)
    for i, pl in enumerate(PL):
        assert np.array_equal(bins[i].cpu().numpy(), want_idx[pl][:, 0].numpy())
    C = 32
    feat = np.random.default_rng(1).normal(size=(B, T, C)).astype(np.float32)
    f = torch.from_numpy(feat)
    # pool_local: scatter_max -> gather, summed over the planes starting from 0
    want = 0
    for pl in PL:
        want = want + tp._scatter_max_gather(f, want_idx[pl], R * R)
    want = want.permute(0, 2, 1).contiguous().numpy()
    src, out = dev(feat), torch.empty((B, T, C), dtype=torch.float32, device="cuda")
    capi.check(capi.lib().ifd_scatter_max_gather(capi.ptr(src), capi.ptr(bins), 3, B, T, C, R * R, capi.ptr(out), capi.stream()))
    assert np.array_equal(out.cpu().numpy(), want)
    # generate_plane_features: scatter_mean (sequential sum / clamp(count, 1)), channels-last output
    for i, pl in enumerate(PL):
        idx = want_idx[pl].expand(-1, C, -1)
        s = f.permute(0, 2, 1)
        plane = s.new_zeros(B, C, R * R).scatter_add_(-1, idx, s)
        cnt = torch.zeros_like(plane).scatter_add_(-1, idx, torch.ones_like(s))
        plane = (plane / cnt.clamp_(min=1)).permute(0, 2, 1).contiguous().numpy()          # [B, R^2, C]
        got = torch.empty((B, R * R, C), dtype=torch.float32, device="cuda")
        capi.check(capi.lib().ifd_scatter_mean_cl(capi.ptr(src), capi.ptr(bins[i].contiguous()), B, T, C, R * R, capi.ptr(got),
                                                  capi.stream()))
        assert np.array_equal(got.cpu().numpy(), plane), pl
    again = torch.empty_like(out)
    capi.check(capi.lib().ifd_scatter_max_gather(capi.ptr(src), capi.ptr(bins), 3, B, T, C, R * R, capi.ptr(again), capi.stream()))
    assert torch.equal(again, out)


def test_product_encoder_on_gpu_matches_oracle_and_feeds_the_loop():
    sd = models.synthetic_state_dict("convonet", 0)
    model = models.build_convonet()
    model.load_state_dict(sd)
    model = model.cuda().eval()
    case = synth.make_case(3, K=256, seed=1)                              # CPU planes (torch encoder on the CPU)
    with torch.no_grad():
        got = model.encode_inputs(case.sel.cuda())
        again = model.encode_inputs(case.sel.cuda())
    want = tp.convonet_encode(sd, case.sel)
    for k in PL:
        assert got[k].shape == (3, 32, 64, 64)
        assert torch.equal(got[k], again[k])                             # reproducible run to run
        d = (got[k].cpu() - want[k]).abs().max().item()
        assert d < 1e-5 * want[k].abs().max().item(), (k, d)             # own kernels, fp32 semantics (3xTF32) vs torch CPU
    # the planes are born channels-last: taken as they are (no copy), and equal to the transposing NCHW route
    pl_born = convonet.planes_to_channels_last(got)
    assert pl_born is got.channels_last
    pl_fast = convonet.planes_to_channels_last({k: v.contiguous(memory_format=torch.channels_last) for k, v in got.items()})
    pl_ref = convonet.planes_to_channels_last({k: v.contiguous() for k, v in got.items()})
    assert torch.equal(pl_fast, pl_ref) and torch.equal(pl_born, pl_ref)


def test_encoder_errors():
    x = dev(cloud(1, 64, 0))
    with pytest.raises(RuntimeError, match="bad arguments"):
        capi.check(capi.lib().ifd_scatter_max_gather(capi.ptr(x), capi.ptr(x), 3, 1, 64, 3, 16, capi.ptr(x), capi.stream()))
    with pytest.raises(RuntimeError, match="nbins"):
        capi.check(capi.lib().ifd_scatter_mean_cl(capi.ptr(x), capi.ptr(x), 1, 64, 3, 1 << 20, capi.ptr(x), capi.stream()))
