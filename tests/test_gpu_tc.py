"""-m gpu: the tcgen05 GEMM engine (csrc/tc_gemm.cuh, tc_ops.cu) through the C ABI: Linear with fused ReLU / bias / residual /
K-concatenated and per-group inputs, conv3x3 (pool-on-read, concat-on-read), ConvTranspose2d(2, 2), the whole U-Net and both
encoders, against float64 torch on the same inputs.  fp32 semantics through 3xTF32: errors are a few 1e-7 of the term scale."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from ifdefense_b200 import capi, models, synth, tc
from oracle import torch_port as tp

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 2, 4], ids=lambda n: "cluster%d" % n)
def cluster(request):
    """CTAs per thread-block cluster of the engine: 1 = no cluster, 2 / 4 = weight chunks TMA-multicast across the cluster."""
    L = capi.lib()
    L.ifd_test_hook(6, request.param)
    yield request.param
    L.ifd_test_hook(6, 2)


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def rel_err(got, want, scale):
    return float(((got.double().cpu() - want.cpu()).abs() / scale.cpu()).max())


@pytest.mark.parametrize("M,K,N", [(1, 32, 32), (77, 64, 32), (128, 96, 40), (1000, 3, 64), (4099, 256, 256), (300, 1024, 512),
                                   (513, 70, 100), (20000, 64, 32)])
def test_linear_vs_float64(M, K, N, cluster):
    x, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    got = tc.linear([(x, K, False)], tc.pack(w), N, bias=b)
    want = x.double() @ w.double().T + b.double()
    scale = x.double().abs() @ w.double().abs().T + b.double().abs()
    assert got.shape == (M, N) and rel_err(got, want, scale) < 2e-6
    again = tc.linear([(x, K, False)], tc.pack(w), N, bias=b)
    assert torch.equal(got, again)                                          # bitwise reproducible


def test_linear_fused_relu_residual_segments_groups():
    """ResnetBlockFC as the engine runs it: fc_0(relu(cat(a, b))) and [fc_1 | shortcut] . [relu(h) | a | b] with b a per-group
    vector expanded on read."""
    G, T, Ha = 5, 60, 64
    a, pooled = rnd(G * T, Ha, seed=1), rnd(G, Ha, seed=2)
    w0, b0 = rnd(32, 2 * Ha, seed=3, scale=0.1), rnd(32, seed=4)
    w1, ws, b1 = rnd(48, 32, seed=5, scale=0.2), rnd(48, 2 * Ha, seed=6, scale=0.1), rnd(48, seed=7)
    xcat = torch.cat([a, pooled.repeat_interleave(T, 0)], 1).double()
    h_want = F.relu(xcat) @ w0.double().T + b0.double()
    h = tc.linear([(a, Ha, True), (pooled, Ha, True, T)], tc.pack(w0, [Ha, Ha]), 32, bias=b0)
    assert (h.double() - h_want).abs().max() < 1e-5
    out = tc.linear([(h, 32, True), (a, Ha, False), (pooled, Ha, False, T)], tc.pack(torch.cat([w1, ws], 1), [32, Ha, Ha]), 48, bias=b1)
    want = F.relu(h.double()) @ w1.double().T + b1.double() + xcat @ ws.double().T
    assert (out.double() - want).abs().max() < 2e-5
    res = rnd(G * T, 48, seed=8)
    out2 = tc.linear([(h, 32, True)], tc.pack(w1), 48, bias=b1, resid=res, relu_out=True)
    want2 = F.relu(F.relu(h.double()) @ w1.double().T + b1.double() + res.double())
    assert (out2.double() - want2).abs().max() < 1e-5
    # strided output view (a column block of a wider matrix)
    wide = torch.zeros((G * T, 96), device="cuda")
    tc.linear([(h, 32, True)], tc.pack(w1), 48, bias=b1, out=wide[:, 32:80])
    assert torch.equal(wide[:, 32:80], tc.linear([(h, 32, True)], tc.pack(w1), 48, bias=b1)) and float(wide[:, :32].abs().max()) == 0.0


@pytest.mark.parametrize("B,H,W,C0,C1,Cout,pool", [(2, 8, 8, 32, 0, 32, False), (3, 16, 16, 64, 64, 64, False), (2, 8, 8, 32, 0, 64, True),
                                                   (1, 64, 64, 32, 0, 32, False), (5, 8, 8, 256, 0, 256, False), (2, 16, 16, 128, 128, 128, False)])
def test_conv3x3_vs_float64(B, H, W, C0, C1, Cout, pool, cluster):
    Hs, Ws = (2 * H, 2 * W) if pool else (H, W)
    a = rnd(B, Hs, Ws, C0, seed=1)
    b = rnd(B, H, W, C1, seed=2) if C1 else None
    w = rnd(Cout, C0 + C1, 3, 3, seed=3, scale=(9 * (C0 + C1)) ** -0.5)
    bias = rnd(Cout, seed=4)
    got = tc.conv3x3(a, tc.pack(w.permute(0, 2, 3, 1).reshape(Cout, -1)), bias, Cout, src1=b, pool=pool, relu=True)
    x = a.permute(0, 3, 1, 2).double()
    if pool:
        x = F.max_pool2d(x, 2, 2)
    if b is not None:
        x = torch.cat((x, b.permute(0, 3, 1, 2).double()), 1)
    want = F.relu(F.conv2d(x, w.double(), bias.double(), padding=1)).permute(0, 2, 3, 1)
    assert got.shape == (B, H, W, Cout)
    scale = F.conv2d(x.abs(), w.double().abs(), bias.double().abs(), padding=1).permute(0, 2, 3, 1)      # sum |a||w|
    err = float(((got.double() - want).abs() / scale).max())
    print("conv K = %d: max error / sum|a||w| = %.2e, max abs error %.2e" % (9 * (C0 + C1), err, float((got.double() - want).abs().max())))
    assert err < 2e-6
    # the same convolution on feature maps in the engine's warp-transposed layout: the same bits
    ab = tc.Blocked.from_rows(a)
    bb = tc.Blocked.from_rows(b) if b is not None else None
    gotb = tc.conv3x3(ab, tc.pack(w.permute(0, 2, 3, 1).reshape(Cout, -1)), bias, Cout, src1=bb, pool=pool, relu=True, blocked_out=True)
    assert gotb.dims == (B, H, W) and torch.equal(gotb.to_rows().view(B, H, W, Cout), got)


def test_convtranspose_and_unet_vs_torch(cluster):
    cin, cout, B, H, W = 64, 32, 3, 8, 8
    x = rnd(B, H, W, cin, seed=1)
    w, bias = rnd(cin, cout, 2, 2, seed=2, scale=0.1), rnd(cout, seed=3)
    got = tc.linear([(x.view(-1, cin), cin, False)], tc.pack(w.permute(2, 3, 1, 0).reshape(4 * cout, cin)), 4 * cout, bias=bias,
                    shuffle=(cout, H, W))
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2).double(), w.double(), bias.double(), stride=2).permute(0, 2, 3, 1)
    assert got.shape == (B, 2 * H, 2 * W, cout) and float((got.double() - want).abs().max()) < 1e-5
    upb = tc.Blocked(B * 4 * H * W, cout, x.device, dims=(B, 2 * H, 2 * W))
    tc.linear([(tc.Blocked.from_rows(x), cin, False)], tc.pack(w.permute(2, 3, 1, 0).reshape(4 * cout, cin)), 4 * cout, bias=bias,
              shuffle=(cout, H, W), out=upb)
    assert torch.equal(upb.to_rows().view(B, 2 * H, 2 * W, cout), got)
    # the whole U-Net of the shipped config (depth 4, 32 -> 256 channels) against the float64 torch forward
    torch.manual_seed(0)
    net = models.UNet(32, in_channels=32, depth=4, start_filts=32).cuda().eval()
    for p in net.parameters():
        p.requires_grad = False
    xin = rnd(4, 64, 64, 32, seed=5)
    got = net.forward_cl(xin)
    want = net.double()(xin.permute(0, 3, 1, 2).double()).permute(0, 2, 3, 1)
    net.float()
    assert got.shape == (4, 64, 64, 32)
    ref32 = net(xin.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)                # torch / cuDNN (whatever its default precision is): for the printout
    err, err32 = float((got.double() - want).abs().max()), float((ref32.double() - want).abs().max())
    print("U-Net 4 x 64^2: max |err| %.2e (torch fp32: %.2e), max |out| %.3f" % (err, err32, float(want.abs().max())))
    assert err < 2e-5 * float(want.abs().max())
    assert torch.equal(got, net.forward_cl(xin))


def test_encoders_match_the_oracle_within_1e5_and_do_not_depend_on_the_batch():
    """Both encoders on the library's kernels against the oracle (stock torch CPU): <= 1e-5 relative, and a
    cloud's code does not depend on the batch it arrives in or on its position in it (the reason encoder_chunk existed)."""
    sd = models.synthetic_state_dict("convonet", 0)
    model = models.build_convonet()
    model.load_state_dict(sd)
    model = model.cuda().eval()
    case = synth.make_case(5, K=64, seed=1)
    with torch.no_grad():
        got = model.encode_inputs(case.sel.cuda())
        want = tp.convonet_encode(sd, case.sel)
        for k in ("xz", "xy", "yz"):
            d = float((got[k].cpu() - want[k]).abs().max())
            assert d < 1e-5 * float(want[k].abs().max()), (k, d)
        assert got.channels_last.shape == (3, 5, 64, 64, 32)
        part = model.encode_inputs(case.sel[3:5].cuda())                  # clouds 3, 4 alone, at positions 0, 1
        rev = model.encode_inputs(case.sel.flip(0).cuda())                # every cloud at another position
        for k in ("xz", "xy", "yz"):
            assert torch.equal(part[k], got[k][3:5]) and torch.equal(rev[k].flip(0), got[k])
    sd2 = models.synthetic_state_dict("onet", 0)
    m2 = models.build_onet()
    m2.load_state_dict(sd2)
    m2 = m2.cuda().eval()
    oc = synth.make_onet_case(4, K=64, seed=2)
    with torch.no_grad():
        c = m2.encode_inputs(oc.sel.cuda())
        wantc = tp.onet_encode(sd2, oc.sel)
        assert c.shape == (4, 512)
        assert float((c.cpu() - wantc).abs().max()) < 1e-5 * float(wantc.abs().max())
        assert torch.equal(m2.encode_inputs(oc.sel[2:3].cuda()), c[2:3])
