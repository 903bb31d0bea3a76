"""Synthetic workload protocol: the reference ships neither data nor weights, so benchmarks
and parity tests run on seeded synthetic clouds of ModelNet40 shape (1024 points) and seeded random-init
weights with the reference's parameter names.  Everything here is deterministic given the seeds."""
import os
import types

import numpy as np
import torch

from . import driver, models

_AIRPLANE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "airplane.npy")


def synth_cloud(i, n=1024):
    """Cloud i: points on a random ellipsoid (even i) or tapered box (odd i) surface, radii U[0.3,1], 5 % of
    the points displaced by N(0, 0.02^2) as 'adversarial' noise."""
    r = np.random.default_rng(1000 + i)
    radii = r.uniform(0.3, 1.0, size=3)
    if i % 2 == 0:
        v = r.normal(size=(n, 3))
        pts = v / np.linalg.norm(v, axis=1, keepdims=True) * radii
    else:
        pts = r.uniform(-1, 1, size=(n, 3))
        ax = r.integers(0, 3, size=n)
        pts[np.arange(n), ax] = np.sign(pts[np.arange(n), ax])
        pts *= radii * (1.0 - 0.3 * (pts[:, 2:3] * 0.5 + 0.5))
    k = n // 20
    sel = r.choice(n, k, replace=False)
    pts[sel] += r.normal(scale=0.02, size=(k, 3))
    return pts.astype(np.float32)


def clouds(B, n=1024, airplane=True):
    """[B,n,3]; cloud 0 is the one real ModelNet40 shape the reference ships (baselines/data/airplane.npy),
    a thin-sheet hard case for kNN repulsion."""
    out = [synth_cloud(i, n) for i in range(B)]
    if airplane and n == 1024 and os.path.exists(_AIRPLANE):
        out[0] = np.load(_AIRPLANE).astype(np.float32)
    return np.stack(out)


def make_case(B, K=1024, seed=0, T=600, device="cpu", sd=None):
    """One ConvONet-Opt batch: weights, encoder input, feature planes `c` (as the reference's encode_inputs
    returns them) and init points.  Encoding runs with the product's torch encoder on `device`."""
    sd = sd if sd is not None else models.synthetic_state_dict("convonet", 0)
    raw = clouds(B)
    proc = [driver.preprocess_pc(raw[i], num_points=T, padding_scale=0.9, rng=np.random.default_rng(2000 + i)) for i in range(B)]
    sel = torch.from_numpy(np.stack([s for _, s in proc]))
    model = models.build_convonet()
    model.load_state_dict(sd)
    model = model.to(device).eval()
    with torch.no_grad():
        c = {k: v.cpu() for k, v in model.encoder(sel.to(device)).items()}
    gen = torch.Generator().manual_seed(3000 + seed)
    p0 = driver.init_points([p for p, _ in proc], npoint=K, sigma=0.01, padding_scale=0.9, gen=gen)
    return types.SimpleNamespace(sd=sd, raw=raw, sel=sel, c=c, p0=p0, B=B, K=K)


def make_onet_case(B, K=1024, seed=0, T=300, device="cpu", sd=None):
    """One ONet-Opt batch: weights, encoder input [B,300,3], latent codes c [B,512] (what the reference's
    encode_inputs returns), init points."""
    sd = sd if sd is not None else models.synthetic_state_dict("onet", 0)
    raw = clouds(B)
    proc = [driver.preprocess_pc(raw[i], num_points=T, padding_scale=0.9, rng=np.random.default_rng(2000 + i)) for i in range(B)]
    sel = torch.from_numpy(np.stack([s for _, s in proc]))
    model = models.build_onet()
    model.load_state_dict(sd)
    model = model.to(device).eval()
    with torch.no_grad():
        c = model.encoder(sel.to(device)).cpu()
    gen = torch.Generator().manual_seed(3000 + seed)
    p0 = driver.init_points([p for p, _ in proc], npoint=K, sigma=0.01, padding_scale=0.9, gen=gen)
    return types.SimpleNamespace(sd=sd, raw=raw, sel=sel, c=c, p0=p0, B=B, K=K)


def onet_with_surface(sd, c, occupied=0.25, resolution=16):
    """Random-init ONet weights put the whole box on one side of the threshold, which makes mesh extraction trivial.  Returns a
    copy of `sd` whose `decoder.fc_out.bias` is shifted so that the logit(0.2) level set encloses about `occupied` of the box
    for the latent code c (a wiggly closed surface: a harder case than a real shape for MISE and marching cubes)."""
    from . import onet as onet_mod
    dec = onet_mod.ONetDecoder(sd)
    g = dec.eval_dense_grid(c, resolution=resolution).flatten().float().cpu()
    thr = float(np.log(0.2) - np.log(0.8))
    shift = thr - float(torch.quantile(g, 1.0 - occupied))
    out = {k: v.clone() for k, v in sd.items()}
    out["decoder.fc_out.bias"] = out["decoder.fc_out.bias"] + shift
    return out


def fill_classifier_state_(model, seed=0):
    """Deterministic non-trivial weights for a victim classifier, a function of the parameter NAMES and shapes only, so that
    the reference model and its mirror receive identical tensors (no victim checkpoint ships: baselines/README.md:23,31).
    Weights ~ N(0, 1/fan_in), BatchNorm scale U(0.5, 1.5), shift and running mean N(0, 0.1^2), running variance U(0.5, 1.5)."""
    import zlib
    sd = model.state_dict()
    for name in sorted(sd):
        t = sd[name]
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
        if name.endswith("num_batches_tracked"):
            continue
        if name.endswith("running_var"):
            t.copy_(torch.rand(t.shape, generator=g) + 0.5)
        elif name.endswith("running_mean"):
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        elif t.dim() == 1 and "bn" in name and name.endswith("weight"):
            t.copy_(torch.rand(t.shape, generator=g) + 0.5)
        elif t.dim() == 1:
            t.copy_(torch.randn(t.shape, generator=g) * 0.1)
        else:
            fan_in = int(np.prod(t.shape[1:]))
            t.copy_(torch.randn(t.shape, generator=g) / np.sqrt(fan_in))
    return model
