"""Generate tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) on CPU.

Run in the build container only (the GPU box has no reference tree):
    python tests/golden/make_golden.py
Single-threaded torch: the multi-threaded CPU reference is not run-to-run reproducible over long
trajectories ("Parity tiers").  Inputs follow the synthetic protocol of ifdefense_b200/synth.py.

Files
  geometry.npz   knn_point / RepulsionLoss fwd+bwd / DGCNN knn / FPS / ball query / SOR on 4 clouds
                 (cloud 0 = the reference's own airplane.npy), plus a duplicate-point cloud
  convonet.npz   B=2, K=256: planes, decoder weights, logits, d(sum gl*logit)/dp, clamp cases,
                 20-step optimize_points trace, a late-state single-step case (Adam state at t=150)
  onet.npz       ONet: B=2, K=256 logits / gradients / 20-step trace, and BASELINE configs[0]
                 (1 cloud x 1024 points, 20 iterations); weights regenerated from the seed, pinned by checksum
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402
from ifdefense_b200 import synth, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(1)


def load_baseline_module(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(ref_import.REF_ROOT, "baselines", "model", name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def with_fixed_randint(values, fn):
    """The reference seeds FPS with torch.randint; return `values` from it for the duration of fn."""
    orig = torch.randint
    torch.randint = lambda *a, **k: values.clone()
    try:
        return fn()
    finally:
        torch.randint = orig


def geometry(ns):
    raw = synth.clouds(4)
    pts = []
    for i in range(4):
        allp, _ = synth.driver.preprocess_pc(raw[i], None, 0.9)
        pts.append(allp)
    base = torch.from_numpy(np.stack(pts))
    g = torch.Generator().manual_seed(11)
    x = torch.clamp(base + torch.randn(base.shape, generator=g) * 0.01, -0.45, 0.45)   # like init_points noise
    out = {"xyz": x.numpy()}
    out["knn5"] = ns.pn_utils.knn_point(5, x).numpy().astype(np.int32)
    xr = x.clone().requires_grad_()
    loss = ns.repulsion.RepulsionLoss()(xr)
    gl = torch.tensor([1.0, 0.5, 2.0, 125.0])
    (loss * gl).sum().backward()
    out["rep_loss"], out["rep_grad_loss"], out["rep_grad"] = loss.detach().numpy(), gl.numpy(), xr.grad.numpy()
    # duplicate points: rows 0/1 and 7/9 of a small cloud are exact copies
    d = x[:1, :64].clone()
    d[0, 1] = d[0, 0]
    d[0, 9] = d[0, 7]
    out["dup_xyz"] = d.numpy()
    dr = d.clone().requires_grad_()
    ld = ns.repulsion.RepulsionLoss()(dr)
    ld.sum().backward()
    out["dup_rep_loss"], out["dup_rep_grad"] = ld.detach().numpy(), dr.grad.numpy()
    # classifier-side ops (baselines/model)
    dg, pn2 = load_baseline_module("dgcnn"), load_baseline_module("pointnet2")
    out["dgcnn_knn20"] = dg.knn(x.transpose(2, 1).contiguous(), 20).numpy().astype(np.int32)
    start = torch.tensor([5, 17, 100, 1023])
    fps1 = with_fixed_randint(start, lambda: pn2.farthest_point_sample(x, 512))
    out["fps_start"], out["fps512"] = start.numpy().astype(np.int32), fps1.numpy().astype(np.int32)
    new_xyz = pn2.index_points(x, fps1)
    out["ball_0.2_32"] = pn2.query_ball_point(0.2, 32, x, new_xyz).numpy().astype(np.int32)
    start2 = torch.tensor([0, 3, 200, 511])
    fps2 = with_fixed_randint(start2, lambda: pn2.farthest_point_sample(new_xyz, 128))
    out["fps2_start"], out["fps128"] = start2.numpy().astype(np.int32), fps2.numpy().astype(np.int32)
    out["ball_0.4_64"] = pn2.query_ball_point(0.4, 64, new_xyz, pn2.index_points(new_xyz, fps2)).numpy().astype(np.int32)
    # defense-side FPS (ConvONet/defense/pn_utils.py) and SOR on the raw (un-noised) clouds with outliers
    fps3 = with_fixed_randint(start, lambda: ns.pn_utils.farthest_point_sample(x, 64))
    out["fps64_defense"] = fps3.numpy().astype(np.int32)
    rawt = torch.from_numpy(raw)
    kept = ns.sor.SORDefense(k=2, alpha=1.1)(rawt)
    mask = np.zeros(raw.shape[:2], dtype=np.uint8)
    for b in range(4):                       # recover the mask from the kept rows (all rows are distinct)
        keys = {tuple(r) for r in kept[b].numpy().tolist()}
        mask[b] = [tuple(r) in keys for r in raw[b].tolist()]
        assert mask[b].sum() == len(kept[b])
    out["sor_xyz"], out["sor_keep"] = raw, mask
    np.savez_compressed(os.path.join(OUT, "geometry.npz"), **out)
    print("geometry.npz", {k: v.shape for k, v in out.items()})


def convonet_case(ns):
    B, K = 2, 256
    case = synth.make_case(B, K=K, seed=1)
    cfg, model = ref_import.build_model(ns)
    model.load_state_dict(case.sd, strict=True)
    with torch.no_grad():
        c = model.encode_inputs(case.sel)                       # the reference's own encoder output
    for k in c:
        assert torch.equal(c[k], case.c[k]), "product encoder shell != reference encoder on CPU"
    out = {"planes_nchw": np.stack([c[k].numpy() for k in ("xz", "xy", "yz")]),
           "dec_blob": weights.pack_convonet_decoder(case.sd), "p0": case.p0.numpy(), "sel": case.sel.numpy()}
    for k, v in case.sd.items():
        if k.startswith("decoder."):
            out["sd/" + k] = v.numpy()
    # decode forward / backward
    gl = torch.randn(B, K, generator=torch.Generator().manual_seed(5))
    p = case.p0.clone().requires_grad_()
    logits = model.decode(p, c).logits
    (logits * gl).sum().backward()
    out["logits"], out["gl"], out["grad_p"] = logits.detach().numpy(), gl.numpy(), p.grad.numpy()
    # clamp / border cases (normalize_coordinate overwrites, grid_sample border)
    pc = case.p0.clone()
    pc[:, :40] *= 1.6
    pc[:, 40:48] = 0.55
    pc[:, 48:56] = -0.55
    pc[:, 56:60] = 0.5500055
    pc[:, 60:64] = -0.5500055
    pc = pc.requires_grad_()
    lc = model.decode(pc, c).logits
    (lc * gl).sum().backward()
    out["clamp_p"], out["clamp_logits"], out["clamp_grad_p"] = pc.detach().numpy(), lc.detach().numpy(), pc.grad.numpy()

    # optimize_points, restated verbatim around the reference's own objects (opt_defense.py:182-239)
    def run(points, iterations, m0=None, v0=None, t0=0, hook=None):
        opt_points = points.clone().float()
        opt_points.requires_grad_()
        Bc, Kc = opt_points.shape[:2]
        thr = torch.ones((Bc, Kc)).float() * 0.2
        opt = torch.optim.Adam([opt_points], lr=0.001)
        if m0 is not None:
            opt.state[opt_points] = {"step": torch.tensor(float(t0)), "exp_avg": m0.clone(), "exp_avg_sq": v0.clone()}
        stats = []
        for i in range(iterations + 1):
            occ_value = model.decode(opt_points, c).logits
            occ_loss = F.binary_cross_entropy_with_logits(occ_value, thr, reduction='none')
            occ_loss = torch.mean(occ_loss) * Kc
            rep_loss = torch.mean(ns.repulsion.repulsion_loss(opt_points)) * 500.
            loss = occ_loss + rep_loss
            opt.zero_grad()
            loss.backward()
            g = opt_points.grad.detach().clone()
            opt.step()
            if hook:
                hook(i, opt_points, g, opt.state[opt_points])
            if i % 100 == 0:
                stats.append([loss.item(), occ_loss.item(), rep_loss.item(), torch.sigmoid(occ_value).mean().item()])
        return opt_points.detach(), np.asarray(stats)

    trace = {}

    def hook(i, p_, g, st):
        if i in (0, 1, 9, 19):
            trace["xyz_%d" % i] = p_.detach().clone().numpy()
            trace["grad_%d" % i] = g.numpy()
        if i == 149:
            trace["late_xyz"], trace["late_m"], trace["late_v"] = (p_.detach().clone().numpy(), st["exp_avg"].clone().numpy(),
                                                                  st["exp_avg_sq"].clone().numpy())
        if i == 150:
            trace["late_xyz_next"] = p_.detach().clone().numpy()

    final, stats = run(case.p0, 200, hook=hook)
    out.update({"trace/" + k: v for k, v in trace.items()})
    out["stats_201"] = stats
    out["final_201_raw"] = final.numpy().copy()
    f20, _ = run(case.p0, 19)
    cen = torch.mean(f20, dim=1)
    f20 = f20 - cen[:, None, :]
    f20 = f20 / torch.max(torch.sum(f20 ** 2, dim=2) ** 0.5, dim=1)[0][:, None, None]
    out["final_20_normalized"] = f20.numpy()
    np.savez_compressed(os.path.join(OUT, "convonet.npz"), **out)
    print("convonet.npz", {k: v.shape for k, v in out.items() if not k.startswith("sd/")})


def onet_case():
    """ONet-Opt: B=2, K=256 through the reference's OccupancyNetwork (eval mode, z_dim = 0), plus BASELINE.json
    configs[0] (1 cloud x 1024 points, 20 iterations, CPU).  The 3.5 M decoder weights are NOT stored: they are
    regenerated from the seed (ifdefense_b200.models.synthetic_state_dict) and pinned by a checksum."""
    ns = ref_import.load("ONet")
    _, model = ref_import.build_model(ns)
    out = {}
    for tag, (B, K, iters) in {"b2": (2, 256, 19), "cfg0": (1, 1024, 20)}.items():
        case = synth.make_onet_case(B, K=K, seed=4)
        model.load_state_dict(case.sd, strict=True)
        with torch.no_grad():
            c = model.encode_inputs(case.sel)
        assert torch.equal(c, case.c), "product ONet encoder shell != reference encoder on CPU"
        z = model.get_z_from_prior((B,), sample=False)
        blob = weights.pack_onet_decoder(case.sd)
        out["weights_checksum"] = np.array([np.sum(blob.astype(np.float64)), np.sum(np.abs(blob).astype(np.float64)), blob[12345]])
        out[tag + "/c"], out[tag + "/p0"], out[tag + "/sel"] = c.numpy(), case.p0.numpy(), case.sel.numpy()
        gl = torch.randn(B, K, generator=torch.Generator().manual_seed(6))
        p = case.p0.clone().requires_grad_()
        logits = model.decode(p, z, c).logits
        (logits * gl).sum().backward()
        out[tag + "/logits"], out[tag + "/gl"], out[tag + "/grad_p"] = logits.detach().numpy(), gl.numpy(), p.grad.numpy()
        pts = case.p0.clone().float().requires_grad_()
        thr = torch.ones((B, K)).float() * 0.2
        opt = torch.optim.Adam([pts], lr=0.001)
        for i in range(iters + 1):
            occ = model.decode(pts, z, c).logits
            loss = torch.mean(F.binary_cross_entropy_with_logits(occ, thr, reduction='none')) * K + \
                torch.mean(ns.repulsion.repulsion_loss(pts)) * 500.
            opt.zero_grad()
            loss.backward()
            if i == 0:
                out[tag + "/grad_0"] = pts.grad.detach().clone().numpy()
            opt.step()
            if i in (0, 1, 9):
                out[tag + "/xyz_%d" % i] = pts.detach().clone().numpy()
        f = pts.detach()
        out[tag + "/final_raw"] = f.numpy().copy()
        f = f - torch.mean(f, dim=1)[:, None, :]
        f = f / torch.max(torch.sum(f ** 2, dim=2) ** 0.5, dim=1)[0][:, None, None]
        out[tag + "/final_normalized"] = f.numpy()
    np.savez_compressed(os.path.join(OUT, "onet.npz"), **out)
    print("onet.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    which = sys.argv[1:] or ["geometry", "convonet", "onet"]
    if "geometry" in which or "convonet" in which:
        ns = ref_import.load("ConvONet")
        if "geometry" in which:
            geometry(ns)
        if "convonet" in which:
            convonet_case(ns)
    if "onet" in which:
        onet_case()
