"""TEST INFRASTRUCTURE ONLY -- the oracle.  Not imported by the product package.

CPU restatement, in stock PyTorch ops, of the reference's optimisation-based restoration path.
Every function cites the reference lines it restates.  It uses the *same ATen op sequence* as the
reference so that, in this container, it is pinned bit-for-bit against the unmodified reference
classes (tests/test_oracle_vs_reference.py, skipped where /root/reference is absent) and against the
committed fixtures in tests/golden/ (generated from the real reference by tests/golden/make_golden.py).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import it.

Weights are passed as a plain dict keyed by the reference's state_dict names,
e.g. 'decoder.fc_p.weight'.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------------------------
# geometry helpers: ConvONet/defense/pn_utils.py
# ----------------------------------------------------------------------------------------------


def index_points(points, idx):
    """pn_utils.py:6-23 -- batched gather points[b, idx[b, ...], :]."""
    B = points.shape[0]
    shape = [B] + [1] * (idx.dim() - 1)
    b = torch.arange(B, dtype=torch.long, device=points.device).view(shape).expand_as(idx)
    return points[b, idx, :]


def knn_point(k, points):
    """pn_utils.py:64-83 -- expanded-form squared distances, topk(k+1) of -dist, column 0 dropped."""
    pc = points.clone().detach().transpose(2, 1)                    # [B,3,K]
    inner = -2. * torch.matmul(pc.transpose(2, 1), pc)              # [B,K,K]
    xx = torch.sum(pc ** 2, dim=1, keepdim=True)                    # [B,1,K]
    dist = xx + inner + xx.transpose(2, 1)
    assert dist.min().item() >= -1e-4
    _, top = (-dist).topk(k=k + 1, dim=-1)
    return top[:, :, 1:]


def knn_dist_matrix(points):
    """The [B,K,K] matrix knn_point ranks (pn_utils.py:76-78); exposed for tie analysis in tests."""
    pc = points.clone().detach().transpose(2, 1)
    inner = -2. * torch.matmul(pc.transpose(2, 1), pc)
    xx = torch.sum(pc ** 2, dim=1, keepdim=True)
    return xx + inner + xx.transpose(2, 1)


def farthest_point_sample(xyz, num_point, start):
    """pn_utils.py:26-48 / baselines/model/pointnet2.py:53-74.  The reference draws the first index with
    torch.randint; here it is an input (`start`, [B] long) so that both paths see the same value."""
    B, N, _ = xyz.shape
    cent = torch.zeros(B, num_point, dtype=torch.long)
    distance = torch.ones(B, N) * 1e10
    far = start.clone().long()
    bi = torch.arange(B, dtype=torch.long)
    for i in range(num_point):
        cent[:, i] = far
        c = xyz[bi, far, :].view(B, 1, 3)
        d = torch.sum((xyz - c) ** 2, -1)
        m = d < distance
        distance[m] = d[m]
        far = torch.max(distance, -1)[1]
    return cent


def square_distance(src, dst):
    """baselines/model/pointnet2.py:9-30 -- -2 src dst^T, then += |src|^2, += |dst|^2 in that order."""
    B, N, _ = src.shape
    M = dst.shape[1]
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


def query_ball_point(radius, nsample, xyz, new_xyz):
    """baselines/model/pointnet2.py:77-98 -- first `nsample` ascending indices with d2 <= r2, padded with
    the first hit."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    gi = torch.arange(N, dtype=torch.long).view(1, 1, N).repeat([B, S, 1])
    sq = square_distance(new_xyz, xyz)
    gi[sq > radius ** 2] = N
    gi = gi.sort(dim=-1)[0][:, :, :nsample]
    first = gi[:, :, 0].view(B, S, 1).repeat([1, 1, nsample])
    mask = gi == N
    gi[mask] = first[mask]
    return gi


def dgcnn_knn(x, k):
    """baselines/model/dgcnn.py:7-13 -- x is [B,C,N]; top-k (self included) of -|xi-xj|^2 in the
    reference's association: -xx - inner - xx^T."""
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    pd = -xx - inner - xx.transpose(2, 1)
    return pd.topk(k=k, dim=-1)[1]


def dgcnn_pd_matrix(x):
    inner = -2 * torch.matmul(x.transpose(2, 1), x)
    xx = torch.sum(x ** 2, dim=1, keepdim=True)
    return -xx - inner - xx.transpose(2, 1)


# ----------------------------------------------------------------------------------------------
# losses / pre-step: defense/repulsion_loss.py, defense/SOR.py
# ----------------------------------------------------------------------------------------------

REP_NN, REP_RADIUS, REP_H, REP_EPS = 5, 0.07, 0.03, 1e-12          # repulsion_loss.py:9-10


def repulsion_loss(pred, idx=None):
    """repulsion_loss.py:18-54 -> [B].  idx may be supplied (e.g. from the C oracle) to decouple the
    float part from topk's tie order."""
    if idx is None:
        with torch.no_grad():
            idx = knn_point(REP_NN, pred)
    g = index_points(pred, idx) - pred.unsqueeze(-2)               # [B,N,k,3]
    d2 = torch.sum(g ** 2, dim=-1)
    d2 = torch.max(d2, torch.tensor(REP_EPS))
    d = torch.sqrt(d2)
    w = torch.exp(-((d / REP_H) ** 2))
    return torch.mean((REP_RADIUS - d) * w, dim=[1, 2])


def sor_outlier_removal(x, k=2, alpha=1.1):
    """SOR.py:22-49 -- float64 kNN-k mean distance; keep value <= mean + alpha*std (unbiased)."""
    pc = x.clone().detach().double().transpose(2, 1)
    inner = -2. * torch.matmul(pc.transpose(2, 1), pc)
    xx = torch.sum(pc ** 2, dim=1, keepdim=True)
    dist = xx + inner + xx.transpose(2, 1)
    assert dist.min().item() >= -1e-6
    neg, _ = (-dist).topk(k=k + 1, dim=-1)
    value = torch.mean(-(neg[..., 1:]), dim=-1)
    thr = torch.mean(value, dim=-1) + alpha * torch.std(value, dim=-1)
    mask = value <= thr[:, None]
    return [x[i][mask[i]] for i in range(x.shape[0])], mask, value


# ----------------------------------------------------------------------------------------------
# ConvONet decoder: src/common.py:235-258, src/conv_onet/models/decoder.py:50-95, src/layers.py:39-48
# ----------------------------------------------------------------------------------------------

PLANE_AXES = {"xz": [0, 2], "xy": [0, 1], "yz": [1, 2]}             # common.py:243-248


def normalize_coordinate(p, padding=0.1, plane="xz"):
    """common.py:235-258 (in-place clamps; the masked assignment kills the gradient there)."""
    xy = p[:, :, PLANE_AXES[plane]]
    xy = xy / (1 + padding + 10e-6)
    xy = xy + 0.5
    if xy.max() >= 1:
        xy[xy >= 1] = 1 - 10e-6
    if xy.min() < 0:
        xy[xy < 0] = 0.0
    return xy


def normalize_3d_coordinate(p, padding=0.1):
    """common.py:260-276 (the 'grid' feature volume; in-place clamps as in the 2-D function, constants 10e-4)."""
    p_nor = p / (1 + padding + 10e-4)
    p_nor = p_nor + 0.5
    if p_nor.max() >= 1:
        p_nor[p_nor >= 1] = 1 - 10e-4
    if p_nor.min() < 0:
        p_nor[p_nor < 0] = 0.0
    return p_nor


def coordinate2index(x, reso, coord_type="2d"):
    """common.py:300-315."""
    x = (x * reso).long()
    if coord_type == "2d":
        index = x[:, :, 0] + reso * x[:, :, 1]
    else:                                                           # '3d': the grid volume
        index = x[:, :, 0] + reso * (x[:, :, 1] + reso * x[:, :, 2])
    return index[:, None, :]


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def _resblock_fc(sd, name, x):
    """layers.py:39-48."""
    net = _lin(sd, name + ".fc_0", F.relu(x))
    dx = _lin(sd, name + ".fc_1", F.relu(net))
    xs = _lin(sd, name + ".shortcut", x) if (name + ".shortcut.weight") in sd else x
    return xs + dx


def convonet_decode(sd, p, c_plane, padding=0.1, n_blocks=5):
    """LocalDecoder.forward (decoder.py:69-95) -> logits [B,K].  c_plane: dict plane -> [B,C,R,R]."""
    c = 0
    if "grid" in c_plane:                                           # decoder.py:59-67, 72-73: trilinear sample of the volume
        p_nor = normalize_3d_coordinate(p.clone(), padding=padding)
        vgrid = 2.0 * p_nor[:, :, None, None].float() - 1.0
        c = c + F.grid_sample(c_plane["grid"], vgrid, padding_mode="border", align_corners=True,
                              mode="bilinear").squeeze(-1).squeeze(-1)
    for plane in ("xz", "xy", "yz"):                                # decoder.py:75-80 order
        if plane in c_plane:
            xy = normalize_coordinate(p.clone(), padding=padding, plane=plane)
            vgrid = 2.0 * xy[:, :, None].float() - 1.0
            c = c + F.grid_sample(c_plane[plane], vgrid, padding_mode="border", align_corners=True,
                                  mode="bilinear").squeeze(-1)
    c = c.transpose(1, 2)
    net = _lin(sd, "decoder.fc_p", p.float())
    for i in range(n_blocks):
        net = net + _lin(sd, "decoder.fc_c.%d" % i, c)
        net = _resblock_fc(sd, "decoder.blocks.%d" % i, net)
    return _lin(sd, "decoder.fc_out", F.relu(net)).squeeze(-1)


# ----------------------------------------------------------------------------------------------
# ONet decoder (eval mode): im2mesh/onet/models/decoder.py:115-133, im2mesh/layers.py:98-107,226-242
# ----------------------------------------------------------------------------------------------


def _cbn(sd, name, x, c):
    """CBatchNorm1d.forward in eval mode (layers.py:226-242)."""
    c3 = c.unsqueeze(2) if c.dim() == 2 else c
    gamma = F.conv1d(c3, sd[name + ".conv_gamma.weight"], sd[name + ".conv_gamma.bias"])
    beta = F.conv1d(c3, sd[name + ".conv_beta.weight"], sd[name + ".conv_beta.bias"])
    net = F.batch_norm(x, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], None, None,
                       False, 0.1, 1e-5)
    return gamma * net + beta


def _cresblock(sd, name, x, c):
    """CResnetBlockConv1d.forward (layers.py:98-107), size_in == size_out so no shortcut."""
    net = F.conv1d(F.relu(_cbn(sd, name + ".bn_0", x, c)), sd[name + ".fc_0.weight"], sd[name + ".fc_0.bias"])
    dx = F.conv1d(F.relu(_cbn(sd, name + ".bn_1", net, c)), sd[name + ".fc_1.weight"], sd[name + ".fc_1.bias"])
    return x + dx


def onet_decode(sd, p, c):
    """DecoderCBatchNorm.forward with z_dim == 0 (decoder.py:115-133) -> logits [B,K]."""
    net = F.conv1d(p.transpose(1, 2), sd["decoder.fc_p.weight"], sd["decoder.fc_p.bias"])
    for i in range(5):
        net = _cresblock(sd, "decoder.block%d" % i, net, c)
    out = F.conv1d(F.relu(_cbn(sd, "decoder.bn", net, c)), sd["decoder.fc_out.weight"],
                   sd["decoder.fc_out.bias"])
    return out.squeeze(1)


# ----------------------------------------------------------------------------------------------
# encoders (once per batch)
# ----------------------------------------------------------------------------------------------


def _scatter_max_gather(c, index, dim_size):
    """pool_local's scatter_max -> gather round trip (encoder/pointnet.py:104-122) for one plane.
    c: [B,T,F], index: [B,1,T] -> [B,F,T]."""
    src = c.permute(0, 2, 1)
    idx = index.expand(-1, src.shape[1], -1)
    val = src.new_full((src.shape[0], src.shape[1], dim_size), float("-inf"))
    val.scatter_reduce_(-1, idx, src, reduce="amax", include_self=True)
    return val.gather(2, idx)


def _unet(sd, x, depth=4):
    """encoder/unet.py:225-239 (concat merge, transpose-conv up, relu, 2x2 maxpool)."""
    pre = "encoder.unet."
    skips = []
    for i in range(depth):
        x = F.relu(F.conv2d(x, sd[pre + "down_convs.%d.conv1.weight" % i], sd[pre + "down_convs.%d.conv1.bias" % i], padding=1))
        x = F.relu(F.conv2d(x, sd[pre + "down_convs.%d.conv2.weight" % i], sd[pre + "down_convs.%d.conv2.bias" % i], padding=1))
        skips.append(x)
        if i < depth - 1:
            x = F.max_pool2d(x, kernel_size=2, stride=2)
    for i in range(depth - 1):
        skip = skips[-(i + 2)]
        up = F.conv_transpose2d(x, sd[pre + "up_convs.%d.upconv.weight" % i], sd[pre + "up_convs.%d.upconv.bias" % i], stride=2)
        x = torch.cat((up, skip), 1)
        x = F.relu(F.conv2d(x, sd[pre + "up_convs.%d.conv1.weight" % i], sd[pre + "up_convs.%d.conv1.bias" % i], padding=1))
        x = F.relu(F.conv2d(x, sd[pre + "up_convs.%d.conv2.weight" % i], sd[pre + "up_convs.%d.conv2.bias" % i], padding=1))
    return F.conv2d(x, sd[pre + "conv_final.weight"], sd[pre + "conv_final.bias"])


def convonet_encode(sd, p, reso=64, padding=0.1, c_dim=32, planes=("xz", "xy", "yz"), n_blocks=5):
    """LocalPoolPointnet.forward (encoder/pointnet.py:124-168) for the shipped 3-plane config."""
    index = {}
    for pl in planes:
        index[pl] = coordinate2index(normalize_coordinate(p.clone(), padding=padding, plane=pl), reso)
    net = _lin(sd, "encoder.fc_pos", p)
    net = _resblock_fc(sd, "encoder.blocks.0", net)
    for i in range(1, n_blocks):
        pooled = 0
        for pl in planes:
            pooled = pooled + _scatter_max_gather(net, index[pl], reso ** 2)
        net = torch.cat([net, pooled.permute(0, 2, 1)], dim=2)
        net = _resblock_fc(sd, "encoder.blocks.%d" % i, net)
    c = _lin(sd, "encoder.fc_c", net)
    fea = {}
    for pl in planes:
        src = c.permute(0, 2, 1)
        idx = index[pl].expand(-1, c_dim, -1)
        plane = src.new_zeros(p.shape[0], c_dim, reso ** 2).scatter_add_(-1, idx, src)
        cnt = torch.zeros_like(plane).scatter_add_(-1, idx, torch.ones_like(src))
        plane = (plane / cnt.clamp_(min=1)).reshape(p.shape[0], c_dim, reso, reso)
        fea[pl] = _unet(sd, plane)
    return fea


def convonet_grid_features(c, p, reso=32, padding=0.1):
    """generate_grid_features up to the 3-D U-Net (encoder/pointnet.py:88-99): scatter_mean of the point features c [B,T,C]
    into the reso^3 volume -> [B,C,reso,reso,reso] (indexed [z][y][x]: coordinate2index '3d' is x + reso (y + reso z))."""
    index = coordinate2index(normalize_3d_coordinate(p.clone(), padding=padding), reso, coord_type="3d")
    src = c.permute(0, 2, 1)
    idx = index.expand(-1, src.shape[1], -1)
    vol = src.new_zeros(p.shape[0], src.shape[1], reso ** 3).scatter_add_(-1, idx, src)
    cnt = torch.zeros_like(vol).scatter_add_(-1, idx, torch.ones_like(src))
    return (vol / cnt.clamp_(min=1)).reshape(p.shape[0], src.shape[1], reso, reso, reso), index


def onet_encode(sd, p):
    """ResnetPointnet.forward (im2mesh/encoder/pointnet.py:85-113)."""
    net = _lin(sd, "encoder.fc_pos", p)
    net = _resblock_fc(sd, "encoder.block_0", net)
    for i in range(1, 5):
        pooled = net.max(dim=1, keepdim=True)[0].expand(net.size())
        net = torch.cat([net, pooled], dim=2)
        net = _resblock_fc(sd, "encoder.block_%d" % i, net)
    net = net.max(dim=1)[0]
    return _lin(sd, "encoder.fc_c", F.relu(net))


# ----------------------------------------------------------------------------------------------
# driver pieces: ConvONet/opt_defense.py
# ----------------------------------------------------------------------------------------------


def normalize_batch_pc(points):
    """opt_defense.py:76-83 (in place)."""
    points -= torch.mean(points, dim=1)[:, None, :]
    dist = torch.sum(points ** 2, dim=2) ** 0.5
    points /= torch.max(dist, dim=1)[0][:, None, None]
    return points


def preprocess_pc_np(pc, num_points=None, padding_scale=1., rng=None):
    """opt_defense.py:114-146 numpy part: centre, divide by the largest bbox extent, times padding_scale;
    optional subset without replacement (reference: np.random.choice; here an explicit Generator)."""
    c = pc - np.mean(pc, axis=0)
    scale = (np.max(c, axis=0) - np.min(c, axis=0)).max()
    allp = c / scale * padding_scale
    if num_points is not None and allp.shape[0] > num_points:
        idx = (rng or np.random).choice(allp.shape[0], num_points, replace=False)
        sel = allp[idx]
    else:
        sel = allp
    return allp.astype(np.float32), sel.astype(np.float32)


def init_points(pcs, npoint=1024, sigma=0.01, padding_scale=0.9, gen=None):
    """opt_defense.py:149-179 with an explicit torch.Generator (host RNG on both paths)."""
    pts = []
    for pc in pcs:
        pc = torch.as_tensor(pc).float()
        idx = torch.randint(0, pc.shape[0], (npoint,), generator=gen)
        pts.append(pc[idx])
    pts = torch.stack(pts, 0)
    noise = torch.randn(pts.shape, generator=gen) * sigma
    return torch.clamp(pts + noise, min=-0.5 * padding_scale, max=0.5 * padding_scale)


def optimize_points(decode_fn, opt_points, rep_weight=500., iterations=200, lr=1e-3, threshold=0.2,
                    normalize=True, trace=None, trace_steps=()):
    """opt_defense.py:182-239.  decode_fn(p) -> logits [B,K].  Returns float32 numpy [B,K,3].
    `trace` (a dict) collects per-step tensors at `trace_steps` for step-level goldens."""
    opt_points = opt_points.clone().float()
    opt_points.requires_grad_()
    B, K = opt_points.shape[:2]
    target = torch.ones((B, K)).float() * threshold
    opt = torch.optim.Adam([opt_points], lr=lr)
    stats = []
    for i in range(iterations + 1):
        occ_value = decode_fn(opt_points)
        occ_loss = F.binary_cross_entropy_with_logits(occ_value, target, reduction="none")
        occ_loss = torch.mean(occ_loss) * K
        rep_loss = torch.tensor(0.).float()
        if rep_weight > 0.:
            rep_loss = torch.mean(repulsion_loss(opt_points)) * rep_weight
        loss = occ_loss + rep_loss
        opt.zero_grad()
        loss.backward()
        if trace is not None and i in trace_steps:
            trace.setdefault("grad", {})[i] = opt_points.grad.detach().clone().numpy()
            trace.setdefault("logits", {})[i] = occ_value.detach().clone().numpy()
        opt.step()
        if trace is not None and i in trace_steps:
            trace.setdefault("xyz", {})[i] = opt_points.detach().clone().numpy()
        if i % 100 == 0:
            stats.append([loss.item(), occ_loss.item(), rep_loss.item(),
                          torch.sigmoid(occ_value).mean().item()])
    opt_points = opt_points.detach()
    if normalize:
        opt_points = normalize_batch_pc(opt_points)
    if trace is not None:
        trace["stats"] = np.asarray(stats, dtype=np.float64)
    return opt_points.numpy()


# ----------------------------------------------------------------------------------------------
# synthetic-weights / synthetic-cloud protocol -- shared by tests and bench
# ----------------------------------------------------------------------------------------------


def synth_cloud(i, n=1024):
    """Seeded ellipsoid / box surface cloud with 5 % displaced points, before preprocess_pc."""
    r = np.random.default_rng(1000 + i)
    radii = r.uniform(0.3, 1.0, size=3)
    if i % 2 == 0:
        v = r.normal(size=(n, 3))
        v /= np.linalg.norm(v, axis=1, keepdims=True)
        pts = v * radii
    else:
        pts = r.uniform(-1, 1, size=(n, 3))
        ax = r.integers(0, 3, size=n)
        pts[np.arange(n), ax] = np.sign(pts[np.arange(n), ax])
        pts *= radii * (1.0 - 0.3 * (pts[:, 2:3] * 0.5 + 0.5))
    k = n // 20
    sel = r.choice(n, k, replace=False)
    pts[sel] += r.normal(scale=0.02, size=(k, 3))
    return pts.astype(np.float32)


def randomize_zero_init(sd, seed=0):
    """Re-initialise the tensors the reference zero-initialises (layers.py:37, ONet layers.py:96,220-224)
    and the BN running stats, so that synthetic weights exercise every term.  In place on a state_dict."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if k.endswith("fc_1.weight"):
            fan_in = v.shape[1] * (v.shape[2] if v.dim() > 2 else 1)
            bound = 1.0 / math.sqrt(fan_in)                      # kaiming_uniform_(a=sqrt(5))
            v.copy_((torch.rand(v.shape, generator=g) * 2 - 1) * bound)
        elif k.endswith("conv_gamma.weight") or k.endswith("conv_beta.weight"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.02)
        elif k.endswith("running_mean"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.1)
        elif k.endswith("running_var"):
            v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    return sd
