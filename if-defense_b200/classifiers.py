"""The consumer side of the defense: the victim classifiers of baselines/inference.py evaluated on the
restored npz, with their index-producing geometry ops running on the library's kernels.

  DGCNN              baselines/model/dgcnn.py:43-129   edge features from `ifd_knn` (k = 20, self included, feature space)
  PointNet2ClsSsg    baselines/model/pointnet2.py:342-367, set abstraction :153-198, sample_and_group :101-130
                     FPS from `ifd_fps` (the start index the reference draws with torch.randint is an argument),
                     grouping from `ifd_ball_query`
  load_npz / test_normal / test_target    baselines/dataset/ModelNet40.py:9-16,31-38, baselines/inference.py:31-83

Parameter names equal the reference's `state_dict` keys (with or without the `module.` prefix nn.DataParallel adds, inference.py:
179-185), so a reference checkpoint loads unchanged.  The 1x1 convolutions, BatchNorm and Linear layers stay torch: they
are not on the hot path.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .defense import pn_utils


def strip_data_parallel(state_dict):
    """Keys of a checkpoint saved from nn.DataParallel(model) (inference.py:179-185) -> keys of the bare model."""
    return {(k[len("module."):] if k.startswith("module.") else k): v for k, v in state_dict.items()}


# ------------------------------------------------------------------------------------------------ DGCNN
def edge_features(x, k, idx=None):
    """get_graph_feature (dgcnn.py:16-40): x [B,C,N] -> [B,2C,N,k] = (x_j - x_i, x_i) over the k nearest j of i
    (nearest in this layer's feature space, self included)."""
    B, C, N = x.shape
    if idx is None:
        idx = pn_utils.dgcnn_knn(x, k)                                   # [B,N,k] int64
    pts = x.transpose(2, 1)                                              # [B,N,C]
    nbr = pn_utils.index_points(pts, idx)                                # [B,N,k,C]
    ctr = pts.unsqueeze(2).expand(B, N, k, C)
    return torch.cat((nbr - ctr, ctr), dim=3).permute(0, 3, 1, 2)


class DGCNN(nn.Module):
    WIDTHS = ((6, 64), (128, 64), (128, 128), (256, 256))

    def __init__(self, emb_dims=1024, k=20, output_channels=40):
        super().__init__()
        self.k, self.emb_dims = k, emb_dims
        for n, (cin, cout) in enumerate(self.WIDTHS, start=1):
            bn = nn.BatchNorm2d(cout)
            setattr(self, "bn%d" % n, bn)                                # the reference registers the norm twice (bnN and convN.1)
            setattr(self, "conv%d" % n, nn.Sequential(nn.Conv2d(cin, cout, kernel_size=1, bias=False), bn, nn.LeakyReLU(0.2)))
        self.bn5 = nn.BatchNorm1d(emb_dims)
        self.conv5 = nn.Sequential(nn.Conv1d(512, emb_dims, kernel_size=1, bias=False), self.bn5, nn.LeakyReLU(0.2))
        self.bn6 = nn.BatchNorm1d(512)
        self.bn7 = nn.BatchNorm1d(256)
        self.linear1 = nn.Sequential(nn.Linear(emb_dims * 2, 512, bias=False), self.bn6)
        self.linear2 = nn.Sequential(nn.Linear(512, 256), self.bn7)
        self.dp1 = nn.Dropout(p=0.5)
        self.dp2 = nn.Dropout(p=0.5)
        self.linear3 = nn.Linear(256, output_channels)

    def forward(self, x):
        """x [B,3,N] -> logits [B,40]."""
        B = x.size(0)
        feats = []
        for n in range(1, 5):
            x = getattr(self, "conv%d" % n)(edge_features(x, self.k)).max(dim=-1)[0]
            feats.append(x)
        x = self.conv5(torch.cat(feats, dim=1))
        x = torch.cat((F.adaptive_max_pool1d(x, 1).view(B, -1), F.adaptive_avg_pool1d(x, 1).view(B, -1)), 1)
        x = self.dp1(F.leaky_relu(self.linear1(x), negative_slope=0.2))
        x = self.dp2(F.leaky_relu(self.linear2(x), negative_slope=0.2))
        return self.linear3(x)


# ------------------------------------------------------------------------------------------------ PointNet++ (SSG)
class PointNetSetAbstraction(nn.Module):
    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint, self.radius, self.nsample, self.group_all = npoint, radius, nsample, group_all
        self.mlp_convs = nn.ModuleList()
        self.mlp_bns = nn.ModuleList()
        for cout in mlp:
            self.mlp_convs.append(nn.Conv2d(in_channel, cout, 1))
            self.mlp_bns.append(nn.BatchNorm2d(cout))
            in_channel = cout

    def forward(self, xyz, points, start=None):
        """xyz [B,3,N], points [B,D,N] or None -> (new_xyz [B,3,S], features [B,D',S]).  `start` [B]: the FPS seed the
        reference draws inside farthest_point_sample (pointnet2.py:64)."""
        xyz = xyz.permute(0, 2, 1).contiguous()
        pts = points.permute(0, 2, 1) if points is not None else None
        B, N, C = xyz.shape
        if self.group_all:                                               # sample_and_group_all, :133-150
            new_xyz = torch.zeros(B, 1, C, device=xyz.device)
            grouped = xyz.view(B, 1, N, C)
            if pts is not None:
                grouped = torch.cat([grouped, pts.reshape(B, 1, N, -1)], dim=-1)
        else:                                                            # sample_and_group, :101-130
            fps_idx = pn_utils.farthest_point_sample(xyz, self.npoint, start)
            new_xyz = pn_utils.index_points(xyz, fps_idx)
            idx = pn_utils.query_ball_point(self.radius, self.nsample, xyz, new_xyz)
            grouped = pn_utils.index_points(xyz, idx) - new_xyz.view(B, self.npoint, 1, C)
            if pts is not None:
                grouped = torch.cat([grouped, pn_utils.index_points(pts, idx)], dim=-1)
        f = grouped.permute(0, 3, 2, 1)                                  # [B, C+D, nsample, S]
        for conv, bn in zip(self.mlp_convs, self.mlp_bns):
            f = F.relu(bn(conv(f)))
        return new_xyz.permute(0, 2, 1), torch.max(f, 2)[0]


class PointNet2ClsSsg(nn.Module):
    def __init__(self, num_classes=40):
        super().__init__()
        self.sa1 = PointNetSetAbstraction(512, 0.2, 32, 3, [64, 64, 128], False)
        self.sa2 = PointNetSetAbstraction(128, 0.4, 64, 128 + 3, [128, 128, 256], False)
        self.sa3 = PointNetSetAbstraction(None, None, None, 256 + 3, [256, 512, 1024], True)
        self.fc1 = nn.Linear(1024, 512)
        self.bn1 = nn.BatchNorm1d(512)
        self.drop1 = nn.Dropout(0.4)
        self.fc2 = nn.Linear(512, 256)
        self.bn2 = nn.BatchNorm1d(256)
        self.drop2 = nn.Dropout(0.4)
        self.fc3 = nn.Linear(256, num_classes)

    def forward(self, xyz, starts=None):
        """xyz [B,3,N] -> logits.  starts = (start1 [B], start2 [B]) makes the two FPS seeds explicit (else drawn like the
        reference: torch.randint per call)."""
        B = xyz.shape[0]
        s1, s2 = starts if starts is not None else (None, None)
        l1_xyz, l1 = self.sa1(xyz, None, s1)
        l2_xyz, l2 = self.sa2(l1_xyz, l1, s2)
        _, l3 = self.sa3(l2_xyz, l2)
        x = self.drop1(F.relu(self.bn1(self.fc1(l3.view(B, 1024)))))
        x = self.drop2(F.relu(self.bn2(self.fc2(x))))
        return self.fc3(x)


# ------------------------------------------------------------------------------------------------ inference.py
def normalize_points_np(points):
    """baselines/util/pointnet_utils.py:107-113."""
    points = points - np.mean(points, axis=0)[None, :]
    return points / np.max(np.sqrt(np.sum(points ** 2, axis=1)), 0)


def load_npz(path, partition="test"):
    """dataset/ModelNet40.py:9-16: the arrays the defense writes (driver.defend_npz_test_data)."""
    npz = np.load(path, allow_pickle=True)
    if partition == "attack":
        return npz["test_pc"], npz["test_label"], npz["target_label"]
    return npz["test_pc"], npz["test_label"]


@torch.no_grad()
def predict(model, clouds, batch_size=32, num_points=1024, normalize=True):
    """The loop body of test_normal / test_target (inference.py:44-51, 71-79) over an array of clouds [n,K,>=3]:
    first num_points points, optional normalisation (ModelNet40.__getitem__), [B,3,N] in, argmax out."""
    model.eval()
    dev = next(model.parameters()).device
    preds = []
    for lo in range(0, len(clouds), batch_size):
        pcs = [np.asarray(pc)[:num_points, :3] for pc in clouds[lo:lo + batch_size]]
        pcs = [normalize_points_np(pc) if normalize else pc for pc in pcs]
        data = torch.from_numpy(np.stack(pcs)).float().to(dev).transpose(1, 2).contiguous()
        preds.append(torch.argmax(model(data), dim=-1).cpu())
    return torch.cat(preds).numpy()


def test_normal(model, path, **kw):
    """inference.py:58-83 -> overall accuracy."""
    pc, label = load_npz(path)
    return float((predict(model, pc, **kw) == label.astype(np.int64)).mean())


def test_target(model, path, **kw):
    """inference.py:31-55 -> (accuracy, attack success rate)."""
    pc, label, target = load_npz(path, "attack")
    p = predict(model, pc, **kw)
    return float((p == label.astype(np.int64)).mean()), float((p == target.astype(np.int64)).mean())
