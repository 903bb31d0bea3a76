"""Generates tests/golden/classifiers.npz by running the REFERENCE's victim classifiers (baselines/model/dgcnn.py,
pointnet2.py, imported unmodified from /root/reference) on the CPU.  The modules hard-code torch.device('cuda')
(dgcnn.py:22) and draw the FPS seeds inside the forward pass (pointnet2.py:64); they are imported with a proxy `torch` in
their namespace that maps the device to the CPU and records the torch.randint draws, nothing else changes.
Weights: synth.fill_classifier_state_ (a function of the state_dict names, so the mirrors get the same tensors).
    python tests/golden/make_classifier_golden.py"""
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from ifdefense_b200 import synth  # noqa: E402

REF = "/root/reference/baselines/model"


class TorchProxy:
    """`torch` as the reference module sees it: device('cuda') is the CPU, randint draws are recorded."""

    def __init__(self):
        self.draws = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def device(self, *a, **k):
        return torch.device("cpu")

    def randint(self, *a, **k):
        r = torch.randint(*a, **k)
        self.draws.append(r.clone())
        return r


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    proxy = TorchProxy()
    mod.torch = proxy
    return mod, proxy


def clouds():
    rng = np.random.default_rng(77)
    pcs = [np.load(os.path.join(HERE, "airplane.npy"))[:, :3]]
    for i in range(3):
        p = rng.standard_normal((1024, 3)) * rng.uniform(0.3, 1.0, size=3)
        pcs.append(p / np.linalg.norm(p, axis=1, keepdims=True) * rng.uniform(0.6, 1.0, size=(1024, 1)))
    out = []
    for p in pcs:
        p = p - p.mean(axis=0)
        out.append((p / np.sqrt((p ** 2).sum(axis=1)).max()).astype(np.float32))
    return np.stack(out)


def main():
    torch.set_num_threads(1)
    pcs = clouds()
    x = torch.from_numpy(pcs).transpose(1, 2).contiguous()
    out = {"clouds": pcs}
    dg, _ = load("dgcnn")
    m = synth.fill_classifier_state_(dg.DGCNN(1024, 20, output_channels=40), seed=1).eval()
    with torch.no_grad():
        out["dgcnn_logits"] = m(x).numpy()
        out["dgcnn_knn0"] = dg.knn(x, 20).numpy().astype(np.int32)
    out["dgcnn_keys"] = np.array(sorted(m.state_dict().keys()))
    p2, proxy = load("pointnet2")
    m = synth.fill_classifier_state_(p2.PointNet2ClsSsg(num_classes=40), seed=2).eval()
    torch.manual_seed(5)
    with torch.no_grad():
        out["pointnet2_logits"] = m(x).numpy()
    assert len(proxy.draws) == 2
    out["pointnet2_start1"] = proxy.draws[0].numpy()
    out["pointnet2_start2"] = proxy.draws[1].numpy()
    out["pointnet2_keys"] = np.array(sorted(m.state_dict().keys()))
    for k, v in out.items():
        print(k, v.shape, v.dtype)
    print("dgcnn argmax", out["dgcnn_logits"].argmax(1), "pointnet2 argmax", out["pointnet2_logits"].argmax(1))
    np.savez_compressed(os.path.join(HERE, "classifiers.npz"), **out)


if __name__ == "__main__":
    main()
