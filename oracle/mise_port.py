"""TEST INFRASTRUCTURE ONLY: numpy restatement of MISE (ONet/im2mesh/utils/libmise/mise.pyx:33-369) as the loop of
Generator3D.generate_from_latent drives it (ONet/im2mesh/onet/generation.py:113-130), with dense state instead of the
reference's vectors + hash map (every decision of the reference is a set decision).  Pinned against the reference's own
Cython build by tests/golden/mise.npz (tests/golden/make_mise_golden.py)."""
import numpy as np


def analytic_field(pts, res, kind=0):
    """A reproducible field on integer lattice points [N,3] (only exactly rounded float64 operations)."""
    p = pts.astype(np.float64) / res - 0.5
    x, y, z = p[:, 0], p[:, 1], p[:, 2]
    if kind == 0:
        return (0.33 - np.sqrt(x * x + 0.6 * y * y + 1.4 * z * z)) * 12.0
    if kind == 1:     # two blobs and a thin plate: late activation of coarse voxels, thin features MISE can miss
        a = 0.18 - np.sqrt((x - 0.2) * (x - 0.2) + y * y + z * z)
        b = 0.12 - np.sqrt((x + 0.25) * (x + 0.25) + (y - 0.1) * (y - 0.1) + z * z)
        c = 0.012 - np.abs(z + 0.3) - 0.2 * np.maximum(np.abs(x) - 0.3, 0) - 0.2 * np.maximum(np.abs(y) - 0.3, 0)
        return np.maximum(np.maximum(a, b), c) * 20.0
    return np.round((x + y - z) * 4.0)          # plateaus: many values exactly equal to the threshold


def mise_loop(eval_fn, resolution0, depth, threshold):
    """-> (dense [n,n,n] float64 == MISE.to_dense(), [number of points queried per round])."""
    res = resolution0 << depth
    n = res + 1
    val = np.full((n, n, n), np.nan)
    exists = np.zeros((n, n, n), dtype=bool)
    known = np.zeros((n, n, n), dtype=bool)
    s0 = 1 << depth
    exists[::s0, ::s0, ::s0] = True                                   # mise.pyx:69-81
    sub = [np.zeros((resolution0 << L,) * 3, dtype=bool) for L in range(depth)]
    rounds = []
    while True:
        pts = np.argwhere(exists & ~known)                            # query(), :107-126
        if pts.shape[0] == 0:
            break
        rounds.append(int(pts.shape[0]))
        val[tuple(pts.T)] = np.asarray(eval_fn(pts), dtype=np.float64)  # update(), :83-105
        known[tuple(pts.T)] = True
        for L in reversed(range(depth)):                              # subdivide_voxels(), :180-237
            s = 1 << (depth - L)
            r = resolution0 << L
            present = np.ones((r, r, r), dtype=bool) if L == 0 else sub[L - 1].repeat(2, 0).repeat(2, 1).repeat(2, 2)
            leaf = present & ~sub[L]
            pos = np.zeros((r, r, r), dtype=bool)
            neg = np.zeros((r, r, r), dtype=bool)
            with np.errstate(invalid="ignore"):
                for a in range(s + 1):
                    for b in range(s + 1):
                        for c in range(s + 1):
                            k = known[a::s, b::s, c::s][:r, :r, :r]
                            v = val[a::s, b::s, c::s][:r, :r, :r]
                            pos |= k & (v >= threshold)
                            neg |= k & (v <= threshold)
            act = np.argwhere(leaf & pos & neg)
            sub[L][tuple(act.T)] = True
            h = s >> 1
            for a in range(3):                                        # subdivide_voxel(), :239-282
                for b in range(3):
                    for c in range(3):
                        exists[act[:, 0] * s + a * h, act[:, 1] * s + b * h, act[:, 2] * s + c * h] = True
    dense = val.copy()                                                # to_dense(), :128-163
    for i in range(1, n):
        m = np.isnan(dense[i])
        dense[i][m] = dense[i - 1][m]
    for j in range(1, n):
        m = np.isnan(dense[:, j])
        dense[:, j][m] = dense[:, j - 1][m]
    for k in range(1, n):
        m = np.isnan(dense[:, :, k])
        dense[:, :, k][m] = dense[:, :, k - 1][m]
    return dense, rounds
