"""ONet-Mesh: mesh extraction and re-sampling on the device.

Mirrors, with the reference's names and argument meaning:
  libmcubes.marching_cubes(volume, isovalue)          ONet/im2mesh/utils/libmcubes/mcubes.pyx:20-25
  Generator3D.generate_from_latent / extract_mesh     ONet/im2mesh/onet/generation.py:88-221
  trimesh.sample.sample_surface(mesh, count)          call site ONet/remesh_defense.py:157
  reconstruct_mesh / resample_points                  ONet/remesh_defense.py:126-171
  libmise.MISE                                        ONet/im2mesh/utils/libmise/mise.pyx:33-369

Differences: the octree state and the occupancy lattice stay on the device (Generator3D(dense=True) evaluates the whole
(resolution0 * 2^steps + 1)^3 lattice instead of refining); a mesh is a pair of cuda tensors (vertices [V,3] float64, faces [F,3] int64) instead of a trimesh.Trimesh; random numbers take
an explicit numpy Generator.  There is no CPU fallback.
"""
import ctypes

import numpy as np
import torch

from . import capi


def _as_volume(volume):
    if isinstance(volume, torch.Tensor):
        v = volume.detach()
        if v.dtype not in (torch.float32, torch.float64):
            v = v.double()
    else:
        a = np.asarray(volume)
        v = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32 if a.dtype == np.float32 else np.float64))
    if v.dim() != 3:
        raise RuntimeError("Only three-dimensional arrays are supported.")          # pywrapper.cpp:92-93
    return v.cuda().contiguous()


def marching_cubes_device(volume, isovalue, pad=False, pad_value=-1e6, box_size=None):
    """-> (verts [V,3] float64 cuda, faces [F,3] int64 cuda).  pad / box_size: the padding and the vertex transform of
    Generator3D.extract_mesh (generation.py:172-183) fused into the extraction."""
    capi.require_gpu()
    L = capi.lib()
    v = _as_volume(volume)
    nx, ny, nz = v.shape
    dtype = 1 if v.dtype == torch.float64 else 0
    nb = L.ifd_mc_workspace_bytes(nx, ny, nz, int(pad))
    ws = torch.empty(max(int(nb), 256), dtype=torch.uint8, device=v.device)
    nv, nf = ctypes.c_longlong(), ctypes.c_longlong()
    capi.check(L.ifd_mc_count(capi.ptr(v), dtype, nx, ny, nz, int(pad), float(pad_value), float(isovalue), capi.ptr(ws), ws.numel(),
                              ctypes.byref(nv), ctypes.byref(nf), capi.stream()), "ifd_mc_count")
    verts = torch.empty((nv.value, 3), dtype=torch.float64, device=v.device)
    faces = torch.empty((nf.value, 3), dtype=torch.int64, device=v.device)
    if nv.value > 0 or nf.value > 0:
        # empty outputs have null data pointers; the library wants real ones
        vo = verts if nv.value > 0 else torch.empty((1, 3), dtype=torch.float64, device=v.device)
        fo = faces if nf.value > 0 else torch.empty((1, 3), dtype=torch.int64, device=v.device)
        capi.check(L.ifd_mc_emit(capi.ptr(v), dtype, nx, ny, nz, int(pad), float(pad_value), float(isovalue),
                                 0 if box_size is None else 1, 0.0 if box_size is None else float(box_size), capi.ptr(ws),
                                 ws.numel(), capi.ptr(vo), capi.ptr(fo), capi.stream()), "ifd_mc_emit")
    return verts, faces


def marching_cubes(volume, isovalue):
    """libmcubes.marching_cubes: numpy (verts [V,3] float64, faces [F,3] uint64-valued int64), the reference's order."""
    verts, faces = marching_cubes_device(volume, isovalue)
    return verts.cpu().numpy(), faces.cpu().numpy()


def extract_mesh(occ_hat, threshold=0.2, padding=0.1):
    """Generator3D.extract_mesh (generation.py:160-186) up to the Trimesh constructor; simplification and refinement are
    off in the shipped config (configs/default.yaml: simplify_nfaces null, refinement_step 0)."""
    thr = np.log(threshold) - np.log(1. - threshold)
    return marching_cubes_device(occ_hat, thr, pad=True, pad_value=-1e6, box_size=1 + padding)


def sample_surface_device(verts, faces, count, rng=None, uniforms=None, return_index=False):
    """trimesh.sample.sample_surface: `count` area-weighted points on the mesh (float64 cuda [count,3]).  The draws are
    rng.random(count) for the faces, then rng.random((count, 2)) for the edge lengths (trimesh's order)."""
    capi.require_gpu()
    if faces.shape[0] == 0:
        raise IndexError("cannot sample an empty mesh")      # what np.cumsum(...)[-1] raises inside trimesh
    if uniforms is None:
        rng = rng if rng is not None else np.random.default_rng()
        uniforms = np.concatenate([rng.random(count)[:, None], rng.random((count, 2))], axis=1)
    u = torch.as_tensor(np.ascontiguousarray(uniforms, dtype=np.float64)).to(verts.device)
    if u.shape != (count, 3):
        raise RuntimeError("uniforms must be [count, 3]")
    L = capi.lib()
    ws = torch.empty(int(L.ifd_sample_surface_workspace_bytes(faces.shape[0])), dtype=torch.uint8, device=verts.device)
    out = torch.empty((count, 3), dtype=torch.float64, device=verts.device)
    fidx = torch.empty(count, dtype=torch.int64, device=verts.device) if return_index else None
    capi.check(L.ifd_sample_surface(capi.ptr(verts.contiguous()), verts.shape[0], capi.ptr(faces.contiguous()), faces.shape[0],
                                    capi.ptr(u), count, capi.ptr(out), capi.ptr(fidx), capi.ptr(ws), ws.numel(), capi.stream()),
               "ifd_sample_surface")
    return (out, fidx) if return_index else out


class MISE:
    """libmise.MISE (ONet/im2mesh/utils/libmise/mise.pyx) with its state on the device; points and values are cuda tensors."""

    def __init__(self, resolution_0, depth, threshold):
        capi.require_gpu()
        self.resolution_0, self.depth, self.threshold = int(resolution_0), int(depth), float(threshold)
        self.resolution = self.resolution_0 << self.depth
        L = capi.lib()
        nb = int(L.ifd_mise_workspace_bytes(self.resolution_0, self.depth))
        if nb == 0:
            raise RuntimeError("MISE: unsupported resolution_0 / depth")
        self._ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        capi.check(L.ifd_mise_init(self.resolution_0, self.depth, capi.ptr(self._ws), nb, capi.stream()), "ifd_mise_init")

    def query(self):
        """-> int64 cuda [N, 3]: the points to evaluate next (empty when the extraction is complete)."""
        L = capi.lib()
        n = ctypes.c_longlong()
        capi.check(L.ifd_mise_query(self.resolution_0, self.depth, capi.ptr(self._ws), self._ws.numel(), None, 0, ctypes.byref(n),
                                    capi.stream()), "ifd_mise_query")
        pts = torch.empty((n.value, 3), dtype=torch.int64, device="cuda")
        if n.value:
            capi.check(L.ifd_mise_query(self.resolution_0, self.depth, capi.ptr(self._ws), self._ws.numel(), capi.ptr(pts), n.value,
                                        ctypes.byref(n), capi.stream()), "ifd_mise_query")
        return pts

    def update(self, points, values):
        points = torch.as_tensor(points).to("cuda", torch.int64).contiguous()
        values = torch.as_tensor(values).to("cuda", torch.float64).contiguous()
        assert points.shape[0] == values.shape[0] and points.shape[1] == 3            # mise.pyx:85-86
        L = capi.lib()
        rc = L.ifd_mise_update(self.resolution_0, self.depth, self.threshold, capi.ptr(self._ws), self._ws.numel(),
                               capi.ptr(points) if points.shape[0] else None, capi.ptr(values) if points.shape[0] else None,
                               points.shape[0], capi.stream())
        if rc != 0:
            msg = (L.ifd_last_error() or b"").decode()
            if "Point not in grid" in msg:
                raise ValueError("Point not in grid!")                                  # mise.pyx:102
            capi.check(rc, "ifd_mise_update")

    def to_dense(self):
        n = self.resolution + 1
        out = torch.empty((n, n, n), dtype=torch.float64, device="cuda")
        capi.check(capi.lib().ifd_mise_to_dense(self.resolution_0, self.depth, capi.ptr(self._ws), self._ws.numel(), capi.ptr(out),
                                                capi.stream()), "ifd_mise_to_dense")
        return out


class Generator3D:
    """The part of im2mesh.onet.generation.Generator3D that remesh_defense.py uses (generate_from_latent with z of
    width 0).  `decoder` is an onet.ONetDecoder."""

    def __init__(self, decoder, threshold=0.2, resolution0=32, upsampling_steps=2, padding=0.1, points_batch_size=100000,
                 dense=False):
        self.decoder = decoder
        self.dense = dense            # True: evaluate the whole fine lattice instead of refining with MISE
        self.points_evaluated = 0
        self.threshold = threshold
        self.resolution0 = resolution0
        self.upsampling_steps = upsampling_steps
        self.padding = padding
        self.points_batch_size = points_batch_size

    def value_grid(self, c):
        """The lattice of logits generate_from_latent hands to extract_mesh (generation.py:104-130)."""
        if self.upsampling_steps == 0:
            # generation.py:105-111: make_3d_grid((-0.5,)*3, (0.5,)*3, (nx,)*3): nx points from -0.5 to 0.5 per axis,
            # i.e. the lattice of resolution nx - 1
            return self.decoder.eval_dense_grid(c, self.resolution0 - 1, self.padding, self.points_batch_size)
        if self.dense:
            return self.decoder.eval_dense_grid(c, self.resolution0 * 2 ** self.upsampling_steps, self.padding, self.points_batch_size)
        threshold = np.log(self.threshold) - np.log(1. - self.threshold)
        cc = torch.as_tensor(c).detach().float().reshape(1, -1).cuda().contiguous()
        extractor = MISE(self.resolution0, self.upsampling_steps, threshold)            # generation.py:113-130
        ws = self.decoder._prepare_eval(cc, self.points_batch_size)                   # once per shape, not per round
        self.points_evaluated = 0
        points = extractor.query()
        while points.shape[0] != 0:
            values = self.decoder.eval_lattice_points(points, extractor.resolution, cc, ws, self.padding, self.points_batch_size)
            self.points_evaluated += int(points.shape[0])
            extractor.update(points, values)
            points = extractor.query()
        return extractor.to_dense()

    def generate_from_latent(self, z, c=None):
        return extract_mesh(self.value_grid(c), self.threshold, self.padding)


def resample_points(generator, encode_inputs, ori_pc, num_points=1024, input_npoint=300, padding_scale=0.9, rng=None):
    """remesh_defense.py:126-171: pre-process, encode, reconstruct the mesh, sample `num_points` from its surface; when
    the reconstruction is empty fall back to a random subset / zero padding of the input, as the reference does."""
    from .driver import preprocess_pc
    rng = rng if rng is not None else np.random.default_rng()
    ori_pc = np.asarray(ori_pc)[:, :3]
    _, sel = preprocess_pc(ori_pc, num_points=input_npoint, padding_scale=padding_scale, rng=rng)
    c = encode_inputs(torch.from_numpy(sel).float().cuda().unsqueeze(0))
    verts, faces = generator.generate_from_latent(None, c)
    try:
        return sample_surface_device(verts, faces, num_points, rng=rng).cpu().numpy()
    except IndexError:
        pc = np.zeros((num_points, 3), dtype=np.float32)
        if ori_pc.shape[0] > num_points:
            return ori_pc[rng.choice(ori_pc.shape[0], num_points, replace=False)]
        pc[:ori_pc.shape[0]] = ori_pc
        return pc


def resample_points_sharded(generator, encode_inputs, clouds, num_points=1024, seed=0, rank=None, world=None, device=None,
                            resample=None, **kw):
    """The file loop of remesh_defense.py:189-225 over the ranks of torch.distributed (one process per GPU): clouds are
    independent, so rank r re-meshes a contiguous block and one all_gather returns all [n, num_points, 3] clouds in input
    order on every rank.  Cloud i draws from its own stream (seed, i), so the result does not depend on the number of ranks.
    `resample` (default resample_points) is the per-cloud function."""
    from . import shard
    fn = resample if resample is not None else resample_points

    def block(lo, hi, _n):
        return np.stack([np.asarray(fn(generator, encode_inputs, clouds[i], num_points=num_points,
                                       rng=np.random.default_rng([int(seed), i]), **kw), dtype=np.float32) for i in range(lo, hi)])

    return shard.restore_sharded(block, len(clouds), max(len(clouds), 1), rank=rank, world=world, device=device)
