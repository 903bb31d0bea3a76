"""The defense driver: the functions of ConvONet/opt_defense.py (and its ONet twin) as a library.

Reference call stack mirrored:
  defend_npz_test_data :317-344 -> defend_point_cloud :255-314 -> sor_process :86-111 -> preprocess_pc :114-146
  -> model.encode_inputs :300 -> init_points :149-179 -> optimize_points :182-239 -> normalize_batch_pc :76-83

Differences, all forced by turning a script with module-level state into a library: the globals `args`,
`generator`, `device` become fields of `Defender`; random draws take an explicit generator so that the same
draw can be fed to the oracle.  Flag names and defaults are the reference's (opt_defense.py:21-53).
"""
import os

import numpy as np
import torch

from . import convonet, shard
from .defense import SORDefense


class Args:
    """argparse defaults of ConvONet/opt_defense.py:21-53."""
    sample_npoint = 1024
    padding_scale = 0.9
    init_sigma = 0.01
    iterations = 200
    batch_size = 192
    lr = 0.001
    rep_weight = 500.
    sor = True
    sor_k = 2
    sor_alpha = 1.1
    threshold = 0.2       # cfg['test']['threshold']
    input_npoint = 600    # cfg['data']['pointcloud_n'] (300 for ONet)
    device_preprocess = True   # SOR selection + preprocess_pc + init gather on the device (same bits as the numpy path)
    encoder_chunk = 0     # sharded path only: 0 = encode a slice in one call (the library's own encoder kernels do not depend on
                          # the batch a cloud arrives in); n > 0 = chunks of exactly n clouds aligned to the job's cloud
                          # indices and padded to full size, for encoders that run on torch / cuDNN (their results were
                          # measured to depend on a sample's position in the batch)

    def __init__(self, **over):
        for k, v in over.items():
            if not hasattr(type(self), k):
                raise RuntimeError("unknown option %r" % k)
            setattr(self, k, v)


def preprocess_pc(pc, num_points=None, padding_scale=1., rng=None):
    """opt_defense.py:114-146 (numpy part): centre, scale by the largest bbox extent, pad; optional random
    subset of num_points (np.random.choice without replacement; `rng` makes the draw explicit).
    Returns (all_points [K_i,3] float32, selected [min(K_i,num_points),3] float32)."""
    centered = pc - np.mean(pc, axis=0)
    scale = (np.max(centered, axis=0) - np.min(centered, axis=0)).max()
    allp = centered / scale * padding_scale
    if num_points is not None and allp.shape[0] > num_points:
        idx = (rng if rng is not None else np.random).choice(allp.shape[0], num_points, replace=False)
        sel = allp[idx]
    else:
        sel = allp
    return allp.astype(np.float32), sel.astype(np.float32)


def init_points(pcs, npoint=1024, sigma=0.01, padding_scale=0.9, gen=None):
    """opt_defense.py:149-179: K indices with replacement per cloud, + N(0, sigma^2), clamp to the padded
    cube.  Host RNG (the kernels never draw random numbers); returns a CPU tensor [B,npoint,3]."""
    pts = []
    for pc in pcs:
        pc = torch.as_tensor(pc).float().cpu()
        pts.append(pc[torch.randint(0, pc.shape[0], (npoint,), generator=gen)])
    pts = torch.stack(pts, 0)
    noise = torch.randn(pts.shape, generator=gen) * sigma
    return torch.clamp(pts + noise, min=-0.5 * padding_scale, max=0.5 * padding_scale)


class Defender:
    """ConvONet-Opt end to end on one GPU.  `model` is a models.ConvolutionalOccupancyNetwork (or any object
    with encode_inputs() returning the 3-plane dict and a `decoder` holding the reference's decoder.* keys)."""
    loops_in_flight = 4        # sharded driver: segments whose loops may run at the same time (each call owns its workspace)

    def __init__(self, model, args=None, device="cuda"):
        self.args = args or Args()
        from . import capi
        self.device = capi.use_device(device)
        self.model = model.to(self.device).eval()
        for p in self.model.parameters():
            p.requires_grad = False
        sd = {"decoder." + k: v for k, v in self.model.decoder.state_dict().items()}
        self.decoder = convonet.ConvONetDecoder(sd, padding=self.model.decoder.padding, device=self.device)
        self.restorer = convonet.Restorer(self.decoder, threshold=self.args.threshold, lr=self.args.lr)

    def _encode(self, sel):
        """encode_inputs (opt_defense.py:300) in the layout the loop wants: the 3 planes, channels-last."""
        return convonet.planes_to_channels_last(self.model.encode_inputs(sel))

    def sor_process(self, pc):
        """opt_defense.py:86-111: SOR in batches of 32 -> list of [K_i,3] float32 arrays."""
        sor = SORDefense(k=self.args.sor_k, alpha=self.args.sor_alpha)
        out = []
        for i in range(0, len(pc), 32):
            x = torch.from_numpy(np.asarray(pc[i:i + 32])).float().to(self.device)
            out += [o.detach().cpu().numpy().astype(np.float32) for o in sor(x)]
        return out

    def _to_device(self, host, tag):
        """Host array / tensor -> device tensor through a pinned staging buffer this object keeps per (tag, shape, dtype): a
        copy from pageable memory is staged by the driver and blocks the host for 3-100 ms when the GPU is busy with the
        loops of earlier batches (measured); from pinned memory it is one asynchronous DMA.  The staging buffer is reused by
        the next call with the same tag: the stream is synchronised on the copy before this function returns (a few MB)."""
        t = torch.from_numpy(host) if isinstance(host, np.ndarray) else host
        key = (tag, tuple(t.shape), t.dtype)
        pins = self.__dict__.setdefault("_pins", {})
        buf = pins.get(key)
        if buf is None:
            buf = pins[key] = torch.empty(t.shape, dtype=t.dtype).pin_memory()
        buf.copy_(t)
        out = buf.to(self.device, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return out

    def prepare_batch_device(self, raw, rng=None, gen=None):
        """sor_process + preprocess_pc + the encoder subset + init_points (opt_defense.py:86-179) for one batch [B,K,3] with the
        arrays on the device: SOR mask (`ifd_sor`), ragged selection + normalisation (`ifd_preprocess_pc`), gathers by the
        indices the host RNGs draw (the same draws, in the same order, as the numpy path).  `rng` / `gen` may be lists with
        one generator per cloud (the sharded path: draws that do not depend on how the job is cut).
        -> (sel [B,T,3], init [B,npoint,3])."""
        from . import capi
        a = self.args
        x = self._to_device(np.ascontiguousarray(np.asarray(raw)[..., :3], dtype=np.float32), "raw")
        B, K, _ = x.shape
        L = capi.lib()
        keep = None
        if a.sor:
            keep = torch.empty((B, K), dtype=torch.uint8, device=self.device)
            capi.check(L.ifd_sor(capi.ptr(x), B, K, a.sor_k, float(a.sor_alpha), capi.ptr(keep), None, capi.stream()), "ifd_sor")
        allp = torch.empty_like(x)
        counts = torch.empty(B, dtype=torch.int32, device=self.device)
        capi.check(L.ifd_preprocess_pc(capi.ptr(x), capi.ptr(keep), B, K, float(a.padding_scale), capi.ptr(allp), capi.ptr(counts),
                                       capi.stream()), "ifd_preprocess_pc")
        n = counts.cpu().numpy()
        T = a.input_npoint
        # preprocess_pc :134-141: a cloud with more than input_npoint points is subsampled, otherwise ALL its points are the
        # encoder input and no random number is drawn.  The reference then np.stack()s the batch, so the no-subsample case
        # only exists when every cloud of the batch kept the same number of points (e.g. --sor=False and K <= input_npoint).
        small = T is None or bool((n <= T).all())
        if small and len(set(int(k) for k in n)) != 1:
            raise RuntimeError("clouds of one batch keep different numbers of points (<= input_npoint) after SOR: the "
                               "reference cannot stack them either (opt_defense.py:296)")
        if not small and (n <= T).any():
            raise RuntimeError("some clouds of the batch have <= input_npoint points after SOR and some more: the reference "
                               "cannot stack them either (opt_defense.py:296)")
        per_cloud = isinstance(rng, (list, tuple))                                             # one RNG stream per cloud
        if small:
            sel_idx = np.tile(np.arange(int(n[0])), (B, 1))
        elif per_cloud:
            sel_idx = np.stack([r.choice(int(k), T, replace=False) for r, k in zip(rng, n)])
        else:
            draw = rng if rng is not None else np.random
            sel_idx = np.stack([draw.choice(int(k), T, replace=False) for k in n])             # preprocess_pc :134-141
        if isinstance(gen, (list, tuple)):                                                     # init_points per cloud
            ini_idx = torch.stack([torch.randint(0, int(k), (a.sample_npoint,), generator=g) for g, k in zip(gen, n)])
            noise = torch.stack([torch.randn((a.sample_npoint, 3), generator=g) * a.init_sigma for g in gen])
        else:
            ini_idx = torch.stack([torch.randint(0, int(k), (a.sample_npoint,), generator=gen) for k in n])   # init_points :163-167
            noise = torch.randn((B, a.sample_npoint, 3), generator=gen) * a.init_sigma
        rows = torch.arange(B, device=self.device).view(B, 1)
        sel = allp[rows, self._to_device(sel_idx, "sel")]
        pts = allp[rows, self._to_device(ini_idx, "ini")] + self._to_device(noise, "noise")
        return sel, torch.clamp(pts, min=-0.5 * a.padding_scale, max=0.5 * a.padding_scale)

    def defend_point_cloud(self, pc, rng=None, gen=None, printing=False):
        """opt_defense.py:255-314.  pc: [N,K,3] array.  Returns float32 [N,sample_npoint,3]."""
        a = self.args
        if a.device_preprocess and isinstance(pc, np.ndarray) and pc.ndim == 3:
            # pipelined (see _pipelined); the random draws happen in batch order, as in the serial code
            segs = [(lo, min(lo + a.batch_size, len(pc)), min(lo + a.batch_size, len(pc)) - lo) for lo in range(0, len(pc), a.batch_size)]
            if not segs:
                return np.zeros((0, a.sample_npoint, 3), dtype=np.float32)
            parts = self._pipelined(pc, segs, lambda lo, hi: (rng, gen), lambda sel, lo: self.model.encode_inputs(sel), printing)
            return np.concatenate(parts, axis=0).astype(np.float32)
        pcs = self.sor_process(pc) if a.sor else [np.asarray(p, dtype=np.float32) for p in pc]
        out = np.zeros((len(pcs), a.sample_npoint, 3), dtype=np.float32)
        for lo in range(0, len(pcs), a.batch_size):
            batch = pcs[lo:lo + a.batch_size]
            proc = [preprocess_pc(p, num_points=a.input_npoint, padding_scale=a.padding_scale, rng=rng) for p in batch]
            sel = torch.from_numpy(np.stack([s for _, s in proc])).float().to(self.device)
            with torch.no_grad():
                c = self.model.encode_inputs(sel)
            pts = init_points([p for p, _ in proc], a.sample_npoint, a.init_sigma, a.padding_scale, gen)
            out[lo:lo + a.batch_size] = self.restorer.optimize_points(pts, None, c, rep_weight=a.rep_weight,
                                                                      iterations=a.iterations, printing=printing)
        return out

    def encode_chunked(self, sel, first=0):
        """encode_inputs for clouds first .. first + len(sel) - 1 of the job, in chunks of exactly `encoder_chunk` clouds that
        are aligned to the JOB's cloud indices: cloud i always sits at position i % encoder_chunk of a full-size batch
        (positions outside this slice are filled with a copy of one of its clouds).  cuDNN / cuBLAS kernels are deterministic
        functions of (shape, position in the batch, the sample's own data) -- measured: at some batch sizes the position
        matters in the last bit -- so a cloud's planes do not depend on where the slice boundaries fall."""
        if not self.args.encoder_chunk:
            with torch.no_grad():
                return self.model.encode_inputs(sel)
        step = self.args.encoder_chunk
        n = sel.shape[0]
        out = None
        with torch.no_grad():
            g0 = first - first % step
            for g in range(g0, first + n, step):
                idx = torch.arange(g, g + step)
                inside = (idx >= first) & (idx < first + n)
                src = torch.where(inside, idx - first, torch.zeros_like(idx)).to(sel.device)
                c = self.model.encode_inputs(sel[src].contiguous())
                pos = torch.nonzero(inside).flatten().to(sel.device)
                if not isinstance(c, dict):                               # ONet: one latent code per cloud
                    c = {"c": c}
                if out is None:
                    out = {k: [] for k in c}
                for k, v in c.items():
                    out[k].append(v[pos])
        out = {k: torch.cat(v) for k, v in out.items()}
        return out["c"] if list(out) == ["c"] else out

    def _slice_streams(self, lo, hi, seed):
        """One numpy and one torch RNG stream per cloud, a function of (seed, cloud index) only."""
        rngs = [np.random.default_rng([int(seed), i]) for i in range(lo, hi)]
        gens = [torch.Generator().manual_seed((int(seed) * 1000003 + i) % (2 ** 63 - 1)) for i in range(lo, hi)]
        return rngs, gens

    def restore_slices(self, pc, segments, seed):
        """restore_slice for a list of (lo, hi, B_ref) segments, as a two-stage pipeline (the device pre-processing path):
        SOR + preprocess + encoder of segment j + 1 are enqueued on a side stream while the loop of segment j runs, and the
        restored clouds come back with one D2H per segment after the next stage has been enqueued.  Same bits as
        restore_slice segment by segment (every draw comes from the per-cloud streams)."""
        a = self.args
        if not (a.device_preprocess and isinstance(pc, np.ndarray) and pc.ndim == 3) or not segments:
            return [self.restore_slice(pc, lo, hi, n, seed) for lo, hi, n in segments]
        def draws(lo, hi):
            return self._slice_streams(lo, hi, seed)

        return self._pipelined(pc, segments, draws, lambda sel, lo: self.encode_chunked(sel, lo))

    def _pipelined(self, pc, segments, draws, encode, printing=False):
        """The batch loop as a pipeline on the device pre-processing path: SOR + preprocess + encoder of a segment are enqueued
        on a side stream while the loops of up to `depth` earlier segments run, each loop on its own stream (a loop is a chain
        of dependent launches that does not fill the GPU -- least of all the small slices of a many-rank job -- loops of
        different segments do); one D2H per segment.  draws(lo, hi) -> (rng, gen) for prepare_batch_device; stages run in
        segment order, so shared generators see the draws in the serial code's order."""
        a = self.args
        main = torch.cuda.current_stream(self.device)
        depth = 4 if max(hi - lo for lo, hi, _ in segments) <= 64 else 2
        depth = 1 if printing else min(depth, int(self.loops_in_flight))
        # the streams live as long as the object: torch's allocator keeps one block pool per stream, and a fresh set of streams
        # per call would start every call with empty pools (cudaMalloc inside the pipeline)
        pool = self.__dict__.setdefault("_streams", [])
        while len(pool) < 1 + depth:
            pool.append(torch.cuda.Stream(device=self.device))
        side, loops = pool[0], pool[1:1 + depth]
        side.wait_stream(main)
        for S in loops:
            S.wait_stream(main)

        def stage(seg):
            lo, hi, _ = seg
            rng, gen = draws(lo, hi)
            with torch.cuda.stream(side), torch.no_grad():
                sel, pts = self.prepare_batch_device(np.asarray(pc[lo:hi])[..., :3], rng, gen)
                c = encode(sel, lo)
                done = torch.cuda.Event()
                done.record(side)
            return pts, c, done

        out, flying = [], []

        def land():
            x, keep, ev = flying.pop(0)
            ev.synchronize()                               # (the stage's tensors in `keep` are released only now)
            out.append(x.cpu().numpy())

        for j, (lo, hi, n) in enumerate(segments):
            pts, c, done = stage((lo, hi, n))              # enqueued while up to `depth` earlier loops run
            if len(flying) >= depth:
                land()
            S = loops[j % depth]
            S.wait_event(done)
            with torch.cuda.stream(S):
                x = self.restorer.optimize_points(pts, None, c, rep_weight=a.rep_weight, iterations=a.iterations, B_ref=n,
                                                  printing=printing, return_tensor=True)
                ev = torch.cuda.Event()
                ev.record(S)
            flying.append((x, (pts, c), ev))
        while flying:
            land()
        main.wait_stream(side)
        for S in loops:
            main.wait_stream(S)
        return out

    def restore_slice(self, pc, lo, hi, B_ref, seed):
        """Clouds lo..hi-1 of a reference batch of B_ref clouds, every random draw taken from a per-cloud stream
        (seed, cloud index) so that the result does not depend on how the job is cut into slices."""
        a = self.args
        raw = np.asarray(pc[lo:hi])[..., :3]
        rngs, gens = self._slice_streams(lo, hi, seed)
        if a.device_preprocess and raw.ndim == 3:
            sel, pts = self.prepare_batch_device(raw, rngs, gens)
        else:
            pcs = self.sor_process(raw) if a.sor else [np.asarray(p, dtype=np.float32) for p in raw]
            proc, pts = [], []
            for p, r, g in zip(pcs, rngs, gens):
                allp, s_ = preprocess_pc(p, num_points=a.input_npoint, padding_scale=a.padding_scale, rng=r)
                proc.append(s_)
                pts.append(init_points([allp], a.sample_npoint, a.init_sigma, a.padding_scale, g)[0])
            sel = torch.from_numpy(np.stack(proc)).float().to(self.device)
            pts = torch.stack(pts)
        c = self.encode_chunked(sel, lo)
        return self.restorer.optimize_points(pts, None, c, rep_weight=a.rep_weight, iterations=a.iterations, B_ref=B_ref)

    def defend_point_cloud_sharded(self, pc, seed=0, rank=None, world=None):
        """defend_point_cloud across the ranks of torch.distributed (one process per GPU): every reference batch is
        cut into `world` contiguous slices (shard.plan), each rank restores its slices with the batch's B_ref, one
        all_gather returns all N restored clouds in input order on every rank."""
        return shard.restore_sharded(lambda lo, hi, n: self.restore_slice(pc, lo, hi, n, seed), len(pc), self.args.batch_size,
                                     rank=rank, world=world, device=self.device,
                                     restore_many=lambda segs: self.restore_slices(pc, segs, seed))


class ONetDefender(Defender):
    """ONet-Opt end to end (ONet/opt_defense.py, the twin of the ConvONet script: same functions and line numbers, 300 encoder
    points, latent code c [B,512] instead of feature planes, DecoderCBatchNorm).  `model` is a models.OccupancyNetwork."""
    loops_in_flight = 1        # the ONet restorer keeps ONE cached workspace per shape (1 GB at B = 64): one loop at a time; its
                               # layer GEMMs are persistent kernels that fill the GPU anyway

    def __init__(self, model, args=None, device="cuda"):
        from . import onet
        self.args = args or Args(input_npoint=300)
        from . import capi
        self.device = capi.use_device(device)
        self.model = model.to(self.device).eval()
        for p in self.model.parameters():
            p.requires_grad = False
        sd = {"decoder." + k: v for k, v in self.model.decoder.state_dict().items()}
        self.decoder = onet.ONetDecoder(sd, device=self.device)
        self.restorer = onet.ONetRestorer(self.decoder, threshold=self.args.threshold, lr=self.args.lr)

    def _encode(self, sel):
        return self.model.encode_inputs(sel)


def get_save_name(path, tag="convonet_opt-", folder="ConvONet-Opt"):
    """opt_defense.py:242-252."""
    name = path.split('/')[-1]
    save_folder = os.path.join(path[:path.rindex(name)], folder)
    os.makedirs(save_folder, exist_ok=True)
    return os.path.join(save_folder, tag + name)


def defend_npz_test_data(defender, path, **kw):
    """opt_defense.py:317-344: npz in (test_pc, test_label[, target_label]) -> npz out, same keys, test_pc
    float32 [N,1024,3], labels uint8."""
    npz = np.load(path)
    test_pc = npz['test_pc'][..., :3]
    out = defender.defend_point_cloud(test_pc, **kw)
    if isinstance(defender, ONetDefender):
        save_path = get_save_name(path, tag="onet_opt-", folder="ONet-Opt")          # ONet/opt_defense.py:242-252
    else:
        save_path = get_save_name(path)
    fields = dict(test_pc=out.astype(np.float32), test_label=npz['test_label'].astype(np.uint8))
    if 'target_label' in npz.files:
        fields['target_label'] = npz['target_label'].astype(np.uint8)
    np.savez(save_path, **fields)
    print('defense result saved to {}'.format(save_path))
    return save_path
