"""TEST INFRASTRUCTURE ONLY (container-side): import the *unmodified* reference classes from
/root/reference on CPU so that golden vectors can be generated from the reference itself.

/root/reference does not exist on the GPU box.  There, the only user is bench.py's reference arm
(`--impl reference`), which drives the unmodified copy vendored into oracle/_ref/pyref by `make -C oracle pyref`.
Otherwise it is used by tests/golden/make_golden.py and by the (skipped-when-absent) tests that pin
oracle/torch_port.py bit-for-bit against the real reference.

Shim recipe:
  * `trimesh` and the three compiled mesh extensions are stubbed (they are imported at package
    import time by src/conv_onet/generation.py:6-10 but never used on the opt_defense path);
  * `torch_scatter` (un-vendored third party, pinned 2.0.5 in requirements.txt:59) is replaced by
    oracle.scatter_shim, a pure-torch restatement of its published semantics;
  * `yaml.load` gets a default Loader (src/config.py:20 calls it without one);
  * `Tensor.cuda` becomes the identity so RepulsionLoss (defense/repulsion_loss.py:47) runs on CPU.
"""
import importlib
import os
import sys
import types

_VENDORED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "pyref")


def _default_root():
    """/root/reference in the build container; on the GPU box the UNMODIFIED copy of the hot path's Python files that
    `make -C oracle pyref` placed in the git-ignored oracle/_ref/pyref (it travels with the snapshot)."""
    if os.path.isdir("/root/reference/ConvONet/src"):
        return "/root/reference"
    return _VENDORED


REF_ROOT = os.environ.get("IFD_REFERENCE_ROOT") or _default_root()


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "ConvONet", "src"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _purge(prefixes):
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + ".") for p in prefixes):
            del sys.modules[k]


def _common_shims():
    import torch
    import yaml
    from oracle import scatter_shim

    sys.modules["torch_scatter"] = scatter_shim
    if "trimesh" not in sys.modules:
        _stub("trimesh")
    if not getattr(yaml.load, "_ifd_patched", False):
        _orig = yaml.load

        def _load(stream, Loader=None, **kw):
            return _orig(stream, Loader=Loader or yaml.FullLoader, **kw)

        _load._ifd_patched = True
        yaml.load = _load
    if not getattr(torch.Tensor.cuda, "_ifd_patched", False):
        _orig_cuda = torch.Tensor.cuda

        def _cuda(self, *a, **k):
            return self

        _cuda._ifd_patched = True
        _cuda._ifd_orig = _orig_cuda
        torch.Tensor.cuda = _cuda


def restore_cuda():
    """Undo the Tensor.cuda identity patch (a process that also drives the GPU calls this when it is done with the
    reference classes: RepulsionLoss needs the patch at every forward, defense/repulsion_loss.py:47)."""
    import torch
    orig = getattr(torch.Tensor.cuda, "_ifd_orig", None)
    if orig is not None:
        torch.Tensor.cuda = orig


def load(which: str):
    """which in {'ConvONet', 'ONet'} -> namespace with the reference modules (CPU)."""
    assert which in ("ConvONet", "ONet")
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF_ROOT)
    _common_shims()
    pkg = "src" if which == "ConvONet" else "im2mesh"
    root = os.path.join(REF_ROOT, which)
    _purge(["src", "im2mesh", "defense"])
    # stub the compiled extensions of this package before anything imports them
    importlib.invalidate_caches()
    sys.path.insert(0, root)
    try:
        for ext, attrs in (("libmcubes", {}), ("libsimplify", {"simplify_mesh": None}),
                           ("libmise", {"MISE": None})):
            _stub("%s.utils.%s" % (pkg, ext), **attrs)
        # ONet imports a few more third-party bits at package import time
        for name in ("skimage", "skimage.transform", "imageio", "plyfile", "tensorboardX", "h5py",
                     "PIL", "PIL.Image", "torchvision", "torchvision.models", "torchvision.transforms",
                     "torchvision.utils"):
            if name not in sys.modules:
                try:
                    importlib.import_module(name)
                except Exception:
                    _stub(name)
        ns = types.SimpleNamespace()
        ns.root = root
        ns.config = importlib.import_module(pkg + ".config")
        ns.defense = importlib.import_module("defense")
        ns.pn_utils = importlib.import_module("defense.pn_utils")
        ns.repulsion = importlib.import_module("defense.repulsion_loss")
        ns.sor = importlib.import_module("defense.SOR")
        if which == "ConvONet":
            ns.decoder = importlib.import_module("src.conv_onet.models.decoder")
            ns.encoder = importlib.import_module("src.encoder.pointnet")
            ns.common = importlib.import_module("src.common")
            ns.layers = importlib.import_module("src.layers")
            ns.yaml = os.path.join(root, "configs", "convonet_3plane_mn40.yaml")
        else:
            ns.decoder = importlib.import_module("im2mesh.onet.models.decoder")
            ns.encoder = importlib.import_module("im2mesh.encoder.pointnet")
            ns.layers = importlib.import_module("im2mesh.layers")
            ns.yaml = os.path.join(root, "configs", "onet_mn40.yaml")
        ns.default_yaml = os.path.join(root, "configs", "default.yaml")
    finally:
        sys.path.remove(root)
    return ns


def build_model(ns):
    """config.load_config -> get_model(cpu) exactly as opt_defense.py:56-73 does (minus torch.load)."""
    import torch
    cwd = os.getcwd()
    os.chdir(ns.root)
    try:
        cfg = ns.config.load_config(ns.yaml, ns.default_yaml)
        model = ns.config.get_model(cfg, device=torch.device("cpu"), dataset=None)
    finally:
        os.chdir(cwd)
    model.eval()
    for p in model.parameters():
        p.requires_grad = False
    return cfg, model
