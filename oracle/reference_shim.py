"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

Imports the *unmodified* reference POPCORN model from /root/reference on a CPU-only
host so that (a) oracle/popcorn_oracle.py can be validated against it and (b) golden
vectors under tests/golden/ can be generated (oracle/make_golden.py).  /root/reference
does not exist on the GPU box, so nothing at test/bench run time may import this file.

Shims:
  1. stub modules pylab / matplotlib(.pyplot)   (imported at utils/utils.py:16-18, unused on the path)
  2. stub fvcore.common.config.CfgNode          (model/DDA_model/utils/experiment_manager.py:5, type hint)
  3. os.path.isdir() answers True for one hard-coded data root during import
                                                (utils/constants.py:16-25 leaves large_file_path unbound otherwise)
  4. chdir to the reference root                (relative checkpoint dir, utils/constants.py:172)
  5. force device="cpu" in load_checkpoint, no-op Module.cuda   (model/popcorn.py:57,96-97)
"""
import os
import sys
import types
import contextlib

REF_ROOT = os.environ.get("POPCORN_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "model", "popcorn.py"))


class _AnyAttrModule(types.ModuleType):
    """A module whose every missing attribute resolves to a harmless callable."""

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return lambda *a, **k: None


def _stub(name, **attrs):
    m = _AnyAttrModule(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules.setdefault(name, m)
    return sys.modules[name]


_loaded = {}


def load_reference():
    """Returns a namespace with the reference's POPCORN class, model_dict, get_model_kwargs, Args."""
    if _loaded:
        return _loaded["ns"]
    import torch
    import torch.nn as nn

    _stub("pylab")
    mpl = _stub("matplotlib")
    plt = _stub("matplotlib.pyplot")
    mpl.pyplot = plt
    fv = _stub("fvcore")
    fvc = _stub("fvcore.common")
    fvcc = _stub("fvcore.common.config", CfgNode=type("CfgNode", (dict,), {}))
    fv.common = fvc
    fvc.config = fvcc
    # shim 3: utils/constants.py:21-25 only needs os.path.isdir() to be true for one hard-coded root
    real_isdir = os.path.isdir
    os.path.isdir = lambda p: True if str(p) == "/scratch/metzgern/HAC/data" else real_isdir(p)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    try:
        import model.DDA_model.utils.networks as networks
        orig_load = networks.load_checkpoint

        def cpu_load(epoch, cfg, device):
            return orig_load(epoch, cfg, "cpu" if not torch.cuda.is_available() else device)

        networks.load_checkpoint = cpu_load
        import model.popcorn as popcorn_mod
        popcorn_mod.load_checkpoint = cpu_load
        import model.get_model as get_model_mod
    finally:
        os.chdir(cwd)
        os.path.isdir = real_isdir

    ns = types.SimpleNamespace(POPCORN=popcorn_mod.POPCORN, model_dict=get_model_mod.model_dict,
                               get_model_kwargs=get_model_mod.get_model_kwargs, Args=get_model_mod.Args,
                               networks=networks, root=REF_ROOT)
    _loaded["ns"] = ns
    return ns


@contextlib.contextmanager
def reference_cwd():
    """The reference builds its checkpoint path relative to cwd (utils/constants.py:172)."""
    import torch
    import torch.nn as nn
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    orig_cuda = nn.Module.cuda
    orig_empty = torch.cuda.empty_cache
    if not torch.cuda.is_available():
        nn.Module.cuda = lambda self, device=None: self
        torch.cuda.empty_cache = lambda: None
    try:
        yield
    finally:
        nn.Module.cuda = orig_cuda
        torch.cuda.empty_cache = orig_empty
        os.chdir(cwd)


def build_reference_model(input_channels=6, occupancymodel=True, pretrained=False, biasinit=0.9407,
                          sentinelbuildings=True, seed=1600):
    """POPCORN(...) exactly as run_eval.py:51-52 / run_train.py:60-61 build it (on CPU here)."""
    import torch
    ns = load_reference()
    torch.manual_seed(seed)
    with reference_cwd():
        m = ns.POPCORN(input_channels, feature_extractor="DDA", occupancymodel=occupancymodel,
                       pretrained=pretrained, biasinit=biasinit, sentinelbuildings=sentinelbuildings)
    return m


def reference_forward(model, inputs, **kw):
    with reference_cwd():
        return model(inputs, **kw)
