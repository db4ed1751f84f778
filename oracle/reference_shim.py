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


def load_reference_dataset_class(boundaries: dict):
    """The reference's ``Population_Dataset`` class (data/PopulationDataset.py) with a stub ``rasterio`` whose
    ``open(path).read(1)`` serves the numpy boundary rasters in ``boundaries`` (path -> array).  Instances are made with
    ``object.__new__`` (the real __init__ needs the on-disk dataset); only methods that are pure functions of a few
    attributes are called: get_patch_indices (:294-334), _create_mask (:656-672), convert_popmap_to_census (:675-729),
    adjust_map_to_census (:823-852)."""
    load_reference()

    class _Src:
        def __init__(self, arr):
            self.arr = arr

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def read(self, band, window=None):
            return self.arr

    rio = _stub("rasterio")
    rio.open = lambda path, mode="r", **kw: _Src(boundaries[path])
    cwd = os.getcwd()
    os.chdir(REF_ROOT)
    real_isdir = os.path.isdir
    os.path.isdir = lambda p: True if str(p) == "/scratch/metzgern/HAC/data" else real_isdir(p)
    try:
        import data.PopulationDataset as pds
    finally:
        os.chdir(cwd)
        os.path.isdir = real_isdir
    pds.rasterio = rio
    return pds.Population_Dataset


def load_reference_run_eval():
    """The reference's ``run_eval`` module (its ``Trainer.test_target`` is the evaluation loop, run_eval.py:83-203).
    Needs: a stub ``configargparse`` (absent here; arguments/eval.py parses at import, so sys.argv is emptied meanwhile),
    the rasterio stub of load_reference_dataset_class, cwd = reference root.  The caller patches ``run_eval.torch`` /
    ``run_eval.ips`` / ``run_eval.wandb`` for a CPU run at small tile sizes (tests/test_oracle_vs_reference.py)."""
    import argparse
    load_reference_dataset_class({})

    class _AP(argparse.ArgumentParser):
        def add_argument(self, *a, **k):
            k.pop("is_config_file", None)
            return super().add_argument(*a, **k)

        def format_values(self):
            return ""

    cap = _stub("configargparse")
    cap.ArgumentParser = _AP
    argv, cwd, real_isdir = sys.argv, os.getcwd(), os.path.isdir
    sys.argv = ["run_eval.py"]
    os.chdir(REF_ROOT)
    os.path.isdir = lambda p: True if str(p) == "/scratch/metzgern/HAC/data" else real_isdir(p)
    try:
        import run_eval
    finally:
        sys.argv = argv
        os.chdir(cwd)
        os.path.isdir = real_isdir
    return run_eval
