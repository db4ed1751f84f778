"""CPU, only where /root/reference exists (not on the GPU box): oracle == imported reference on fresh inputs."""
import pytest
import torch

from oracle import popcorn_oracle as po
from oracle import reference_shim as rs
from util import max_rel

pytestmark = pytest.mark.skipif(not rs.available(), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_model():
    return rs.build_reference_model(seed=1601).eval()


@pytest.mark.parametrize("shape,padding", [((64, 64), False), ((45, 83), False), ((40, 56), True)])
def test_dense_forward(ref_model, shape, padding):
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    x = po.synthetic_input(*shape, seed=5)
    a, b = {"input": x.clone()}, {"input": x.clone()}
    with torch.no_grad():
        ref = rs.reference_forward(ref_model, a, padding=padding)
        ora = po.forward(sd, b, padding=padding)
    assert max_rel(b["building_counts"], a["building_counts"]) < 1e-3
    assert max_rel(ora["popdensemap"], ref["popdensemap"]) < 1e-3


def test_sparse_mask_is_region(ref_model):
    """with sigmoid scores builtup>0 is always true, so the index set equals admin==idx."""
    sd = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
    x = po.synthetic_input(64, 64, seed=9)
    admin = torch.zeros(1, 64, 64)
    admin[0, 10:40, 20:50] = 3
    inp = {"input": x, "admin_mask": admin, "census_idx": torch.tensor([3])}
    with torch.no_grad():
        out = po.forward(sd, inp, padding=False, sparse=True)
    assert torch.equal(out["mask"], admin == 3)


def test_read_and_normalize_matches_reference_apply_normalize():
    """oracle.read_and_normalize == the reference's own apply_normalize + concatenation (utils/utils.py:105-127, 162-171)
    on the band-reordered float32 cast of the raw window (PopulationDataset.py:565-567, 594-604), with the reference's
    dataset_stats.json.  (Tensor.cuda is a no-op here: apply_normalize calls .cuda() on the statistics.)"""
    import json
    import os
    rs.load_reference()
    import utils.utils as ru                      # the reference module (sys.path set by the shim)
    with open(os.path.join(rs.REF_ROOT, "data", "config", "dataset_stats.json")) as f:
        stats = json.load(f)
    for k in stats:
        for kk in stats[k]:
            stats[k][kk] = torch.tensor(stats[k][kk])
    assert tuple(float("%.4f" % v) for v in stats["sen2springNIR"]["mean"]) == po.REF_STATS["sen2springNIR"]["mean"]
    s2_file, s1 = po.synthetic_raw(40, 56, seed=3)
    sample = {"S2": s2_file[[2, 1, 0, 3]].float()[None], "S1": s1.float()[None]}
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        ref = ru.apply_transformations_and_normalize(sample, None, stats)["input"]
    finally:
        torch.Tensor.cuda = orig
    assert torch.equal(po.read_and_normalize(s2_file, s1), ref)
