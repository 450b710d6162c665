"""The UNMODIFIED reference engine driving the b200 plugin: INTEGRATION.md §1's three-line patch is
applied to the reference's own hbird_eval.py (read from baseline/_ref or /root/reference, patched in
memory, never written back) and `HbirdEvaluation(..., nn_method="b200")` is run on a GPU.  Everything
except the search — bank construction, neighbour gather, cross-attention, upsample, argmax, PredsmIoU
— is then the reference's code on the CPU, and the result must equal the golden run of the same
engine with its exact faiss index.  Skipped when no reference tree travelled to the box."""
import os
import sys
import types

import numpy as np
import pytest
import torch

from helpers import load_golden
from hbird_b200.data import SyntheticSegmentationData

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.path.join(ROOT, "baseline", "_ref"), os.environ.get("HBIRD_REFERENCE", "/root/reference")]
REF = next((d for d in REF_DIRS if os.path.isfile(os.path.join(d, "hbird", "hbird_eval.py"))), None)

PATCH = [
    # hbird/hbird_eval.py:121
    ('assert self.nn_method in ["faiss", "scann"], "Only faiss and scann are supported"',
     'assert self.nn_method in ["faiss", "scann", "b200"], "Only faiss, scann and b200 are supported"'),
    # hbird/hbird_eval.py:270-281: one more branch in _create_nn
    ('        else:\n            raise ValueError("Unsupported NN method. Choose from {\'faiss\',\'scann\'}.")',
     '        elif nn_method == "b200":\n'
     '            from hbird_b200 import NearestNeighborSearchB200\n'
     '            self.NN_algorithm = NearestNeighborSearchB200(self.feature_memory, n_neighbors=n_neighbours, **kwargs)\n'
     '        else:\n            raise ValueError("Unsupported NN method. Choose from {\'faiss\',\'scann\'}.")'),
]


@pytest.fixture(scope="module")
def patched_reference():
    if REF is None:
        pytest.skip("no reference tree (baseline/_ref or /root/reference) on this box")
    from oracle.make_golden import install_shims

    install_shims()  # pytorch_lightning stub so that hbird.data imports; the faiss shim is not used
    if REF not in sys.path:
        sys.path.insert(0, REF)
    src = open(os.path.join(REF, "hbird", "hbird_eval.py")).read()
    for old, new in PATCH:
        assert src.count(old) == 1, "INTEGRATION.md §1 no longer matches the reference source"
        src = src.replace(old, new)
    mod = types.ModuleType("hbird_eval_patched")
    mod.__file__ = os.path.join(REF, "hbird", "hbird_eval.py")
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod


@pytest.mark.parametrize("name", ["voc_tiny", "ade_tiny"])
def test_reference_engine_with_b200_plugin_reproduces_its_faiss_run(patched_reference, name):
    from hbird.models import FeatureExtractorSimple  # the reference's own wrapper

    cfg, g = load_golden(name)
    data = SyntheticSegmentationData(**cfg)
    fe = FeatureExtractorSimple(data.model, ftr_extr_fn=data.ftr_extr_fn, eval_spatial_resolution=data.S, d_model=data.d)
    calls = []
    import hbird_b200

    orig = hbird_b200.NearestNeighborSearchB200.find_nearest_neighbors

    def spy(self, q, k=None):
        out = orig(self, q, k)
        calls.append(out)
        return out

    hbird_b200.NearestNeighborSearchB200.find_nearest_neighbors = spy
    try:
        ev = patched_reference.HbirdEvaluation(fe, data.train_dataloader(), num_classes=data.C, n_neighbours=30,
                                               device="cpu", nn_method="b200", nn_params={"k_prime": 64})
        miou, details = ev.evaluate(data.val_dataloader(), data.S, return_knn_details=True,
                                    ignore_index=data.ignore_index)
    finally:
        hbird_b200.NearestNeighborSearchB200.find_nearest_neighbors = orig
    assert isinstance(ev.NN_algorithm, hbird_b200.NearestNeighborSearchB200) and calls
    assert abs(float(miou) - float(g["miou"])) <= 5e-4
    idx = np.concatenate([c[0] for c in calls])
    dist = np.concatenate([c[1] for c in calls])
    assert idx.dtype == np.int64 and dist.dtype == np.float32
    hit = (idx[:, :, None] == g["knn_idx"][:, None, :]).any(axis=2).mean()
    assert hit >= 0.999
    assert (np.abs(dist - g["knn_dist"]) / np.abs(g["knn_dist"])).max() <= 1e-3
    lh = details["knns_ca_labels"].numpy().reshape(g["label_hat"].shape)
    np.testing.assert_allclose(lh, g["label_hat"], rtol=0, atol=2e-5)
