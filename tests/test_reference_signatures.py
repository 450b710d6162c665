"""The drop-in boundary, pinned mechanically: the public entry points of hbird_b200 have the
reference's parameter names, order and defaults.  Runs wherever a reference tree is
available (/root/reference in the dev container, baseline/_ref on a GPU box) and is skipped
otherwise.  The reference is imported with the two sys.modules shims oracle/make_golden.py uses
(pytorch_lightning stub, exact-IP faiss) — nothing of it is executed here, only inspected."""
import inspect
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIRS = [os.environ.get("HBIRD_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref")]
REF = next((d for d in REF_DIRS if os.path.isdir(os.path.join(d, "hbird"))), None)
pytestmark = pytest.mark.skipif(REF is None, reason="no reference tree (/root/reference or baseline/_ref)")


@pytest.fixture(scope="module")
def ref():
    from oracle.make_golden import install_shims

    install_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import hbird.hbird_eval as r_eval
    import hbird.nn.search_base as r_base
    import hbird.nn.search_faiss as r_faiss
    import hbird.utils.eval_metrics as r_metrics

    return {"eval": r_eval, "base": r_base, "faiss": r_faiss, "metrics": r_metrics}


def params(fn):
    """[(name, kind, default)] — annotations are presentation, not contract."""
    return [(p.name, p.kind, p.default) for p in inspect.signature(fn).parameters.values()]


def test_engine_and_entry_point_signatures_equal_the_reference(ref):
    import hbird_b200

    assert params(hbird_b200.HbirdEvaluation.__init__) == params(ref["eval"].HbirdEvaluation.__init__)
    assert params(hbird_b200.HbirdEvaluation.evaluate) == params(ref["eval"].HbirdEvaluation.evaluate)
    assert params(hbird_b200.hbird_evaluation) == params(ref["eval"].hbird_evaluation)
    # the methods third-party code overrides or calls
    for name in ("_create_nn", "_create_memory", "load_memory"):
        ours, theirs = params(getattr(hbird_b200.HbirdEvaluation, name)), params(getattr(ref["eval"].HbirdEvaluation, name))
        if name == "_create_nn":  # default backend name of a private helper: 'b200' here, 'faiss' there
            ours = [(n, k, d if n != "nn_method" else "faiss") for n, k, d in ours]
        assert ours == theirs, name


def test_plugin_base_class_and_metric_signatures_equal_the_reference(ref):
    import hbird_b200
    from hbird_b200.nn.search_base import NearestNeighborSearchBase

    for name in ("__init__", "_initialize_index", "_add_features_to_index", "find_nearest_neighbors"):
        assert params(getattr(NearestNeighborSearchBase, name)) == params(getattr(ref["base"].NearestNeighborSearchBase, name)), name
    R, M = ref["metrics"].PredsmIoU, hbird_b200.PredsmIoU
    for name in ("__init__", "update", "compute", "reset", "compute_miou"):
        assert params(getattr(M, name)) == params(getattr(R, name)), name


def test_b200_plugin_accepts_the_faiss_backend_arguments_in_the_same_positions(ref):
    """NearestNeighborSearchB200(feature_memory, n_neighbors, distance_measure, idx_shard, use_fp16,
    gpu_ids, ...) — the faiss backend's parameters first, same order and defaults
    (search_faiss.py:7), then its own keyword-only extras."""
    from hbird_b200 import NearestNeighborSearchB200

    theirs = [p for p in params(ref["faiss"].NearestNeighborSearchFaiss.__init__) if p[0] != "kwargs"]
    ours = params(NearestNeighborSearchB200.__init__)
    assert ours[:len(theirs)] == theirs
    assert ours[-1][0] == "kwargs"  # present so that unknown names can be rejected with a clear message
