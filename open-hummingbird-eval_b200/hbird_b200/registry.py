"""nn_method registry.  The reference hard-codes its two backends in three places
(hbird_eval.py:121, :270-281; eval.py:409); here a name -> factory table replaces them so a new
backend is one `register_nn_backend` call.  "faiss" and "scann" keep the reference behaviour by
importing the reference's own classes when that package and its dependency are installed."""
from __future__ import annotations

from typing import Callable, Dict


def _b200(feature_memory, n_neighbors=30, **kw):
    from .nn.search_b200 import NearestNeighborSearchB200

    return NearestNeighborSearchB200(feature_memory, n_neighbors=n_neighbors, **kw)


def _reference_backend(module: str, cls: str) -> Callable:
    def make(feature_memory, n_neighbors=30, **kw):
        import importlib

        try:
            mod = importlib.import_module(module)
        except ImportError as e:  # the legacy backends live in the reference package
            raise ValueError(
                f"nn_method backed by {module} needs the reference package and its dependency "
                f"installed ({e}); use nn_method='b200' for the native backend."
            ) from e
        return getattr(mod, cls)(feature_memory, n_neighbors=n_neighbors, **kw)

    return make


NN_BACKENDS: Dict[str, Callable] = {
    "b200": _b200,
    "faiss": _reference_backend("hbird.nn.search_faiss", "NearestNeighborSearchFaiss"),
    "scann": _reference_backend("hbird.nn.search_scann", "NearestNeighborSearchScaNN"),
}


def register_nn_backend(name: str, factory: Callable) -> None:
    NN_BACKENDS[name] = factory


def create_nn_backend(name: str, feature_memory, n_neighbors: int = 30, **kw):
    if name not in NN_BACKENDS:
        # same exception type and wording as hbird_eval.py:281
        raise ValueError(f"Unsupported NN method. Choose from {set(NN_BACKENDS)}.")
    return NN_BACKENDS[name](feature_memory, n_neighbors=n_neighbors, **kw)


def parse_nn_params(kv_list) -> Dict[str, object]:
    """`--nn-param KEY=VALUE` (repeatable) -> ctor kwargs, with the reference CLI's coercion order
    (eval.py:444-462): true/false -> bool, then int, then float, else the raw string.  Raises
    ValueError for an item without '='."""
    out: Dict[str, object] = {}
    for item in kv_list or []:
        key, sep, val = str(item).partition("=")
        if not sep:
            raise ValueError(f"Invalid --nn-param '{item}'. Use KEY=VALUE.")
        key, val = key.strip(), val.strip()
        low = val.lower()
        if low in ("true", "false"):
            out[key] = low == "true"
            continue
        for cast in (int, float):
            try:
                out[key] = cast(val)
                break
            except ValueError:
                pass
        else:
            out[key] = val
    return out


def nn_method_choices():
    """What `--nn-method` should offer (eval.py:409 lists its two names literally)."""
    return sorted(NN_BACKENDS)
