"""Makes `from src.models.blocks.sageblock import SageBlock` resolve to this package.

The reference imports SageBlock at src/models/grusage.py:7 and
src/models/map/mapencoder.py:4 (`from ..blocks.sageblock import SageBlock`).  Calling
install_reference_shim() before importing the reference's `src.models` registers a
module object under the name `src.models.blocks.sageblock` whose SageBlock is ours, so
grusage.py / mapencoder.py / main.py / test.py / rcv.py run unchanged.
The alternative is the one-line file replacement shown there.

`install_reference_shim(extras=True)` also registers the components either side of the block:
`src.models.map.mapattention.MapSpatialAttention` (imported at src/models/grusage.py:9 as `..map.mapattention`) and
a minimal `torch_geometric.nn` exposing `global_mean_pool` / `global_max_pool` (imported at src/models/grusage.py:5)
when PyG itself is not installed -- which is what lets the reference's GruSage import and run in a PyG-less image.
"""
from __future__ import annotations

import sys
import types


def _register(name: str, **attrs) -> types.ModuleType:
    mod = types.ModuleType(name)
    mod.__doc__ = "sldm_gnn_b200 drop-in"
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    parent, _, leaf = name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod


def install_reference_shim(module_name: str = "src.models.blocks.sageblock", extras: bool = False) -> types.ModuleType:
    from .sageblock import SageBlock

    if extras:
        from .map_attention import MapSpatialAttention
        from .readout import global_mean_pool, global_max_pool
        _register("src.models.map.mapattention", MapSpatialAttention=MapSpatialAttention)
        try:
            import torch_geometric  # noqa: F401  (the real package wins when it is installed)
        except ImportError:
            pkg = _register("torch_geometric")
            pkg.__path__ = []       # mark as package so that `import torch_geometric.nn` resolves
            _register("torch_geometric.nn", global_mean_pool=global_mean_pool, global_max_pool=global_max_pool)

    mod = types.ModuleType(module_name)
    mod.__doc__ = "sldm_gnn_b200 drop-in for the reference's SageBlock"
    mod.SageBlock = SageBlock
    sys.modules[module_name] = mod
    # if the parent package is already imported, bind the attribute too
    parent, _, leaf = module_name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod
