"""Makes `from src.models.blocks.sageblock import SageBlock` resolve to this package.

The reference imports SageBlock at src/models/grusage.py:7 and
src/models/map/mapencoder.py:4 (`from ..blocks.sageblock import SageBlock`).  Calling
install_reference_shim() before importing the reference's `src.models` registers a
module object under the name `src.models.blocks.sageblock` whose SageBlock is ours, so
grusage.py / mapencoder.py / main.py / test.py / rcv.py run unchanged.
The alternative is the one-line file replacement shown there.
"""
from __future__ import annotations

import sys
import types


def install_reference_shim(module_name: str = "src.models.blocks.sageblock") -> types.ModuleType:
    from .sageblock import SageBlock

    mod = types.ModuleType(module_name)
    mod.__doc__ = "sldm_gnn_b200 drop-in for the reference's SageBlock"
    mod.SageBlock = SageBlock
    sys.modules[module_name] = mod
    # if the parent package is already imported, bind the attribute too
    parent, _, leaf = module_name.rpartition(".")
    if parent in sys.modules:
        setattr(sys.modules[parent], leaf, mod)
    return mod
