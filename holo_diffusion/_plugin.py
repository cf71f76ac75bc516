"""Glue between the Implicitron config system (when present) and the holo_diffusion_b200 implementation."""
from __future__ import annotations

from typing import Iterable

try:  # the real library, or tests' stand-in (oracle/pt3d_stub)
    from pytorch3d.implicitron.tools.config import (Configurable, ReplaceableBase, registry,  # noqa: F401
                                                    run_auto_creation)
    HAVE_CONFIG = True
except ImportError:  # plain re-exports
    Configurable = ReplaceableBase = object
    registry = None
    run_auto_creation = None
    HAVE_CONFIG = False


def adopt(facade, impl, children: Iterable[str] = ()):
    """Make `facade` (an object the config system constructed from its fields) stand for `impl`: the named
    sub-modules of `impl` are registered on the facade under the same names -- so parameters, state-dict keys,
    ``.to()`` / ``.cuda()`` behave as in the reference -- and `impl` itself is kept OUTSIDE the module tree (instance
    dict), where ``holo_diffusion_b200.renderer.impl_of`` finds it."""
    for name in children:
        setattr(facade, name, getattr(impl, name))
    facade.__dict__["_impl"] = impl


def fields(obj, names: Iterable[str]) -> dict:
    return {n: getattr(obj, n) for n in names}


def plain(v):
    """OmegaConf containers / enums -> plain python for the implementation's constructors."""
    import enum
    if isinstance(v, enum.Enum):
        return v.name
    if isinstance(v, dict) or type(v).__name__ == "DictConfig":
        return {k: plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)) or type(v).__name__ == "ListConfig":
        return tuple(plain(x) for x in v)
    return v
