"""Minimal stand-in for pytorch3d.implicitron.tools.config (TEST INFRASTRUCTURE ONLY): annotated class attributes
become constructor keyword arguments, torch.nn.Module.__init__ runs before the fields are set and __post_init__ after
(what expand_args_fields generates), `<field>_args` dicts exist for Configurable-typed fields, run_auto_creation
calls create_<field>() / instantiates them."""
import copy
import dataclasses

import torch


def _fields(cls):
    """Annotated class attributes of the Configurable classes in the MRO (not torch.nn.Module's own annotations)."""
    hints = {}
    for klass in reversed(cls.__mro__):
        if issubclass(klass, Configurable):
            hints.update(vars(klass).get("__annotations__", {}))
    return hints


class Configurable:
    def __init__(self, **kwargs):
        if isinstance(self, torch.nn.Module):
            torch.nn.Module.__init__(self)
        hints = _fields(type(self))
        for name in hints:
            for klass in type(self).__mro__:
                if name in vars(klass):
                    default = vars(klass)[name]
                    if isinstance(default, dataclasses.Field):   # field(default_factory=...) as the real dataclass does
                        default = default.default_factory() if default.default_factory is not dataclasses.MISSING \
                            else default.default
                    setattr(self, name, copy.deepcopy(default))
                    break
        for name, t in hints.items():
            if isinstance(t, type) and issubclass(t, Configurable) and name + "_args" not in kwargs:
                setattr(self, name + "_args", {})
        unknown = [k for k in kwargs if k not in hints and not k.endswith("_args")]
        if unknown:
            raise TypeError(f"{type(self).__name__}: unexpected arguments {unknown}")
        for k, v in kwargs.items():
            setattr(self, k, v)
        self.__post_init__()

    def __post_init__(self):
        pass


class ReplaceableBase(Configurable):
    pass


class _Registry:
    def __init__(self):
        self.classes = {}

    def register(self, cls):
        # class X(torch.nn.Module, SomeReplaceableBase): Module.__init__ comes first in the MRO; the real
        # expand_args_fields generates a dataclass __init__ on X itself (fields, then __post_init__)
        if issubclass(cls, Configurable) and cls.__init__ is torch.nn.Module.__init__:
            cls.__init__ = Configurable.__init__
        self.classes[cls.__name__] = cls
        return cls

    def get(self, base, name):
        return self.classes[name]


registry = _Registry()


def expand_args_fields(cls):
    return cls


def run_auto_creation(self):
    for name, t in _fields(type(self)).items():
        if isinstance(t, type) and issubclass(t, Configurable):
            creator = getattr(self, "create_" + name, None)
            if creator is not None:
                creator()
            else:
                setattr(self, name, t(**getattr(self, name + "_args")))
