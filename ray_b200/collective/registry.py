"""Backend registry (python/ray/util/collective/backend_registry.py:7-40,46-122).

Same contract as the reference: names are case-insensitive, a backend must subclass
``BaseGroup``, double registration is a ``ValueError``, and registering ``"X"`` also makes
``types.Backend.X`` resolvable.
"""
from __future__ import annotations

from typing import Dict, Type

from . import types
from .base_group import BaseGroup


class BackendRegistry:
    def __init__(self):
        self._classes: Dict[str, Type[BaseGroup]] = {}

    def put(self, name: str, group_cls: Type[BaseGroup]) -> None:
        key = name.upper()
        if not (isinstance(group_cls, type) and issubclass(group_cls, BaseGroup)):
            raise TypeError(f"{group_cls} is not a subclass of BaseGroup")
        if key in self._classes:
            raise ValueError(f"Backend {key} already registered")
        self._classes[key] = group_cls

    def get(self, name: str) -> Type[BaseGroup]:
        key = name.upper()
        try:
            return self._classes[key]
        except KeyError:
            raise ValueError(f"Backend {key} not registered") from None

    def is_registered(self, name: str) -> bool:
        return name.upper() in self._classes

    def check(self, name: str) -> bool:
        """Registered and usable on this machine."""
        try:
            return bool(self.get(name).check_backend_availability())
        except (ValueError, AttributeError):
            return False


_global_registry = BackendRegistry()


def register_collective_backend(name: str, group_cls: Type[BaseGroup]) -> None:
    """Register ``group_cls`` under ``name`` in this process (must be repeated in every
    actor process, backend_registry.py:55-58)."""
    _global_registry.put(name, group_cls)
    key = name.upper()
    if not hasattr(types.Backend, key):
        setattr(types.Backend, key, key)
