"""Config plumbing with the reference's semantics (factorizer/utils/helpers.py:36-147): any
constructor hook may be a class/callable or a ``(callable, {kwargs}, (args...))`` spec."""
from __future__ import annotations

import functools
import inspect
import itertools
import operator
from collections.abc import Sequence
from typing import Any, Callable, Iterable

__all__ = ["as_tuple", "cumprod", "has_args", "partialize", "is_partializable", "Universaltuple"]


class Universaltuple(tuple):
    """A tuple that claims to contain everything (reference helpers.py:21-33)."""

    def __contains__(self, item: Any) -> bool:
        return True


def as_tuple(obj: Any) -> tuple:
    if isinstance(obj, Sequence) and not isinstance(obj, str):
        return tuple(obj)
    return (obj,)


def cumprod(values: Iterable[float]) -> list:
    return list(itertools.accumulate(values, operator.mul))


def has_args(obj: Any, keywords) -> bool:
    if not callable(obj):
        return False
    try:
        params = inspect.signature(obj).parameters
    except ValueError:
        return False
    return all(k in params for k in as_tuple(keywords))


def partialize(spec) -> Callable:
    """``cls`` -> ``cls``;  ``(cls, {kw}, (args,), scalar)`` -> ``functools.partial(cls, *args, **kw)``."""
    if callable(spec):
        return spec
    if isinstance(spec, Sequence) and len(spec) > 0 and callable(spec[0]):
        args, kwargs = [], {}
        for item in spec[1:]:
            if isinstance(item, dict):
                kwargs.update(item)
            elif isinstance(item, Sequence) and not isinstance(item, str):
                args.extend(item)
            else:
                args.append(item)
        return functools.partial(spec[0], *args, **kwargs)
    raise TypeError(f"Expected a callable or valid tuple, got {type(spec).__name__}")


def is_partializable(obj: Any) -> bool:
    if callable(obj):
        return True
    return isinstance(obj, Sequence) and not isinstance(obj, str) and len(obj) > 0 and callable(obj[0])
