"""Search-space algebra for the autotuner (the role of utils/space.py in the reference: `For`, `Table`,
`Unit`, `*` = cartesian product, `+` = concatenation; spaces iterate as tuples of (field, value) pairs and
`iter_dict()` yields dicts).  Re-implemented around one small base class."""
from __future__ import annotations

import itertools
from typing import Any, Dict, Iterable, Iterator, List, Sequence, Tuple, Type

__all__ = ["Space", "For", "Table", "Unit"]

Point = Tuple[Tuple[str, Any], ...]


class Space:
    """An ordered, finite set of configurations.  Subclasses provide `_points()` and `fields()`."""

    def _points(self) -> Iterator[Point]:
        raise NotImplementedError

    def fields(self) -> Tuple[str, ...]:
        raise NotImplementedError

    def __iter__(self) -> Iterator[Point]:
        return self._points()

    def __len__(self) -> int:
        return sum(1 for _ in self._points())

    def __mul__(self, other: "Space") -> "Space":
        return _Product(self, other)

    def __add__(self, other: "Space") -> "Space":
        return _Concat(self, other)

    def iter_dict(self, dict_class: Type = dict):
        for point in self:
            yield dict_class(point)


class Unit(Space):
    """The space with exactly one, empty, configuration (neutral element of `*`)."""

    def _points(self):
        yield ()

    def fields(self):
        return ()

    def __len__(self):
        return 1


class For(Space):
    """One field ranging over a list of values."""

    def __init__(self, name: str, values: Iterable[Any]):
        self.name, self.values = name, list(values)

    def _points(self):
        for v in self.values:
            yield ((self.name, v),)

    def fields(self):
        return (self.name,)

    def __len__(self):
        return len(self.values)


class Table(Space):
    """Explicit rows over named columns."""

    def __init__(self, headers: Iterable[str], rows: Iterable[Iterable[Any]]):
        self.headers = list(headers)
        self.rows = [list(r) for r in rows]
        for r in self.rows:
            if len(r) != len(self.headers):
                raise RuntimeError("row width does not match the headers")

    @classmethod
    def from_dicts(cls, dicts: Iterable[Dict[str, Any]]) -> "Table":
        dicts = list(dicts)
        headers: Sequence[str] = tuple(dicts[0].keys()) if dicts else ()
        for d in dicts:
            if tuple(d.keys()) != tuple(headers):
                raise RuntimeError("All dicts must have same keys")
        return cls(headers, [list(d.values()) for d in dicts])

    def _points(self):
        for r in self.rows:
            yield tuple(zip(self.headers, r))

    def fields(self):
        return tuple(self.headers)

    def __len__(self):
        return len(self.rows)


class _Product(Space):
    def __init__(self, a: Space, b: Space):
        if set(a.fields()) & set(b.fields()):
            raise RuntimeError("Cannot have duplicated fields")
        self.a, self.b = a, b

    def _points(self):
        for x, y in itertools.product(self.a, self.b):
            yield (*x, *y)

    def fields(self):
        return self.a.fields() + self.b.fields()

    def __len__(self):
        return len(self.a) * len(self.b)


class _Concat(Space):
    def __init__(self, a: Space, b: Space):
        if set(a.fields()) != set(b.fields()):
            raise RuntimeError("Fields must be equal")
        self.a, self.b = a, b

    def _points(self):
        return itertools.chain(self.a, self.b)

    def fields(self):
        return self.a.fields()

    def __len__(self):
        return len(self.a) + len(self.b)
