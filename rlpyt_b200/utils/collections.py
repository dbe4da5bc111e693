"""Tuple-of-arrays containers of the sample buffers (host-side mirror of
``rlpyt/utils/collections.py:16-133`` ``namedarraytuple`` and ``:206`` ``AttrDict``).

Same observable behaviour as the reference (documented example at collections.py:25-45):
fields by name, ``x[loc]`` indexes every field and returns the same type, ``x[loc] = v``
assigns field-by-field (same structure) or broadcasts ``v``, ``None`` fields are skipped,
``"name" in x`` tests field names, ``x.get(i)`` is plain tuple indexing and ``x.items()``
yields ``(name, value)``.  Works on numpy arrays and on torch tensors of any device, which is
what lets one ``Samples`` layout describe pinned-host step buffers and device-resident
``[T,B]`` buffers alike.  Classes keep the reference's shape (namedtuple subclass, MRO depth 4)
so the reference's own ``is_namedarraytuple`` heuristics and ``buffer_from_example`` accept them.
"""
import sys
from collections import namedtuple

_RESERVED = ("get", "items")
_registry = {}


def _index_all(self, loc):
    picked = []
    for name, field in zip(self._fields, self):
        if field is None:
            picked.append(None)
            continue
        try:
            picked.append(field[loc])
        except IndexError as err:
            raise Exception(f"Occured in {type(self)} at field '{name}'.") from err
    return type(self)(*picked)


def _assign_all(self, loc, value):
    same_layout = isinstance(value, tuple) and getattr(value, "_fields", None) == self._fields
    for k, (name, field) in enumerate(zip(self._fields, self)):
        v = tuple.__getitem__(value, k) if same_layout else value
        if field is None:
            if same_layout and v is not None:
                raise Exception(f"Occured in {type(self)} at field '{name}': "
                                "cannot assign into a None field.")
            continue
        try:
            field[loc] = v
        except (ValueError, IndexError, TypeError) as err:
            raise Exception(f"Occured in {type(self)} at field '{name}'.") from err


def _has_field(self, key):
    return key in self._fields


def _tuple_get(self, index):
    return tuple.__getitem__(self, index)


def _pairs(self):
    return zip(self._fields, self)


def namedarraytuple(typename, field_names, return_namedtuple_cls=False, classname_suffix=False):
    """Class factory; see module docstring.  Mirrors the call signature of
    rlpyt/utils/collections.py:16."""
    nt_name = typename
    if classname_suffix:
        nt_name, typename = typename + "_nt", typename + "_nat"
    try:  # module of the caller, so instances pickle (needed by mp workers)
        module = sys._getframe(1).f_globals.get("__name__", "__main__")
    except (AttributeError, ValueError):
        module = None
    base = namedtuple(nt_name, field_names, module=module)
    for name in base._fields:
        if name in _RESERVED:
            raise ValueError(f"Disallowed field name: {name}.")
    body = {
        "__slots__": (),
        "__doc__": f"{typename}({', '.join(base._fields)})",
        "__getitem__": _index_all,
        "__setitem__": _assign_all,
        "__contains__": _has_field,
        "get": _tuple_get,
        "items": _pairs,
    }
    for pos, name in enumerate(base._fields):
        body[name] = property(lambda self, _p=pos: tuple.__getitem__(self, _p),
                              doc=f"Alias for field number {pos}")
    cls = type(typename, (base,), body)
    cls.__module__ = base.__module__
    _registry[(cls.__module__, typename)] = cls
    return (cls, base) if return_namedtuple_cls else cls


def is_namedtuple_class(obj):
    return (isinstance(obj, type) and len(obj.mro()) == 3 and obj.mro()[1] is tuple
            and all(hasattr(obj, a) for a in ("_fields", "_asdict", "_make", "_replace")))


def is_namedarraytuple_class(obj):
    return (isinstance(obj, type) and len(obj.mro()) == 4 and is_namedtuple_class(obj.mro()[1])
            and all(hasattr(obj, a) for a in _RESERVED))


def is_namedtuple(obj):
    return is_namedtuple_class(type(obj))


def is_namedarraytuple(obj):
    return is_namedarraytuple_class(type(obj))


def namedarraytuple_like(namedtuple_or_class, classname_suffix=False):
    """namedarraytuple class with the name/fields of a namedtuple (class or instance);
    namedarraytuples pass through (rlpyt/utils/collections.py:175-203)."""
    x = namedtuple_or_class
    if is_namedarraytuple(x):
        return type(x)
    if is_namedarraytuple_class(x):
        return x
    if is_namedtuple(x):
        return namedarraytuple(type(x).__name__, x._fields, classname_suffix=classname_suffix)
    if is_namedtuple_class(x):
        return namedarraytuple(x.__name__, x._fields, classname_suffix=classname_suffix)
    raise TypeError(f"Input must be namedtuple or namedarraytuple instance or class, got {type(x)}.")


class AttrDict(dict):
    """dict with attribute access (rlpyt/utils/collections.py:206-226)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self

    def copy(self):
        return type(self)(**{k: (v.copy() if isinstance(v, AttrDict) else v) for k, v in self.items()})
