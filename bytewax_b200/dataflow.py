"""Dataflow data model: the host-side mirror of ``bytewax.dataflow``.

Same observable contract as the reference (pysrc/bytewax/dataflow.py) because
the engine identifies work by it: a ``Dataflow`` owns a tree of frozen step
objects; a step's class is named after its operator function
(src/dataflow.rs:72-74), core steps derive from ``_CoreOperator``
(src/dataflow.rs:84-90), arguments are read back by attribute
(src/dataflow.rs:68-70) and ports carry ``stream_id`` strings
(src/dataflow.rs:91-113).  Step ids are ``parent.child`` paths
(dataflow.py:560-566, 603).  The implementation is this repository's own:
port-ness is decided from the *values* passed at call time, not from type
annotations, and step classes are made with ``type()``.
"""

from __future__ import annotations

import dataclasses
import functools
import inspect
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Generic, Iterable, List, Optional, TypeVar

X = TypeVar("X")

__all__ = ["Dataflow", "DataflowId", "MultiPort", "Operator", "Port", "SinglePort", "Stream", "f_repr", "operator"]


def f_repr(f: Callable) -> str:
    """Short printable name of a callable (dataflow.py:54-75)."""
    if hasattr(f, "__qualname__"):
        mod = getattr(f, "__module__", None)
        return f"<function {mod + '.' if mod else ''}{f.__qualname__}>"
    return repr(f)


@dataclass(frozen=True)
class SinglePort:
    """Reference to one stream crossing a step boundary."""

    port_id: str
    stream_id: str

    @property
    def stream_ids(self) -> Dict[str, str]:
        return {"stream": self.stream_id}


@dataclass(frozen=True)
class MultiPort:
    """Reference to several streams (``*args`` / ``**kwargs`` ports)."""

    port_id: str
    stream_ids: Dict[Any, str]


Port = (SinglePort, MultiPort)


@dataclass(frozen=True)
class DataflowId:
    flow_id: str


@dataclass(frozen=True)
class Operator:
    """Base of every generated step class."""

    step_name: str
    step_id: str
    substeps: List["Operator"]
    ups_names: tuple = ()
    dwn_names: tuple = ()


@dataclass(frozen=True)
class _CoreOperator(Operator):
    """Marker base: the engine executes these; everything else only nests."""


class _Scope:
    """Where new steps get appended while an operator function runs."""

    __slots__ = ("parent_id", "substeps", "flow")

    def __init__(self, parent_id: str, substeps: list, flow: "Dataflow"):
        self.parent_id, self.substeps, self.flow = parent_id, substeps, flow


@dataclass(frozen=True)
class Dataflow:
    """Dataflow definition; add steps with :mod:`bytewax_b200.operators`."""

    flow_id: str
    substeps: List[Operator] = field(default_factory=list)
    _scope: Any = field(default=None, compare=False, repr=False)

    def __post_init__(self):
        if "." in self.flow_id:
            raise ValueError("flow ID can't contain a period `.`")
        if self._scope is None:
            object.__setattr__(self, "_scope", _Scope(self.flow_id, self.substeps, self))

    def _rescoped(self, scope: _Scope) -> "Dataflow":
        return dataclasses.replace(self, _scope=scope)


@dataclass(frozen=True)
class Stream(Generic[X]):
    """Handle on a stream of items; pass it to operator functions."""

    stream_id: str
    _scope: Any = field(compare=False, repr=False)

    def flow(self) -> Dataflow:
        return self._scope.flow

    def then(self, op_fn: Callable, step_id: str, *args, **kwargs):
        """Fluent chaining: ``s.then(op.map, "id", f)`` == ``op.map("id", s, f)``."""
        return op_fn(step_id, self, *args, **kwargs)

    def _rescoped(self, scope: _Scope) -> "Stream":
        return dataclasses.replace(self, _scope=scope)


def _scopes_of(val) -> Iterable[_Scope]:
    if isinstance(val, (Stream, Dataflow)):
        yield val._scope
    elif isinstance(val, (tuple, list)):
        for v in val:
            if isinstance(v, Stream):
                yield v._scope
    elif isinstance(val, dict):
        for v in val.values():
            if isinstance(v, Stream):
                yield v._scope


def _rescope(val, scope: _Scope):
    if isinstance(val, (Stream, Dataflow)):
        return val._rescoped(scope)
    if isinstance(val, tuple) and any(isinstance(v, Stream) for v in val):
        return tuple(_rescope(v, scope) for v in val)
    if isinstance(val, dict) and any(isinstance(v, Stream) for v in val.values()):
        return {k: _rescope(v, scope) for k, v in val.items()}
    if dataclasses.is_dataclass(val) and not isinstance(val, type) and not isinstance(val, (Stream, Dataflow)):
        changed = {
            f.name: getattr(val, f.name)._rescoped(scope)
            for f in dataclasses.fields(val)
            if isinstance(getattr(val, f.name), Stream)
        }
        if changed:
            return dataclasses.replace(val, **changed)
    return val


def _to_ref(val, ref_id: str):
    """Scoped handles become plain references inside step objects."""
    if isinstance(val, Stream):
        return SinglePort(ref_id, val.stream_id)
    if isinstance(val, Dataflow):
        return DataflowId(val.flow_id)
    if isinstance(val, tuple) and val and all(isinstance(v, Stream) for v in val):
        return MultiPort(ref_id, {i: v.stream_id for i, v in enumerate(val)})
    if isinstance(val, dict) and val and all(isinstance(v, Stream) for v in val.values()):
        return MultiPort(ref_id, {k: v.stream_id for k, v in val.items()})
    return val


_CLASS_CACHE: Dict[Any, type] = {}


def _step_class(builder: Callable, core: bool, field_names: tuple) -> type:
    key = (builder.__module__, builder.__qualname__, core, field_names)
    cls = _CLASS_CACHE.get(key)
    if cls is None:
        base = _CoreOperator if core else Operator
        cls = dataclasses.make_dataclass(
            builder.__name__, [(n, Any, field(default=None)) for n in field_names], bases=(base,), frozen=True, eq=False
        )
        cls.__module__ = builder.__module__
        cls.__doc__ = f"`{builder.__name__}` operator data model."
        _CLASS_CACHE[key] = cls
    return cls


def operator(builder=None, *, _core: bool = False):
    """Decorator turning a builder function into an operator.

    Calling the decorated function inside a dataflow runs the builder in a
    nested scope (its own operator calls become ``substeps``), then appends one
    step object describing the call to the enclosing scope.
    """

    def deco(fn: Callable) -> Callable:
        sig = inspect.signature(fn)
        if "step_id" not in sig.parameters:
            raise TypeError("builder function requires a 'step_id' parameter")

        ups_params = tuple(n for n in sig.parameters if n in ("up", "ups", "sides", "left", "right"))

        @functools.wraps(fn)
        def call(*args, **kwargs):
            try:
                bound = sig.bind(*args, **kwargs)
            except TypeError as ex:
                raise TypeError(f"operator {fn.__name__!r} called incorrectly; see cause above") from ex
            bound.apply_defaults()
            step_name = bound.arguments["step_id"]
            if not isinstance(step_name, str):
                raise TypeError("'step_id' must be a `str`")
            if "." in step_name:
                raise ValueError("'step_id' can't contain any periods '.'")
            # upstream-looking parameters must really be streams (dataflow.py:550-558)
            for pname in ups_params:
                v = bound.arguments.get(pname)
                vals = v if isinstance(v, tuple) else (list(v.values()) if isinstance(v, dict) else [v])
                for one in vals:
                    if not isinstance(one, Stream):
                        raise TypeError(
                            f"{pname!r} argument must be a `Stream`; got a {type(one)!r} instead; did you forget "
                            "to unpack the result of an operator that returns multiple streams?"
                        )
            scopes = {id(s): s for v in bound.arguments.values() for s in _scopes_of(v)}
            if len(scopes) != 1:
                raise AssertionError(
                    f"inconsistent stream scoping; found multiple scopes {list(scopes.values())!r}; expected one"
                )
            outer = next(iter(scopes.values()))
            inner_id = f"{outer.parent_id}.{step_name}"
            inner = _Scope(inner_id, [], None)
            inner.flow = outer.flow._rescoped(inner)
            call_args = {k: _rescope(v, inner) for k, v in bound.arguments.items()}
            call_args["step_id"] = inner_id
            recorded = dict(call_args)
            pos, kw = [], {}
            for pname, param in sig.parameters.items():
                v = call_args[pname]
                if param.kind == param.VAR_POSITIONAL:
                    pos.extend(v)
                elif param.kind == param.VAR_KEYWORD:
                    kw.update(v)
                elif param.kind == param.KEYWORD_ONLY:
                    kw[pname] = v
                else:
                    pos.append(v)
            out = fn(*pos, **kw)
            outs: Dict[str, Any] = {}
            if isinstance(out, Stream) or (isinstance(out, (tuple, dict)) and _to_ref(out, "") is not out):
                outs["down"] = out
            elif out is None:
                pass
            elif dataclasses.is_dataclass(out) and not isinstance(out, type):
                for f in dataclasses.fields(out):
                    outs[f.name] = getattr(out, f.name)
            else:
                outs["down"] = out
            clash = set(outs) & set(recorded)
            if clash:
                raise TypeError(f"{sorted(clash)!r} are both a builder parameter and a return field name")
            ups = tuple(k for k, v in recorded.items() if k != "step_id" and _to_ref(v, "") is not v and not isinstance(v, Dataflow))
            dwn = tuple(k for k, v in outs.items() if _to_ref(v, "") is not v)
            fields = {k: _to_ref(v, f"{inner_id}.{k}") for k, v in {**recorded, **outs}.items() if k != "step_id"}
            cls = _step_class(fn, _core, tuple(fields))
            step = cls(step_name=step_name, step_id=inner_id, substeps=inner.substeps, ups_names=ups, dwn_names=dwn, **fields)
            if any(s.step_id == inner_id for s in outer.substeps):
                raise ValueError(f"step {inner_id!r} already exists; do you have two steps with the same ID?")
            outer.substeps.append(step)
            return _rescope(out, outer)

        call._op_cls_name = fn.__name__  # type: ignore[attr-defined]
        return call

    return deco(builder) if builder is not None else deco
