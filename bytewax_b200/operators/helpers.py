"""Small helpers for writing operator callbacks (mirror of ``bytewax.operators.helpers``,
pysrc/bytewax/operators/helpers.py:9-80)."""

from typing import Callable, Dict, TypeVar

K = TypeVar("K")
V = TypeVar("V")


def map_dict_value(key: K, mapper: Callable[[V], V]) -> Callable[[Dict[K, V]], Dict[K, V]]:
    """A mapper for ``op.map`` that rewrites one entry of a dict item in place and hands the dict on."""

    def apply(item: Dict[K, V]) -> Dict[K, V]:
        item[key] = mapper(item[key])
        return item

    return apply
