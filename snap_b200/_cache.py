"""Device-side caches keyed on a parameter tree.

The launch plans / weight banks built from a Flax parameter tree are cached per tree OBJECT.  `id(params)` alone is not a
key: CPython recycles ids, so a tree loaded for step N + 3 can get the id of the (collected) tree of step N and would
silently run with the old weights.  An entry therefore keeps a strong reference to its tree (the id cannot be recycled
while the entry lives), a hit requires `entry.params is params`, and the cache is a small LRU so that a host that reloads
checkpoints in a loop does not accumulate device copies of every tree it has ever seen.

A tree that is MODIFIED IN PLACE between calls is not detected (hashing 48 M parameters per call would cost more than the
forward): pass a new tree, or call `clear()` on the module's cache (`module.clear_cache()`).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Callable, Hashable


class ParamCache:
    def __init__(self, max_trees: int = 2):
        self.max_trees = max_trees
        self._d: "OrderedDict[Any, tuple]" = OrderedDict()

    def lookup(self, params: Any, key: Hashable, build: Callable[[], Any]) -> Any:
        k = (id(params), key)
        hit = self._d.get(k)
        if hit is not None and hit[0] is params:
            self._d.move_to_end(k)
            return hit[1]
        val = build()
        self._d[k] = (params, val)
        self._evict()
        return val

    def _evict(self) -> None:
        # keep the entries of the `max_trees` most recently used trees (a tree may own several entries: one per shape)
        trees = []
        for (pid, _), _v in reversed(self._d.items()):
            if pid not in trees:
                trees.append(pid)
        keep = set(trees[: self.max_trees])
        for k in [k for k in self._d if k[0] not in keep]:
            del self._d[k]

    def clear(self) -> None:
        self._d.clear()

    def __len__(self) -> int:
        return len(self._d)
