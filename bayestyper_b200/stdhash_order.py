"""Iteration order of libstdc++'s `std::unordered_set<uint>` / `std::unordered_map<uint, T>`.

The reference keeps the clusters of a group in an `unordered_map<uint, VariantCluster*>` and the pending cluster merges in
`unordered_set<uint>`s (VariantFileParser.cpp:257,261,1000-1040,1064-1090); the order in which it walks them fixes the order of
the clusters inside a variant-cluster group, which cluster survives a merge (and with it the cluster index that seeds the path
search, VariantClusterGroup.cpp:138) and the order of the group's dependency edges.  Those orders are part of the result, so
the host-side builder reproduces them: this class restates the container's observable order for integer keys (identity hash,
singly linked node list with per-bucket insertion at the bucket's head, prime bucket counts, load factor 1).
tests/test_graph_builder.py pins it against the toolchain's own containers.
"""
from __future__ import annotations

import bisect

# bucket counts the rehash policy can pick (prefix of libstdc++'s prime table; groups never get near the end of it)
_PRIMES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97, 103, 109, 113, 127, 137, 139,
           149, 157, 167, 179, 193, 199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503, 541, 577, 619, 661, 709,
           761, 823, 887, 953, 1031, 1109, 1193, 1289, 1381, 1493, 1613, 1741, 1879, 2029, 2179, 2357, 2549, 2753, 2971, 3209, 3469,
           3739, 4027, 4349, 4703, 5087, 5503, 5953, 6427, 6949, 7517, 8123, 8783, 9497, 10273)
_FAST = (2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13)


def _next_bkt(n: int) -> int:
    if n < len(_FAST):
        return _FAST[n] if n else 1
    i = bisect.bisect_left(_PRIMES, n)
    if i == len(_PRIMES):
        raise ValueError("container larger than the restated prime table")
    return _PRIMES[i]


class UnorderedUInt:
    """Keys in the order a range-for over the libstdc++ container would visit them; optional mapped values."""

    def __init__(self):
        self._nb = 1
        self._next_resize = 0
        self._order: list[int] = []
        self._val: dict[int, object] = {}

    @staticmethod
    def _link(order, nb, key):
        b = key % nb
        for i, k in enumerate(order):
            if k % nb == b:
                order.insert(i, key)        # head of its bucket's run
                return
        order.insert(0, key)                # empty bucket: head of the whole list

    def _rehash(self, nb):
        new = []
        for k in self._order:
            self._link(new, nb, k)
        self._order, self._nb, self._next_resize = new, nb, nb

    def insert(self, key: int, value=None) -> bool:
        if key in self._val:
            return False
        n = len(self._order)
        if n + 1 > self._next_resize:
            min_bkts = max(n + 1, 0 if self._next_resize else 11)
            if min_bkts >= self._nb:
                self._rehash(_next_bkt(max(min_bkts + 1, self._nb * 2)))
            else:
                self._next_resize = self._nb
        self._link(self._order, self._nb, key)
        self._val[key] = value
        return True

    def erase(self, key: int) -> None:
        del self._val[key]
        self._order.remove(key)

    def __contains__(self, key):
        return key in self._val

    def __getitem__(self, key):
        return self._val[key]

    def __len__(self):
        return len(self._order)

    def __iter__(self):
        return iter(list(self._order))

    def items(self):
        return [(k, self._val[k]) for k in self._order]
