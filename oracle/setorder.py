"""ORACLE (test infrastructure only).

Iteration order of ``set(range(m)).difference(used)`` on CPython 3.12 -- the order in which
CentroidTracker registers several new detections in one frame (/root/reference/ysmr/tracker.py:193,216-217).

The algorithm lives in CPython (Objects/setobject.c, 3.12.3 here), not in the reference: small ints hash to
themselves; ``set_difference`` walks the left operand in table order (ascending for ``set(range(m))``) and
``set_add_key``s survivors into a fresh table of 8 slots: slot ``h & mask``, then up to LINEAR_PROBES = 9 further
consecutive slots when they fit below the mask, then ``perturb >>= 5; i = (i*5 + 1 + perturb) & mask``; after an
insert, if ``fill*5 >= mask*3`` the table is rebuilt at the smallest power of two > used*4 (used*2 beyond 50,000
entries), reinserting in old slot order.  Iteration is by slot.

Pinned by tests/test_oracle_setorder.py against the running interpreter's real ``set``.
"""
from __future__ import annotations

LINEAR_PROBES = 9
MIN_SIZE = 8


def _insert(table, mask, key):
    perturb = key
    i = key & mask
    while True:
        last = i + LINEAR_PROBES if i + LINEAR_PROBES <= mask else i
        for j in range(i, last + 1):
            if table[j] is None:
                table[j] = key
                return
        perturb >>= 5
        i = (i * 5 + 1 + perturb) & mask


def set_order(keys_in_insertion_order):
    """Slot order of a CPython set built by adding distinct non-negative ints one at a time."""
    mask = MIN_SIZE - 1
    table = [None] * MIN_SIZE
    fill = 0
    for key in keys_in_insertion_order:
        _insert(table, mask, key)
        fill += 1
        if fill * 5 >= mask * 3:
            want = fill * 2 if fill > 50000 else fill * 4
            size = MIN_SIZE
            while size <= want:
                size <<= 1
            old = table
            table = [None] * size
            mask = size - 1
            for k in old:
                if k is not None:
                    _insert(table, mask, k)
    return [k for k in table if k is not None]


def unused_cols_order(m, used_cols):
    return set_order([c for c in range(m) if c not in used_cols])
