"""``jax.lax`` control flow as plain Python."""

import numpy as _np

from ._array import wrap


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val


def cond(pred, true_fun, false_fun, *operands):
    return true_fun(*operands) if bool(pred) else false_fun(*operands)


def _tree_index(tree, i):
    if tree is None:
        return None
    if isinstance(tree, tuple):
        vals = [_tree_index(t, i) for t in tree]
        return type(tree)(*vals) if hasattr(tree, "_fields") else tuple(vals)
    return tree[i]


def _tree_len(tree):
    if isinstance(tree, tuple):
        for t in tree:
            n = _tree_len(t)
            if n is not None:
                return n
        return None
    return None if tree is None else len(tree)


def _tree_stack(items):
    first = items[0]
    if first is None:
        return None
    if isinstance(first, tuple):
        vals = [_tree_stack([it[k] for it in items]) for k in range(len(first))]
        return type(first)(*vals) if hasattr(first, "_fields") else tuple(vals)
    return wrap(_np.stack([_np.asarray(it) for it in items]))


def scan(f, init, xs, length=None, reverse=False):
    n = length if xs is None else _tree_len(xs)
    order = range(n - 1, -1, -1) if reverse else range(n)
    carry, ys = init, [None] * n
    for i in order:
        carry, y = f(carry, _tree_index(xs, i))
        ys[i] = y
    return carry, _tree_stack(ys)
