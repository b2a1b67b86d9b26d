"""Minimal pytrees: tuples / NamedTuples / lists / dicts are nodes, None is an empty node, the rest leaves."""
from __future__ import annotations

import functools


class Partial(functools.partial):
    pass


def _is_namedtuple(x):
    return isinstance(x, tuple) and hasattr(x, "_fields")


def tree_flatten(tree, is_leaf=None):
    leaves = []

    def rec(x):
        if is_leaf is not None and is_leaf(x):
            leaves.append(x)
            return ("leaf",)
        if x is None:
            return ("none",)
        if _is_namedtuple(x):
            return ("nt", type(x), [rec(v) for v in x])
        if isinstance(x, tuple):
            return ("tuple", [rec(v) for v in x])
        if isinstance(x, list):
            return ("list", [rec(v) for v in x])
        if isinstance(x, dict):
            return ("dict", list(x.keys()), [rec(x[k]) for k in x])
        leaves.append(x)
        return ("leaf",)

    return leaves, rec(tree)


def tree_unflatten(treedef, leaves):
    it = iter(leaves)

    def rec(d):
        kind = d[0]
        if kind == "leaf":
            return next(it)
        if kind == "none":
            return None
        if kind == "nt":
            return d[1](*[rec(c) for c in d[2]])
        if kind == "tuple":
            return tuple(rec(c) for c in d[1])
        if kind == "list":
            return [rec(c) for c in d[1]]
        if kind == "dict":
            return {k: rec(c) for k, c in zip(d[1], d[2])}
        raise TypeError(kind)

    return rec(treedef)


def tree_map(f, tree, *rest, is_leaf=None):
    leaves, treedef = tree_flatten(tree, is_leaf)
    others = [tree_flatten(r, is_leaf)[0] for r in rest]
    for o in others:
        assert len(o) == len(leaves), "tree structure mismatch"
    return tree_unflatten(treedef, [f(*xs) for xs in zip(leaves, *others)])


def tree_leaves(tree):
    return tree_flatten(tree)[0]
