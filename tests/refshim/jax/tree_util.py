"""`jax.tree_util` / `jax.tree` subset (TEST INFRASTRUCTURE)."""
from ._core import register_pytree_node, tree_flatten, tree_map, tree_unflatten  # noqa: F401

map = tree_map


def tree_leaves(tree):
    return tree_flatten(tree)[0]


def tree_structure(tree):
    return tree_flatten(tree)[1]
