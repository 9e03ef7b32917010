from .tree_util import tree_map as map, tree_flatten as flatten, tree_unflatten as unflatten, tree_leaves as leaves, tree_structure as structure  # noqa
