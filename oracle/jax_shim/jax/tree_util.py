"""Registry-based pytrees: leaves are anything not registered / tuple / list / dict / None."""
_REG = {}


def register_pytree_node(cls, flatten, unflatten):
  _REG[cls] = (flatten, unflatten)


def register_pytree_node_class(cls):
  _REG[cls] = (lambda x: x.tree_flatten(),
               lambda aux, ch: cls.tree_unflatten(aux, ch))
  return cls


class _Def:
  def __init__(self, kind, aux, children):
    self.kind, self.aux, self.children = kind, aux, children

  def __eq__(self, other):
    if not isinstance(other, _Def) or self.kind != other.kind:
      return False
    if self.kind == 'leaf':
      return True
    try:
      same_aux = bool(self.aux == other.aux)
    except Exception:  # pylint: disable=broad-except
      same_aux = self.aux is other.aux
    return same_aux and self.children == other.children

  def __hash__(self):
    return hash(self.kind)

  @property
  def num_leaves(self):
    return 1 if self.kind == 'leaf' else sum(c.num_leaves for c in self.children)


def _flatten(x, leaves):
  if x is None:
    return _Def('none', None, [])
  t = type(x)
  if t in _REG:
    ch, aux = _REG[t][0](x)
    return _Def(t, aux, [_flatten(c, leaves) for c in ch])
  if isinstance(x, tuple) and hasattr(x, '_fields'):
    return _Def(('namedtuple', t), None, [_flatten(c, leaves) for c in x])
  if t is tuple:
    return _Def('tuple', None, [_flatten(c, leaves) for c in x])
  if t is list:
    return _Def('list', None, [_flatten(c, leaves) for c in x])
  if t is dict:
    keys = sorted(x)
    return _Def('dict', tuple(keys), [_flatten(x[k], leaves) for k in keys])
  leaves.append(x)
  return _Def('leaf', None, [])


def _unflatten(d, it):
  if d.kind == 'leaf':
    return next(it)
  if d.kind == 'none':
    return None
  ch = [_unflatten(c, it) for c in d.children]
  if d.kind == 'tuple':
    return tuple(ch)
  if d.kind == 'list':
    return ch
  if d.kind == 'dict':
    return dict(zip(d.aux, ch))
  if isinstance(d.kind, tuple):
    return d.kind[1](*ch)
  return _REG[d.kind][1](d.aux, ch)


def tree_flatten(x, is_leaf=None):
  leaves = []
  d = _flatten(x, leaves)
  return leaves, d


def tree_unflatten(d, leaves):
  return _unflatten(d, iter(leaves))


def tree_leaves(x):
  return tree_flatten(x)[0]


def tree_structure(x):
  return tree_flatten(x)[1]


def tree_map(f, x, *rest, is_leaf=None):
  leaves, d = tree_flatten(x)
  others = []
  for r in rest:
    l, dr = tree_flatten(r)
    if len(l) != len(leaves):
      raise ValueError('tree_map: mismatched tree structures')
    others.append(l)
  return tree_unflatten(d, [f(*a) for a in zip(leaves, *others)])
