"""Import shim for `immutabledict` (bsuite/sweep.py:62,134-150): a read-only dict."""


class immutabledict(dict):  # pylint: disable=invalid-name
  def _readonly(self, *args, **kwargs):
    raise TypeError('immutabledict is read-only')

  __setitem__ = __delitem__ = clear = pop = popitem = setdefault = update = _readonly

  def __hash__(self):
    return hash(frozenset(self.items()))
