"""Import shim for `immutabledict` (bsuite/sweep.py:62,134-150): a read-only mapping that copies like one."""
import collections.abc


class immutabledict(collections.abc.Mapping):  # pylint: disable=invalid-name

  def __init__(self, *args, **kwargs):
    self._data = dict(*args, **kwargs)

  def __getitem__(self, key):
    return self._data[key]

  def __iter__(self):
    return iter(self._data)

  def __len__(self):
    return len(self._data)

  def __repr__(self):
    return f'immutabledict({self._data!r})'

  def __hash__(self):
    return hash(frozenset(self._data.items()))

  def __reduce__(self):
    return (immutabledict, (self._data,))
