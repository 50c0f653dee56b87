def resize(*args, **kwargs):
  raise NotImplementedError('skimage is not installed; ImageObservation is outside the hot path')
