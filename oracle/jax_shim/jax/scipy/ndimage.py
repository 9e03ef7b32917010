def map_coordinates(*a, **k):
  raise NotImplementedError('map_coordinates is off the hot path; not provided by the stand-in')
