class ShapedArray:  # only used in isinstance checks
  pass
