class _Xla:
  class DeviceArray:
    pass
xla = _Xla()
