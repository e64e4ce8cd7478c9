"""Size-string parsing ('2GB' -> bytes).  Parity: reference python/utils/units.py:18-36."""
UNITS = {'KB': 2 ** 10, 'MB': 2 ** 20, 'GB': 2 ** 30, 'TB': 2 ** 40,
         'K': 2 ** 10, 'M': 2 ** 20, 'G': 2 ** 30, 'T': 2 ** 40, 'B': 1}


def parse_size(sz) -> int:
  if isinstance(sz, (int, float)):
    return int(sz)
  s = str(sz).strip().upper()
  for suffix in ('KB', 'MB', 'GB', 'TB', 'K', 'M', 'G', 'T', 'B'):
    if s.endswith(suffix):
      return int(float(s[:-len(suffix)]) * UNITS[suffix])
  return int(float(s))
