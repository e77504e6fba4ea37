"""Stub of dominate.tags; tests only."""


class _Tag:
    def __init__(self, *a, **k):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def add(self, *a):
        return self


meta = h3 = table = tr = td = p = br = a = img = _Tag
