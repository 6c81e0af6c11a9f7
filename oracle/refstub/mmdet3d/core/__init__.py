def bbox3d2result(*a, **k):
    raise NotImplementedError
