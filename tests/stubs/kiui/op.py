def recenter(*a, **k):
    raise NotImplementedError('image conditioning is outside the tested path')
