def __getattr__(name):
    def _missing(*a, **k):
        raise NotImplementedError(name)
    return _missing
