"""No-op pyplot: every attribute is a function that accepts anything and returns a harmless object."""


class _Noop:
    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return self

    def __iter__(self):
        return iter(())


_noop = _Noop()


def __getattr__(name):
    return _noop
