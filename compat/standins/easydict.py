"""Stand-in for the ``easydict`` package (used by the reference's utils/misc.py:6 and CLI scripts): a dict whose items are
also attributes, recursively.  Only used when the real package is not installed (offline boxes)."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super().__init__()
        for k, v in dict(d or {}, **kwargs).items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __delattr__(self, k):
        del self[k]

    def update(self, e=None, **f):
        for k, v in dict(e or {}, **f).items():
            self[k] = v
