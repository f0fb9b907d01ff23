"""Minimal stand-in for the `addict` package (absent from this image), only so that the
UNMODIFIED reference at /root/reference can be imported by tests/golden/make_golden.py
(yolov6/utils/config.py:12 does `from addict import Dict`).  Not product code."""


class Dict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if a is None:
                continue
            for k, v in (a.items() if isinstance(a, dict) else a):
                self[k] = self._wrap(v)
        for k, v in kwargs.items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, Dict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(i) for i in v)
        return v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            return self.__missing__(name)

    def __missing__(self, name):
        v = type(self)()
        self[name] = v
        return v

    def __setattr__(self, name, value):
        self[name] = self._wrap(value)

    def __delattr__(self, name):
        del self[name]

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Dict) else v) for k, v in self.items()}
