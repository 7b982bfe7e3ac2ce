"""Minimal pyspark.ml.param: keyword params with defaults, setX()/getX() accessors (setters return self)."""


class Params:
    _defaults = {}

    def __init__(self, **kwargs):
        self._paramMap = {}
        self._set(**kwargs)

    def _set(self, **kwargs):
        for k, v in kwargs.items():
            if k not in self._all_defaults():
                raise TypeError("%s got an unexpected keyword argument %r" % (type(self).__name__, k))
            if v is not None:
                self._paramMap[k] = v
        return self

    @classmethod
    def _all_defaults(cls):
        d = {}
        for klass in reversed(cls.__mro__):
            d.update(klass.__dict__.get("_defaults", {}))
        return d

    def getOrDefault(self, name):
        if name in self._paramMap:
            return self._paramMap[name]
        return self._all_defaults()[name]

    def isSet(self, name):
        return name in self._paramMap

    def setParams(self, **kwargs):
        return self._set(**kwargs)

    def copy(self, extra=None):
        import copy as _copy
        c = _copy.copy(self)
        c._paramMap = dict(self._paramMap)
        if extra:
            c._set(**extra)
        return c

    def explainParams(self):
        return "\n".join("%s: (default: %r%s)" % (k, v, ", current: %r" % self._paramMap[k] if k in self._paramMap else "")
                         for k, v in sorted(self._all_defaults().items()))

    def __getattr__(self, name):
        # setFoo(value) / getFoo() for every declared param
        if name.startswith("set") and len(name) > 3:
            p = name[3].lower() + name[4:]
            if p in self._all_defaults():
                return lambda value: self._set(**{p: value})
        if name.startswith("get") and len(name) > 3:
            p = name[3].lower() + name[4:]
            if p in self._all_defaults():
                return lambda: self.getOrDefault(p)
        raise AttributeError("%s has no attribute %r" % (type(self).__name__, name))
