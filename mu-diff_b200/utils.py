"""Model registry decorators (backbones/utils.py:10-30).  Unlike the reference, registering
the same name twice replaces the entry instead of raising, so the main and the 'healthy'
generator files can live in one process."""
_MODELS = {}


def register_model(cls=None, *, name=None):
    def _register(cls):
        _MODELS[name if name is not None else cls.__name__] = cls
        return cls
    return _register if cls is None else _register(cls)


def get_model(name):
    return _MODELS[name]
