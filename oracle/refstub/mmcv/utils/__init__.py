"""mmcv.utils.Registry stand-in: name -> class map with cfg-dict construction."""


class Registry:
    def __init__(self, name, build_func=None, parent=None, scope=None):
        self.name = name
        self._module_dict = {}

    def get(self, key):
        return self._module_dict.get(key)

    def register_module(self, name=None, force=False, module=None):
        def _register(cls):
            self._module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return _register(module)
        return _register

    def build(self, cfg, *args, **kwargs):
        cfg = dict(cfg)
        typ = cfg.pop('type')
        cls = self._module_dict[typ] if isinstance(typ, str) else typ
        return cls(*args, **cfg, **kwargs)

    def __contains__(self, key):
        return key in self._module_dict
