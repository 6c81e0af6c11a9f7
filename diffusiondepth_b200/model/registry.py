"""Name -> class registries the plugin surface exposes (stand-ins for mmcv.utils.Registry /
mmdet3d HEADS as used by reference src/model/diffusion_dcbase_model.py:77-91 and
src/model/ops/depth_transform.py:7)."""


class Registry:
    def __init__(self, name):
        self.name, self._classes = name, {}

    def register_module(self, name=None):
        def deco(cls):
            self._classes[name or cls.__name__] = cls
            return cls
        return deco

    def get(self, key):
        return self._classes.get(key)

    def __contains__(self, key):
        return key in self._classes

    def build(self, cfg, **extra):
        cfg = dict(cfg)
        kind = cfg.pop("type")
        if kind not in self._classes:
            raise KeyError(f"{kind!r} is not registered in {self.name}; known: {sorted(self._classes)}")
        return self._classes[kind](**cfg, **extra)


HEADS = Registry("heads")
DEPTH_TRANSFORM = Registry("depth_transforms")
