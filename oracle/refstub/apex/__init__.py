"""apex placeholder so `import apex` in reference src/main.py resolves; training is out of scope."""
from . import parallel  # noqa: F401


class amp:  # noqa: N801
    @staticmethod
    def initialize(model, optimizer=None, opt_level='O0', **k):
        return (model, optimizer) if optimizer is not None else model
