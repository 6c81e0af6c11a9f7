from .weight_init import *  # noqa: F401,F403
