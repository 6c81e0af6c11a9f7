from . import multi_scale_deform_attn  # noqa: F401
