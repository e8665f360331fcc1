from . import normalization, attention, CNN, linear, hypermixing, losses  # noqa: F401
