from .gaussian_render import GaussianRenderer  # noqa: F401  (reference renderers/__init__.py:4)
