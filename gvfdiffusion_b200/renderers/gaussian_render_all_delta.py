"""The reference keeps a (dead) duplicate of the renderer under this name
(renderers/gaussian_render_all_delta.py, same render logic); both module paths resolve to the
one implementation."""
from .gaussian_render import GaussianRenderer, edict  # noqa: F401
