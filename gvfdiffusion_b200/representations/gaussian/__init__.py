from .gaussian_model import GaussianModel  # noqa: F401
