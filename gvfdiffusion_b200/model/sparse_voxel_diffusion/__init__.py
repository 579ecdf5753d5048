from .sparse_vae import SparseVAE  # noqa: F401
