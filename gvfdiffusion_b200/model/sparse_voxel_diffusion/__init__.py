from .sparse_vae import SparseVAE  # noqa: F401
from .sparse_transformer_vae import SparseTransformerVAE  # noqa: F401
