from .structured_latent_flow import SLatFlowModel, SparseResBlock3d  # noqa: F401
from .structured_latent_vae import SLatGaussianDecoder  # noqa: F401
