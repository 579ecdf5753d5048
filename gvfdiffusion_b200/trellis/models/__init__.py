from .sparse_structure_flow import SparseStructureFlowModel  # noqa: F401
from .sparse_structure_vae import SparseStructureDecoder  # noqa: F401
from .structured_latent_flow import SLatFlowModel, SparseResBlock3d  # noqa: F401
from .structured_latent_vae import SLatGaussianDecoder  # noqa: F401
