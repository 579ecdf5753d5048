from .structured_latent_flow import SLatFlowModel, SparseResBlock3d  # noqa: F401
