from .windowed_attn import calc_window_partition, sparse_windowed_scaled_dot_product_self_attention  # noqa: F401
from .serialized_attn import (SerializeMode, SerializeModes, calc_serialization,  # noqa: F401
                              sparse_serialized_scaled_dot_product_self_attention)
from .full_attn import sparse_scaled_dot_product_attention  # noqa: F401
