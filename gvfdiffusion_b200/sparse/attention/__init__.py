from .windowed_attn import calc_window_partition, sparse_windowed_scaled_dot_product_self_attention  # noqa: F401
