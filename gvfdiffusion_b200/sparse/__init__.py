"""Sparse-voxel operators of the static-VAE path (reference sparse/)."""

_partition_cache = {}
