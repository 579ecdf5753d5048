"""Mirrors of the TRELLIS stage in front of the path."""
