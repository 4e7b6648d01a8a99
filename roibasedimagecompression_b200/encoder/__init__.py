"""Mirror of the reference's ``encoder`` package for the hot path (encoder/compression only)."""
