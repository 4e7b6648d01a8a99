"""Segmentation slot of the encoder: the label map stage 1 quantises segment by segment."""
from .slic import DbscanSegmenter, enhanced_slic_with_texture

__all__ = ["DbscanSegmenter", "enhanced_slic_with_texture"]
