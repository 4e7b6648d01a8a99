"""Drop-in counterparts of /root/reference/encoder/compression/{clustering,merging,subregions,regions,image}.py.

Same function names, argument meaning, dict contracts and degenerate-input behaviour as the
reference; the arithmetic runs in librhccq.so on the GPU.
"""
from .clustering import (get_all_unique_colors, compute_clustering_params,                 # noqa: F401
                         cluster_palette_colors_parallel)
from .merging import merge_region_components_simple                                        # noqa: F401
from .subregions import subregion_quantization                                             # noqa: F401
from .regions import region_quantization                                                   # noqa: F401
from .image import quantize_image                                                          # noqa: F401
from .image import optimize_compressed_dtype                                               # noqa: F401
from .compression import (compress_palette, compress_indices_simple_optimized, lossless_compress_optimized,   # noqa: F401
                          save_compressed, save_compression, save_encoded)
