from .uncompression import (load_compressed, lossless_decompress, decompress_color_quantization,   # noqa: F401
                            quality_metrics)
