def subregion_quantization(image_rgb, subregions, quality=10, subregion_type=None, debug=False):
    raise NotImplementedError
