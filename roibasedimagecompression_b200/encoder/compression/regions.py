def region_quantization(regions_components, original_image_height, original_image_width, quality=50):
    raise NotImplementedError
