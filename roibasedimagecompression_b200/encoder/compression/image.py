def quantize_image(image_components, original_image_height, original_image_width, quality=100):
    raise NotImplementedError
