def merge_region_components_simple(region_components, roi_bbox):
    raise NotImplementedError
