from ..data_types import RasterConfig
from .function import rasterize, rasterize_with_tiles, RasterOut, set_raster_options

__all__ = ['RasterConfig', 'rasterize', 'rasterize_with_tiles', 'RasterOut', 'set_raster_options']
