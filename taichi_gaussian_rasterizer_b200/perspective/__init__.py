from .params import CameraParams
from .projection import project_to_image, apply

__all__ = ['CameraParams', 'project_to_image', 'apply']
