from .image_projector import ImageProjector  # noqa: F401
