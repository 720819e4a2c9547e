from .resnet import Model  # noqa: F401
