"""Reference import path `src.models`: the B200-native generator plus the two tensor-library modules of the adversarial /
perceptual branch (same three names the reference package exports)."""
from .discriminator import DiscriminatorN_IN  # noqa: F401
from .generator import GeneratorJ  # noqa: F401
from .perception import PerceptualVGG19  # noqa: F401

__all__ = ["GeneratorJ", "DiscriminatorN_IN", "PerceptualVGG19"]
