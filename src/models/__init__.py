"""Reference import path `src.models` -> the B200-native generator."""
from .generator import GeneratorJ  # noqa: F401
