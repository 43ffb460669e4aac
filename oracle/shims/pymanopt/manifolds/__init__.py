from .manifold import Manifold  # noqa: F401
