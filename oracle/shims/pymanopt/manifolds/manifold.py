import numpy as np


class Manifold:
    """Base class: only `zerovec` is used by the reference (trust_region.py:187,445)."""

    def zerovec(self, X):
        return np.zeros(np.shape(X))
