"""Stand-in for the slice of pymanopt==0.2.5 that GraphIK's Riemannian path uses.

TEST INFRASTRUCTURE ONLY.  pymanopt 0.2.5 (reference setup.py:19) is not
vendored under /root/reference and cannot be installed offline.  The
trust-region algorithm itself IS vendored by the reference
(graphik/solvers/trust_region.py), so only glue is restated here, from
pymanopt 0.2.5's public behaviour as the reference calls it
(riemannian_solver.py:207-209, trust_region.py:61,103,177,414,433):

  Problem.grad(x)    = manifold.egrad2rgrad(x, egrad(x))
  Problem.hess(x, a) = manifold.ehess2rhess(x, egrad(x), ehess(x, a), a)
  Problem.precon     = identity

RESTATED, NOT VERIFIED AGAINST THE pymanopt SOURCE (unavailable offline).
"""
from . import tools  # noqa: F401
from .core import Problem  # noqa: F401
