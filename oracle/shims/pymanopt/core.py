class Problem:
    def __init__(self, manifold, cost, egrad=None, ehess=None, grad=None,
                 hess=None, arg=None, precon=None, verbosity=2):
        self.manifold = manifold
        self.cost = cost
        self.egrad = egrad
        self.ehess = ehess
        self._grad = grad
        self._hess = hess
        self.verbosity = verbosity
        self.precon = precon if precon is not None else (lambda x, d: d)

    @property
    def grad(self):
        if self._grad is None:
            egrad, man = self.egrad, self.manifold
            self._grad = lambda x: man.egrad2rgrad(x, egrad(x))
        return self._grad

    @property
    def hess(self):
        if self._hess is None:
            egrad, ehess, man = self.egrad, self.ehess, self.manifold
            # pymanopt 0.2.5 evaluates egrad(x) on every Hessian-vector call
            self._hess = lambda x, a: man.ehess2rhess(x, egrad(x), ehess(x, a), a)
        return self._hess
