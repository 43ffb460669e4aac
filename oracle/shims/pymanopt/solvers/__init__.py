from .solver import Solver  # noqa: F401


class ConjugateGradient(Solver):
    """Entirely third-party in the reference (line search + beta rules live in
    pymanopt); cannot be restated from the tree.  Parity unpinned -> refuse."""

    def __init__(self, *args, **kwargs):
        for k in ("beta_type", "orth_value", "linesearch"):
            kwargs.pop(k, None)
        super().__init__(*args, **kwargs)

    def solve(self, *args, **kwargs):
        raise NotImplementedError("pymanopt ConjugateGradient is not restated in the oracle shim")
