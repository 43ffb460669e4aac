from .._groups import SE3Matrix
