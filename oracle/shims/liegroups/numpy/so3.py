from .._groups import SO3Matrix
