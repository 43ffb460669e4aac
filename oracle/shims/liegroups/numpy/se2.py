from .._groups import SE2Matrix
