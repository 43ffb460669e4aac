from .._groups import SO2Matrix
