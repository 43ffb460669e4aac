from .._groups import SO2Matrix as SO2, SO3Matrix as SO3, SE2Matrix as SE2, SE3Matrix as SE3
