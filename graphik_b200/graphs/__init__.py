from graphik_b200.graphs.graph_revolute import GoalGraph, ProblemGraphRevolute  # noqa: F401
