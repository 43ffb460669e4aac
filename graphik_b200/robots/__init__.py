from graphik_b200.robots.robot_revolute import RobotRevolute  # noqa: F401
