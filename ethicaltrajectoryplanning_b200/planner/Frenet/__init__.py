"""Mirror of ``planner/Frenet`` (frenet_planner.py and utils/) for the scored path."""
