"""Mirror of the reference's ``planner`` package for the scored path."""
