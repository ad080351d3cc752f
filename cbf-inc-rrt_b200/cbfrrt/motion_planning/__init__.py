"""Mirror of the planner-side callers of the hot path (motion_planning/baseline/{tsa,batch_tsa}.py)."""
