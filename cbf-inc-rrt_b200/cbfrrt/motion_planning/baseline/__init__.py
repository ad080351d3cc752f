from .tsa import steer, steer_fused, edge_checking, edge_checking_batch, edge_checking_eps, edge_checking_reference_loop
from .batch_tsa import (steer as batch_steer, steer_fused as batch_steer_fused, steer_fused_reference as batch_steer_fused_reference, steer_segments_fused,
                        explore_nearest, explore_nearest_each)
from .batch_rrt import SearchTree, insert_new_state, global_explore, batch_RRT_plan

__all__ = ["SearchTree", "insert_new_state", "global_explore", "batch_RRT_plan", "steer", "steer_fused", "edge_checking", "edge_checking_batch", "edge_checking_eps",
           "edge_checking_reference_loop", "batch_steer", "batch_steer_fused", "batch_steer_fused_reference", "steer_segments_fused", "explore_nearest", "explore_nearest_each"]
