from .problems import (propose_test_cases, generate_refined_test_cases, append_problem, read_problem_stream,
                       collect_problems, save_collected, get_test_cases)

from .eval_rrt import eval_batchrrt_vanilla, eval_batchrrt_mindis, eval_rrt_all, summarize, seed_everything

__all__ = ["eval_batchrrt_vanilla", "eval_batchrrt_mindis", "eval_rrt_all", "summarize", "seed_everything",
           "propose_test_cases", "generate_refined_test_cases", "append_problem", "read_problem_stream",
           "collect_problems", "save_collected", "get_test_cases"]
