from .problems import (propose_test_cases, generate_refined_test_cases, append_problem, read_problem_stream,
                       collect_problems, save_collected, get_test_cases)

__all__ = ["propose_test_cases", "generate_refined_test_cases", "append_problem", "read_problem_stream",
           "collect_problems", "save_collected", "get_test_cases"]
