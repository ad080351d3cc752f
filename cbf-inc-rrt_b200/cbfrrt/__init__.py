"""cbfrrt -- B200-native batched safety evaluation and control of manipulator states (CBF-INC-RRT hot path).

Host-side mirror of the reference's API (same module names under this package):
    cbfrrt.environment              ArmEnv, BasicRobot, FrankaPanda, Magician
    cbfrrt.neural_cbf.systems       ControlAffineSystem, ArmDynamics, ArmMindis
    cbfrrt.neural_cbf.controllers   Controller, CLFController, CBFController, NeuralObsCBFController,
                                    NeuralMindisCBFController
    cbfrrt.motion_planning.baseline steer, batch_steer, edge_checking, ...
    cbfrrt.ops                      torch custom ops (namespace cbfrrt::) over the C-ABI of libcbfrrt.so
    cbfrrt.sharded                  one-process-per-GPU sharding + NCCL all-gather of masks and controls
The CUDA library is mandatory; there is no CPU path.
"""
from . import _lib  # noqa: F401  (fails loudly when libcbfrrt.so is missing)

__all__ = ["ops", "environment", "neural_cbf", "motion_planning", "sharded"]
__version__ = "0.1.0"
