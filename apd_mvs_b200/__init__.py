"""apd_mvs_b200 — B200-native PatchMatch engine behind the APD-MVS `class APD` surface.

Only the hot path lives here (csrc/ = sm_100a kernels + C-ABI, engine.py = host mirror of the
reference interface, scene.py = synthetic inputs). The CUDA library is the product; importing
`engine` without it raises.
"""
from .engine import (APD, Problem, PatchMatchParams, default_params, ProcessProblem, ApdError,  # noqa: F401
                     FIRST_INIT, REFINE_INIT, REFINE_ITER, WEAK, STRONG, UNKNOWN)
