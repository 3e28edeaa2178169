"""B200-native batched evaluator for the planar-quadruped landing NLP.

The package holds only what the hot path needs: the CUDA kernels + C ABI (``csrc/``,
``include/qlnlp.h``), the in-tree build, the host-side problem tables, and ``HybridNLP``, the
host-side mirror of the reference's ``MOI.AbstractNLPEvaluator`` surface (src/nlp.jl, src/moi.jl).
"""
from .problem import (NX, NU, PlanarQuadruped, QuadraticCost, LQRCost, ProblemData, reference_trajectory,
                      packZ, unpackZ, default_states, build_problem, default_problem, initial_guess,
                      sweep_initial_states, initial_guess_batch, save_solution_csv, load_solution_csv, solution_table)
from .evaluator import HybridNLP, QlnlpError, load_library, even_ld, EXPORTED_SYMBOLS, host_alloc, host_free
from .ragged import RaggedEvaluator
from .sharding import shard_bounds, evaluate_sharded

__all__ = [
    "NX", "NU", "PlanarQuadruped", "QuadraticCost", "LQRCost", "ProblemData", "reference_trajectory",
    "packZ", "unpackZ", "default_states", "build_problem", "default_problem", "initial_guess",
    "HybridNLP", "QlnlpError", "load_library", "even_ld", "EXPORTED_SYMBOLS", "RaggedEvaluator",
    "shard_bounds", "evaluate_sharded", "sweep_initial_states", "initial_guess_batch", "save_solution_csv",
    "load_solution_csv", "solution_table", "host_alloc", "host_free",
]
