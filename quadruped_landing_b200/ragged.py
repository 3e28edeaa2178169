"""Ragged batches: problems of mixed horizon / contact schedule in one call.

The reference has a single problem size per `HybridNLP`; "varying horizon" sweeps therefore mix several
(N, k_trans, init_mode) classes.  Each class gets its own handle (its own segment plan and templates) and is
evaluated with one fused launch on its own CUDA stream, so classes overlap on the GPU.  Inputs and outputs
are flat arrays with offset tables, the layout C4 asks for.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .evaluator import HybridNLP
from .problem import ProblemData


class RaggedEvaluator:
    """Evaluates a batch whose b-th problem belongs to class ``class_of[b]`` (index into ``probs``)."""

    def __init__(self, probs: Sequence[ProblemData], device: int = 0):
        self.probs = list(probs)
        self.nlps: List[HybridNLP] = [HybridNLP.from_problem(p, device=device) for p in self.probs]
        self.device = device
        self.n = np.array([e.n_nlp for e in self.nlps], dtype=np.int64)
        self.m = np.array([e.m_nlp for e in self.nlps], dtype=np.int64)
        self.nnz = np.array([e.nnz_block for e in self.nlps], dtype=np.int64)
        self._streams = None

    def offsets(self, class_of: np.ndarray) -> Dict[str, np.ndarray]:
        """Exclusive prefix sums: problem b owns Z[z_off[b]:z_off[b+1]], g[g_off[b]:...], jac[j_off[b]:...]."""
        class_of = np.asarray(class_of, dtype=np.int64)
        cs = lambda w: np.concatenate([[0], np.cumsum(w[class_of])])
        return {"z_off": cs(self.n), "g_off": cs(self.m), "j_off": cs(self.nnz)}

    def eval(self, class_of: np.ndarray, Z_flat, want=("f", "grad", "g", "jac")):
        """``Z_flat``: 1-D float64 CUDA tensor holding the decision vectors back to back (see ``offsets``).
        Returns flat CUDA tensors ``f[B]``, ``grad`` (Z layout), ``g``, ``jac`` and the offset tables."""
        import torch

        class_of = np.asarray(class_of, dtype=np.int64)
        B = class_of.shape[0]
        off = self.offsets(class_of)
        dev = Z_flat.device
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in self.nlps]
        out = {}
        if "f" in want:
            out["f"] = torch.empty(B, dtype=torch.float64, device=dev)
        if "grad" in want:
            out["grad"] = torch.empty(int(off["z_off"][-1]), dtype=torch.float64, device=dev)
        if "g" in want:
            out["g"] = torch.empty(int(off["g_off"][-1]), dtype=torch.float64, device=dev)
        if "jac" in want:
            out["jac"] = torch.empty(int(off["j_off"][-1]), dtype=torch.float64, device=dev)
        cur = torch.cuda.current_stream(dev)
        keep = []
        for c, nlp in enumerate(self.nlps):
            idx = np.nonzero(class_of == c)[0]
            if idx.size == 0:
                continue
            s = self._streams[c]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                def rows(offs, width):
                    base = torch.from_numpy(offs[idx]).to(dev)
                    return base[:, None] + torch.arange(width, device=dev)[None, :]
                zi = rows(off["z_off"], nlp.n_nlp)
                res = nlp.eval_batch(Z_flat[zi], want=want)
                if "f" in want:
                    out["f"][torch.from_numpy(idx).to(dev)] = res["f"]
                if "grad" in want:
                    out["grad"][zi] = res["grad"]
                if "g" in want:
                    out["g"][rows(off["g_off"], nlp.m_nlp)] = res["g"]
                if "jac" in want:
                    out["jac"][rows(off["j_off"], nlp.nnz_block)] = res["jac"]
                keep.append(res)
            cur.wait_stream(s)
        out.update(off)
        return out
