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
        """Row starts (in doubles): problem b owns Z[z_off[b] : z_off[b] + n_b], g[g_off[b] : ...], jac[j_off[b] : ...].
        Every row is padded to an even length, so all rows are 16-byte aligned (TMA load / store paths); the last
        entry of each table is the total length to allocate."""
        class_of = np.asarray(class_of, dtype=np.int64)
        even = lambda w: (w + 1) & ~1
        cs = lambda w: np.concatenate([[0], np.cumsum(even(w)[class_of])])
        return {"z_off": cs(self.n), "g_off": cs(self.m), "j_off": cs(self.nnz)}

    def pack(self, class_of: np.ndarray, vectors) -> np.ndarray:
        """Lay decision vectors (one per problem) out in the padded flat Z layout of ``offsets``."""
        off = self.offsets(class_of)["z_off"]
        Z = np.zeros(int(off[-1]))
        for b, v in enumerate(vectors):
            Z[off[b]:off[b] + len(v)] = v
        return Z

    def eval(self, class_of: np.ndarray, Z_flat, want=("f", "grad", "g", "jac")):
        """``Z_flat``: 1-D float64 CUDA tensor holding the decision vectors back to back (see ``offsets``).
        Returns flat CUDA tensors ``f[B]``, ``grad`` (Z layout), ``g``, ``jac`` and the offset tables.

        Every class is one ``qlnlp_eval_ragged_device`` launch on its own stream: the kernel addresses each
        problem's rows through the offset tables, so nothing is gathered or scattered."""
        import torch

        class_of = np.asarray(class_of, dtype=np.int64)
        B = class_of.shape[0]
        off = self.offsets(class_of)
        dev = Z_flat.device
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in self.nlps]
        flat = {"Z": Z_flat}
        if "f" in want:
            flat["f"] = torch.empty(B, dtype=torch.float64, device=dev)
        if "grad" in want:
            flat["grad"] = torch.empty(int(off["z_off"][-1]), dtype=torch.float64, device=dev)
        if "g" in want:
            flat["g"] = torch.empty(int(off["g_off"][-1]), dtype=torch.float64, device=dev)
        if "jac" in want:
            flat["jac"] = torch.empty(int(off["j_off"][-1]), dtype=torch.float64, device=dev)
        doff = {k: torch.from_numpy(v).to(dev) for k, v in off.items()}
        cur = torch.cuda.current_stream(dev)
        keep = []
        for c, nlp in enumerate(self.nlps):
            idx = np.nonzero(class_of == c)[0]
            if idx.size == 0:
                continue
            s = self._streams[c]
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                index = torch.from_numpy(idx).to(dev)
                nlp.eval_ragged(index, flat, doff, stream=s, z_padded=True)
                keep.append(index)
            cur.wait_stream(s)
        out = {k: v for k, v in flat.items() if k != "Z"}
        out.update(off)
        self._keep = (keep, doff)          # alive until the next call: the launches are asynchronous
        return out
