"""Ragged batches: problems of mixed horizon / contact schedule in one call.

The reference has a single problem size per `HybridNLP`; "varying horizon" sweeps therefore mix several
(N, k_trans, init_mode) classes.  Each class gets its own handle (its own segment plan and cost tables); the
whole mixed batch is evaluated by ONE fused launch (`qlnlp_eval_ragged_classes`): problems are handed out
ordered by class, largest horizon first, and a warp swaps class records when the class changes.  Inputs and
outputs are flat arrays with offset tables, the layout C4 asks for.
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
        self._plan = None
        self.launches_per_eval = 1                     # one fused launch for the whole mixed batch
        self.single_launch = True                      # False: one launch per class on its own stream (the older path)

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

    def plan(self, class_of: np.ndarray, device=None) -> "RaggedPlan":
        """Everything about a batch that does not depend on the decision vectors: the offset tables and each class's
        problem list, uploaded once.  ``eval`` caches the plan of the last ``class_of`` it saw, so repeated
        evaluations of one batch (solver iterations) do no per-call host work."""
        import torch

        class_of = np.ascontiguousarray(class_of, dtype=np.int64)
        dev = torch.device("cuda", self.device) if device is None else device
        off = self.offsets(class_of)
        doff = {k: torch.from_numpy(v).to(dev) for k, v in off.items()}
        index = []
        for c in range(len(self.nlps)):
            idx = np.nonzero(class_of == c)[0]
            index.append(torch.from_numpy(idx).to(dev) if idx.size else None)
        # single launch: all problems, ordered by class, largest horizon first (the long evaluations start first)
        by_class = sorted(range(len(self.nlps)), key=lambda c: -self.nlps[c].N)
        rank_of = np.empty(len(self.nlps), dtype=np.int64)
        rank_of[by_class] = np.arange(len(self.nlps))
        order = np.argsort(rank_of[class_of], kind="stable")
        plan = RaggedPlan(class_of.copy(), off, doff, index)
        plan.order = torch.from_numpy(order.astype(np.int64)).to(dev)
        plan.dclass = torch.from_numpy(class_of.astype(np.int32)).to(dev)
        return plan

    def eval(self, class_of, Z_flat, want=("f", "grad", "g", "jac"), out=None):
        """``Z_flat``: 1-D float64 CUDA tensor holding the decision vectors back to back (see ``offsets``);
        ``class_of``: the class index of every problem, or a ``RaggedPlan``.  Returns flat CUDA tensors ``f[B]``,
        ``grad`` (Z layout), ``g``, ``jac`` and the offset tables; pass the returned dict back as ``out`` to reuse
        the output arrays.

        The whole mixed batch is ONE ``qlnlp_eval_ragged_classes`` launch (``single_launch=False``: one
        ``qlnlp_eval_ragged_device`` launch per class on its own stream); the kernel addresses each problem's rows
        through the offset tables, so nothing is gathered or scattered."""
        import torch

        dev = Z_flat.device
        if isinstance(class_of, RaggedPlan):
            plan = class_of
        else:
            class_of = np.ascontiguousarray(class_of, dtype=np.int64)
            plan = self._plan
            if plan is None or plan.class_of.shape != class_of.shape or not np.array_equal(plan.class_of, class_of):
                plan = self._plan = self.plan(class_of, dev)
        B = plan.class_of.shape[0]
        off = plan.off
        if self._streams is None:
            self._streams = [torch.cuda.Stream(device=dev) for _ in self.nlps]
        flat = {"Z": Z_flat}
        sizes = {"f": B, "grad": int(off["z_off"][-1]), "g": int(off["g_off"][-1]), "jac": int(off["j_off"][-1])}
        for name in ("f", "grad", "g", "jac"):
            if name in want:
                t = None if out is None else out.get(name)
                if t is None or t.numel() != sizes[name] or t.device != dev:
                    t = torch.empty(sizes[name], dtype=torch.float64, device=dev)
                flat[name] = t
        cur = torch.cuda.current_stream(dev)
        if self.single_launch:
            HybridNLP.eval_ragged_classes(self.nlps, plan.order, plan.dclass, flat, plan.doff, z_padded=True)
        else:
            for c, nlp in enumerate(self.nlps):
                if plan.index[c] is None:
                    continue
                s = self._streams[c]
                s.wait_stream(cur)
                nlp.eval_ragged(plan.index[c], flat, plan.doff, stream=s, z_padded=True)
                cur.wait_stream(s)
        res = {k: v for k, v in flat.items() if k != "Z"}
        res.update(off)
        self._keep = plan                  # alive until the next call: the launches are asynchronous
        return res


class RaggedPlan:
    """Offset tables (host + device) and per-class problem lists of one mixed batch (``RaggedEvaluator.plan``)."""

    def __init__(self, class_of, off, doff, index):
        self.class_of, self.off, self.doff, self.index = class_of, off, doff, index
