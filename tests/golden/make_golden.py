"""Regenerates tests/golden/ref_solutions.npz from the reference checkout.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py

Contents
--------
data_1 .. data_6 : the six solved decision vectors the reference ships as
    src/data_{1..6}.csv (written by `writedlm`, src/main.ipynb:881); N=61,
    k_trans=21, init_mode=1.  data_6 is the notebook's recorded `Z_sol`.
The recorded scalar outputs that pin eval_f / eval_c on data_6 live in
tests/test_oracle_kat.py next to their main.ipynb line numbers.
"""
import os

import numpy as np

REF = "/root/reference/src"
HERE = os.path.dirname(os.path.abspath(__file__))

if __name__ == "__main__":
    out = {}
    for i in range(1, 7):
        z = np.loadtxt(os.path.join(REF, f"data_{i}.csv"), delimiter=",")
        assert z.shape == (1215,)
        out[f"data_{i}"] = z
    np.savez_compressed(os.path.join(HERE, "ref_solutions.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_solutions.npz"))
