"""Diagnostic run of the device static-condensation path on the GPU box: prints the
measured errors against the oracle / golden vectors for every SC golden case
(the pass/fail version is tests/test_gpu_condensed.py).

    python tests/condensed_check.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

import sem_oracle as so  # noqa: E402
from conftest import build_package_case, golden_case_names, load_case, rel_l2  # noqa: E402


def main():
    for name in [n for n in golden_case_names() if "_sc" in n]:
        g = load_case(name)
        ref = so.condensed_system(int(g["p"]), g["invJ"], g["JxW"], g["l2g"])
        for tier in ("T1", "T2"):
            mesh, mngr = build_package_case(g["kind"], g["nx"], g["ny"], g["p"], g["sc"], g["rcm"])
            kw = {"geometric_factors": (g["invJ"], g["JxW"])} if tier == "T1" else {}
            sc = mngr.condensed_poisson_operator(dirichlet=g["on_ebc"], **kw)
            u = np.random.default_rng(0).standard_normal(sc.n_ext)
            ud = torch.from_numpy(u).cuda()
            y = sc.apply(ud, flags=0).cpu().numpy()
            sol, info = sc.solve(1.0, g["ebc_vals"], rtol=1e-13)
            print("%-12s %s  S %.2e  apply %.2e  diag %.2e  rhs %.2e  solution %.2e  its %d"
                  % (name, tier, rel_l2(sc.local_schur(), ref["S"]), rel_l2(y, ref["Sg"] @ u),
                     rel_l2(sc.diagonal(masked=False).cpu().numpy(), ref["Sg"].diagonal()),
                     rel_l2(sc.rhs(1.0).cpu().numpy(), ref["grhs"]),
                     rel_l2(sol.cpu().numpy(), g["solution"]), info.iterations), flush=True)


if __name__ == "__main__":
    main()
