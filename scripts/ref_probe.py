"""Run the rebuilt reference (oracle/_ref) on one synthetic input and compare with the oracle."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mh_spgemm_b200  # noqa
from mh_spgemm_b200 import generators as G
from oracle import Oracle, Reference

name = sys.argv[1] if len(sys.argv) > 1 else "poisson32"
A = {"poisson32": lambda: G.poisson2d(32), "poisson256": lambda: G.poisson2d(256),
     "fem-small": lambda: G.fem3d(4, 4, 10, 3, seed=5), "F": lambda: G.fem3d(),
     "rmat14": lambda: G.rmat(14, 16000, 60000, seed=6), "R": lambda: G.rmat(),
     "uniform": lambda: G.uniform_random(3000, 3000, 30000, seed=8)}[name]()
o = Oracle()
Cp, Cc, Cv = o.spgemm(A, A)
print(name, "M", A.M, "nnz", A.nnz, "oracle nnzC", Cp[-1], flush=True)
R = Reference().spgemm(A, A, reps=int(os.environ.get("REPS", "1")), warmup=0, e2e_reps=0)
print("ref nnzC", R["nnz"], "ptr ok", np.array_equal(R["ptr"].astype(np.int64), Cp),
      "col ok", np.array_equal(R["col"], Cc), "ms", R["ms_device"], "stages", R["stage_ms"])
