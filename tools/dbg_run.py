import sys; sys.path.insert(0,'.')
import numpy as np, percept_b200 as pb
from tests import meshes
from tests.orc import oracle
kind, nele, eps, niter = "sgrid", 8, 0.3, 12
X,W = meshes.perturbed_coords(nele, eps)
for strict in (0,1):
    g = meshes.build(pb.api, kind, nele, X, W, True, {"strict_fp": strict})
    o = meshes.build(oracle, kind, nele, X, W, True, {"real_is_double": 1})
    g.run_algorithm(niter, 0.0, "/tmp/g.txt"); o.run_algorithm(niter, 0.0, "/tmp/o.txt")
    print("strict", strict, "maxdiff coords", np.max(np.abs(g.download("coordinates")-o.download("coordinates"))))
    for a,b in zip(open("/tmp/g.txt").read().splitlines(), open("/tmp/o.txt").read().splitlines()):
        if a!=b: print("G:",a); print("O:",b)
