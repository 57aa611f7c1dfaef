import sys, os, time
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np
from multiprocessing import Pool
from hostsim import HostSim
from curiousrl_b200.utils.synthetic import synthetic_batch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x0, tgt = synthetic_batch("ThreeLinkPlanarManipulator", 65536, seed=1234)
def work(rng):
    hs = HostSim("ThreeLinkPlanarManipulator")
    out = []
    for b in range(*rng):
        r = hs.solve(x0[b], tgt[b], 100)
        Js = r["Js"]
        out.append((r["iters"], Js[:64].copy() if len(Js) >= 64 else np.pad(Js, (0, 64 - len(Js)), constant_values=np.nan), r["alphas"][:64].copy() if len(Js)>=64 else np.pad(r["alphas"], (0, 64-len(Js)), constant_values=-2)))
    return out
if __name__ == "__main__":
    t = time.time()
    chunks = [(i, min(i + 32, B)) for i in range(0, B, 32)]
    with Pool(8) as p:
        res = p.map(work, chunks)
    flat = [r for c in res for r in c]
    iters = np.array([r[0] for r in flat]); J = np.stack([r[1] for r in flat]); A = np.stack([r[2] for r in flat])
    np.savez("exp/pred_%d.npz" % B, iters=iters, J=J, A=A)
    print("time", time.time() - t)
