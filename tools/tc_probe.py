import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import build_pair, rel_err
from lagrangebench_b200 import GNS
from oracle import gns as ogns
for name, mp in [("tgv2d", 1), ("tgv2d", 10), ("ldc3d", 10)]:
    c, ours, orac = build_pair(name, "float32")
    sample = (c["positions"], c["particle_type"])
    f_gpu, _ = ours.allocate_eval(sample)
    f_cpu, _ = orac.allocate_eval(sample)
    d = c["metadata"]["dim"]
    node_in = sum(f_cpu[k].reshape(f_cpu[k].shape[0], -1).shape[1] for k in ("vel_hist", "bound", "force") if k in f_cpu)
    params = ogns.init_params(node_in, d + 1, d, num_mp_steps=mp, seed=11)
    model = GNS(d, 128, 2, mp, 16)
    ref64 = ogns.forward(params, f_cpu, c["particle_type"], mp, np.float64)["acc"]
    for impl in ("simt", "tc"):
        model.edge_impl = impl
        out, _ = model.apply(params, {}, (f_gpu, c["particle_type"]))
        torch.cuda.synchronize()
        got = out["acc"].cpu().numpy()
        print(name, mp, impl, "rel_err vs f64 oracle: %.3e" % rel_err(got, ref64), "finite", np.isfinite(got).all(), flush=True)
