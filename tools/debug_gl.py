import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from tests.gpu_util import R50_360, make_inputs, oracle_all, to_cuda
from mdqe_cvpr2023_b200 import ops, _lib
inp = make_inputs(1, R50_360, 8, 32, 4, dist="local", seed=1)
want = oracle_all(inp)
d = to_cuda(inp)
a = (d["value"], d["shapes"], d["level_start"], d["loc"], d["aw"])
for variant in (0, 2, 1):
    _lib.set_option("bwd_variant", variant); _lib.set_option("fwd_variant", variant)
    out = ops.ms_deform_attn_forward(*a, 64)
    gv, gl, ga = ops.ms_deform_attn_backward(*a, d["grad_out"], 64)
    for name, g, w in (("out", out, want[0]), ("gv", gv, want[1]), ("gl", gl, want[2]), ("ga", ga, want[3])):
        g = g.cpu().numpy().reshape(w.shape)
        err = np.abs(g - w); e = err.max() / np.abs(w).max()
        print("variant", variant, name, "nerr %.3e" % e, "n_bad", int((err > 1e-4 * np.abs(w).max()).sum()))
        if name == "gl" and e > 1e-4:
            idx = np.unravel_index(np.argsort(err.ravel())[-5:], err.shape)
            for k in range(5):
                i = tuple(ix[k] for ix in idx)
                n, q, m, l, p, c = i
                H, W = R50_360[l]
                lx, ly = inp["loc"][n, q, m, l, p].tolist()
                print("   idx", i, "got", g[i], "want", w[i], "loc", (lx, ly), "pix", (lx * W - 0.5, ly * H - 0.5), "aw", float(inp["aw"][n, q, m, l, p]))
