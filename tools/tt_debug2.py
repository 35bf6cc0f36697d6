"""Debug: compare the internal training arrays of the tensor-core path with the fp32 CUDA-core path (GPU box)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from helpers import Golden, build_module
from particle_fm_b200.training import fm_loss_autograd

name = sys.argv[1] if len(sys.argv) > 1 else "c1_jetnet30"
g = Golden(name)
kind = "FM-OT"
gen = torch.Generator().manual_seed(31)
x = g.x * 5.0 * g.mask
B, N = x.shape[0], x.shape[1]
t = torch.rand(B, generator=gen); n0 = torch.randn(x.shape, generator=gen)
out = {}
for mode in ("cuda_cores", "auto"):
    m = build_module(g.ctor, g.sd, loss_type=kind, device="cuda:0")
    eng = m.flows[0].net.engine()
    eng.set_train_mode(mode)
    loss = fm_loss_autograd(m.flows[0], kind, x.cuda(), g.mask.cuda(), None, t.cuda(), n0.cuda(), None, 1e-4)
    loss.backward()
    lib = eng.lib
    lib.pfm_debug_copy.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_longlong]
    H, L = g.cfg.hid, g.cfg.layers
    rows = int(g.mask.sum())
    SS = B * N * H
    bstride = sum(((o + 3) // 4) * 4 for o, i in eng.linear_shapes())
    arrs = {}
    for which, nm, n in ((0, "act", (2 + 2 * L) * SS), (1, "dact", (2 + 2 * L) * SS), (2, "dbeff", B * bstride)):
        buf = np.zeros(n, dtype=np.float32)
        assert lib.pfm_debug_copy(eng._h, which, buf.ctypes.data_as(C.c_void_p), n) == 0
        arrs[nm] = buf
    out[mode] = (arrs, rows, SS, bstride, float(loss))
a, rows, SS, bstride, l0 = out["cuda_cores"]; b, _, _, _, l1 = out["auto"]
print("loss", l0, l1, "rows", rows)
H = g.cfg.hid
for nm in ("act", "dact"):
    for s in range(2 + 2 * g.cfg.layers):
        x0 = a[nm][s * SS:s * SS + rows * H]; x1 = b[nm][s * SS:s * SS + rows * H]
        print(f"{nm}[{s:2d}] rel {np.linalg.norm(x1 - x0) / max(np.linalg.norm(x0), 1e-30):.3e}  |ref| {np.linalg.norm(x0):.3e}")
offs = np.cumsum([0] + [((o + 3) // 4) * 4 for o, i in eng.linear_shapes()])
d0 = a["dbeff"].reshape(B, bstride); d1 = b["dbeff"].reshape(B, bstride)
for i in range(len(offs) - 1):
    s0 = d0[:, offs[i]:offs[i + 1]]; s1 = d1[:, offs[i]:offs[i + 1]]
    print(f"dbeff lin {i:2d} rel {np.linalg.norm(s1 - s0) / max(np.linalg.norm(s0), 1e-30):.3e}")

# pattern of the worst stage
worst = max(range(2 + 2 * g.cfg.layers), key=lambda s_: np.linalg.norm(b["dact"][s_ * SS:s_ * SS + rows * H] - a["dact"][s_ * SS:s_ * SS + rows * H]) / max(np.linalg.norm(a["dact"][s_ * SS:s_ * SS + rows * H]), 1e-30))
for st_ in sorted({worst, min(worst + 1, 2 * g.cfg.layers + 1)}):
    x0 = a["dact"][st_ * SS:st_ * SS + rows * H].reshape(rows, H); x1 = b["dact"][st_ * SS:st_ * SS + rows * H].reshape(rows, H)
    d = np.abs(x1 - x0)
    tol = 1e-4 * np.abs(x0).max()
    bad_rows = np.where((d > tol).any(axis=1))[0]
    bad_cols = np.where((d > tol).any(axis=0))[0]
    print(f"stage {st_}: {len(bad_rows)} bad rows of {rows}: {bad_rows[:40]}  {len(bad_cols)} bad cols: {bad_cols[:40]}")
    if len(bad_rows):
        r = bad_rows[0]
        print("  row", r, "ref", x0[r, :8], "got", x1[r, :8], "ratio", (x1[r, :8] / x0[r, :8]))
n_real = g.mask.sum(dim=(1, 2)).int().tolist()
print("n_real", n_real)
