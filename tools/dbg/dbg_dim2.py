import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [os.path.join(ROOT, "delta-prox_b200"), os.path.join(ROOT, "oracle")]
import dprox_b200 as dp
import dprox_oracle as orc
g = dict(np.load(os.path.join(ROOT, "tests/golden/admm_grad_dim2.npz")))
def rel(a, b):
    a = a.detach().cpu().double().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float64)
    b = b.detach().cpu().double().numpy() if isinstance(b, torch.Tensor) else np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
b = torch.from_numpy(g["b"])
psi = [orc.Term("norm1", orc.Grad(d, orc.Identity())) for d in (2, 1, 0)]
data = orc.Term("sum_squares", orc.Conv(g["psf"], orc.Identity()), c=b)
for T_ in (1, 2, 6):
    want = orc.Solver([data] + psi, "admm").solve(b.clone(), rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=T_, return_full_states=True)
    x = dp.Variable()
    bd = b.cuda()
    fns = dp.sum_squares(dp.conv(x, g["psf"]) - bd) + dp.norm1(dp.grad(x, dim=2)) + dp.norm1(dp.grad(x, dim=1)) + dp.norm1(dp.grad(x, dim=0))
    s = dp.compile(fns, method="admm", device="cuda")
    st = s.solve(x0=bd, rhos=float(g["rho"]), lams=float(g["lam"]), max_iter=T_, return_full_states=True)
    print(T_, s.spec.tier, s.spec.xupdate, "x", rel(st[0], want[0]), "v", [rel(a, c) for a, c in zip(st[1], want[1])], "u", [rel(a, c) for a, c in zip(st[2], want[2])])
# pieces
x = dp.Variable()
op = dp.grad(x, dim=2)
t = torch.rand(2, 3, 32, 48)
o = orc.Grad(2, orc.Identity())
print("fwd", rel(op.forward(t.cuda()), o.fwd(t)), "adj", rel(op.adjoint(t.cuda()), o.adj(t)))
low = op.lower()
gh = low.gram_fn((2, 3, 32, 48))
print("gram", gh.shape, gh[0, :, 0, 0], "oracle diag", o.diag(t, True)[0, :, 0, 0])
