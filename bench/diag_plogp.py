import numpy as np, torch, sys
sys.path.insert(0, '.')
import diffuncertainty_b200 as vu
p = np.concatenate([np.logspace(-37, 0, 400000), 1 - np.logspace(-8, -0.31, 400000), np.linspace(0.4, 1.0, 400000), 1 + np.logspace(-7, 0, 40000)]).astype(np.float32)
x = torch.zeros(2, 1, 2, p.size); x[:, 0, 0] = torch.from_numpy(p)
tu = vu.fused_pass(x.cuda()).maps["TU"][0].cpu().numpy().astype(np.float64)
p64 = p.astype(np.float64); want = -p64*np.log(p64)
rel = np.abs(tu-want)/np.maximum(np.abs(want),1e-300)
for lo, hi in [(0,1e-30),(1e-30,1e-10),(1e-10,0.1),(0.1,0.5),(0.5,0.9),(0.9,0.96),(0.96,1-1/64),(1-1/64,1),(1,1+1/64),(1+1/64,1.1),(1.1,2.01)]:
    m = (p64>=lo)&(p64<hi)&(want!=0)
    if m.any():
        i = np.argmax(np.where(m, rel, 0)); print(f"[{lo:g},{hi:g}) n={m.sum()} max rel {rel[i]:.3e} at p={p64[i]!r}")
