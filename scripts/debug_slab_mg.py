"""single-GPU emulation of a 2-slab multigrid (manual halo copies) compared stage by stage with the
single-domain hierarchy -- development aid"""
import sys
sys.path.insert(0, ".")
import numpy as np, torch
from pyro2_b200.mg_handle import MGHandle

class FakeDecomp:
    def __init__(self, rank, size): self.rank, self.size, self.group = rank, size, None

n, split, R = 512, 256, 2
bc = ("dirichlet",)*4
ref = MGHandle(n, bc, 0.0, -1.0, 0., 1., 0., 1., 10, 50)
hs = [MGHandle(n, bc, 0.0, -1.0, 0., 1., 0., 1., 10, 50, decomposition=FakeDecomp(r, R), split_n=split) for r in range(R)]
L = ref.nlevels - 1
print("levels", ref.nlevels, [hs[0].info(l) for l in (L, L-1, L-2)])
x = (torch.arange(n+2, device="cuda", dtype=torch.float64) - 0.5)/n
X, Y = torch.meshgrid(x, x, indexing="ij")
F = -2.0*((1-6*X**2)*Y**2*(1-Y**2) + (1-6*Y**2)*X**2*(1-X**2))
ref.plane(L, "f").copy_(F)
for r, h in enumerate(hs):
    ni = h.info(L)["ni"]
    h.plane(L, "f")[1:ni+1].copy_(F[1 + r*ni: 1 + (r+1)*ni])

def exchange(level, which, depth):
    g = hs[0].info(level)
    if not g["slab"]: return
    ni = g["ni"]
    views = [h.halo_rows(level, which, depth) for h in hs]
    for r in range(R - 1):
        views[r+1][0:depth].copy_(views[r][ni:ni+depth])            # my top owned rows -> upper neighbour's low halo
        views[r][depth+ni:depth+ni+depth].copy_(views[r+1][depth:2*depth])

def check(level, which, tag):
    g = hs[0].info(level)
    a = ref.plane(level, which).cpu().numpy()
    worst = 0
    for r, h in enumerate(hs):
        b = h.plane(level, which).cpu().numpy()
        if g["slab"]:
            ni = g["ni"]
            d = np.abs(b[1:ni+1, 1:-1] - a[1 + r*ni: 1 + (r+1)*ni, 1:-1])
        else:
            d = np.abs(b[1:-1, 1:-1] - a[1:-1, 1:-1])
        worst = max(worst, d.max())
        if d.max() > 0:
            idx = np.argwhere(d > 0)
            print(f"   rank {r}: {len(idx)} cells differ, rows {idx[:,0].min()}..{idx[:,0].max()} cols {idx[:,1].min()}..{idx[:,1].max()}")
    print(f"{tag:40s} level {level} {which}: max diff {worst:.3e}")

def smooth_slabs(level, ns):
    src, dst = "v", "w"
    left = ns
    while left > 0:
        it = min(left, hs[0].tb_iters)
        exchange(level, src, hs[0].tb_halo)
        for h in hs: h.tb_pass(level, src, dst, it)
        left -= it; src, dst = dst, src
    assert src == "v"

exchange(L, "f", hs[0].tb_halo)
split_level = hs[0].info(L)["split_level"]
def vcycle(level):
    if level < split_level:
        for h in hs: h.vcycle_level(level)
        ref_v(level); check(level, "v", "after replicated sub-vcycle"); return
    smooth_slabs(level, 10); ref.smooth(level, 10); check(level, "v", "after pre-smooth")
    exchange(level, "v", 1)
    for h in hs: h.residual(level)
    ref.residual(level); check(level, "r", "after residual")
    for h in hs: h.restrict(level)
    ref.restrict(level)
    if level - 1 < split_level:
        g = hs[0].info(level-1); chunk = g["ni"] // R
        fs = [h.halo_rows(level-1, "f", 0) for h in hs]
        for r in range(R):
            for q in range(R):
                if q != r: fs[q][r*chunk:(r+1)*chunk].copy_(fs[r][r*chunk:(r+1)*chunk])
    else:
        exchange(level-1, "f", hs[0].tb_halo)
    check(level-1, "f", "after restrict")
    vcycle(level-1)
    if level - 1 >= split_level: exchange(level-1, "v", 1)
    for h in hs: h.prolong_correct(level)
    ref.prolong_correct(level); check(level, "v", "after prolong")
    smooth_slabs(level, 10); ref.smooth(level, 10); check(level, "v", "after post-smooth")

def ref_v(level):
    # the reference hierarchy runs the same sub-V-cycle with the library call
    import ctypes
    from pyro2_b200 import _lib
    _lib.check(_lib.lib().p2b_mg_vcycle_level(ref._h, level, ref._s()))

for cyc in range(1):
    for h in hs + [ref]: h.zero_coarse()
    vcycle(L)
