import sys, os
sys.path.insert(0, os.getcwd())
import torch
import materialmodels_jl_b200 as mm
from materialmodels_jl_b200 import synth as sy
N = int(2e6)
m = mm.CrystalViscoPlastic(slipsystems=mm.slipsystems(mm.FCC(), sy.CRYSTAL_RODRIGUES), **sy.CRYSTAL_PARAMS)
dim, nc, amp = mm.PlaneStress(), 3, 4e-3
e = mm.synth_uniform(N, nc, 51, -amp, amp)
st0 = mm.initial_material_state(m, N)
for _ in range(2):
    _, _, st1 = mm.material_response(dim, m, e, st0, 1.0, options={"defer_check": True})
torch.cuda.synchronize()
print("iters mean", st1.iters.float().mean().item(), "outer", st1.outer_iters.float().mean().item())
