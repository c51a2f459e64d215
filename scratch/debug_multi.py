import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch, torch.distributed as dist
import pysph_b200 as pb
from pysph_b200 import geometry as geo
from pysph_b200.parallel import make_slab_solver
from test_gpu_multi import _perturb, DX
NS = int(os.environ.get('NS', '25'))
rank = int(os.environ['RANK']); world = int(os.environ['WORLD_SIZE'])
torch.cuda.set_device(rank)
dist.init_process_group('nccl', device_id=torch.device('cuda', rank))
params = geo.dam_break_3d_params(DX)
# reference on every rank (same device)
pas = geo.dam_break_3d_particles(dx=DX); _perturb(pas)
ref = pb.make_wcsph_solver(pas, dict(params), pb.CubicSpline(dim=3), device=rank, adaptive_timestep=False, n_damp=0)
solver, pm, mpas = make_slab_solver(DX, params, pb.CubicSpline(dim=3), rank, world, device=rank, adaptive_timestep=False, n_damp=0)
_perturb(mpas); solver.backend.push_all()
solver.backend.use_torch_stream()
f = pas[0]
for step in range(NS):
    ref.step(); solver.step()
    ref.pull(['x','u','rho','gid']); solver.pull(['x','y','z','u','rho','gid','tag'])
    m = mpas[0]; nr = m.get_number_of_particles(real=True)
    order = np.argsort(f.gid); pos = np.searchsorted(f.gid[order], m.gid[:nr]); idx = order[pos]
    ex = np.abs(m.x[:nr] - f.x[idx]); er = np.abs(m.rho[:nr] - f.rho[idx])
    worst = np.argmax(ex)
    nb = mpas[1]; nrb = nb.get_number_of_particles(real=True)
    print('rank %d step %d nreal %d nghost %d full %d refresh %d max|dx| %.3e (x=%.4f cut lo=%.4f hi=%.4f) max|drho| %.3e nan %d' % (
        rank, step, nr, m.get_number_of_particles()-nr, pm.n_full, pm.n_refresh, ex.max(), m.x[:nr][worst], pm.lo, pm.hi, er.max(), int(np.isnan(m.x).sum())), flush=True)
dist.destroy_process_group()
