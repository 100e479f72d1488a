'''
Where does a sharded backup spend its time on one rank?  Emulates rank 0 of an 8-rank job on ONE GPU: Gamma build + selection on
B / 8 belief points, then assembly of ALL B new vectors (what parallel.backup_sharded(exchange='indices') does after its tiny
all-gather).  CUDA-event time per stage.
'''
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200._lib import lib, ptr, stream, check
torch.cuda.set_device(0); dev = torch.device('cuda', 0)
c = onb.synthetic.config('C2'); env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.QMDP_Agent(env, thresholds=c['thresholds']); m = agent.model
Sp, A, O = m.padded_states, m.action_count, m.observation_count
rng = np.random.default_rng(0)
B, G, W = 10_000, 512, int(os.environ.get('WORLD', 8))
bs = onb.BeliefSet(m, B)
for _ in range(6):
    bs = bs.update(rng.integers(0, 4, B), (rng.random(B) < 0.15).astype(int), raise_on_impossible_belief=False)
alphas = torch.as_tensor(np.abs(rng.standard_normal((G, Sp))).astype(np.float32), device=dev)
part = onb.BeliefSet(m, _buffers=(bs._b[:B // W], bs._scale[:B // W], B // W), device=0)
h = m.handle(0)
def ev(): return torch.cuda.Event(enable_timing=True)
for rep in range(3):
    e = [ev() for _ in range(6)]
    e[0].record()
    gam = torch.empty((A * O * G, Sp), dtype=torch.float32, device=dev)
    best_g = torch.empty((B // W, A, O), dtype=torch.int32, device=dev); best_a = torch.empty(B // W, dtype=torch.int32, device=dev); best_v = torch.empty(B // W, dtype=torch.float32, device=dev)
    e[1].record()
    check(lib.olfnav_backup_gamma(h, ptr(alphas), alphas.stride(0), G, 0.99, ptr(gam), Sp, stream()))
    e[2].record()
    check(lib.olfnav_backup_select(h, ptr(part._b), part.ld, B // W, ptr(part._scale), ptr(gam), Sp, G, ptr(best_g), ptr(best_a), ptr(best_v), 2, stream()))
    e[3].record()
    bg_all = best_g.repeat(W, 1, 1).contiguous(); ba_all = best_a.repeat(W).contiguous()      # stands for the all-gather
    e[4].record()
    new_alpha = onb.solvers.backup_assemble_device(m, gam, G, bg_all, ba_all, 0)
    e[5].record()
    torch.cuda.synchronize()
names = ['alloc', 'gamma build', 'select (split + GEMM + pick)', 'gather stand-in', 'assemble all B']
print(json.dumps({n: round(e[i].elapsed_time(e[i + 1]), 3) for i, n in enumerate(names)} | {'total_ms': round(e[0].elapsed_time(e[5]), 3), 'world': W}))
