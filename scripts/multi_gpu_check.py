'''
Multi-GPU check (run under torchrun, one rank per GPU): the sharded paths give the single-GPU results.
  * backup_sharded: belief points split across ranks + NCCL all-gather == the single-GPU backup of all points
  * run_test_sharded: agents split across ranks, histories gathered == the single-GPU run of all agents
Usage: python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/multi_gpu_check.py
'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import olfactory_navigation_b200 as onb
from olfactory_navigation_b200 import parallel

local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
dist.init_process_group('nccl', device_id=dev)
rank, ws = parallel.world()

c = onb.synthetic.config('C1')
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
agent.train(expansions=6, print_progress=False, print_stats=False)
m = agent.model
rng = np.random.default_rng(5)

# ---- backup ----
B = 301
bs = onb.BeliefSet(m, B)
for _ in range(4):
    bs = bs.update(rng.integers(0, m.action_count, B), np.zeros(B, dtype=int))
alphas = agent.value_function.device_alphas()
full_alpha, full_a, _, _ = onb.solvers.backup_device(m, bs, alphas, 0.99)
for exchange in ('indices', 'alphas'):
    sh_alpha, sh_a = parallel.backup_sharded(m, bs, alphas, 0.99, exchange=exchange)
    assert sh_alpha.shape == full_alpha.shape, (sh_alpha.shape, full_alpha.shape)
    assert torch.equal(sh_a, full_a), f'sharded backup ({exchange}): actions differ'
    assert torch.equal(sh_alpha, full_alpha), f'sharded backup ({exchange}): alpha vectors differ'

# ---- training through the Agent API: train() under the process group shards every backup; must equal the unsharded training ----
os.environ['OLFNAV_SHARDED_TRAIN'] = '0'
ref_agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
ref_agent.train(expansions=8, print_progress=False, print_stats=False)
os.environ['OLFNAV_SHARDED_TRAIN'] = '1'
sh_agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=onb.synthetic.SEED)
sh_agent.train(expansions=8, print_progress=False, print_stats=False)
assert np.array_equal(sh_agent.value_function.actions, ref_agent.value_function.actions), 'sharded train(): actions differ'
assert torch.equal(sh_agent.value_function.device_alphas(), ref_agent.value_function.device_alphas()), 'sharded train(): alpha vectors differ'
for cls in (onb.agents.PBVI_SSRA_Agent, onb.agents.PBVI_GER_Agent):
    os.environ['OLFNAV_SHARDED_TRAIN'] = '0'
    r = cls(env, thresholds=c['thresholds'], rng=3)
    r.train(expansions=4, max_belief_growth=6, print_progress=False, print_stats=False)
    os.environ['OLFNAV_SHARDED_TRAIN'] = '1'
    t = cls(env, thresholds=c['thresholds'], rng=3)
    t.train(expansions=4, max_belief_growth=6, print_progress=False, print_stats=False)
    assert torch.equal(t.value_function.device_alphas(), r.value_function.device_alphas()), f'sharded train() of {cls.__name__}: alpha vectors differ'
# different seeds on the ranks are refused
try:
    bad = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'], rng=100 + rank)
    bad.train(expansions=1, print_progress=False, print_stats=False)
    raise SystemExit('sharded train() accepted rank-dependent seeds')
except RuntimeError:
    pass

# ---- simulation ----
n = 203
starts = env.random_start_points(n, np.random.default_rng(11))
tshift = np.random.default_rng(12).integers(0, env.timesteps, n)
env.reference_time_quirk = False          # every agent observes its own time: shard-invariant (the reference's sorted-time indexing is not)
single = onb.run_test(agent, n=n, start_points=starts, time_shift=tshift, horizon=60, print_progress=False, print_stats=False, batches=1)
sharded = parallel.run_test_sharded(agent, n, starts, tshift, horizon=60, batches=1)
if rank == 0:
    assert np.array_equal(sharded.done_at_step, single.done_at_step)
    assert np.array_equal(sharded.reached_source, single.reached_source)
    assert np.array_equal(np.array(sharded.positions), np.array(single.positions))
    assert np.array_equal(np.array(sharded.actions), np.array(single.actions))
    print(f'multi-GPU check OK on {ws} ranks: backup ({B} points, G={alphas.shape[0]}) and run_test ({n} agents, {int(single.reached_source.sum())} reached) identical to 1 GPU; FSVI / SSRA / GER train() sharded == unsharded')
dist.barrier()
dist.destroy_process_group()
