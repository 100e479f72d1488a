'''Where do the device and oracle FSVI trainings at C2 part ways when both follow the SAME MDP policy and select the same beliefs?
Prints, per expansion, the alpha counts and, at the first difference, the float64 values new - old of the new belief points.'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import numpy as np, torch
import olfactory_navigation_b200 as onb
import replay_check
sim, setup = replay_check.oracle_setup('C2')
import pomdp_toolkit as tk
om = setup['model']
c = onb.synthetic.config('C2')
env = onb.Environment(**c['env_kwargs'])
agent = onb.agents.FSVI_Agent(env, thresholds=c['thresholds'])
m = agent.model
vi_o, _ = tk.solvers.VI.solve(om, horizon=1000, gamma=0.99, eps=1e-6)
Qo = vi_o.alpha_vector_array
shared = onb.ValueFunction(m, Qo, np.arange(4))
for E in range(1, 11):
    kw = dict(expansions=E, gamma=0.99, eps=1e-6, print_progress=False, print_stats=False)
    r_vf, r_h = tk.solvers.FSVI.solve(om, rng=np.random.default_rng(onb.synthetic.SEED), mdp_policy=vi_o, **kw)[:2]
    g_vf, g_h = onb.solvers.FSVI.solve(m, rng=np.random.default_rng(onb.synthetic.SEED), mdp_policy=shared, **kw)[:2]
    print(E, 'oracle', r_h.alpha_vector_counts[-1], 'device', g_h.alpha_vector_counts[-1], flush=True)
    if r_h.alpha_vector_counts[-1] != g_h.alpha_vector_counts[-1]:
        # redo the last expansion's backup in float64 from the oracle state of expansion E - 1
        kw1 = dict(kw, expansions=E - 1)
        p_vf, p_h = tk.solvers.FSVI.solve(om, rng=np.random.default_rng(onb.synthetic.SEED), mdp_policy=vi_o, **kw1)[:2] if E > 1 else (tk.ValueFunction(om, om.expected_rewards_table.T.copy(), om.actions), None)
        nb = r_h.final_belief_set.belief_array[(p_h.beliefs_counts[-1] if p_h else 1):]
        _, det = tk.solvers.backup(om, tk.BeliefSet(om, nb), p_vf, gamma=0.99, append=True, belief_dominance_prune=True, return_details=True)
        old = np.max(nb @ p_vf.alpha_vector_array.T, axis=1)
        print('new beliefs', len(nb), 'new - old (float64):', det['new_value'] - old)
        print('relative:', (det['new_value'] - old) / np.abs(old))
        print('oracle keep:', det['keep'])
        A_d, A_o = g_vf.alpha_vector_array, r_vf.alpha_vector_array
        twin = [bool(np.any(np.all(np.isclose(A_o, a[None, :], rtol=1e-5, atol=1e-9), axis=1))) for a in A_d]
        print('device alphas without an oracle twin:', [i for i, t in enumerate(twin) if not t])
        twin_o = [bool(np.any(np.all(np.isclose(A_d, a[None, :], rtol=1e-5, atol=1e-9), axis=1))) for a in A_o]
        print('oracle alphas without a device twin:', [i for i, t in enumerate(twin_o) if not t])
        # is the extra device vector a near-duplicate of another device vector?
        for i, t in enumerate(twin):
            pass
        D = A_d[:, None, :] - A_d[None, :, :]
        close = np.max(np.abs(D), axis=2) / (np.max(np.abs(A_d), axis=1)[:, None] + 1e-300)
        np.fill_diagonal(close, 1.0)
        i, j = np.unravel_index(np.argmin(close), close.shape)
        print('closest pair of device alphas:', i, j, close[i, j], 'actions', g_vf.actions[i], g_vf.actions[j])
        # the device's selection for the same points against the oracle's, with the float64 margins of the differing choices
        bs_d = onb.BeliefSet(m, nb.astype(np.float32).astype(np.float64))
        vf_prev = onb.ValueFunction(m, p_vf.alpha_vector_array, p_vf.actions)
        new_alpha, best_a, best_v, best_g = onb.solvers.backup_device(m, bs_d, vf_prev.device_alphas(0), 0.99, 0)
        best_a, best_g = best_a.cpu().numpy(), best_g.cpu().numpy()
        gam = tk.solvers.gamma_a_o(om, p_vf.alpha_vector_array, 0.99)
        sc = np.einsum('bs,aogs->baog', nb, gam)
        print('best_a equal:', np.array_equal(best_a, det['best_a']))
        for b in np.flatnonzero(best_a != det['best_a']):
            q = det['q'][b]
            print(f'  belief {b}: oracle a {det["best_a"][b]} device a {best_a[b]} q = {q!r} rel gap {(q[det["best_a"][b]] - q[best_a[b]]) / abs(q[det["best_a"][b]]):.3e}')
        for (b, a, o) in np.argwhere(best_g != det['best_g']):
            so, sd = sc[b, a, o, det['best_g'][b, a, o]], sc[b, a, o, best_g[b, a, o]]
            used = (a == det['best_a'][b])
            print(f'  belief {b} action {a} obs {o}: oracle g {det["best_g"][b, a, o]} ({so:.12e}) device g {best_g[b, a, o]} ({sd:.12e}) rel gap {(so - sd) / max(abs(so), 1e-300):.3e} used={used}')
        break
