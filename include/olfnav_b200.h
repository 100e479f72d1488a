/*
 * olfnav_b200.h — C ABI of the B200-native POMDP core for PimLb/olfactory-navigation.
 *
 * The reference has no FFI layer: its hot path is Python calling the third-party `pomdp_toolkit`
 * (reference requirements.txt:7).  Each entry point below replaces one `pomdp_toolkit` call (or one
 * in-repo per-step function) that the reference makes; the call site is cited as
 * `file:line` relative to /root/reference/olfactory_navigation/.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; `stream` is a cudaStream_t passed as void*;
 *   - every pointer marked [dev] is device memory owned by the CALLER (e.g. a torch tensor's
 *     data_ptr()); [host] pointers are host memory read synchronously during the call;
 *   - all work is enqueued on `stream`; no call synchronises the device unless it says so;
 *   - every function returns an olfnav_status; olfnav_last_error() gives a message (per thread);
 *   - there is no CPU fallback: without a CUDA device every compute entry returns
 *     OLFNAV_ERR_CUDA.
 *
 * Belief layout in HBM ("padded grid rows")
 *   A belief over the H x W grid is stored as H rows of Wp = olfnav_padded_width(W) floats
 *   (W rounded up to a multiple of 4, padding columns always 0), i.e. Sp = H*Wp floats per agent,
 *   agents `ld` floats apart (ld >= Sp, ld % 4 == 0).  State s = y*W + x of the reference
 *   (agents/model_based_util/environment_converter.py:43,111,118) lives at y*Wp + x.
 *   Beliefs are stored UNNORMALISED with a per-agent scale: b_true[n,:] = b[n,:] * scale[n]
 *   (deferred normalisation; the argmax of b.alpha is scale invariant).
 *   Alpha vectors use the same padded state indexing.
 */
#ifndef OLFNAV_B200_H
#define OLFNAV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OLFNAV_ABI_VERSION 1
#define OLFNAV_MAX_ACTIONS 16
#define OLFNAV_MAX_OBSERVATIONS 32

typedef enum {
    OLFNAV_OK = 0,
    OLFNAV_ERR_INVALID = 1,      /* bad argument */
    OLFNAV_ERR_CUDA = 2,         /* CUDA runtime error (message in olfnav_last_error) */
    OLFNAV_ERR_UNSUPPORTED = 3,  /* valid request outside what the kernels cover */
    OLFNAV_ERR_NOMEM = 4
} olfnav_status;

/* environment.py:711-769 boundary conditions */
typedef enum {
    OLFNAV_BOUNDARY_STOP = 0,            /* clip both axes ('stop', 'clip') */
    OLFNAV_BOUNDARY_WRAP = 1,            /* wrap both axes */
    OLFNAV_BOUNDARY_WRAP_VERTICAL = 2,   /* wrap rows (y), clip columns */
    OLFNAV_BOUNDARY_WRAP_HORIZONTAL = 3  /* wrap columns (x), clip rows */
} olfnav_boundary;

typedef struct olfnav_model olfnav_model;      /* POMDP tables on one device   */
typedef struct olfnav_alphas olfnav_alphas;    /* value function on one device */
typedef struct olfnav_env olfnav_env;          /* stepping environment on one device */

int         olfnav_abi_version(void);
const char* olfnav_last_error(void);
int         olfnav_padded_width(int W);                 /* Wp under the default rule: W rounded up to 4, to 64 from 128 columns up */
int         olfnav_padded_width_for(int W, int row_pad);/* Wp under an explicit rule: row_pad = 0 (by width), 4 or 64 */
int64_t     olfnav_padded_states(int H, int W);         /* Sp = H * Wp (default rule) */
uint64_t    olfnav_kernel_launches(void);               /* kernels this library has launched so far (process wide) */

/* ------------------------------------------------------------------------------------------------
 * Model  — replaces pomdp_toolkit.POMDP(...) built at
 *          agents/model_based_util/environment_converter.py:125-133 (grid model of exact_converter).
 *   moves        [host] A x 2 (dy, dx), |dy|,|dx| <= 1            (agent.py:239-250)
 *   obs_table    [host] S x A x O float64, P(o | arrive s', a)    (environment_converter.py:77-94)
 *   end_mask     [host] S bytes, 1 = end state                    (environment_converter.py:44-45)
 *   start        [host] S float64 start probabilities             (environment_converter.py:132)
 * Rewards follow the toolkit default: r_a(s) = 1 if R_a(s) is an end state.
 * ---------------------------------------------------------------------------------------------- */
int  olfnav_model_create(olfnav_model** out, int device, int H, int W, int A, int O,
                         const int32_t* moves, int boundary,
                         const double* obs_table, const uint8_t* end_mask, const double* start);
/* Same with an explicit row padding rule (see olfnav_padded_width_for); olfnav_model_create uses row_pad = 0. */
int  olfnav_model_create_padded(olfnav_model** out, int device, int H, int W, int A, int O,
                                const int32_t* moves, int boundary,
                                const double* obs_table, const uint8_t* end_mask, const double* start, int row_pad);
/* Dense stochastic-transition model — POMDP(states, actions, observations, transition_table = T[s, a, s'], observation_table,
 * end_states, start_probabilities) as built by the reference's minimal_converter
 * (agents/model_based_util/environment_converter.py:137-292, call :282-290).  Every entry point that takes an olfnav_model
 * (belief step, evaluate, backup, VI sweep, infotaxis, sim step) accepts it; the state space is a flat list (H = 1, W = S). */
int olfnav_model_create_dense(olfnav_model** out, int device, int S, int A, int O,
                              const double* transition_table,   /* [host] S x A x S */
                              const double* obs_table,          /* [host] S x A x O */
                              const uint8_t* end_mask,          /* [host] S */
                              const double* start);             /* [host] S */
void olfnav_model_destroy(olfnav_model* m);
int  olfnav_model_dims(const olfnav_model* m, int* H, int* W, int* Wp, int* A, int* O);

/* Move a dense reference-layout array (n x S, state-major rows) into / out of the padded layout.
 * dense_f64 / dense_f32: exactly one non-NULL.  `scale` may be NULL (treated as 1). */
int olfnav_pack_rows(const olfnav_model* m, const double* dense_f64, const float* dense_f32,      /* [dev] n x S */
                     float* padded, int64_t ld, int64_t n, void* stream);                       /* [dev] */
int olfnav_unpack_rows(const olfnav_model* m, const float* padded, int64_t ld, int64_t n,
                       const float* scale,                                                      /* [dev] n or NULL */
                       double* dense_f64, float* dense_f32, void* stream);                      /* [dev] n x S */

/* ------------------------------------------------------------------------------------------------
 * Beliefs — BeliefSet(model, [Belief(model) for _ in range(n)])  agents/pbvi_agent.py:717,
 *           agents/infotaxis_agent.py:232
 * ---------------------------------------------------------------------------------------------- */
int olfnav_belief_init(const olfnav_model* m, float* b, int64_t ld, int64_t n, float* scale, void* stream);

/* BeliefSet.update(actions, observations, raise_on_impossible_belief=False, use_reachability=True,
 *                  return_provenance=True)      agents/pbvi_agent.py:783-787, agents/infotaxis_agent.py:335-339
 * For i in [0,n):  u = O[., a_i, o_i] * push_{a_i}( b_in[src_i,:] * scale_in[src_i] );  Z = sum(u)
 *                  b_out[dst_i,:] = u ; scale_out[dst_i] = Z > 0 ? 1/Z : 0 ; ok[i] = Z > 0
 * src_rows / dst_rows may be NULL (identity).  In place (b_out == b_in) is allowed iff
 * src_i == dst_i for every i.  Writing dst_i = i with src_i = surviving rows compacts for free.
 * `ok` replaces the provenance array: ok[i] <=> row i is in provenance[:,0]. */
int olfnav_belief_step(const olfnav_model* m,
                       const float* b_in, float* b_out, int64_t ld, int64_t n,
                       const int32_t* act, const int32_t* obs,          /* [dev] n */
                       const int32_t* src_rows, const int32_t* dst_rows,/* [dev] n or NULL */
                       const float* scale_in, float* scale_out,         /* [dev] */
                       uint8_t* ok,                                     /* [dev] n */
                       void* stream);

/* One-pass  BeliefSet.update (agents/pbvi_agent.py:783-787)  +  ValueFunction.evaluate_at of the UPDATED beliefs
 * (agents/pbvi_agent.py:741, the choose_action of the next run_test iteration, simulation.py:1454): 8*S bytes of HBM traffic per
 * agent instead of 12*S.  Same arithmetic as olfnav_belief_step followed by olfnav_evaluate(b_out, scale_out):
 *   value[i] = max_g b'_i . alpha_g (normalised), arg[i] = first argmax, action[i] = alpha_actions[arg[i]]  (any may be NULL)
 * but the updated rows are written GROUPED BY ACTION, 128 to a tile (the kernel's unit of work): the belief of input i lands in
 * row row_out[i] of b_out, scale_out is indexed by that row, and b_out / scale_out must hold olfnav_fused_out_rows(m, n) rows.
 * src_rows[i] (NULL = i) is the row of input i in b_in; b_in holds rows_in rows.  b_out != b_in.
 * Supported when olfnav_can_fuse_policy(m, a) == 1 (<= 80 alpha vectors; grid rows padded to 64 floats; vertical moves wrap,
 * horizontal moves clip: the BASELINE plume configurations); OLFNAV_ERR_UNSUPPORTED otherwise (callers issue the two calls). */
int     olfnav_can_fuse_policy(const olfnav_model* m, const olfnav_alphas* a);
int64_t olfnav_fused_out_rows(const olfnav_model* m, int64_t n);
int olfnav_belief_step_evaluate(const olfnav_model* m, const olfnav_alphas* a,
                                const float* b_in, int64_t rows_in, float* b_out, int64_t rows_out, int64_t ld, int64_t n,
                                const int32_t* act, const int32_t* obs,  /* [dev] n */
                                const int32_t* src_rows,                 /* [dev] n or NULL */
                                const float* scale_in, float* scale_out, /* [dev] rows_in / rows_out */
                                uint8_t* ok,                             /* [dev] n */
                                float* value, int32_t* arg, int32_t* action, /* [dev] n or NULL */
                                int32_t* row_out,                        /* [dev] n or NULL */
                                void* stream);

/* Belief compaction — `belief_array[~kill]`, agents/pbvi_agent.py:810-811, agents/infotaxis_agent.py:362-363.
 * dst[i,:] = src[src_rows[i],:] and scale_dst[i] = scale_src[src_rows[i]] for i in [0,n); src != dst. */
int olfnav_copy_rows(const float* src, float* dst, int64_t ld, int64_t row_len,
                     const int32_t* src_rows, const float* scale_src, float* scale_dst, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Value function — ValueFunction(model, alpha_vectors, action_list)  agents/pbvi_agent.py:692
 *   alpha   [dev] G x ld_alpha floats in the padded state layout (padding columns ignored)
 *   actions [dev] G int32
 * The handle keeps its own packed copy (split-precision operand planes for the tensor-core path).
 * ---------------------------------------------------------------------------------------------- */
int  olfnav_alphas_create(olfnav_alphas** out, const olfnav_model* m, const float* alpha, int64_t ld_alpha,
                          const int32_t* actions, int G, void* stream);
void olfnav_alphas_destroy(olfnav_alphas* a);

/* ValueFunction.evaluate_at(BeliefSet) -> (values, actions)      agents/pbvi_agent.py:741
 * value[i] = scale[i] * max_g b[i,:].alpha_g ; arg[i] = first argmax ; action[i] = actions[arg[i]].
 * Any of value/arg/action may be NULL.  `path`: 0 = auto, 1 = FP32 SIMT kernel,
 * 2 = tcgen05 tensor-core kernel (split precision with FP32 accuracy: 3xTF32 below G = 64, 2-way FP16 split with power-of-two
 *     block scales from G = 64 up; OLFNAV_GEMM_F16=0 forces 3xTF32). */
int olfnav_evaluate(const olfnav_model* m, const olfnav_alphas* a,
                    const float* b, int64_t ld, int64_t n, const float* scale,
                    float* value, int32_t* arg, int32_t* action, int path, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Infotaxis — Infotaxis_Agent.choose_action           agents/infotaxis_agent.py:244-301
 * action[i] = first argmax_a [ H(b_i) - sum_o P(o|b_i,a) H(b_i^{a,o}) ]; dH (n x A) optional.
 * ---------------------------------------------------------------------------------------------- */
int olfnav_infotaxis_choose(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                            int32_t* action, float* dH, void* stream);
/* Agents of the last olfnav_infotaxis_choose call on this model whose best two actions were closer than the FP32 error bound and
 * were therefore re-evaluated in float64 (host int; synchronises `stream`). */
int olfnav_infotaxis_last_unsure(const olfnav_model* m, int32_t* count, void* stream);

/* BeliefSet.entropies   agents/infotaxis_agent.py:257 : H[i] = -sum b log b of the normalised row */
int olfnav_entropies(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                     float* H_out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * PBVI backup — inside pomdp_toolkit.solvers.<X>.solve        agents/fsvi_agent.py:201-224 (+7 siblings)
 *   gamma_build : Gamma[a][o][g][:] = gamma * O_o(R_a(.)) * alpha_g(R_a(.))      (A*O*G x ld)
 *   select      : g*[b,a,o] = argmax_g b.Gamma_{a,o}^g ; a*[b] = argmax_a ( b.r_a + sum_o max_g ... )
 *   assemble    : alpha_new[b,:] = r_{a*} + sum_o Gamma_{a*,o}^{g*[b,a*,o]}
 * ---------------------------------------------------------------------------------------------- */
int olfnav_backup_gamma(const olfnav_model* m, const float* alpha, int64_t ld_alpha, int G, float discount,
                        float* gamma_ao, int64_t ld_gamma, void* stream);
int olfnav_backup_select(const olfnav_model* m, const float* b, int64_t ld, int64_t n_beliefs, const float* scale,
                         const float* gamma_ao, int64_t ld_gamma, int G,
                         int32_t* best_g /* n x A x O */, int32_t* best_a /* n */, float* best_value /* n */,
                         int path, void* stream);
int olfnav_backup_assemble(const olfnav_model* m, const float* gamma_ao, int64_t ld_gamma, int G,
                           const int32_t* best_g, const int32_t* best_a, int64_t n_beliefs,
                           float* alpha_new, int64_t ld_new, void* stream);
/* The same backup without Gamma in memory (grid models, tensor path): the selection builds the FP16 operand planes of Gamma straight
 * from the alpha vectors, the assembly recomputes the selected Gamma rows from them (bit-identical vectors to the three calls above;
 * at G = 512 the FP32 Gamma is 783 MB).  alpha [dev] G x ld_alpha in the padded layout. */
int olfnav_backup_select_alphas(const olfnav_model* m, const float* b, int64_t ld, int64_t n, const float* scale,
                                const float* alpha, int64_t ld_alpha, int G, float discount,
                                int32_t* best_g, int32_t* best_a, float* best_value, void* stream);
int olfnav_backup_assemble_alphas(const olfnav_model* m, const float* alpha, int64_t ld_alpha, int G, float discount,
                                  const int32_t* best_g, const int32_t* best_a, int64_t n, float* alpha_new, int64_t ld_new,
                                  void* stream);

/* Row-pair reductions used by the belief-point selection heuristics and the pruning of solvers.*.solve
 * (agents/pbvi_ssea_agent.py:170, agents/pbvi_ger_agent.py:170, prune_level of every train(), agents/pbvi_agent.py:584-658):
 * out[i * ny + j] = reduce_s f(x_i(s), y_j(s)) with x_i = X[i,:] * sx[i], y_j = Y[j,:] * sy[j] (sx / sy NULL = 1), s < row_len:
 *   OLFNAV_PAIR_L1         f = |x - y|                                         (L1 distance between beliefs: SSEA)
 *   OLFNAV_PAIR_GER        f = (x - y) * ((x >= y ? vmax : vmin) - W[j,s])     (Pineau's error bound of belief x_i w.r.t. belief y_j
 *                                                                               whose best alpha vector is W[j,:]: GER)
 *   OLFNAV_PAIR_DOMINATED  out = 1 if y_j >= x_i everywhere and > somewhere, else 0   (point-wise dominance pruning of alpha vectors) */
#define OLFNAV_PAIR_L1 0
#define OLFNAV_PAIR_GER 1
#define OLFNAV_PAIR_DOMINATED 2
int olfnav_pair_reduce(int mode, const float* X, int64_t ldx, const float* sx, int64_t nx,
                       const float* Y, int64_t ldy, const float* sy, int64_t ny,
                       const float* W, int64_t ldw, float vmax, float vmin, int64_t row_len,
                       float* out /* [dev] nx x ny */, void* stream);

/* solvers.VI.solve sweep      agents/qmdp_agent.py:145-153
 * Q_out[a][s] = r_a(s) + discount * max_a' Q_in[a'][R_a(s)] ; max_change = max_s |V_out - V_in|.
 * float64 on purpose: V grows to 1/(1-discount) while the stopping rule looks at changes of 1e-4. */
int olfnav_vi_sweep(const olfnav_model* m, const double* Q_in, double* Q_out, int64_t ld_q, double discount,
                    double* max_change /* [dev] 1 double, atomically maxed; caller zeroes */, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Environment stepping — environment.py:711-769 (move), :554-652 (get_observation),
 *                        :655-680 (source_reached); agent.py:375-439 (discretize_observations)
 *   data [host] T x h x w float64 odor frames; margins (my, mx) = low margins of the data zone.
 * ---------------------------------------------------------------------------------------------- */
int  olfnav_env_create(olfnav_env** out, int device, int H, int W, int boundary,
                       const double* data, int T, int h, int w, int margin_y, int margin_x,
                       double source_y, double source_x, double source_radius);
/* Layered data (environment.py:202-235): data [host] L x T x h x w; olfnav_sim_step picks the layer of each agent's action. */
int  olfnav_env_create_layered(olfnav_env** out, int device, int H, int W, int boundary,
                               const double* data, int L, int T, int h, int w, int margin_y, int margin_x,
                               double source_y, double source_x, double source_radius);
void olfnav_env_destroy(olfnav_env* e);

/* One simulation step for n agents (simulation.py:1457-1466 + agent.py:401-439):
 *   pos <- move(pos, moves[act]); odor = data[(time_shift + step) % T, pos - margins] or 0 outside;
 *   reached = |pos - source|^2 <= r^2; obs_id = threshold bin of odor, or n_bins when reached.
 *   thresholds [dev] n_thr float64 ascending, thresholds[0] = -inf, thresholds[n_thr-1] = +inf. */
int olfnav_env_step(const olfnav_env* e, int64_t n,
                    int32_t* pos /* [dev] n x 2 (y, x), in/out */, const int32_t* act /* [dev] n */,
                    const int32_t* moves /* [dev] A x 2 */, const int64_t* time_shift /* [dev] n */, int64_t step,
                    const double* thresholds, int n_thr,
                    double* odor /* [dev] n */, int32_t* obs_id /* [dev] n */, uint8_t* reached /* [dev] n */,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * One whole simulation step — the body of run_test's loop, simulation.py:1452-1497 — with no host
 * synchronisation inside.  All pointers [dev]; arrays hold >= n entries unless noted.
 *   policy      alphas != NULL: argmax_g b.alpha_g (agents/pbvi_agent.py:741); NULL: infotaxis (agents/infotaxis_agent.py:244-301)
 *   env         move / observe / source check / discretise                      (environment.py:554-769, agent.py:401-439)
 *   update      survivors (agents that did not reach the source) are updated into rows 0..m-1 of b_out / scale_out,
 *               pos_out / ts_out hold their positions / time shifts in the same order (replaces arrays[~to_terminate] and
 *               belief_array[~kill], simulation.py:1495-1497, agents/pbvi_agent.py:810-811)
 *   counts      [0] = m survivors, [1] = failed updates among them, [2] = ended recordings among them (time_loop = 0), [3] = n
 *   records     indexed like the INPUT agents: act (action id), rec_pos (n x 2, after the move), rec_odor, rec_reached,
 *               rec_term (reached | failed | ended).  term_c[j] (j < m) flags survivors that must still be dropped.
 * ---------------------------------------------------------------------------------------------- */
typedef struct olfnav_sim_step_args {
    const float* b_in; float* b_out; int64_t ld;
    const float* scale_in; float* scale_out;
    const int32_t* pos_in; int32_t* pos_out;          /* n x 2 (y, x) */
    const int64_t* ts_in; int64_t* ts_out;            /* time shifts */
    int64_t n; int64_t step;
    const int32_t* moves;                             /* A x 2 */
    const double* thresholds; int32_t n_thr;
    int32_t time_mode;                                /* 0: each agent its own time; 1: the reference's sorted-time indexing */
    int32_t time_loop;                                /* 1: recordings loop for ever (run_test default) */
    int32_t eval_path;                                /* see olfnav_evaluate */
    int32_t have_act;                                 /* 1: `act` already holds this step's action ids (the caller ran the policy
                                                         on b_in ahead of time, overlapping it with its read-back of the previous
                                                         step's counts): the policy kernels are skipped */
    int32_t* act; int32_t* obs_id; int32_t* src_rows; int32_t* act_c; int32_t* obs_c;
    uint8_t* ok; uint8_t* at_end; uint8_t* term_c;
    int32_t* hist;                                    /* T + 1 ints, time_mode 1 only */
    float* eval_val; int32_t* eval_idx;               /* alphas != NULL only */
    int32_t* rec_pos; double* rec_odor; uint8_t* rec_reached; uint8_t* rec_term;
    int32_t* counts;                                  /* 4 ints */
    /* optional cudaEvent_t handles (NULL = none), recorded on `stream`: after the policy kernels, before and after the
     * fused belief update — lets a caller time the dominant kernels live without breaking the single call apart */
    void* ev_policy_end; void* ev_update_begin; void* ev_update_end;
    /* optional: act_next != NULL (needs alphas != NULL and olfnav_can_fuse_policy(m, alphas) == 1) => the belief update of the
     * survivors also evaluates the value function on the updated rows (olfnav_belief_step_evaluate) and writes the NEXT step's
     * action ids of the m survivors to act_next[0..m): the caller skips its own policy call for the next step.
     * The one-pass kernel groups the updated rows by action: brow_out[j] (j < m) receives the row of b_out that holds survivor j's
     * belief, brow_in[i] (NULL = i) tells where agent i's belief sits in b_in, scale_in / scale_out are indexed by belief row, and
     * both belief buffers hold belief_rows >= olfnav_fused_out_rows(m, n) rows. */
    int32_t* act_next;
    const int32_t* brow_in; int32_t* brow_out; int64_t belief_rows;
    /* inplace = 1: ONE belief buffer (b_out == b_in, scale_out == scale_in): the survivors are updated in the rows they occupy and
     * brow_out[j] = brow_in[src] carries where they are (10^6 agents x 30 793 states are 127 GB: two buffers do not fit one GPU).
     * The caller runs the policy itself (have_act = 1), over all rows, and gathers the actions through the row map. */
    int32_t inplace;
    /* layered environments (environment.py:202-235; simulation.py:1456-1463): layer of every action id (A ints, NULL without
     * layers), thresholds laid out [L][n_thr] when thr_per_layer = 1 (agent.py:414-417), L + 1 ints of scratch for time_mode 1 */
    const int32_t* act_layer; int32_t thr_per_layer; int32_t* hist_layer;
    /* optional cudaEvent_t recorded on `stream` right after the compaction plan, i.e. as soon as counts[0] (the survivors of this step)
     * is final and before the belief update starts: a caller that copies counts[0] out on another stream behind this event knows the
     * size of the NEXT step while this one is still running and can enqueue it without ever leaving the GPU idle */
    void* ev_plan_done;
} olfnav_sim_step_args;

int olfnav_sim_step(const olfnav_model* m, const olfnav_alphas* alphas /* NULL = infotaxis */, const olfnav_env* e,
                    const olfnav_sim_step_args* args /* [host] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OLFNAV_B200_H */
