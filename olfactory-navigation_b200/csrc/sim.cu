// olfnav_sim_step — one whole simulation step behind one C-ABI call, no host synchronisation inside.
//
// Replaces the body of the reference's run_test loop (olfactory_navigation/simulation.py:1452-1497):
//   action = agent.choose_action()                       (agents/pbvi_agent.py:729-749 or agents/infotaxis_agent.py:244-301)
//   agent_position = environment.move(...)               (environment.py:711-769)
//   observation = environment.get_observation(...)       (environment.py:554-652)
//   source_reached = environment.source_reached(...)     (environment.py:655-680)
//   agent.update_state(action, observation, reached)     (agent.py:375-439 discretisation + BeliefSet.update)
//   to_terminate = source_reached | sims_at_end | ~update_succeeded ; arrays[~to_terminate] ; agent.kill(...)
// The agents that reached the source are dropped by the belief step itself (survivors are written to rows 0..m-1 of
// the other buffer; positions and time shifts are compacted the same way).  Failed updates and end-of-recording
// terminations are flagged in rec_term and counted; the (rare) second compaction is left to the caller.
#include "common.cuh"

namespace {

__device__ __forceinline__ int bmove_sim(int v, int d, int n, int wrap) {
    v += d;
    if (wrap) { if (v < 0) v += n; if (v >= n) v -= n; }
    else { v = max(0, min(n - 1, v)); }
    return v;
}

// time histogram for the reference's sorted-time indexing (Environment.reference_time_quirk).  Recordings of up to TH_SMEM_BINS frames
// are counted in shared memory first (100 000 agents on a few hundred frames are 100 000 atomics on a few hundred addresses: 25 us
// of a 4.8 ms step when they all go to global memory), one global atomic per non-empty bin and CTA afterwards.
constexpr int TH_SMEM_BINS = 4096;
constexpr int TH_ITEMS = 4;               // agents per thread
__global__ void __launch_bounds__(1024) time_hist_kernel(int n, const long long* __restrict__ ts, long long step, int T, int* __restrict__ hist) {
    __shared__ int bins[TH_SMEM_BINS];
    const bool local = T <= TH_SMEM_BINS;
    if (local) {
        for (int t = threadIdx.x; t < T; t += blockDim.x) bins[t] = 0;
        __syncthreads();
    }
    const int base = blockIdx.x * blockDim.x * TH_ITEMS;
#pragma unroll
    for (int k = 0; k < TH_ITEMS; ++k) {
        const int i = base + k * blockDim.x + threadIdx.x;
        if (i < n) {
            long long t = (ts[i] + step) % T;
            if (t < 0) t += T;
            if (local) atomicAdd(&bins[t], 1);
            else atomicAdd(&hist[t + 1], 1);
        }
    }
    if (local) {
        __syncthreads();
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
            const int c = bins[t];
            if (c) atomicAdd(&hist[t + 1], c);
        }
    }
}

// inclusive scan of hist[1..T] in place (hist[0] = 0): hist[t] = number of agents with time < t
__global__ void time_scan_kernel(int T, int* __restrict__ hist) {
    __shared__ int carry;
    __shared__ int wsum[32];
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 1; base <= T; base += 1024) {
        const int idx = base + threadIdx.x;
        int v = (idx <= T) ? hist[idx] : 0;
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
        if (lane == 31) wsum[w] = v;
        __syncthreads();
        if (w == 0) {
            int s = wsum[lane];
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, s, o); if (lane >= o) s += u; }
            wsum[lane] = s;
        }
        __syncthreads();
        const int total = v + (w > 0 ? wsum[w - 1] : 0) + carry;
        if (idx <= T) hist[idx] = total;
        __syncthreads();
        if (threadIdx.x == 1023) carry = total;
        __syncthreads();
    }
}

// layer histogram for the reference's sorted-layer indexing (same quirk as for the times, environment.py:594-595)
__global__ void layer_hist_kernel(int n, const int* __restrict__ act, const int* __restrict__ act_layer, int* __restrict__ hist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    atomicAdd(&hist[act_layer[act[i]] + 1], 1);
}

struct EnvArgs {
    const int* act_layer; const int* hist_layer; int thr_per_layer;
    int n;
    const int* pos_in; const int* act; const int* moves; const long long* ts; long long step;
    const double* thr; int n_thr; const int* hist; int time_mode; int time_loop;
    int* rec_pos; double* rec_odor; int* obs_id; unsigned char* rec_reached; unsigned char* rec_term; unsigned char* at_end;
};

__global__ void sim_env_kernel(const olfnav_env e, const EnvArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const int act = a.act[i];
    const int y = bmove_sim(a.pos_in[2 * i], a.moves[2 * act], e.H, e.wrap_y);
    const int x = bmove_sim(a.pos_in[2 * i + 1], a.moves[2 * act + 1], e.W, e.wrap_x);
    a.rec_pos[2 * i] = y;
    a.rec_pos[2 * i + 1] = x;
    long long t;
    if (a.time_mode == 1) {
        // agent i observes the frame of the i-th smallest effective time: largest t with hist[t] <= i
        int lo = 0, hi = e.T - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.hist[mid] <= i) lo = mid; else hi = mid - 1; }
        t = lo;
    } else {
        t = (a.ts[i] + a.step) % e.T;
        if (t < 0) t += e.T;
    }
    // layered environment: the layer is the first component of the action (simulation.py:1461-1463); under the reference's
    // indexing quirk agent i observes the i-th smallest layer
    const int own_layer = a.act_layer ? a.act_layer[act] : 0;
    int layer = own_layer;
    if (a.act_layer && a.time_mode == 1) {
        int lo = 0, hi = e.L - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (a.hist_layer[mid] <= i) lo = mid; else hi = mid - 1; }
        layer = lo;
    }
    const int dyp = y - e.my, dxp = x - e.mx;
    double v = 0.0;
    if (dyp >= 0 && dyp < e.h && dxp >= 0 && dxp < e.w) v = e.data[(((size_t)layer * e.T + t) * e.h + dyp) * e.w + dxp];
    a.rec_odor[i] = v;
    const double ddy = (double)y - e.sy, ddx = (double)x - e.sx;
    const bool hit = (ddy * ddy + ddx * ddx) <= e.r2;
    int id = -1;
    const double* thr = a.thr + (a.thr_per_layer ? (size_t)own_layer * a.n_thr : 0);      // per-layer thresholds (agent.py:414-417)
    for (int k = 0; k + 1 < a.n_thr; ++k)
        if (v >= thr[k] && v < thr[k + 1]) { id = k; break; }
    a.obs_id[i] = hit ? (a.n_thr - 1) : id;
    a.rec_reached[i] = hit ? 1 : 0;
    a.rec_term[i] = hit ? 1 : 0;
    a.at_end[i] = (!a.time_loop && (a.ts[i] + a.step + 1) >= e.T) ? 1 : 0;
}

// Order-preserving compaction of the agents that did not reach the source: src_rows[j] = j-th such agent; counts[0] = m.
// Two launches over tiles of 4096 flags (one CTA each): per-tile survivor counts, then every tile adds up the counts of the tiles
// before it (a few hundred values at most), scans its own flags and scatters.  (A single-CTA scan took 53 us per step at 100 000
// agents: 0.9 % of the step on the critical path between the environment step and the belief update.)
__global__ void __launch_bounds__(1024) sim_plan_count_kernel(int n, const unsigned char* __restrict__ reached, int* __restrict__ tile_counts) {
    __shared__ int wsum[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i0 = blockIdx.x * 4096 + 4 * threadIdx.x;
    int mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) mine += (i0 + k < n && !reached[i0 + k]) ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if (lane == 0) wsum[w] = mine;
    __syncthreads();
    if (w == 0) {
        int v = wsum[lane];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) tile_counts[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(1024) sim_plan_kernel(int n, const unsigned char* __restrict__ reached, const int* __restrict__ tile_counts,
                                                         int* __restrict__ src_rows, int* __restrict__ counts) {
    __shared__ int wsum[32];
    __shared__ int base_s;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    // survivors in the tiles before this one (and, for the last tile, the total)
    int before = 0;
    for (int j = threadIdx.x; j < (int)blockIdx.x; j += 1024) before += tile_counts[j];
    for (int o = 16; o > 0; o >>= 1) before += __shfl_xor_sync(0xffffffffu, before, o);
    if (lane == 0) wsum[w] = before;
    __syncthreads();
    if (w == 0) {
        int v = wsum[lane];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) base_s = v;
    }
    __syncthreads();
    const int base = base_s;
    const int i0 = blockIdx.x * 4096 + 4 * threadIdx.x;
    unsigned char f[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) f[k] = (i0 + k < n) ? (reached[i0 + k] ? 0 : 1) : 0;
    const int mine = f[0] + f[1] + f[2] + f[3];
    int v = mine;
    for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += u; }
    __syncthreads();                       // wsum is reused
    if (lane == 31) wsum[w] = v;
    __syncthreads();
    if (w == 0) {
        int t = wsum[lane];
        for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, t, o); if (lane >= o) t += u; }
        wsum[lane] = t;
    }
    __syncthreads();
    int off = base + (w > 0 ? wsum[w - 1] : 0) + v - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k) if (f[k]) src_rows[off++] = i0 + k;
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) { counts[0] = base + wsum[31]; counts[1] = 0; counts[2] = 0; counts[3] = n; }
}

__global__ void sim_gather_kernel(int n, const int* __restrict__ counts, const int* __restrict__ src_rows, const int* __restrict__ act,
                                  const int* __restrict__ obs_id, const int* __restrict__ rec_pos, const long long* __restrict__ ts_in,
                                  int* __restrict__ act_c, int* __restrict__ obs_c, int* __restrict__ pos_out, long long* __restrict__ ts_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || j >= counts[0]) return;
    const int r = src_rows[j];
    act_c[j] = act[r];
    obs_c[j] = obs_id[r];
    pos_out[2 * j] = rec_pos[2 * r];
    pos_out[2 * j + 1] = rec_pos[2 * r + 1];
    ts_out[j] = ts_in[r];
}

// in-place update: survivor j keeps the belief row of the agent it was (rows never move)
__global__ void sim_rows_kernel(int n, const int* __restrict__ counts, const int* __restrict__ src_rows, const int* __restrict__ brow_in,
                                int* __restrict__ brow_out) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || j >= counts[0]) return;
    const int r = src_rows[j];
    brow_out[j] = brow_in ? brow_in[r] : r;
}

// failed updates / end of recording among the survivors: flag and count them (the caller does the rare second compaction)
__global__ void sim_finish_kernel(int n, int* __restrict__ counts, const int* __restrict__ src_rows, const unsigned char* __restrict__ ok,
                                  const unsigned char* __restrict__ at_end, unsigned char* __restrict__ rec_term, unsigned char* __restrict__ term_c) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || j >= counts[0]) return;
    const int r = src_rows[j];
    const bool failed = !ok[j];
    const bool ended = at_end[r] != 0;
    term_c[j] = (failed || ended) ? 1 : 0;
    if (failed || ended) {
        rec_term[r] = 1;
        atomicAdd(&counts[failed ? 1 : 2], 1);
    }
}

}  // namespace

extern "C" int olfnav_sim_step(const olfnav_model* m, const olfnav_alphas* alphas, const olfnav_env* e,
                               const olfnav_sim_step_args* a, void* stream) {
    OLF_REQUIRE(m && e && a, "sim_step: null argument");
    OLF_REQUIRE(a->b_in && a->b_out && a->scale_in && a->scale_out, "sim_step: null belief buffer");
    OLF_REQUIRE(a->b_in != a->b_out || (a->inplace && a->brow_out), "sim_step: one belief buffer needs inplace = 1 and brow_out (the rows stay where they are)");
    OLF_REQUIRE(!a->inplace || (a->b_in == a->b_out && a->scale_in == a->scale_out && !a->act_next), "sim_step: inplace = 1 takes b_out == b_in, scale_out == scale_in and no one-pass policy");
    OLF_REQUIRE(!a->inplace || a->have_act, "sim_step: inplace = 1 needs the caller's policy (have_act = 1): the actions are indexed by agent, the rows are not");
    OLF_REQUIRE(a->pos_in && a->pos_out && a->ts_in && a->ts_out && a->moves && a->thresholds, "sim_step: null state pointer");
    OLF_REQUIRE(a->act && a->obs_id && a->src_rows && a->act_c && a->obs_c && a->ok && a->at_end && a->term_c && a->counts, "sim_step: null scratch pointer");
    OLF_REQUIRE(a->rec_pos && a->rec_odor && a->rec_reached && a->rec_term, "sim_step: null record pointer");
    OLF_REQUIRE(a->ld >= m->g.Sp && (a->ld % 4) == 0 && a->n >= 0 && a->n <= 0x7fffffffLL && a->n_thr >= 2, "sim_step: bad ld, n or n_thr");
    OLF_REQUIRE(alphas == nullptr || (a->eval_val && a->eval_idx), "sim_step: evaluation scratch missing");
    OLF_REQUIRE(a->time_mode == 0 || a->hist, "sim_step: time_mode 1 needs the hist scratch (T + 1 ints)");
    OLF_REQUIRE(m->device == e->device, "sim_step: model and environment live on different devices");
    const int n = (int)a->n;
    if (n == 0) return OLFNAV_OK;
    OLF_CUDA(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = olf_div_up(n, 256);
    int rc;

    // 1. policy (skipped when the caller already ran it on b_in)
    if (a->have_act) rc = OLFNAV_OK;
    else if (alphas) rc = olf_evaluate_into(m, alphas, a->b_in, a->ld, n, a->scale_in, a->eval_val, a->eval_idx, nullptr, nullptr, a->act, a->eval_path, st);
    else        rc = olfnav_infotaxis_choose(m, a->b_in, a->ld, n, a->scale_in, a->act, nullptr, stream);
    if (rc != OLFNAV_OK) return rc;
    if (a->ev_policy_end) OLF_CUDA(cudaEventRecord((cudaEvent_t)a->ev_policy_end, st));

    // 2. environment step (+ the reference's sorted-time indexing when asked for)
    if (a->time_mode == 1) {
        OLF_CUDA(cudaMemsetAsync(a->hist, 0, (size_t)(e->T + 1) * sizeof(int), st));
        time_hist_kernel<<<olf_div_up(n, 1024 * TH_ITEMS), 1024, 0, st>>>(n, (const long long*)a->ts_in, (long long)a->step, e->T, a->hist);
        time_scan_kernel<<<1, 1024, 0, st>>>(e->T, a->hist);
        olf_count_launch(2);
    }
    OLF_REQUIRE(e->L == 1 || a->act_layer, "sim_step: a layered environment needs act_layer (layer of every action id)");
    OLF_REQUIRE(!(a->act_layer && a->time_mode == 1) || a->hist_layer, "sim_step: time_mode 1 with layers needs the hist_layer scratch (L + 1 ints)");
    if (a->act_layer && a->time_mode == 1) {
        OLF_CUDA(cudaMemsetAsync(a->hist_layer, 0, (size_t)(e->L + 1) * sizeof(int), st));
        layer_hist_kernel<<<grid, 256, 0, st>>>(n, a->act, a->act_layer, a->hist_layer);
        time_scan_kernel<<<1, 1024, 0, st>>>(e->L, a->hist_layer);
        olf_count_launch(2);
    }
    EnvArgs ea;
    ea.act_layer = a->act_layer; ea.hist_layer = a->hist_layer; ea.thr_per_layer = a->thr_per_layer;
    ea.n = n; ea.pos_in = a->pos_in; ea.act = a->act; ea.moves = a->moves; ea.ts = (const long long*)a->ts_in; ea.step = (long long)a->step;
    ea.thr = a->thresholds; ea.n_thr = a->n_thr; ea.hist = a->hist; ea.time_mode = a->time_mode; ea.time_loop = a->time_loop;
    ea.rec_pos = a->rec_pos; ea.rec_odor = a->rec_odor; ea.obs_id = a->obs_id; ea.rec_reached = a->rec_reached; ea.rec_term = a->rec_term;
    ea.at_end = a->at_end;
    sim_env_kernel<<<grid, 256, 0, st>>>(*e, ea);
    OLF_LAUNCH_CHECK();

    // 3. survivors and their compacted bookkeeping
    {
        // per-tile counts live behind the compacted action ids (act_c holds n ints and is only written by the gather that follows...
        // which runs after the plan): no extra scratch in the ABI
        const int tiles = olf_div_up(n, 4096);
        OLF_REQUIRE(tiles <= n, "sim_step: internal scratch");
        int* tile_counts = a->act_c;
        sim_plan_count_kernel<<<tiles, 1024, 0, st>>>(n, a->rec_reached, tile_counts);
        sim_plan_kernel<<<tiles, 1024, 0, st>>>(n, a->rec_reached, tile_counts, a->src_rows, a->counts);
        olf_count_launch(1);
    }
    OLF_LAUNCH_CHECK();
    if (a->ev_plan_done) OLF_CUDA(cudaEventRecord((cudaEvent_t)a->ev_plan_done, st));
    sim_gather_kernel<<<grid, 256, 0, st>>>(n, a->counts, a->src_rows, a->act, a->obs_id, a->rec_pos, (const long long*)a->ts_in,
                                            a->act_c, a->obs_c, a->pos_out, (long long*)a->ts_out);
    OLF_LAUNCH_CHECK();

    // 4. fused belief update of the survivors into rows 0..m-1 of the other buffer (grid = n, exact count on the device)
    if (a->ev_update_begin) OLF_CUDA(cudaEventRecord((cudaEvent_t)a->ev_update_begin, st));
    if (a->act_next && alphas) {     // ... fused with the evaluation of the updated rows = the next step's policy
        OLF_REQUIRE(a->brow_out && a->belief_rows >= olfnav_fused_out_rows(m, n), "sim_step: the one-pass update needs brow_out and belief_rows >= olfnav_fused_out_rows");
        rc = olf_fused_step_launch(m, alphas, a->b_in, a->belief_rows, a->b_out, a->belief_rows, a->ld, n, a->act_c, a->obs_c, a->src_rows, a->brow_in,
                                   a->scale_in, a->scale_out, a->ok, nullptr, nullptr, a->act_next, a->brow_out, a->counts, st);
    } else if (a->inplace) {
        // one buffer: survivor j is updated in the row it already occupies (b_out == b_in with src == dst is safe, see belief.cu)
        sim_rows_kernel<<<grid, 256, 0, st>>>(n, a->counts, a->src_rows, a->brow_in, a->brow_out);
        OLF_LAUNCH_CHECK();
        rc = olf_belief_step_launch(m, a->b_in, a->b_out, a->ld, n, a->act_c, a->obs_c, a->brow_out, a->brow_out, a->scale_in, a->scale_out,
                                    a->ok, a->counts, stream);
    } else
        rc = olf_belief_step_launch(m, a->b_in, a->b_out, a->ld, n, a->act_c, a->obs_c, a->src_rows, nullptr, a->scale_in, a->scale_out,
                                    a->ok, a->counts, stream);
    if (rc != OLFNAV_OK) return rc;
    if (a->ev_update_end) OLF_CUDA(cudaEventRecord((cudaEvent_t)a->ev_update_end, st));

    // 5. failed updates / end of recording
    sim_finish_kernel<<<grid, 256, 0, st>>>(n, a->counts, a->src_rows, a->ok, a->at_end, a->rec_term, a->term_c);
    OLF_LAUNCH_CHECK();
    return OLFNAV_OK;
}
