// Model / value-function / environment handles and error plumbing of the C ABI.
#include "common.cuh"

#include <cmath>
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <atomic>

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void olf_count_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
extern "C" uint64_t olfnav_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

void olfnav_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* olfnav_last_error(void) { return g_err; }
extern "C" int olfnav_abi_version(void) { return OLFNAV_ABI_VERSION; }
// Row padding of the belief layout.  Narrow grids: Wp = W rounded up to 4 (rows start 16-byte aligned).  Grids of 128 columns and
// more: Wp = W rounded up to 64 floats, so that a grid row is a whole number of 128-byte cache lines (a vertical move shifts a
// belief by whole lines: measured on B200, 256-byte row segments written at a 16-byte offset stream at 4.8 TB/s, line-aligned ones at
// 5.9 TB/s, scripts/write_pattern_bench.cu) and the state space a whole number of 64-state K blocks of the tcgen05 kernels
// (C2: 371 -> 384 columns, +3.2 % bytes).  row_pad = 0 picks by width; 4 / 64 force a rule (OLFNAV_ROW_PAD in the Python layer).
extern "C" int olfnav_padded_width_for(int W, int row_pad) {
    const int pad = row_pad > 0 ? row_pad : (W >= 128 ? 64 : 4);
    return (W + pad - 1) / pad * pad;
}
extern "C" int olfnav_padded_width(int W) { return olfnav_padded_width_for(W, 0); }
extern "C" int64_t olfnav_padded_states(int H, int W) { return (int64_t)H * olfnav_padded_width(W); }

static int set_boundary(int boundary, int* wrap_y, int* wrap_x) {
    switch (boundary) {
        case OLFNAV_BOUNDARY_STOP: *wrap_y = 0; *wrap_x = 0; return 0;
        case OLFNAV_BOUNDARY_WRAP: *wrap_y = 1; *wrap_x = 1; return 0;
        case OLFNAV_BOUNDARY_WRAP_VERTICAL: *wrap_y = 1; *wrap_x = 0; return 0;
        case OLFNAV_BOUNDARY_WRAP_HORIZONTAL: *wrap_y = 0; *wrap_x = 1; return 0;
        default: return -1;
    }
}

static inline int bmove(int v, int d, int n, int wrap) {
    v += d;
    if (wrap) { if (v < 0) v += n; if (v >= n) v -= n; }
    else { if (v < 0) v = 0; if (v > n - 1) v = n - 1; }
    return v;
}

extern "C" int olfnav_model_create(olfnav_model** out, int device, int H, int W, int A, int O, const int32_t* moves, int boundary,
                                   const double* obs_table, const uint8_t* end_mask, const double* start) {
    return olfnav_model_create_padded(out, device, H, W, A, O, moves, boundary, obs_table, end_mask, start, 0);
}

extern "C" int olfnav_model_create_padded(olfnav_model** out, int device, int H, int W, int A, int O,
                                          const int32_t* moves, int boundary,
                                          const double* obs_table, const uint8_t* end_mask, const double* start, int row_pad) {
    OLF_REQUIRE(row_pad == 0 || row_pad == 4 || row_pad == 64, "model_create: row_pad must be 0 (by width), 4 or 64");
    OLF_REQUIRE(out && moves && obs_table && end_mask && start, "model_create: null argument");
    OLF_REQUIRE(H >= 1 && W >= 1 && (int64_t)H * W < (1LL << 30), "model_create: bad grid %d x %d", H, W);
    OLF_REQUIRE(A >= 1 && A <= OLFNAV_MAX_ACTIONS, "model_create: A must be in [1,%d]", OLFNAV_MAX_ACTIONS);
    OLF_REQUIRE(O >= 1 && O <= OLFNAV_MAX_OBSERVATIONS, "model_create: O must be in [1,%d]", OLFNAV_MAX_OBSERVATIONS);
    int wy, wx;
    OLF_REQUIRE(set_boundary(boundary, &wy, &wx) == 0, "model_create: unknown boundary %d", boundary);
    for (int a = 0; a < A; ++a) {
        if (abs(moves[2 * a]) > 1 || abs(moves[2 * a + 1]) > 1) {
            olfnav_set_error("model_create: move %d = (%d,%d): only unit steps are supported", a, moves[2 * a], moves[2 * a + 1]);
            return OLFNAV_ERR_UNSUPPORTED;
        }
    }
    int ndev = 0;
    OLF_CUDA(cudaGetDeviceCount(&ndev));
    OLF_REQUIRE(device >= 0 && device < ndev, "model_create: device %d of %d", device, ndev);
    OLF_CUDA(cudaSetDevice(device));
    // Stream-ordered temporaries (cudaMallocAsync in evaluate / backup_select) must not be handed back to the OS at every
    // synchronisation: with the default release threshold (0) each simulation step would re-map its scratch memory.
    if (!(getenv("OLFNAV_NO_POOL_RETAIN") && getenv("OLFNAV_NO_POOL_RETAIN")[0] == '1')) {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ULL;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }

    olfnav_model* m = new (std::nothrow) olfnav_model();
    if (!m) { olfnav_set_error("model_create: out of host memory"); return OLFNAV_ERR_NOMEM; }
    memset(m, 0, sizeof(*m));
    m->device = device;
    Geom& g = m->g;
    g.H = H; g.W = W; g.Wp = olfnav_padded_width_for(W, row_pad); g.W4 = g.Wp / 4; g.S = H * W; g.Sp = H * g.Wp;
    g.A = A; g.O = O; g.wrap_y = wy; g.wrap_x = wx;
    g.div_w4 = (g.W4 > 1) ? (unsigned)(((1ULL << 32) + g.W4 - 1) / g.W4) : 0u;
    for (int a = 0; a < A; ++a) { g.moves[a][0] = moves[2 * a]; g.moves[a][1] = moves[2 * a + 1]; }

    // Is the observation table action dependent?  (layered environments only; environment_converter.py:79-85)
    const int S = g.S, Sp = g.Sp;
    int per_action = 0;
    for (int s = 0; s < S && !per_action; ++s)
        for (int a = 1; a < A && !per_action; ++a)
            for (int o = 0; o < O; ++o)
                if (obs_table[((size_t)s * A + a) * O + o] != obs_table[((size_t)s * A) * O + o]) { per_action = 1; break; }
    g.obs_per_action = per_action;
    const int Aeff = per_action ? A : 1;

    std::vector<float> h_obs((size_t)Aeff * O * Sp, 0.f), h_ent((size_t)Aeff * Sp, 0.f), h_rew((size_t)A * Sp, 0.f), h_start(Sp, 0.f);
    std::vector<double> h_ent64((size_t)Aeff * Sp, 0.0);
    std::vector<uint8_t> h_end(Sp, 0);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const int s = y * W + x, sp = y * g.Wp + x;
            h_start[sp] = (float)start[s];
            h_end[sp] = end_mask[s] ? 1 : 0;
            for (int a = 0; a < Aeff; ++a) {
                double c = 0.0;
                for (int o = 0; o < O; ++o) {
                    const double pr = obs_table[((size_t)s * A + a) * O + o];
                    h_obs[((size_t)a * O + o) * Sp + sp] = (float)pr;
                    if (pr > 0.0) c += pr * log(pr);
                }
                h_ent[(size_t)a * Sp + sp] = (float)c;
                h_ent64[(size_t)a * Sp + sp] = c;
            }
            for (int a = 0; a < A; ++a) {
                const int y2 = bmove(y, g.moves[a][0], H, wy), x2 = bmove(x, g.moves[a][1], W, wx);
                h_rew[(size_t)a * Sp + sp] = end_mask[y2 * W + x2] ? 1.f : 0.f;
            }
        }

    // infotaxis pull tables (TF32 hi / lo planes): row a*(O+1)+o = O_o(R_a(s)), row a*(O+1)+O = C(R_a(s))
    m->pull_rows = A * (O + 1);
    m->pull_rows_p = (m->pull_rows + 15) & ~15;
    m->pull_ld = (Sp + 31) & ~31;
    std::vector<float> h_phi((size_t)m->pull_rows_p * m->pull_ld, 0.f), h_plo((size_t)m->pull_rows_p * m->pull_ld, 0.f);
    auto split = [](float v, float& h, float& l) {
        uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; memcpy(&h, &u, 4);
        float d = v - h; memcpy(&u, &d, 4); u &= 0xffffe000u; memcpy(&l, &u, 4);
    };
    for (int a = 0; a < A; ++a) {
        const int ae = per_action ? a : 0;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int sp = y * g.Wp + x;
                const int sp2 = bmove(y, g.moves[a][0], H, wy) * g.Wp + bmove(x, g.moves[a][1], W, wx);
                for (int o = 0; o <= O; ++o) {
                    const float v = (o < O) ? h_obs[((size_t)ae * O + o) * Sp + sp2] : h_ent[(size_t)ae * Sp + sp2];
                    const size_t at = (size_t)(a * (O + 1) + o) * m->pull_ld + sp;
                    split(v, h_phi[at], h_plo[at]);
                }
            }
    }

    // x-move masks of the one-pass update + policy kernel (see olfnav_model::coef)
    std::vector<float> h_coef((size_t)A * 2 * Sp, 0.f);
    for (int a = 0; a < A; ++a) {
        const int dx = g.moves[a][1];
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int sp = y * g.Wp + x;
                const int xs = x - dx;                                        // source column of the move
                h_coef[((size_t)a * 2 + 0) * Sp + sp] = (wx || (xs >= 0 && xs < W)) ? 1.f : 0.f;
                h_coef[((size_t)a * 2 + 1) * Sp + sp] = (!wx && dx != 0 && bmove(x, dx, W, 0) == x) ? 1.f : 0.f;
            }
    }

    cudaError_t e = cudaSuccess;
    auto up = [&](void** dptr, const void* src, size_t bytes) {
        if (e != cudaSuccess) return;
        e = cudaMalloc(dptr, bytes);
        if (e == cudaSuccess) e = cudaMemcpy(*dptr, src, bytes, cudaMemcpyHostToDevice);
    };
    up((void**)&m->obs, h_obs.data(), h_obs.size() * sizeof(float));
    up((void**)&m->obs_ent, h_ent.data(), h_ent.size() * sizeof(float));
    up((void**)&m->obs_ent64, h_ent64.data(), h_ent64.size() * sizeof(double));
    { const int zero = 0; up((void**)&m->it_unsure, &zero, sizeof(int)); }
    up((void**)&m->reward, h_rew.data(), h_rew.size() * sizeof(float));
    up((void**)&m->start, h_start.data(), h_start.size() * sizeof(float));
    up((void**)&m->end_mask, h_end.data(), h_end.size());
    up((void**)&m->pull_hi, h_phi.data(), h_phi.size() * sizeof(float));
    up((void**)&m->pull_lo, h_plo.data(), h_plo.size() * sizeof(float));
    up((void**)&m->coef, h_coef.data(), h_coef.size() * sizeof(float));
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&m->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) {
        olfnav_set_error("model_create: %s", cudaGetErrorString(e));
        olfnav_model_destroy(m);
        return OLFNAV_ERR_CUDA;
    }
    *out = m;
    return OLFNAV_OK;
}

extern "C" void olfnav_model_destroy(olfnav_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->obs); cudaFree(m->obs_ent); cudaFree(m->obs_ent64); cudaFree(m->it_unsure); cudaFree(m->reward); cudaFree(m->start); cudaFree(m->end_mask);
    cudaFree(m->pull_hi); cudaFree(m->pull_lo); cudaFree(m->coef); cudaFree(m->t_pull); cudaFree(m->t_push);
    delete m;
}

extern "C" int olfnav_model_dims(const olfnav_model* m, int* H, int* W, int* Wp, int* A, int* O) {
    OLF_REQUIRE(m, "model_dims: null model");
    if (H) *H = m->g.H;
    if (W) *W = m->g.W;
    if (Wp) *Wp = m->g.Wp;
    if (A) *A = m->g.A;
    if (O) *O = m->g.O;
    return OLFNAV_OK;
}

// ------------------------------------------------------------------------------------------------------
extern "C" int olfnav_env_create_layered(olfnav_env** out, int device, int H, int W, int boundary,
                                         const double* data, int L, int T, int h, int w, int margin_y, int margin_x,
                                         double source_y, double source_x, double source_radius) {
    OLF_REQUIRE(out && data, "env_create: null argument");
    OLF_REQUIRE(H >= 1 && W >= 1 && L >= 1 && T >= 1 && h >= 1 && w >= 1, "env_create: bad sizes");
    int wy, wx;
    OLF_REQUIRE(set_boundary(boundary, &wy, &wx) == 0, "env_create: unknown boundary %d", boundary);
    OLF_CUDA(cudaSetDevice(device));
    olfnav_env* e = new (std::nothrow) olfnav_env();
    if (!e) { olfnav_set_error("env_create: out of host memory"); return OLFNAV_ERR_NOMEM; }
    memset(e, 0, sizeof(*e));
    e->device = device; e->H = H; e->W = W; e->wrap_y = wy; e->wrap_x = wx;
    e->L = L; e->T = T; e->h = h; e->w = w; e->my = margin_y; e->mx = margin_x;
    e->sy = source_y; e->sx = source_x; e->r2 = source_radius * source_radius;
    const size_t bytes = (size_t)L * T * h * w * sizeof(double);
    cudaError_t ce = cudaMalloc((void**)&e->data, bytes);
    if (ce == cudaSuccess) ce = cudaMemcpy(e->data, data, bytes, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        olfnav_set_error("env_create: %s", cudaGetErrorString(ce));
        olfnav_env_destroy(e);
        return OLFNAV_ERR_CUDA;
    }
    *out = e;
    return OLFNAV_OK;
}

extern "C" int olfnav_env_create(olfnav_env** out, int device, int H, int W, int boundary,
                                 const double* data, int T, int h, int w, int margin_y, int margin_x,
                                 double source_y, double source_x, double source_radius) {
    return olfnav_env_create_layered(out, device, H, W, boundary, data, 1, T, h, w, margin_y, margin_x, source_y, source_x, source_radius);
}

extern "C" void olfnav_env_destroy(olfnav_env* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    cudaFree(e->data);
    delete e;
}
