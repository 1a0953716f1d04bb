// bio_capi.cu -- extern "C" entry points declared in include/bio_b200.h.
//
// A handle owns: the model block in the kernel's scalar type (device), the
// reference tables (device), the SoA env state (device) and the statistics
// buffer.  I/O buffers of bio_step/bio_reset are caller-owned device pointers;
// nothing is allocated per step.  There is no CPU fallback: every compute entry
// point launches a kernel or fails.
#include <cuda_runtime.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <new>
#include <stdlib.h>
#include <string>
#include <vector>

#include "bio_launch.cuh"


namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(-2, std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + \
                                std::to_string(__LINE__) + ")");                                   \
    } while (0)

struct HandleBase {
    int precision = 0, device = 0, n = 0;
    int64_t launches = 0;
    BioModelTables model;
    BioTaskConfig task;
    unsigned long long seed = 0;
    long long env_offset = 0;
    int block = 32;
    int n_sms = 148;
    int coop_grid = 0;
    int coop_ctas = 1;            // resident CTAs of the cooperative kernel per SM (occupancy query)
    int coop_threads = 256;       // launch shape of the cooperative kernel
    std::vector<void*> allocs;
    double* stats = nullptr;          // 16 doubles
    // *_host entry points: the small device->host copies run on a side stream next to the observation copy
    cudaStream_t side = nullptr;
    cudaEvent_t ev_step = nullptr;
    // page-locked host buffers are read (actions) and written (obs, reward, done, terms) by the kernel in
    // place; BIO_HOST_ZEROCOPY=1: actions only, 0: staged copies for everything
    int host_zero_copy = 2;
    // stream of the *_host entry points (bio_set_host_stream; default: the legacy stream, which orders
    // itself against every blocking stream)
    cudaStream_t host_stream = nullptr;
    virtual ~HandleBase() {}
};

template <typename T>
struct Handle : HandleBase {
    bio::DevModel<T>* d_model = nullptr;
    bio::DevTask<T> task_d;
    bio::EnvState<T> st;
    size_t smem = 0;
    int coop_cls = -1;        // -1: thread-per-env kernel, 0/1: cooperative kernel size class
    size_t coop_smem = 0;
    // device scratch for the *_host entry points
    T* h_actions = nullptr; T* h_obs = nullptr; T* h_reward = nullptr; T* h_terms = nullptr;
    uint8_t* h_done = nullptr; uint8_t* h_mask = nullptr;
};

template <typename T>
int dev_alloc(HandleBase* h, T** p, size_t count) {
    void* q = nullptr;
    CU(cudaMalloc(&q, (count ? count : 1) * sizeof(T)));
    CU(cudaMemset(q, 0, (count ? count : 1) * sizeof(T)));
    h->allocs.push_back(q);
    *p = (T*)q;
    return 0;
}

template <typename T>
int upload_ref(HandleBase* h, const double* src, size_t count, const T** dst) {
    std::vector<T> tmp(count ? count : 1);
    for (size_t i = 0; i < count; i++) tmp[i] = (T)src[i];
    T* d = nullptr;
    int rc = dev_alloc(h, &d, count);
    if (rc) return rc;
    CU(cudaMemcpy(d, tmp.data(), count * sizeof(T), cudaMemcpyHostToDevice));
    *dst = d;
    return 0;
}

int pick_block(int n) {
    // one thread per env: keep every SM sub-partition busy before stacking
    // warps in a CTA (148 SMs x 4 schedulers)
    if (n <= 148 * 4 * 32) return 32;
    if (n <= 148 * 8 * 64) return 64;
    return 128;
}

template <typename T>
int set_kernel_attrs(Handle<T>* h) {
    h->smem = sizeof(bio::DevModel<T>);
    CU(bio::thread_set_smem<T>((int)h->smem));
    return 0;
}

// Launch shape of the cooperative kernel by the number of item rounds per SM (see COOP_THREADS_LO / _HI in
// bio_coop.cuh), over the SMs the handle launches on (bio_set_grid: a group of a quarter of the batch on a quarter
// of the SMs has the item count per SM of the whole batch, not a quarter of it); BIO_COOP_THREADS = lo | hi |
// <thread count of one of the two shapes> forces one (tests, experiments).  Sets coop_threads / coop_smem / coop_ctas.
template <typename T>
int pick_coop_shape(Handle<T>* h) {
    if (h->coop_cls < 0) return 0;
    const int cls = h->coop_cls;
    const int G = cls == 0 ? (int)bio::CoopCls<0>::G : (int)bio::CoopCls<1>::G;
    const int lo = COOP_THREADS_LO(T), hi = cls == 1 ? COOP_THREADS_HI(T, 1) : COOP_THREADS_HI(T, 0);
    int threads = 0;
    if (const char* e = getenv("BIO_COOP_THREADS")) {
        if (!strcmp(e, "lo")) threads = lo;
        else if (!strcmp(e, "hi")) threads = hi;
        else { const int v = atoi(e); if (v == lo || v == hi) threads = v; }
    }
    if (!threads) {
        const long sms = h->coop_grid > 0 ? h->coop_grid : h->n_sms;
        const long items = (h->n + (32 / G) - 1) / (32 / G);
        const long ipsm = (items + sms - 1) / sms;
        const double cost_lo = (double)((ipsm + lo / 32 - 1) / (lo / 32));
        const double cost_hi = (double)((ipsm + hi / 32 - 1) / (hi / 32)) * (cls == 1 ? COOP_SHAPE_COST(1) : COOP_SHAPE_COST(0));
        threads = cost_hi < cost_lo ? hi : lo;
    }
    const size_t base = ((sizeof(bio::DevModel<T>) + 15) / 16) * 16;
    h->coop_threads = threads;
    if (cls == 0) {
        h->coop_smem = base + (threads / G) * sizeof(bio::EnvWork<T, 0>);
        CU((bio::coop_set_smem<T, 0>(threads, (int)h->coop_smem)));
        h->coop_ctas = bio::coop_ctas_per_sm<T, 0>(threads, (int)h->coop_smem);
    } else {
        h->coop_smem = base + (threads / G) * sizeof(bio::EnvWork<T, 1>);
        CU((bio::coop_set_smem<T, 1>(threads, (int)h->coop_smem)));
        h->coop_ctas = bio::coop_ctas_per_sm<T, 1>(threads, (int)h->coop_smem);
    }
    if (h->coop_ctas < 1) return fail(-2, "the cooperative step kernel does not fit an SM");
    return 0;
}

template <typename T>
int create_impl(const BioModelTables* model, const BioTaskConfig* task, const BioRefTables* ref, int n, int device,
                unsigned long long seed, long long env_offset, Handle<T>* h) {
    CU(cudaSetDevice(device));
    h->device = device;
    h->n = n;
    h->model = *model;
    h->task = *task;
    h->seed = seed;
    h->env_offset = env_offset;
    h->block = pick_block(n);
    CU(cudaDeviceGetAttribute(&h->n_sms, cudaDevAttrMultiProcessorCount, device));
    if (const char* e = getenv("BIO_COOP_GRID")) h->coop_grid = atoi(e);
    if (const char* e = getenv("BIO_BLOCK")) { int b = atoi(e); if (b >= 32 && b <= 256 && b % 32 == 0) h->block = b; }
    int rc;
    if ((rc = set_kernel_attrs(h))) return rc;
    bool prog_ok = false, gpath_ok = false;
    int prog_src = 0;
    // model block
    {
        bio::DevModel<T>* hm = new bio::DevModel<T>();
        bio::convert_model(*model, *hm);
        if (bio::build_obs_desc(*model, *task, *hm) != task->obs_dim) {
            delete hm;
            return fail(-1, "obs_dim of the task config does not match the observation layout");
        }
        rc = dev_alloc(h, (unsigned char**)&h->d_model, sizeof(bio::DevModel<T>));
        if (rc) { delete hm; return rc; }
        cudaError_t e = cudaMemcpy(h->d_model, hm, sizeof(bio::DevModel<T>), cudaMemcpyHostToDevice);
        prog_ok = hm->prog.ok != 0;
        prog_src = hm->prog.n_src;
        gpath_ok = hm->prog.gpath_ok != 0;
        delete hm;
        CU(e);
    }
    // cooperative kernel: size class by model, BIO_KERNEL=thread forces the thread-per-env kernel
    {
        const char* kv = getenv("BIO_KERNEL");
        const bool want_coop = !(kv && strcmp(kv, "thread") == 0);
        typedef bio::CoopCls<0> C0;
        typedef bio::CoopCls<1> C1;
        int max_anc = 0, max_mov = 0;
        for (int i = 0; i < model->n_dof; i++) {
            int c = 0;
            for (int j = 0; j < i; j++) c += (model->dof_anc_mask[i] >> j) & 1u;
            if (c > max_anc) max_anc = c;
        }
        for (int i = 0; i < model->n_muscles; i++) {
            int c = 0;
            for (int p = model->mus_pt_begin[i]; p < model->mus_pt_begin[i] + model->mus_pt_count[i]; p++)
                c += model->pt_kind[p] == BIO_PT_MOVING;
            if (c > max_mov) max_mov = c;
        }
        auto fits = [&](int G, int ND, int NM, int NP, int NAX) {
            // factorisation step: (ancestors choose 2) + ancestors pairs on G lanes x (1 | 2) rounds
            if (max_anc * (max_anc + 1) / 2 > G * (G == 16 ? 1 : 2) || max_mov > 1) return false;
            if (model->n_spheres * (max_anc + 1) > 64 || model->n_spheres > 8) return false;
            return model->n_dof <= ND && model->n_muscles <= NM && model->n_act <= NM && model->n_pathpts <= NP &&
                   model->n_axes <= NAX && model->n_bodies + model->n_dof <= G &&
                   model->n_bodies + model->n_obspts <= G && model->n_spheres + model->n_limits <= G &&
                   model->n_obspts <= COOP_MAXOBS && model->n_coords <= 2 * G && task->n_pd <= G;
        };
        // class 0 (half-warp per env) runs the planar program only
        // (the cooperative kernels write the observation row from a 256-slot descriptor table)
        if (task->obs_dim > BIO_COOP_MAX_OBS_DIM) {
            h->coop_cls = -1;
        } else if (want_coop && prog_ok && prog_src <= P2_MAXSRC && fits(C0::G, C0::ND, C0::NM, C0::NP, C0::NAX)) {
            h->coop_cls = 0;
            if ((rc = pick_coop_shape<T>(h))) return rc;
        } else if (want_coop && fits(C1::G, C1::ND, C1::NM, C1::NP, C1::NAX) &&
                   (prog_ok ? prog_src <= P2_MAXSRC : (model->n_muscles == 0 || (gpath_ok && prog_src <= COOP_MAXSRC6)))) {
            // class 1 (warp per env): planar program, or the general evaluation with compiled muscle paths
            h->coop_cls = 1;
            if ((rc = pick_coop_shape<T>(h))) return rc;
        }
        if (h->coop_cls >= 0 && h->coop_ctas < 1) return fail(-2, "the cooperative step kernel does not fit an SM");
    }
    bio::convert_task(*task, h->task_d);
    const int nd = model->n_dof, nm = model->n_muscles, na = model->n_act;
    // reference tables
    h->task_d.ref_rows = ref->n_rows; h->task_d.ref_coords = ref->n_coords; h->task_d.ref_bodies = ref->n_refbodies;
    if ((rc = upload_ref<T>(h, ref->q, (size_t)ref->n_rows * ref->n_coords, &h->task_d.ref_q))) return rc;
    if ((rc = upload_ref<T>(h, ref->u, (size_t)ref->n_rows * ref->n_coords, &h->task_d.ref_u))) return rc;
    if ((rc = upload_ref<T>(h, ref->body_pos, (size_t)ref->n_rows * ref->n_refbodies * 3, &h->task_d.ref_body_pos))) return rc;
    if ((rc = upload_ref<T>(h, ref->com_pos, (size_t)ref->n_rows * 3, &h->task_d.ref_com_pos))) return rc;
    T* lm0 = nullptr;
    if ((rc = dev_alloc(h, &lm0, (size_t)ref->n_rows * (nm ? nm : 1)))) return rc;
    h->task_d.ref_lm0 = lm0;
    if (nm > 0) {
        const int blk = 32, grid = (ref->n_rows + blk - 1) / blk;
        bio::launch_lm0<T>(grid, blk, h->smem, h->d_model, h->task_d.ref_q, ref->n_rows, ref->n_coords, lm0);
        h->launches++;
        CU(cudaGetLastError());
    }
    // env state
    const size_t N = (size_t)n;
    h->st.aos = h->coop_cls >= 0 ? 1 : 0;   // env-major records for the cooperative kernels (EnvState)
    if ((rc = dev_alloc(h, &h->st.q, N * nd))) return rc;
    if ((rc = dev_alloc(h, &h->st.u, N * nd))) return rc;
    if ((rc = dev_alloc(h, &h->st.act, N * nm))) return rc;
    if ((rc = dev_alloc(h, &h->st.lm, N * nm))) return rc;
    if ((rc = dev_alloc(h, &h->st.last_action, N * na))) return rc;
    if ((rc = dev_alloc(h, &h->st.history, N * na * task->horizon))) return rc;
    if ((rc = dev_alloc(h, &h->st.old_px, N))) return rc;
    if ((rc = dev_alloc(h, &h->st.ep_return, N))) return rc;
    if ((rc = dev_alloc(h, &h->st.istep, N))) return rc;
    if ((rc = dev_alloc(h, &h->st.first, N))) return rc;
    if ((rc = dev_alloc(h, &h->st.hist_pos, N))) return rc;
    if ((rc = dev_alloc(h, &h->st.ep_len, N))) return rc;
    if ((rc = dev_alloc(h, &h->st.episode, N))) return rc;
    if ((rc = dev_alloc(h, &h->stats, 16))) return rc;
    // scratch for host entry points
    if ((rc = dev_alloc(h, &h->h_actions, N * na))) return rc;
    if ((rc = dev_alloc(h, &h->h_obs, N * task->obs_dim))) return rc;
    if ((rc = dev_alloc(h, &h->h_reward, N))) return rc;
    if ((rc = dev_alloc(h, &h->h_terms, N * task->n_reward_terms))) return rc;
    if ((rc = dev_alloc(h, &h->h_done, N))) return rc;
    if ((rc = dev_alloc(h, &h->h_mask, N))) return rc;
    CU(cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&h->ev_step, cudaEventDisableTiming));
    { const char* z = getenv("BIO_HOST_ZEROCOPY"); if (z && z[0] >= '0' && z[0] <= '9') h->host_zero_copy = z[0] - '0'; }
    CU(cudaDeviceSynchronize());
    // initial reference-state reset of all envs
    const int grid = (n + h->block - 1) / h->block;
    bio::launch_reset<T>(grid, h->block, h->smem, 0, h->d_model, h->task_d, h->st, n, h->seed, h->env_offset, nullptr,
                         nullptr, 0);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaDeviceSynchronize());
    return 0;
}

int validate(const BioModelTables* m, const BioTaskConfig* t, const BioRefTables* r, int n) {
    if (!m || !t || !r) return fail(-1, "null table pointer");
    if (m->abi_version != BIO_ABI_VERSION || t->abi_version != BIO_ABI_VERSION)
        return fail(-1, "table ABI version mismatch");
    if (n <= 0) return fail(-1, "n_envs must be positive");
    if (m->n_bodies < 1 || m->n_bodies > BIO_MAX_BODIES || m->n_dof < 1 || m->n_dof > BIO_MAX_DOF ||
        m->n_muscles < 0 || m->n_muscles > BIO_MAX_MUSCLES || m->n_act < 1 || m->n_act > BIO_MAX_ACT)
        return fail(-1, "model sizes out of range");
    for (int i = 0; i < m->n_muscles; i++)
        if (m->mus_pt_count[i] > BIO_MAX_MUSCLE_PTS) return fail(-1, "muscle with too many path points");
    if (t->horizon < 1 || t->horizon > BIO_MAX_HORIZON) return fail(-1, "horizon out of range");
    if (t->n_substeps < 1) return fail(-1, "n_substeps must be >= 1");
    if (t->integrator == BIO_INT_ADAPTIVE_RKM)
        return fail(-1, "adaptive_rkm is the CPU-baseline scheme of the oracle; the CUDA path runs fixed-step schemes");
    if (t->integrator < 0 || t->integrator > BIO_INT_IMPLICIT_DAMPING) return fail(-1, "unknown integrator");
    if (r->n_coords != m->n_coords) return fail(-1, "reference tables have a different coordinate count");
    if (r->n_rows < 2 || !r->q || !r->u || !r->body_pos || !r->com_pos) return fail(-1, "bad reference tables");
    return 0;
}

template <typename T>
int step_impl(Handle<T>* h, const void* actions, void* obs, void* reward, uint8_t* done, void* terms,
              cudaStream_t s) {
    if (h->coop_cls >= 0) {
        // persistent launch: resident CTAs only, warps walk the env items round-robin over the CTAs
        const int G = h->coop_cls == 0 ? (int)bio::CoopCls<0>::G : (int)bio::CoopCls<1>::G;
        const int items = (h->n + (32 / G) - 1) / (32 / G);
        const int ctas = h->n_sms * (h->coop_ctas > 0 ? h->coop_ctas : 1);
        int grid = items < ctas ? items : ctas;
        if (h->coop_grid > 0) grid = h->coop_grid;   // BIO_COOP_GRID: launch-shape experiments
        if (h->coop_cls == 0)
            bio::launch_coop<T, 0>(h->coop_threads, grid, h->coop_smem, s, h->d_model, h->task_d, h->st, h->n, h->seed, h->env_offset, (const T*)actions, (T*)obs, (T*)reward, done,
                (T*)terms, h->stats);
        else
            bio::launch_coop<T, 1>(h->coop_threads, grid, h->coop_smem, s, h->d_model, h->task_d, h->st, h->n, h->seed, h->env_offset, (const T*)actions, (T*)obs, (T*)reward, done,
                (T*)terms, h->stats);
    } else {
        const int grid = (h->n + h->block - 1) / h->block;
        bio::launch_step<T>(grid, h->block, h->smem, s, h->d_model, h->task_d, h->st, h->n, h->seed, h->env_offset,
                            (const T*)actions, (T*)obs, (T*)reward, done, (T*)terms, h->stats);
    }
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int reset_impl(Handle<T>* h, const uint8_t* mask, void* obs, cudaStream_t s, int bump) {
    const int grid = (h->n + h->block - 1) / h->block;
    bio::launch_reset<T>(grid, h->block, h->smem, s, h->d_model, h->task_d, h->st, h->n, h->seed, h->env_offset, mask,
                         (T*)obs, bump);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int transpose(Handle<T>* h, const T* src, T* dst, int k, int to_soa, cudaStream_t s) {
    const size_t total = (size_t)h->n * k;
    if (!total) return 0;
    const int blk = 256;
    bio::launch_transpose<T>((unsigned)((total + blk - 1) / blk), blk, s, src, dst, h->n, k, to_soa);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int state_impl(Handle<T>* h, const BioStatePtrs* p, cudaStream_t s, bool set) {
    const int nd = h->model.n_dof, nm = h->model.n_muscles, na = h->model.n_act, H = h->task.horizon;
    int rc;
    struct Item { void* user; T* soa; int k; };
    Item items[] = {{p->q, h->st.q, nd}, {p->u, h->st.u, nd}, {p->act, h->st.act, nm}, {p->lm, h->st.lm, nm},
                    {p->last_action, h->st.last_action, na}, {p->history, h->st.history, H * na},
                    {p->old_px, h->st.old_px, 1}};
    for (auto& it : items) {
        if (!it.user || it.k == 0) continue;
        if (h->st.aos || it.k == 1) {            // env-major state: the caller's [N][k] rows are the storage layout
            if (set) CU(cudaMemcpyAsync(it.soa, it.user, sizeof(T) * h->n * it.k, cudaMemcpyDeviceToDevice, s));
            else CU(cudaMemcpyAsync(it.user, it.soa, sizeof(T) * h->n * it.k, cudaMemcpyDeviceToDevice, s));
            continue;
        }
        if (set) rc = transpose<T>(h, (const T*)it.user, it.soa, it.k, 1, s);
        else rc = transpose<T>(h, it.soa, (T*)it.user, it.k, 0, s);
        if (rc) return rc;
    }
    if (p->istep) {
        if (set) CU(cudaMemcpyAsync(h->st.istep, p->istep, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, s));
        else CU(cudaMemcpyAsync(p->istep, h->st.istep, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, s));
    }
    if (p->first) {
        if (set) CU(cudaMemcpyAsync(h->st.first, p->first, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, s));
        else CU(cudaMemcpyAsync(p->first, h->st.first, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, s));
    }
    if (p->hist_pos) {
        if (set) CU(cudaMemcpyAsync(h->st.hist_pos, p->hist_pos, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, s));
        else CU(cudaMemcpyAsync(p->hist_pos, h->st.hist_pos, sizeof(int32_t) * h->n, cudaMemcpyDeviceToDevice, s));
    }
    if (p->episode) {
        if (set) CU(cudaMemcpyAsync(h->st.episode, p->episode, sizeof(int64_t) * h->n, cudaMemcpyDeviceToDevice, s));
        else CU(cudaMemcpyAsync(p->episode, h->st.episode, sizeof(int64_t) * h->n, cudaMemcpyDeviceToDevice, s));
    }
    return 0;
}

template <typename T>
int eval_impl(Handle<T>* h, const void* controls, const BioDebugPtrs* o, cudaStream_t s) {
    bio::DebugOut<T> d;
    d.udot = (T*)o->udot; d.tendon_force = (T*)o->tendon_force; d.fiber_force = (T*)o->fiber_force;
    d.fiber_vel = (T*)o->fiber_vel; d.act_dot = (T*)o->act_dot; d.path_len = (T*)o->path_len;
    d.path_vel = (T*)o->path_vel; d.contact = (T*)o->contact; d.limit_force = (T*)o->limit_force;
    d.mass_matrix = (T*)o->mass_matrix; d.bias = (T*)o->bias;
    const int grid = (h->n + h->block - 1) / h->block;
    bio::launch_eval<T>(grid, h->block, h->smem, s, h->d_model, h->task_d, h->st, h->n, h->seed, h->env_offset,
                        (const T*)controls, d);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int id_impl(Handle<T>* h, int op, const void* x, const void* controls, const void* shift, void* out, cudaStream_t s) {
    const int grid = (h->n + h->block - 1) / h->block;
    bio::launch_id<T>(grid, h->block, h->smem, s, h->d_model, h->task_d, h->st, h->n, h->seed, h->env_offset, op,
                      (const T*)x, (const T*)controls, (const T*)shift, (T*)out);
    h->launches++;
    CU(cudaGetLastError());
    return 0;
}

template <typename T>
int extra_impl(Handle<T>* h, const BioStepExtra* e) {
    bio::StepExtra<T>& x = h->task_d.ex;
    memset(&x, 0, sizeof(x));
    if (!e) return 0;
    x.terminal_obs = (T*)e->terminal_obs; x.done_reason = e->done_reason; x.udot = (T*)e->udot;
    x.tendon_force = (T*)e->tendon_force; x.fiber_force = (T*)e->fiber_force; x.fiber_vel = (T*)e->fiber_vel;
    x.contact = (T*)e->contact; x.limit_force = (T*)e->limit_force;
    x.any = (e->terminal_obs || e->done_reason || e->udot || e->tendon_force || e->fiber_force || e->fiber_vel ||
             e->contact || e->limit_force) ? 1 : 0;
    return 0;
}

template <typename T>
int step_host_impl(Handle<T>* h, const void* actions, void* obs, void* reward, uint8_t* done, void* terms) {
    const size_t N = h->n;
    const int na = h->model.n_act, od = h->task.obs_dim, nt = h->task.n_reward_terms;
    cudaStream_t hs = h->host_stream;
    // actions in page-locked host memory are read by the kernel in place (one 56..88-byte row per env over
    // PCIe instead of a staged copy and its launch); pageable memory goes through the staging buffer
    const void* a_dev = nullptr;
    if (h->host_zero_copy) {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, actions) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            a_dev = at.devicePointer;
        else
            cudaGetLastError();
    }
    if (!a_dev) {
        CU(cudaMemcpyAsync(h->h_actions, actions, N * na * sizeof(T), cudaMemcpyHostToDevice, hs));
        a_dev = h->h_actions;
    }
    // page-locked output buffers: the kernel writes them in place over PCIe (measured 245 us per step of 4096
    // envs against 260 us with the device staging buffers and four copies)
    if (h->host_zero_copy >= 2 && obs && reward && done && terms) {
        void* dv[4] = {nullptr, nullptr, nullptr, nullptr};
        void* hv[4] = {obs, reward, (void*)done, terms};
        bool all = true;
        for (int k = 0; k < 4 && all; k++) {
            cudaPointerAttributes at;
            if (cudaPointerGetAttributes(&at, hv[k]) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
                dv[k] = at.devicePointer;
            else { cudaGetLastError(); all = false; }
        }
        if (all) {
            int rc2 = step_impl<T>(h, a_dev, dv[0], dv[1], (uint8_t*)dv[2], dv[3], hs);
            if (rc2) return rc2;
            CU(cudaStreamSynchronize(hs));
            return 0;
        }
    }
    int rc = step_impl<T>(h, a_dev, h->h_obs, h->h_reward, h->h_done, h->h_terms, hs);
    if (rc) return rc;
    const bool small = reward || done || terms;
    if (small) {
        CU(cudaEventRecord(h->ev_step, hs));
        CU(cudaStreamWaitEvent(h->side, h->ev_step, 0));
    }
    if (obs) CU(cudaMemcpyAsync(obs, h->h_obs, N * od * sizeof(T), cudaMemcpyDeviceToHost, hs));
    if (reward) CU(cudaMemcpyAsync(reward, h->h_reward, N * sizeof(T), cudaMemcpyDeviceToHost, h->side));
    if (done) CU(cudaMemcpyAsync(done, h->h_done, N, cudaMemcpyDeviceToHost, h->side));
    if (terms) CU(cudaMemcpyAsync(terms, h->h_terms, N * nt * sizeof(T), cudaMemcpyDeviceToHost, h->side));
    CU(cudaStreamSynchronize(hs));
    if (small) CU(cudaStreamSynchronize(h->side));
    return 0;
}

// bio_step_host_begin: the zero-copy launch of step_host_impl without the wait
template <typename T>
int step_host_begin_impl(Handle<T>* h, const void* actions, void* obs, void* reward, uint8_t* done, void* terms) {
    void* dv[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    const void* hv[5] = {actions, obs, reward, (const void*)done, terms};
    for (int k = 0; k < 5; k++) {
        cudaPointerAttributes at;
        if (hv[k] && cudaPointerGetAttributes(&at, hv[k]) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer)
            dv[k] = at.devicePointer;
        else { cudaGetLastError(); return fail(-1, "bio_step_host_begin: every buffer must be page-locked host memory"); }
    }
    return step_impl<T>(h, dv[0], dv[1], dv[2], (uint8_t*)dv[3], dv[4], h->host_stream ? h->host_stream : h->side);
}

template <typename T>
int reset_host_impl(Handle<T>* h, const uint8_t* mask, void* obs) {
    const size_t N = h->n;
    cudaStream_t hs = h->host_stream;
    if (mask) CU(cudaMemcpyAsync(h->h_mask, mask, N, cudaMemcpyHostToDevice, hs));
    int rc = reset_impl<T>(h, mask ? h->h_mask : nullptr, obs ? h->h_obs : nullptr, hs, 1);
    if (rc) return rc;
    if (obs) CU(cudaMemcpyAsync(obs, h->h_obs, N * h->task.obs_dim * sizeof(T), cudaMemcpyDeviceToHost, hs));
    CU(cudaStreamSynchronize(hs));
    return 0;
}

// frees everything a handle owns (also the partially built handle of a failed bio_create)
void release(HandleBase* h) {
    for (void* p : h->allocs) cudaFree(p);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->ev_step) cudaEventDestroy(h->ev_step);
    delete h;
}

}  // namespace

#define DISPATCH(h, expr32, expr64)                                  \
    (((HandleBase*)(h))->precision == BIO_PREC_F32 ? (expr32) : (expr64))
#define H32(h) ((Handle<float>*)(HandleBase*)(h))
#define H64(h) ((Handle<double>*)(HandleBase*)(h))

extern "C" {

int bio_abi_version(void) { return BIO_ABI_VERSION; }
uint64_t bio_sizeof_model_tables(void) { return sizeof(BioModelTables); }
uint64_t bio_sizeof_task_config(void) { return sizeof(BioTaskConfig); }
const char* bio_last_error(void) { return g_err.c_str(); }

int bio_create(const BioModelTables* model, const BioTaskConfig* task, const BioRefTables* ref, int32_t n_envs,
               int32_t device, int32_t precision, uint64_t seed, int64_t env_offset, bio_handle* out) {
    if (!out) return fail(-1, "out handle is null");
    *out = nullptr;
    int rc = validate(model, task, ref, n_envs);
    if (rc) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(-3, "no CUDA device: this backend has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(-1, "device index out of range");
    if (precision == BIO_PREC_F32) {
        Handle<float>* h = new (std::nothrow) Handle<float>();
        if (!h) return fail(-4, "out of host memory");
        h->precision = precision;
        rc = create_impl<float>(model, task, ref, n_envs, device, seed, env_offset, h);
        if (rc) { release(h); return rc; }
        *out = (bio_handle)(HandleBase*)h;
    } else if (precision == BIO_PREC_F64) {
        Handle<double>* h = new (std::nothrow) Handle<double>();
        if (!h) return fail(-4, "out of host memory");
        h->precision = precision;
        rc = create_impl<double>(model, task, ref, n_envs, device, seed, env_offset, h);
        if (rc) { release(h); return rc; }
        *out = (bio_handle)(HandleBase*)h;
    } else {
        return fail(-1, "unknown precision");
    }
    return 0;
}

int bio_destroy(bio_handle hh) {
    if (!hh) return 0;
    HandleBase* h = (HandleBase*)hh;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    release(h);
    return 0;
}

#define ENTER(hh)                                         \
    if (!(hh)) return fail(-1, "null handle");            \
    CU(cudaSetDevice(((HandleBase*)(hh))->device))

int bio_reset(bio_handle hh, const uint8_t* mask, void* obs, void* stream) {
    ENTER(hh);
    return DISPATCH(hh, reset_impl<float>(H32(hh), mask, obs, (cudaStream_t)stream, 1),
                    reset_impl<double>(H64(hh), mask, obs, (cudaStream_t)stream, 1));
}

int bio_step(bio_handle hh, const void* actions, void* obs, void* reward, uint8_t* done, void* reward_terms,
             void* stream) {
    ENTER(hh);
    if (!actions || !obs || !reward || !done) return fail(-1, "bio_step: null buffer");
    return DISPATCH(hh, step_impl<float>(H32(hh), actions, obs, reward, done, reward_terms, (cudaStream_t)stream),
                    step_impl<double>(H64(hh), actions, obs, reward, done, reward_terms, (cudaStream_t)stream));
}

int bio_step_host(bio_handle hh, const void* actions, void* obs, void* reward, uint8_t* done, void* reward_terms) {
    ENTER(hh);
    if (!actions) return fail(-1, "bio_step_host: null actions");
    return DISPATCH(hh, step_host_impl<float>(H32(hh), actions, obs, reward, done, reward_terms),
                    step_host_impl<double>(H64(hh), actions, obs, reward, done, reward_terms));
}

int bio_step_host_begin(bio_handle hh, const void* actions, void* obs, void* reward, uint8_t* done, void* reward_terms) {
    ENTER(hh);
    return DISPATCH(hh, step_host_begin_impl<float>(H32(hh), actions, obs, reward, done, reward_terms),
                    step_host_begin_impl<double>(H64(hh), actions, obs, reward, done, reward_terms));
}

int bio_step_host_end(bio_handle hh) {
    ENTER(hh);
    HandleBase* b = (HandleBase*)hh;
    CU(cudaStreamSynchronize(b->host_stream ? b->host_stream : b->side));
    return 0;
}

// bio_groups_run: native send / recv loop over env groups (include/bio_b200.h).  Device pointers of the page-locked
// buffers are resolved once; a group is relaunched the moment its stream is idle (cudaStreamQuery, polled
// round-robin), so the host side of a step is one query and one launch.
int bio_groups_run(const bio_handle* handles, int32_t n_groups, const BioGroupBuffers* bufs, int64_t steps,
                   bio_policy_fn policy, void* user) {
    if (!handles || !bufs || n_groups < 1 || n_groups > 64) return fail(-1, "bio_groups_run: bad arguments");
    if (steps <= 0) return 0;
    struct Grp { HandleBase* h; cudaStream_t s; void* dv[5]; std::vector<void*> ring; int64_t next; bool flying; };
    std::vector<Grp> g((size_t)n_groups);
    auto dev_ptr = [](const void* host, void** out) {
        cudaPointerAttributes at;
        if (host && cudaPointerGetAttributes(&at, host) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer) {
            *out = at.devicePointer;
            return true;
        }
        cudaGetLastError();
        return false;
    };
    for (int k = 0; k < n_groups; k++) {
        if (!handles[k]) return fail(-1, "bio_groups_run: null handle");
        Grp& x = g[(size_t)k];
        x.h = (HandleBase*)handles[k];
        if (x.h->device != g[0].h->device) return fail(-1, "bio_groups_run: the groups must live on one device");
        x.s = x.h->host_stream ? x.h->host_stream : x.h->side;
        const void* hv[5] = {bufs[k].actions, bufs[k].obs, bufs[k].reward, (const void*)bufs[k].done, bufs[k].reward_terms};
        for (int j = 0; j < 5; j++)
            if (!dev_ptr(hv[j], &x.dv[j])) return fail(-1, "bio_groups_run: every buffer must be page-locked host memory");
        if (!policy && bufs[k].ring_len > 0) {
            if (!bufs[k].action_ring) return fail(-1, "bio_groups_run: ring_len > 0 without action_ring");
            x.ring.resize((size_t)bufs[k].ring_len);
            const void* const* ring = (const void* const*)bufs[k].action_ring;
            for (int j = 0; j < bufs[k].ring_len; j++)
                if (!dev_ptr(ring[j], &x.ring[(size_t)j]))
                    return fail(-1, "bio_groups_run: every action buffer must be page-locked host memory");
        }
        x.next = 0;
        x.flying = false;
    }
    CU(cudaSetDevice(g[0].h->device));
    auto launch = [&](int k) -> int {
        Grp& x = g[(size_t)k];
        if (policy) policy(user, k, x.next);
        const void* a = x.ring.empty() ? x.dv[0] : x.ring[(size_t)(x.next % (int64_t)x.ring.size())];
        const int rc = DISPATCH(handles[k],
                                step_impl<float>(H32(handles[k]), a, x.dv[1], x.dv[2], (uint8_t*)x.dv[3], x.dv[4], x.s),
                                step_impl<double>(H64(handles[k]), a, x.dv[1], x.dv[2], (uint8_t*)x.dv[3], x.dv[4], x.s));
        if (rc == 0) { x.next++; x.flying = true; }
        return rc;
    };
    auto drain = [&]() { for (auto& x : g) if (x.flying) { cudaStreamSynchronize(x.s); x.flying = false; } };
    // Start-up: groups launched together finish together, and their row tails then queue on PCIe one behind the
    // other with nothing to hide them (measured: 4 groups in lockstep 21.9 M env-steps/s, no better than one
    // synchronous batch).  So group 0 runs its first step alone to measure the cycle T, and group k starts k T / G
    // later; equal groups keep that spacing, and one group's tail and relaunch overlap the others' substeps.
    typedef std::chrono::steady_clock Clock;
    std::vector<Clock::time_point> start_at((size_t)n_groups, Clock::time_point::min());
    {
        const Clock::time_point t0 = Clock::now();
        int rc = launch(0);
        if (rc) return rc;
        if (n_groups > 1 && steps >= 4) {
            const cudaError_t e = cudaStreamSynchronize(g[0].s);
            g[0].flying = false;
            if (e != cudaSuccess) return fail(-2, std::string("bio_groups_run: ") + cudaGetErrorString(e));
            const Clock::time_point t1 = Clock::now();
            const Clock::duration cycle = t1 - t0;
            for (int k = 1; k < n_groups; k++) start_at[(size_t)k] = t1 + cycle * k / n_groups;
            rc = launch(0);
            if (rc) return rc;
        }
    }
    // Steady state: the spacing does not hold by itself -- a group that ends within a few microseconds of another
    // one shares the PCIe tail and the relaunch with it and stays with it, bunched groups cycle slower, and the
    // remaining ones run into the bunch (traced: four groups 37 us apart are in lockstep after ~12 steps, cycle 155
    // -> 187 us).  So launches of different groups are kept at least T / G apart, T = the shortest cycle any group
    // has shown over its last 16 steps (the uncontended cycle): a group that is ready early waits for its slot
    // (BIO_GROUPS_SPACE scales the slot, BIO_GROUPS_NO_SPACING=1 disables it: experiments).
    int left = n_groups;
    const bool trace = getenv("BIO_GROUPS_TRACE") != nullptr;
    const bool spaced = n_groups > 1 && steps >= 4 && !getenv("BIO_GROUPS_NO_SPACING");
    // (0.8 of T / G: measured 25.98 M env-steps/s with 4 groups of 1024 envs against 24.55 M at 1.0 -- the shortest
    // cycle underestimates the typical one by a few per cent, full slots make late groups wait -- and 21.9 M unspaced)
    const double space_frac = getenv("BIO_GROUPS_SPACE") ? atof(getenv("BIO_GROUPS_SPACE")) : 0.8;
    const Clock::time_point tt0 = Clock::now();
    int n_ev = 0;
    std::vector<Clock::time_point> launched_at((size_t)n_groups, Clock::now());
    std::vector<std::vector<double>> cyc((size_t)n_groups);       // last cycles of every group, us
    Clock::time_point last_launch = launched_at[0];
    double t_min = spaced ? std::chrono::duration<double, std::micro>(start_at[1] - launched_at[0]).count() * n_groups : 0.0;
    auto launch_at = [&](int k, Clock::time_point now) -> int {
        if (g[(size_t)k].next > 0) {
            auto& c = cyc[(size_t)k];
            c.push_back(std::chrono::duration<double, std::micro>(now - launched_at[(size_t)k]).count());
            if (c.size() > 16) c.erase(c.begin());
            double mn = 1e30;
            for (auto& v : cyc) for (double d : v) mn = d < mn ? d : mn;
            if (mn < 1e30) t_min = mn;
        }
        launched_at[(size_t)k] = now;
        last_launch = now;
        return launch(k);
    };
    while (left > 0) {
        for (int k = 0; k < n_groups; k++) {
            Grp& x = g[(size_t)k];
            if (!x.flying && x.next == 0) {               // not started yet
                const Clock::time_point now = Clock::now();
                if (now >= start_at[(size_t)k]) {
                    const int rc = launch_at(k, now);
                    if (rc) { drain(); return rc; }
                }
                continue;
            }
            if (!x.flying && x.next >= steps) continue;   // finished
            if (x.flying) {
                const cudaError_t q = cudaStreamQuery(x.s);
                if (q == cudaErrorNotReady) continue;
                if (q != cudaSuccess) {
                    drain();
                    return fail(-2, std::string("bio_groups_run: ") + cudaGetErrorString(q));
                }
                x.flying = false;
                if (trace && n_ev++ < 64)
                    fprintf(stderr, "g%d done at %.1f us (T %.1f)\n", k, std::chrono::duration<double, std::micro>(Clock::now() - tt0).count(), t_min);
                if (x.next >= steps) { left--; continue; }
            }
            // ready for its next step: wait for the slot
            const Clock::time_point now = Clock::now();
            if (spaced && std::chrono::duration<double, std::micro>(now - last_launch).count() < space_frac * t_min / n_groups) continue;
            const int rc = launch_at(k, now);
            if (rc) { drain(); return rc; }
        }
    }
    return 0;
}

int bio_set_grid(bio_handle hh, int32_t ctas) {
    ENTER(hh);
    ((HandleBase*)hh)->coop_grid = ctas > 0 ? ctas : 0;
    return DISPATCH(hh, pick_coop_shape<float>(H32(hh)), pick_coop_shape<double>(H64(hh)));
}

int bio_reset_host(bio_handle hh, const uint8_t* mask, void* obs) {
    ENTER(hh);
    return DISPATCH(hh, reset_host_impl<float>(H32(hh), mask, obs), reset_host_impl<double>(H64(hh), mask, obs));
}

int bio_get_state(bio_handle hh, const BioStatePtrs* dst, void* stream) {
    ENTER(hh);
    if (!dst) return fail(-1, "null state pointers");
    return DISPATCH(hh, state_impl<float>(H32(hh), dst, (cudaStream_t)stream, false),
                    state_impl<double>(H64(hh), dst, (cudaStream_t)stream, false));
}

int bio_set_state(bio_handle hh, const BioStatePtrs* src, void* stream) {
    ENTER(hh);
    if (!src) return fail(-1, "null state pointers");
    return DISPATCH(hh, state_impl<float>(H32(hh), src, (cudaStream_t)stream, true),
                    state_impl<double>(H64(hh), src, (cudaStream_t)stream, true));
}

int bio_eval_debug(bio_handle hh, const void* controls, const BioDebugPtrs* out, void* stream) {
    ENTER(hh);
    if (!out) return fail(-1, "null debug pointers");
    return DISPATCH(hh, eval_impl<float>(H32(hh), controls, out, (cudaStream_t)stream),
                    eval_impl<double>(H64(hh), controls, out, (cudaStream_t)stream));
}

int bio_set_step_extra(bio_handle hh, const BioStepExtra* extra) {
    ENTER(hh);
    return DISPATCH(hh, extra_impl<float>(H32(hh), extra), extra_impl<double>(H64(hh), extra));
}

int bio_set_host_stream(bio_handle hh, void* stream) {
    ENTER(hh);
    ((HandleBase*)hh)->host_stream = (cudaStream_t)stream;
    return 0;
}

int bio_id_apply(bio_handle hh, int32_t op, const void* x, const void* controls, const void* shift, void* out,
                 void* stream) {
    ENTER(hh);
    if (!x || !out) return fail(-1, "bio_id_apply: null buffer");
    if (op < BIO_ID_MULTIPLY_M || op > BIO_ID_SOLVE_SHIFTED) return fail(-1, "bio_id_apply: unknown operator");
    return DISPATCH(hh, id_impl<float>(H32(hh), op, x, controls, shift, out, (cudaStream_t)stream),
                    id_impl<double>(H64(hh), op, x, controls, shift, out, (cudaStream_t)stream));
}
int bio_id_multiply_m(bio_handle hh, const void* a, void* out, void* stream) {
    return bio_id_apply(hh, BIO_ID_MULTIPLY_M, a, nullptr, nullptr, out, stream);
}
int bio_id_multiply_minv(bio_handle hh, const void* tau, void* out, void* stream) {
    return bio_id_apply(hh, BIO_ID_MULTIPLY_MINV, tau, nullptr, nullptr, out, stream);
}
int bio_id_residual(bio_handle hh, const void* qddot, const void* controls, void* out, void* stream) {
    return bio_id_apply(hh, BIO_ID_RESIDUAL, qddot, controls, nullptr, out, stream);
}

int bio_stats(bio_handle hh, double* out16, int32_t reset_after, void* stream) {
    ENTER(hh);
    HandleBase* h = (HandleBase*)hh;
    if (!out16) return fail(-1, "null stats buffer");
    CU(cudaMemcpyAsync(out16, h->stats, 16 * sizeof(double), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (reset_after) CU(cudaMemsetAsync(h->stats, 0, 16 * sizeof(double), (cudaStream_t)stream));
    return 0;
}

int bio_kernel_shape(bio_handle hh, int32_t* size_class, int32_t* threads, int32_t* ctas_per_sm) {
    if (!hh) return fail(-1, "null handle");
    HandleBase* h = (HandleBase*)hh;
    const int cls = h->precision == BIO_PREC_F32 ? H32(hh)->coop_cls : H64(hh)->coop_cls;
    if (size_class) *size_class = cls;
    if (threads) *threads = cls >= 0 ? h->coop_threads : h->block;
    if (ctas_per_sm) *ctas_per_sm = cls >= 0 ? h->coop_ctas : 0;
    return 0;
}

int64_t bio_launch_count(bio_handle hh) { return hh ? ((HandleBase*)hh)->launches : 0; }
int32_t bio_obs_dim(bio_handle hh) { return hh ? ((HandleBase*)hh)->task.obs_dim : 0; }
int32_t bio_n_act(bio_handle hh) { return hh ? ((HandleBase*)hh)->model.n_act : 0; }

}  // extern "C"

// ---------------------------------------------------------------------------
// FP32 FMA-chain microbenchmark: the roofline denominator of this compute-bound
// path (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
// ---------------------------------------------------------------------------
namespace {
template <typename F>
__global__ void __launch_bounds__(256) fma_peak_kernel(F* out, int iters, F a, F b) {
    F x0 = threadIdx.x * F(1e-3), x1 = x0 + F(1), x2 = x0 + F(2), x3 = x0 + F(3);
    F x4 = x0 + F(4), x5 = x0 + F(5), x6 = x0 + F(6), x7 = x0 + F(7);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

template <typename F>
double measure_fma_peak(int32_t device, int iters) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int blocks = sms * 8, threads = 256;
    F* out = nullptr;
    if (cudaMalloc(&out, (size_t)blocks * threads * sizeof(F)) != cudaSuccess) return -1.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        fma_peak_kernel<F><<<blocks, threads>>>(out, iters, F(0.999), F(0.001));
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    return cudaGetLastError() == cudaSuccess ? best : -1.0;
}
}  // namespace

extern "C" double bio_measure_fp32_peak(int32_t device) { return measure_fma_peak<float>(device, 4096); }
// the same chain in fp64 (DFMA), for the fp64 build's roofline
extern "C" double bio_measure_fp64_peak(int32_t device) { return measure_fma_peak<double>(device, 256); }
