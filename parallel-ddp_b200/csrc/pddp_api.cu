// pddp_api.cu -- host side of libpddp.so: the C-ABI declared in include/pddp.h over the kernels in kernels.cuh.
// The host code is C++ (the reference is compiled C++/CUDA); there is no CPU fallback: every entry point fails
// with PDDP_E_NODEVICE / PDDP_E_CUDA when no device can run the kernels.
#include "../../include/pddp.h"
#include "../../include/pddp_plant.h"
#include "kernels.cuh"
#include "kuka_model_data.inc"

#include <cmath>
#include <nccl.h>
#include <dlfcn.h>
#include <mutex>
#include <cstdio>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

using namespace pddp;

static thread_local std::string g_create_error;

struct pddp_solver {
    pddp_config cfg;
    DevState S;
    cudaStream_t stream = nullptr;          // stream 0: setup, hand-over of results, and problem group 0
    std::vector<cudaStream_t> gstreams;     // one stream per problem group (gstreams[0] == stream)
    std::vector<cudaEvent_t> gev;           // per-group events (fork / join)
    int groups = 1;
    std::vector<void*> allocs;
    std::string err;
    const pddp_plant_ops *ops = nullptr;    // plug-in plant (PLANT 1-3 and registered ones); null = the Kuka kernels of kernels.cuh
    float *w_KT = nullptr, *w_P = nullptr, *w_p = nullptr, *w_d = nullptr;     // warm-start inputs (pddp_set_warm_start)
    bool skip_env = false;
    int next_clear = 1, next_rollout = 0;                                      // loadVarsGPU flags of the next solve
    MpcState mpc{}; int *d_mpc_flags = nullptr; float *d_xActual = nullptr; bool mpc_ready = false;   // receding-horizon state (pddp_mpc_*)
    float *d_mpc_pack = nullptr, *h_mpc_pack = nullptr; size_t mpc_pack_bytes = 0;     // what a step hands back, packed for one copy (mpc_store_kernel)
    std::vector<int> mpc_lss;                                                  // last_successful_solve per problem (MPCHelpers.cuh:63)
    size_t smem_mpc = 0;
    long launches = 0;
    int n, m, num_sms = 148;
    float *d_xout = nullptr, *d_uout = nullptr; int *d_iters = nullptr;
    float *h_stage = nullptr; size_t h_stage_bytes = 0;    // pinned staging
    int *h_nactive = nullptr;
    cudaEvent_t ev[8];
    double last_ms = 0; int last_launches = 0;
    struct GraphEntry { cudaGraphExec_t exec; long kernels; };
    std::map<std::string, GraphEntry> graphs;                                  // captured iteration chunks, keyed by (device state, iterations, groups)
    bool use_graphs = true; int graph_chunk = 10; long graph_launches = 0;
    std::vector<double> it_ms[4];                                              // per-iteration device times of the last timed solve: sim(+selection), sweep, bp, nis
    size_t smem_bp, smem_sweep, smem_sim, smem_sel, smem_nis, smem_udyn, smem_ugrad;
    float *d_xTarget = nullptr;
    int *d_cost_shift = nullptr; bool use_cost_shift = false;
    int sim_lanes = 16;
    // step-size sharding (pddp_alpha_shard_init): NCCL is bound at run time (dlopen), so single-GPU users do not need it
    struct Nccl {
        void *lib = nullptr; ncclComm_t comm = nullptr; int rank = 0, nranks = 1;
        ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
        ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
        ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
        const char *(*GetErrorString)(ncclResult_t) = nullptr;
    } nccl;
    cudaEvent_t xev[4] = {nullptr, nullptr, nullptr, nullptr}; double xchg_ms = 0; long xchg_calls = 0;      // device time spent in the two collectives
    int bp_shape = 0;                                                          // backward pass (env PDDP_BP_SHAPE): 0 = by launch size, 1 = warp chains (bp_warp.cuh), 2 = block-cooperative (kernels.cuh)                                                        // lanes per simulated trajectory (16: throughput shape, 32: latency shape)
    std::map<std::string, std::pair<void*, size_t>> arrays;
};

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess){ h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return PDDP_E_CUDA; } } while (0)

extern "C" void pddp_default_config_kuka(pddp_config *c, int N, int batch){
    std::memset(c, 0, sizeof(*c));
    c->plant = PDDP_PLANT_KUKA; c->N = N; c->n_alpha = 16; c->M = 4; c->max_iter = 100; c->batch = batch; c->device = 0;
    c->integrator = 1; c->alpha_base = 0.5f; c->total_time = 0.5f;
    c->rho_init = (float)12.5; c->rho_min = (float)0.01; c->rho_max = (float)10000000.0; c->rho_factor = (float)1.25;
    c->exp_red_min = (float)0.05; c->exp_red_max = (float)1.25; c->max_defect = (float)1.0; c->tol_cost = 0.0f;
    c->Q1 = (float)0.1; c->Q2 = (float)0.001; c->R = (float)0.0001; c->QF1 = (float)1000.0; c->QF2 = (float)1000.0;
    c->gravity = KUKA_GRAV;
    c->ee_cost = 0;                                                    // plants/cost_arm.cuh:106-117 defaults
    c->Q_EE1 = (float)0.1; c->Q_EE2 = 0.f; c->QF_EE1 = (float)1000.0; c->QF_EE2 = 0.f; c->R_EE = (float)0.0001;
    c->Q_xdEE = (float)0.1; c->QF_xdEE = (float)1000.0; c->Q_xEE = 0.f; c->QF_xEE = 0.f;
    c->use_limits = 0; c->lim_Q_pos = (float)100.0; c->lim_Q_vel = (float)100.0; c->lim_R_tau = (float)100.0;    // plants/cost_arm.cuh:26-30
    c->use_smooth_abs = 0; c->smooth_abs_alpha = 0.2;                                                   // plants/cost_arm.cuh:119-121
}

// ---------------------------------------------------------------------------------------------------- plant plug-ins
// Pendulum, cart-pole and quadrotor are plant translation units (plant_tu.cu) linked into this library; further plants arrive
// through pddp_register_plant / pddp_load_plant_library.  The Kuka arm (PLANT 4) keeps its own kernels (kernels.cuh).
extern "C" const pddp_plant_ops *pddp_plant_entry_1(void);
extern "C" const pddp_plant_ops *pddp_plant_entry_2(void);
extern "C" const pddp_plant_ops *pddp_plant_entry_3(void);
namespace {
constexpr int MAX_PLANTS = 64;
const pddp_plant_ops *g_plants[MAX_PLANTS] = {nullptr};
std::mutex g_plants_mutex;
std::string g_plant_error;
bool plant_ops_ok(const pddp_plant_ops *o){
    return o && o->abi == PDDP_PLANT_ABI && o->state_size_bytes == sizeof(DevState) && o->mpc_size_bytes == sizeof(MpcState) &&
           o->plant_id >= 1 && o->plant_id < MAX_PLANTS && o->plant_id != PDDP_PLANT_KUKA && o->state_size == 2*o->num_pos &&
           o->state_size <= 32 && o->control_size >= 1 && o->control_size <= 32 && o->launch_bp && o->launch_sim && o->launch_nis;
}
const pddp_plant_ops *find_plant(int id){
    std::lock_guard<std::mutex> lk(g_plants_mutex);
    if (!g_plants[1]){ g_plants[1] = pddp_plant_entry_1(); g_plants[2] = pddp_plant_entry_2(); g_plants[3] = pddp_plant_entry_3(); }
    return (id >= 1 && id < MAX_PLANTS) ? g_plants[id] : nullptr;
}
}
extern "C" int pddp_register_plant(const pddp_plant_ops *ops){
    if (!plant_ops_ok(ops)){ g_plant_error = "plant table rejected: ABI / struct size / dimensions do not match this libpddp"; return PDDP_E_INVALID; }
    find_plant(0);
    std::lock_guard<std::mutex> lk(g_plants_mutex);
    g_plants[ops->plant_id] = ops;
    return 0;
}
extern "C" int pddp_load_plant_library(const char *path){
    if (!path){ return PDDP_E_INVALID; }
    void *lib = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!lib){ g_plant_error = std::string("dlopen: ") + dlerror(); return PDDP_E_INVALID; }
    pddp_plant_entry_fn fn = reinterpret_cast<pddp_plant_entry_fn>(dlsym(lib, "pddp_plant_entry"));
    if (!fn){ g_plant_error = "the library does not export pddp_plant_entry"; dlclose(lib); return PDDP_E_INVALID; }
    const int rc = pddp_register_plant(fn());
    return rc ? rc : fn()->plant_id;
}
extern "C" const char *pddp_plant_error(void){ return g_plant_error.c_str(); }
extern "C" int pddp_plant_dims(int plant, int *num_pos, int *state_size, int *control_size){
    int np = 7, n = 14, m = 7;
    if (plant != PDDP_PLANT_KUKA){ const pddp_plant_ops *o = find_plant(plant); if (!o){ return PDDP_E_INVALID; } np = o->num_pos; n = o->state_size; m = o->control_size; }
    if (num_pos){ *num_pos = np; } if (state_size){ *state_size = n; } if (control_size){ *control_size = m; }
    return 0;
}

// config.cuh:21-136 per PLANT, with the cost weights of plants/cost_{pend,cart,quad}.cuh (the pendulum's velocity weight falls through
// QR(i) to R = 0.1; the cart-pole has its own set for N = 512) and the WAFR example's for the arm
extern "C" int pddp_default_config(pddp_config *c, int plant, int N, int batch){
    if (!c){ return PDDP_E_INVALID; }
    if (plant == PDDP_PLANT_KUKA){ pddp_default_config_kuka(c, N, batch); return 0; }
    if (plant < PDDP_PLANT_PEND || plant > PDDP_PLANT_QUAD){ return PDDP_E_INVALID; }
    std::memset(c, 0, sizeof(*c));
    c->plant = plant; c->N = N; c->M = 4; c->max_iter = 100; c->batch = batch; c->device = 0; c->integrator = 3;
    c->n_alpha = 32; c->alpha_base = 0.75f; c->total_time = 4.0f;
    c->rho_min = (float)0.01; c->rho_max = (float)10000000.0; c->rho_factor = (float)1.25;
    c->exp_red_min = (float)0.05; c->exp_red_max = (float)1.25; c->max_defect = (float)1.0; c->tol_cost = 0.0f;
    if (plant == PDDP_PLANT_PEND){ c->rho_init = (float)10.0; c->Q1 = (float)1.0; c->Q2 = (float)0.1; c->R = (float)0.1; c->QF1 = c->QF2 = (float)1000.0; }
    else if (plant == PDDP_PLANT_CART){
        c->rho_init = (float)10.0; c->max_defect = (float)0.75;
        if (N == 512){ c->Q1 = (float)0.01; c->Q2 = (float)0.01; c->R = (float)0.001; c->QF1 = c->QF2 = (float)100000.0; }
        else { c->Q1 = (float)0.01; c->Q2 = (float)0.001; c->R = (float)0.0001; c->QF1 = c->QF2 = (float)1000.0; }
    } else { c->rho_init = (float)1.0; c->n_alpha = 16; c->alpha_base = 0.5f; c->Q1 = (float)0.01; c->Q2 = (float)0.001; c->R = (float)5.0; c->QF1 = c->QF2 = (float)1000.0; }
    return 0;
}

extern "C" const char *pddp_last_error(pddp_handle h){ return h ? h->err.c_str() : g_create_error.c_str(); }

template <typename T>
static int dalloc(pddp_handle h, T **p, size_t count, const char *name = nullptr, bool zero = true){
    CK(cudaMalloc((void**)p, count*sizeof(T)));
    h->allocs.push_back(*p);
    if (zero){ CK(cudaMemset(*p, 0, count*sizeof(T))); }
    if (name){ h->arrays[name] = {(void*)*p, count*sizeof(T)}; }
    return 0;
}

extern "C" int pddp_create(const pddp_config *cfg, pddp_handle *out){
    *out = nullptr;
    auto fail = [&](const std::string &msg, int code){ g_create_error = msg; return code; };
    if (!cfg){ return fail("null config", PDDP_E_INVALID); }
    const pddp_plant_ops *ops = (cfg->plant == PDDP_PLANT_KUKA) ? nullptr : find_plant(cfg->plant);
    if (cfg->plant != PDDP_PLANT_KUKA && !ops){ return fail("unknown PLANT: 1 pendulum, 2 cart-pole, 3 quadrotor, 4 Kuka iiwa14, or one registered with pddp_register_plant", PDDP_E_INVALID); }
    if (!ops && cfg->integrator != 1){ return fail("the Kuka kernels are built for the Euler integrator (INTEGRATOR 1, config.cuh:58)", PDDP_E_INVALID); }
    if (ops && (cfg->integrator < 1 || cfg->integrator > 3)){ return fail("INTEGRATOR must be 1 (Euler), 2 (Midpoint) or 3 (RK3)", PDDP_E_INVALID); }
    if (ops && cfg->ee_cost){ return fail("EE_COST needs an end effector: PLANT 4 only (cost_pend.cuh:9-10)", PDDP_E_INVALID); }
    if (cfg->N < 32 || cfg->N > 1024 || (cfg->N & (cfg->N-1))){ return fail("N must be a power of two in [32,1024]", PDDP_E_INVALID); }
    if (cfg->M < 1 || cfg->M > 8 || cfg->N % cfg->M){ return fail("M must divide N and be <= 8", PDDP_E_INVALID); }
    if (cfg->n_alpha < 1 || cfg->n_alpha > PDDP_MAX_ALPHA){ return fail("n_alpha out of range", PDDP_E_INVALID); }
    if (cfg->batch < 1 || cfg->max_iter < 1){ return fail("batch and max_iter must be positive", PDDP_E_INVALID); }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0){ cudaGetLastError(); return fail("no CUDA device (this library has no CPU path)", PDDP_E_NODEVICE); }
    if (cfg->device < 0 || cfg->device >= ndev){ return fail("bad device ordinal", PDDP_E_INVALID); }
    pddp_handle h = new pddp_solver();
    for (auto &e : h->ev){ e = nullptr; }
    h->cfg = *cfg; h->ops = ops; h->n = ops ? ops->state_size : kuka::NX; h->m = ops ? ops->control_size : kuka::NU;
    auto bail = [&](int code){ g_create_error = h->err; pddp_destroy(h); return code; };      // the same cleanup as a live handle: streams, events, pinned and device memory
    #define CKC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess){ h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return bail(PDDP_E_CUDA); } } while (0)
    CKC(cudaSetDevice(cfg->device));
    { int sms = 0; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device) == cudaSuccess && sms > 0){ h->num_sms = sms; } }
    CKC(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    for (auto &e : h->ev){ CKC(cudaEventCreate(&e)); }
    h->gstreams.push_back(h->stream);
    for (int g = 1; g < 8; g++){ cudaStream_t st; CKC(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); h->gstreams.push_back(st); }
    for (int g = 0; g < 9; g++){ cudaEvent_t e; CKC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); h->gev.push_back(e); }
    { const char *env = std::getenv("PDDP_SKIP_UNCHANGED"); h->skip_env = env && std::atoi(env) != 0; }
    { const char *env = std::getenv("PDDP_GRAPHS"); h->use_graphs = !(env && std::atoi(env) == 0); }
    { const char *env = std::getenv("PDDP_GRAPH_CHUNK"); const int v = env ? std::atoi(env) : 0; if (v >= 1 && v <= 1000){ h->graph_chunk = v; } }
    { const char *env = std::getenv("PDDP_GROUPS"); int g = env ? std::atoi(env) : 2; h->groups = (g >= 1 && g <= 8 && cfg->batch >= 2*g) ? g : 1; }
    DevState &S = h->S; std::memset(&S, 0, sizeof(S)); S.skip_unchanged = h->skip_env ? 1 : 0;
    const int B = cfg->batch, N = cfg->N, A = cfg->n_alpha, M = cfg->M, n = h->n, m = h->m;
    S.B = B; S.N = N; S.A = A; S.M = M; S.n = n; S.m = m; S.max_iter = cfg->max_iter; S.iter_cap = cfg->max_iter;
    S.npos = ops ? ops->num_pos : kuka::NB; S.integrator = cfg->integrator;
    S.ab_stride = ops ? n*(n+m) : AB_STRIDE; S.h_stride = ops ? (n+m)*(n+m) : H_STRIDE; S.g_stride = ops ? (n+m) : G_STRIDE;
    S.dt = (float)((double)cfg->total_time/(double)(N-1));          // (T)TIME_STEP, config.cuh:136
    S.tol_cost = cfg->tol_cost; S.two_tol = (float)(2*(double)cfg->tol_cost);
    S.rho_min = cfg->rho_min; S.rho_max = cfg->rho_max; S.rho_factor = cfg->rho_factor; S.inv_rho_factor = (float)(1.0/(double)cfg->rho_factor);
    S.exp_red_min = cfg->exp_red_min; S.exp_red_max = cfg->exp_red_max; S.max_defect = cfg->max_defect;
    S.Q1 = cfg->Q1; S.Q2 = cfg->Q2; S.R = cfg->R; S.QF1 = cfg->QF1; S.QF2 = cfg->QF2; S.grav = cfg->gravity;
    S.ee = cfg->ee_cost ? 1 : 0;
    S.use_limits = (cfg->use_limits && cfg->plant == PDDP_PLANT_KUKA) ? 1 : 0; S.Q_PL = cfg->lim_Q_pos; S.Q_VL = cfg->lim_Q_vel; S.R_TL = cfg->lim_R_tau;
    S.smooth_abs = (cfg->use_smooth_abs && S.ee) ? 1 : 0; S.sa_alpha = (float)cfg->smooth_abs_alpha; S.sa_alpha2 = (float)(cfg->smooth_abs_alpha*cfg->smooth_abs_alpha);
    S.a_first = 0; S.a_cnt = A;
    S.Q_EE1 = cfg->Q_EE1; S.Q_EE2 = cfg->Q_EE2; S.QF_EE1 = cfg->QF_EE1; S.QF_EE2 = cfg->QF_EE2; S.R_EE = cfg->R_EE;
    S.Q_xdEE = cfg->Q_xdEE; S.QF_xdEE = cfg->QF_xdEE; S.Q_xEE = cfg->Q_xEE; S.QF_xEE = cfg->QF_xEE;
    float *dI, *dTb, *dal;
    #define DA(ptr, count, name) do { if (dalloc(h, &(ptr), (size_t)(count), name)){ return bail(PDDP_E_CUDA); } } while (0)
    const size_t model_floats = ops ? (size_t)36*ops->num_pos : 252;
    DA(dI, model_floats, nullptr); DA(dTb, model_floats, nullptr); DA(dal, PDDP_MAX_ALPHA, nullptr);
    std::vector<float> al(PDDP_MAX_ALPHA, 0.f);
    for (int i = 0; i < A; i++){ al[i] = (float)std::pow((double)cfg->alpha_base, i); }    // nisInitHelpers.cuh:829
    if (ops){
        std::vector<float> hI(model_floats, 0.f), hT(model_floats, 0.f);
        if (ops->init_model){ ops->init_model(hI.data(), hT.data()); }                   // initI / initT of the plant header
        CKC(cudaMemcpy(dI, hI.data(), model_floats*4, cudaMemcpyHostToDevice)); CKC(cudaMemcpy(dTb, hT.data(), model_floats*4, cudaMemcpyHostToDevice));
    } else {
        CKC(cudaMemcpy(dI, KUKA_I_DATA, sizeof(KUKA_I_DATA), cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(dTb, KUKA_TBODY_DATA, sizeof(KUKA_TBODY_DATA), cudaMemcpyHostToDevice));
    }
    CKC(cudaMemcpy(dal, al.data(), al.size()*sizeof(float), cudaMemcpyHostToDevice));
    S.I = dI; S.Tbody = dTb; S.alpha = dal;
    DA(S.x, (size_t)B*A*N*n, "x"); DA(S.u, (size_t)B*A*N*m, "u"); DA(S.d, (size_t)B*A*N*n, "d");
    DA(S.xp, (size_t)B*N*n, "xp"); DA(S.xp2, (size_t)B*N*n, "xp2"); DA(S.up, (size_t)B*N*m, "up"); DA(S.dp, (size_t)B*N*n, "dp");
    DA(S.AB, (size_t)B*N*S.ab_stride, nullptr); DA(S.H, (size_t)B*N*S.h_stride, nullptr); DA(S.g, (size_t)B*N*S.g_stride, nullptr);
    DA(S.Pbuf[0], (size_t)B*N*n*n, nullptr); DA(S.Pbuf[1], (size_t)B*N*n*n, nullptr); DA(S.pbuf[0], (size_t)B*N*n, nullptr); DA(S.pbuf[1], (size_t)B*N*n, nullptr);
    DA(S.KT, (size_t)B*N*n*m, "KT"); DA(S.du, (size_t)B*N*m, "du"); DA(S.ApBK, (size_t)B*N*n*n, "ApBK"); DA(S.Bdu, (size_t)B*N*n, "Bdu");
    DA(S.xGoal, (size_t)B*n, "xGoal"); DA(S.costk, (size_t)B*A*N, "costk");
    DA(S.J, (size_t)B*A, "J"); DA(S.dT, (size_t)B*A, "dT"); DA(S.dJexp, (size_t)B*2*M, "dJexp");
    DA(S.rho, B, "rho"); DA(S.drho, B, "drho"); DA(S.prevJ, B, "prevJ"); DA(S.dJ, B, "dJ"); DA(S.z, B, "z");
    DA(S.init_knot, B, nullptr);
    DA(S.iter, B, "iter"); DA(S.alphaIndex, B, "alphaIndex"); DA(S.ignore_defect, B, "ignore_defect"); DA(S.done, B, "done");
    DA(S.accepted, B, "accepted"); DA(S.final_src, B, "final_src");
    DA(S.Jout, (size_t)B*(cfg->max_iter+1), "Jout"); DA(S.alphaOut, (size_t)B*(cfg->max_iter+1), "alphaOut");
    DA(S.n_active, 1, nullptr);
    DA(S.dbg, 4096, "dbg");
    DA(h->d_xout, (size_t)B*N*n, nullptr); DA(h->d_uout, (size_t)B*N*m, nullptr); DA(h->d_iters, B, nullptr);
    h->h_stage_bytes = ((size_t)B*N*(n+m) + (size_t)B*n + (size_t)2*B*(cfg->max_iter+1) + B)*sizeof(float);
    CKC(cudaMallocHost((void**)&h->h_stage, h->h_stage_bytes));
    CKC(cudaMallocHost((void**)&h->h_nactive, sizeof(int)));
    // dynamic shared memory of each kernel
    h->smem_sel = ((size_t)A*N + 2*A)*sizeof(float);
    CKC(cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sel));
    if (ops){
        if (ops->prepare){ const int rc = ops->prepare(M); if (rc){ h->err = std::string("plant prepare: ") + cudaGetErrorString((cudaError_t)rc); return bail(PDDP_E_CUDA); } }
    } else {
        h->smem_bp = sizeof(BpSmem<kuka::NX, kuka::NU>);
        h->smem_sweep = SWEEP_SLOTS*sizeof(SweepSlot<kuka::NX>);
        h->smem_sim = (36*kuka::NB + 16)*sizeof(float) + (size_t)(32/SIM_LANES)*sizeof(SimGroupSmem);
        h->smem_nis = NIS_CONST_FLOATS*sizeof(float) + NIS_WARPS*(32/NIS_LANES)*sizeof(NisGroupSmem);
        h->smem_udyn = 2*36*kuka::NB*sizeof(float) + (32/SIM_LANES)*sizeof(SimGroupSmem);
        h->smem_ugrad = 2*36*kuka::NB*sizeof(float) + (32/NIS_LANES)*sizeof(NisGroupSmem);
        CKC(cudaFuncSetAttribute(bp_kernel<kuka::NX, kuka::NU>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_bp));
        CKC(cudaFuncSetAttribute(bp_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BPW_WARPS*sizeof(BpWarpSmem))));
        { const char *env = std::getenv("PDDP_BP_SHAPE"); h->bp_shape = env ? std::atoi(env) : 0; }
        CKC(cudaFuncSetAttribute(sweep_kernel<kuka::NX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sweep));
        CKC(cudaFuncSetAttribute(sim_kernel<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sim));
        CKC(cudaFuncSetAttribute(sim_kernel<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sim));
        CKC(cudaFuncSetAttribute(sim_kernel<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sim));
        CKC(cudaFuncSetAttribute(sim_kernel<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_sim));
        // latency shape of the simulation when one warp per (interval, candidate) of the WHOLE batch still leaves a scheduler per warp
        { const char *env = std::getenv("PDDP_SIM_LANES"); const int v = env ? std::atoi(env) : 0;
          h->sim_lanes = (v == 16 || v == 32) ? v : (((long long)B*A*M <= 4LL*h->num_sms) ? 32 : 16); }
        CKC(cudaFuncSetAttribute(nis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_nis));
        CKC(cudaFuncSetAttribute(unit_dynamics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_udyn));
        CKC(cudaFuncSetAttribute(unit_gradient_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_ugrad));
    }
    CKC(cudaDeviceSynchronize());
    *out = h;
    return 0;
}

extern "C" void pddp_destroy(pddp_handle h){
    if (!h){ return; }
    cudaSetDevice(h->cfg.device);
    if (h->stream){ cudaStreamSynchronize(h->stream); }
    for (auto &kv : h->graphs){ cudaGraphExecDestroy(kv.second.exec); }
    if (h->nccl.comm && h->nccl.CommDestroy){ h->nccl.CommDestroy(h->nccl.comm); }
    for (auto &e : h->xev){ if (e){ cudaEventDestroy(e); } }
    for (void *p : h->allocs){ cudaFree(p); }
    if (h->h_stage){ cudaFreeHost(h->h_stage); } if (h->h_nactive){ cudaFreeHost(h->h_nactive); } if (h->h_mpc_pack){ cudaFreeHost(h->h_mpc_pack); }
    for (auto &e : h->ev){ if (e){ cudaEventDestroy(e); } }
    for (auto &e : h->gev){ cudaEventDestroy(e); }
    for (auto st : h->gstreams){ cudaStreamDestroy(st); }
    delete h;
}

// ---------------------------------------------------------------------------------------------------- launches
__global__ void reset_kernel(DevState S, float rho_init, int ignore_first){
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i <= S.max_iter; i += blockDim.x){ S.Jout[(size_t)b*(S.max_iter+1)+i] = nanf(""); S.alphaOut[(size_t)b*(S.max_iter+1)+i] = -99; }
    if (threadIdx.x == 0){
        S.rho[b] = rho_init; S.drho[b] = 1.0f; S.iter[b] = 1; S.alphaIndex[b] = 0; S.ignore_defect[b] = ignore_first; S.done[b] = 0;
        S.accepted[b] = 0; S.final_src[b] = -1; S.dJ[b] = 0.f; S.z[b] = 0.f; S.prevJ[b] = 0.f;
        if (b == 0){ *S.n_active = S.B; }
    }
}

static int launch_reset(pddp_handle h, int ignore_first, int clear){
    DevState &S = h->S; const int B = S.B, N = S.N, A = S.A, n = S.n, m = S.m;
    CK(cudaMemsetAsync(S.init_knot, 0, (size_t)B*sizeof(int), h->stream));       // runiLQR_GPU starts from alphaIndex = 0
    if (clear){
        // loadVarsGPU with clearVarsFlag=1 (nisInitHelpers.cuh:612-620)
        CK(cudaMemsetAsync(S.Pbuf[0], 0, (size_t)B*N*n*n*4, h->stream)); CK(cudaMemsetAsync(S.Pbuf[1], 0, (size_t)B*N*n*n*4, h->stream));
        CK(cudaMemsetAsync(S.pbuf[0], 0, (size_t)B*N*n*4, h->stream)); CK(cudaMemsetAsync(S.pbuf[1], 0, (size_t)B*N*n*4, h->stream));
        CK(cudaMemsetAsync(S.KT, 0, (size_t)B*N*n*m*4, h->stream));
        CK(cudaMemsetAsync(S.d, 0, (size_t)B*A*N*n*4, h->stream)); CK(cudaMemsetAsync(S.dp, 0, (size_t)B*N*n*4, h->stream));
    } else {
        // clearVarsFlag=0 (:622-631): P, Pp <- P0; p, pp <- p0; KT <- KT0; d <- d0 (every candidate is based on dp here)
        if (!h->w_KT){ h->err = "clearVarsFlag=0 needs pddp_set_warm_start first"; return PDDP_E_INVALID; }
        CK(cudaMemcpyAsync(S.Pbuf[0], h->w_P, (size_t)B*N*n*n*4, cudaMemcpyDeviceToDevice, h->stream)); CK(cudaMemcpyAsync(S.Pbuf[1], h->w_P, (size_t)B*N*n*n*4, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(S.pbuf[0], h->w_p, (size_t)B*N*n*4, cudaMemcpyDeviceToDevice, h->stream)); CK(cudaMemcpyAsync(S.pbuf[1], h->w_p, (size_t)B*N*n*4, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(S.KT, h->w_KT, (size_t)B*N*n*m*4, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpyAsync(S.dp, h->w_d, (size_t)B*N*n*4, cudaMemcpyDeviceToDevice, h->stream));
        CK(cudaMemcpy2DAsync(S.d, (size_t)A*N*n*4, h->w_d, (size_t)N*n*4, (size_t)N*n*4, B, cudaMemcpyDeviceToDevice, h->stream));     // candidate 0
    }
    // always cleared (:633-637)
    CK(cudaMemsetAsync(S.du, 0, (size_t)B*N*m*4, h->stream));
    CK(cudaMemsetAsync(S.dT, 0, (size_t)B*A*4, h->stream));
    reset_kernel<<<B, 128, 0, h->stream>>>(S, h->cfg.rho_init, ignore_first);
    h->launches += 1;
    CK(cudaGetLastError());
    return 0;
}
#define CKP(call) do { const int e_ = (call); if (e_){ h->err = std::string(#call) + ": " + cudaGetErrorString((cudaError_t)e_); return PDDP_E_CUDA; } } while (0)
// forward simulation of n_cand candidates of problems [b0, b0+nb): cost variant and lane shape picked here
static void launch_sim_any(pddp_handle h, cudaStream_t st, int b0, int nb, int n_cand, int a_first = 0){
    DevState &S = h->S;
    if (h->ops){ h->ops->launch_sim(&S, st, b0, nb, n_cand); return; } const int gpw = 32 / h->sim_lanes, grid = nb*((n_cand + gpw - 1)/gpw)*S.M, cta = 32;
    if (h->sim_lanes == 32){
        if (S.ee){ sim_kernel<true, 32><<<grid, cta, h->smem_sim, st>>>(S, b0, n_cand, a_first); } else { sim_kernel<false, 32><<<grid, cta, h->smem_sim, st>>>(S, b0, n_cand, a_first); }
    } else {
        if (S.ee){ sim_kernel<true, 16><<<grid, cta, h->smem_sim, st>>>(S, b0, n_cand, a_first); } else { sim_kernel<false, 16><<<grid, cta, h->smem_sim, st>>>(S, b0, n_cand, a_first); }
    }
}
static int launch_init(pddp_handle h, int rollout){       // initAlgGPU (nisInitHelpers.cuh:353-397), trajectory already in xp/up
    DevState &S = h->S; const int B = S.B, N = S.N, A = S.A, n = S.n;
    S.rolled_out = rollout ? 1 : 0;
    if (rollout){
        // loadVarsGPU's forward rollout (nisInitHelpers.cuh:646-651): candidate 0 starts as the given trajectory and is simulated
        // with alpha[0], du = 0 and the feedback gains KT around it; the result (x, u, defects) becomes the start trajectory
        CK(cudaMemcpy2DAsync(S.x, (size_t)A*N*n*4, S.xp, (size_t)N*n*4, (size_t)N*n*4, B, cudaMemcpyDeviceToDevice, h->stream));
        launch_sim_any(h, h->stream, 0, B, 1);
        h->launches += 1;
    }
    if (h->ops){ CKP(h->ops->launch_nis(&S, h->stream, rollout ? 2 : 1, 0, S.B)); CKP(h->ops->launch_init_cost(&S, h->stream)); h->launches += 1; }
    else {
    nis_kernel<<<(S.B*S.N + NIS_WARPS*(32/NIS_LANES) - 1)/(NIS_WARPS*(32/NIS_LANES)), 32*NIS_WARPS, h->smem_nis, h->stream>>>(S, rollout ? 2 : 1, 1, 0, S.B);
    // end-effector cost: the initial per-knot costs come from nis_kernel (they need the tool pose), or from the rollout's partials
    if (!S.ee){ init_cost_kernel<<<S.B, S.N, 0, h->stream>>>(S); h->launches += 1; }
    }
    select_kernel<<<S.B, 32*S.A, h->smem_sel, h->stream>>>(S, 1, 0);
    h->launches += 2;
    CK(cudaGetLastError());
    return 0;
}
// problems [b0, b0+nb) on stream st
static int launch_bp(pddp_handle h, cudaStream_t st, int b0, int nb){
    DevState &S = h->S;
    if (h->ops){ CKP(h->ops->launch_bp(&S, st, b0, nb)); h->launches += 1; return 0; }
    // two shapes with identical arithmetic: warp chains (20 resident chains per SM, no barriers) win once the launch holds several
    // chains per SM; below that a chain's latency is what counts and the block-cooperative kernel (8 warps per chain) is ~2x quicker
    const bool warp_chains = h->bp_shape == 1 || (h->bp_shape == 0 && (long long)nb*S.M >= 4LL*h->num_sms);
    if (warp_chains){
        const int nchains = nb*S.M;
        bp_warp_kernel<<<(nchains + BPW_WARPS - 1)/BPW_WARPS, 32*BPW_WARPS, BPW_WARPS*sizeof(BpWarpSmem), st>>>(S, b0, nchains);
    } else {
        bp_kernel<kuka::NX, kuka::NU><<<nb*S.M, BP_CTA, h->smem_bp, st>>>(S, b0);
    }
    h->launches += 1; CK(cudaGetLastError()); return 0;
}
static int launch_sweep(pddp_handle h, cudaStream_t st, int b0, int nb){
    DevState &S = h->S; if (S.M == 1){ return 0; }
    if (h->ops){ CKP(h->ops->launch_sweep(&S, st, b0, nb, h->num_sms)); h->launches += 1; return 0; }
    // enough CTAs to cover the SMs: split the step sizes of one problem over up to A CTAs (power-of-two divisor of A)
    int splits = 1; while (nb*splits*2 <= h->num_sms && (S.a_cnt % (splits*2)) == 0){ splits *= 2; }
    sweep_kernel<kuka::NX><<<nb*splits, 32*(S.a_cnt/splits), h->smem_sweep, st>>>(S, splits, b0);
    h->launches += 1; CK(cudaGetLastError()); return 0;
}
static int launch_sim(pddp_handle h, cudaStream_t st, int b0, int nb){
    DevState &S = h->S;
    launch_sim_any(h, st, b0, nb, S.a_cnt, S.a_first);
    h->launches += 1; CK(cudaGetLastError()); return 0;
}
static int launch_select(pddp_handle h, cudaStream_t st, int b0, int nb){
    DevState &S = h->S;
    select_kernel<<<nb, 32*S.A, h->smem_sel, st>>>(S, 0, b0);
    h->launches += 1; CK(cudaGetLastError()); return 0;
}
static int launch_nis(pddp_handle h, cudaStream_t st, int b0, int nb){
    DevState &S = h->S;
    if (h->ops){ CKP(h->ops->launch_nis(&S, st, 0, b0, nb)); h->launches += 1; return 0; }
    nis_kernel<<<(nb*S.N + NIS_WARPS*(32/NIS_LANES) - 1)/(NIS_WARPS*(32/NIS_LANES)), 32*NIS_WARPS, h->smem_nis, st>>>(S, 0, 0, b0, nb);
    h->launches += 1; CK(cudaGetLastError()); return 0;
}

// the iteration loop of runiLQR_GPU (DDPWrappers.cuh:52-114) with all host decisions moved to select_kernel.
// The batch is cut into `groups` problem groups, each on its own stream: problems are independent, so the latency-bound
// kernels of one group (backward pass, sweep, selection) run under the throughput-bound ones (sim, nis) of another.
// One iteration of problems [b0, b0+nb) on stream st: the five launches of the hot path
static int launch_iteration(pddp_handle h, cudaStream_t st, int b0, int nb){
    int rc;
    if ((rc = launch_bp(h, st, b0, nb))){ return rc; }
    if ((rc = launch_sweep(h, st, b0, nb))){ return rc; }
    if ((rc = launch_sim(h, st, b0, nb))){ return rc; }
    if ((rc = launch_select(h, st, b0, nb))){ return rc; }
    return launch_nis(h, st, b0, nb);
}
// Device-resident iteration loop: `cnt` iterations of every problem group -- fork to the group streams, 5 x cnt launches per group,
// join -- captured ONCE as a CUDA graph and replayed by every later solve of the same shape.  All decisions of an
// iteration are taken on the device (select_kernel) and a finished problem's kernels return at their first instruction, so the graph
// needs no host in the loop; only the convergence poll between chunks (TOL_COST > 0) comes back to the host.
static int iteration_graph(pddp_handle h, int cnt, int groups, pddp_solver::GraphEntry *out){
    DevState &S = h->S;
    std::string key(reinterpret_cast<const char*>(&S), sizeof(S)); key += "#" + std::to_string(cnt) + "#" + std::to_string(groups);
    auto it = h->graphs.find(key);
    if (it != h->graphs.end()){ *out = it->second; return 0; }
    if (h->graphs.size() >= 32){ for (auto &kv : h->graphs){ cudaGraphExecDestroy(kv.second.exec); } h->graphs.clear(); }
    const long l0 = h->launches; int rc = 0;
    CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
    if (groups > 1){ cudaEventRecord(h->gev[8], h->stream); for (int g = 1; g < groups; g++){ cudaStreamWaitEvent(h->gstreams[g], h->gev[8], 0); } }
    for (int i = 0; i < cnt && !rc; i++){
        for (int g = 0; g < groups && !rc; g++){
            const int b0 = (int)((long)S.B*g/groups), nb = (int)((long)S.B*(g+1)/groups) - b0;
            rc = launch_iteration(h, h->gstreams[g], b0, nb);
        }
    }
    for (int g = 1; g < groups; g++){ cudaEventRecord(h->gev[g], h->gstreams[g]); cudaStreamWaitEvent(h->stream, h->gev[g], 0); }
    cudaGraph_t graph = nullptr;
    const cudaError_t e = cudaStreamEndCapture(h->stream, &graph);             // always ends the capture, also after a failed launch
    pddp_solver::GraphEntry ge{nullptr, h->launches - l0}; h->launches = l0;
    if (rc){ if (graph){ cudaGraphDestroy(graph); } return rc; }
    if (e != cudaSuccess || !graph){ h->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e); return PDDP_E_CUDA; }
    const cudaError_t e2 = cudaGraphInstantiate(&ge.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess){ h->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2); return PDDP_E_CUDA; }
    h->graphs[key] = ge; *out = ge;
    return 0;
}

// Iteration loop with the line search sharded over ranks (north_star): every rank holds the whole problem, runs the backward
// pass and the next-iteration setup itself, but sweeps and simulates only ITS step sizes.  One exchange at selection: an all-gather of the
// (J, defect) pairs of all step sizes, after which every rank runs the reference's sequential scan (fpHelpers.cuh:395-408) on identical
// data and reaches the identical decision; the accepted candidate then travels from its owner to the others (all-reduce of its bits).
#define CKN(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess){ h->err = std::string(#call) + ": " + (h->nccl.GetErrorString ? h->nccl.GetErrorString(r_) : "NCCL error"); return PDDP_E_CUDA; } } while (0)
static int run_iterations_alpha_sharded(pddp_handle h){
    DevState &S = h->S; cudaStream_t st = h->stream; int rc;
    const size_t per = (size_t)S.N*(2*S.n + S.m);
    for (int it = 0; it < S.iter_cap; it++){
        if ((rc = launch_bp(h, st, 0, S.B))){ return rc; }
        if ((rc = launch_sweep(h, st, 0, S.B))){ return rc; }
        if ((rc = launch_sim(h, st, 0, S.B))){ return rc; }
        select_kernel<<<S.B, 32*S.a_cnt, h->smem_sel, st>>>(S, 2, 0); h->launches += 1; CK(cudaGetLastError());
        CK(cudaEventRecord(h->xev[0], st));
        CKN(h->nccl.AllGather(S.xchg_send, S.xchg_recv, (size_t)S.B*S.a_cnt*2, ncclFloat32, h->nccl.comm, st));
        CK(cudaEventRecord(h->xev[1], st));
        select_kernel<<<S.B, 32*S.A, h->smem_sel, st>>>(S, 3, 0); h->launches += 1; CK(cudaGetLastError());
        ashard_pack_kernel<<<S.B, 256, 0, st>>>(S); h->launches += 1; CK(cudaGetLastError());
        CK(cudaEventRecord(h->xev[2], st));
        CKN(h->nccl.AllReduce(S.acc_buf, S.acc_buf, (size_t)S.B*per, ncclInt32, ncclSum, h->nccl.comm, st));
        CK(cudaEventRecord(h->xev[3], st));
        ashard_unpack_kernel<<<S.B, 256, 0, st>>>(S); h->launches += 1; CK(cudaGetLastError());
        if ((rc = launch_nis(h, st, 0, S.B))){ return rc; }
        if (it == S.iter_cap - 1 || (it % 8) == 7){
            // exchange time of this iteration (sampled: reading events synchronises) and, with TOL_COST > 0, the convergence poll
            CK(cudaMemcpyAsync(h->h_nactive, S.n_active, sizeof(int), cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            float a = 0, b = 0; cudaEventElapsedTime(&a, h->xev[0], h->xev[1]); cudaEventElapsedTime(&b, h->xev[2], h->xev[3]);
            h->xchg_ms += a + b; h->xchg_calls += 1;
            if (h->cfg.tol_cost > 0.0f && *h->h_nactive == 0){ break; }
        }
    }
    return 0;
}

static int run_iterations(pddp_handle h, double *times_ms, int groups){
    DevState &S = h->S;
    if (h->nccl.comm){
        if (times_ms){ times_ms[1] = times_ms[2] = times_ms[3] = times_ms[4] = 0.0; }
        return run_iterations_alpha_sharded(h);
    }
    const bool timing = times_ms != nullptr && groups == 1;
    if (!timing && h->use_graphs){
        if (times_ms){ times_ms[1] = times_ms[2] = times_ms[3] = times_ms[4] = 0.0; }
        const bool poll_g = h->cfg.tol_cost > 0.0f;
        for (int it0 = 0; it0 < S.iter_cap; it0 += h->graph_chunk){
            const int cnt = (S.iter_cap - it0 < h->graph_chunk) ? S.iter_cap - it0 : h->graph_chunk;
            pddp_solver::GraphEntry ge; int rc = iteration_graph(h, cnt, groups, &ge); if (rc){ return rc; }
            CK(cudaGraphLaunch(ge.exec, h->stream));
            h->launches += ge.kernels; h->graph_launches += 1;
            if (poll_g && it0 + cnt < S.iter_cap){
                CK(cudaMemcpyAsync(h->h_nactive, S.n_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
                CK(cudaStreamSynchronize(h->stream));
                if (*h->h_nactive == 0){ break; }
            }
        }
        return 0;
    }
    struct EventBag { std::vector<cudaEvent_t> v; ~EventBag(){ for (auto e : v){ cudaEventDestroy(e); } } } bag;     // released on every path out
    std::vector<cudaEvent_t> &evs = bag.v;
    auto mark = [&](){ if (timing){ cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, h->stream); evs.push_back(e); } };
    const bool poll = h->cfg.tol_cost > 0.0f;
    // fork: every group stream starts after the setup on stream 0
    if (groups > 1){ CK(cudaEventRecord(h->gev[8], h->stream)); for (int g = 1; g < groups; g++){ CK(cudaStreamWaitEvent(h->gstreams[g], h->gev[8], 0)); } }
    int it_done = 0;
    for (int it = 0; it < S.iter_cap; it++){
        for (int g = 0; g < groups; g++){
            const int b0 = (int)((long)S.B*g/groups), nb = (int)((long)S.B*(g+1)/groups) - b0; cudaStream_t st = h->gstreams[g]; int rc;
            mark(); if ((rc = launch_bp(h, st, b0, nb))){ return rc; }
            mark(); if ((rc = launch_sweep(h, st, b0, nb))){ return rc; }
            mark(); if ((rc = launch_sim(h, st, b0, nb))){ return rc; }
            mark(); if ((rc = launch_select(h, st, b0, nb))){ return rc; }
            mark(); if ((rc = launch_nis(h, st, b0, nb))){ return rc; }
        }
        it_done = it + 1;
        if (poll && ((it % 4) == 3 || (long long)S.B*S.A*S.M <= 4LL*h->num_sms)){       // small batches (the latency path): every iteration
            // convergence-driven early exit: read the active-problem counter once all groups reached this iteration
            for (int g = 1; g < groups; g++){ CK(cudaStreamSynchronize(h->gstreams[g])); }
            CK(cudaMemcpyAsync(h->h_nactive, S.n_active, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            if (*h->h_nactive == 0){ break; }
        }
    }
    mark();
    // join: stream 0 continues after every group
    for (int g = 1; g < groups; g++){ CK(cudaEventRecord(h->gev[g], h->gstreams[g])); CK(cudaStreamWaitEvent(h->stream, h->gev[g], 0)); }
    if (times_ms){ times_ms[1] = times_ms[2] = times_ms[3] = times_ms[4] = 0.0; }
    if (timing){
        CK(cudaStreamSynchronize(h->stream));
        double acc[5] = {0,0,0,0,0};
        for (auto &v : h->it_ms){ v.assign(it_done, 0.0); }
        for (int it = 0; it < it_done; it++){
            float ph_ms[5];
            for (int ph = 0; ph < 5; ph++){ ph_ms[ph] = 0; cudaEventElapsedTime(&ph_ms[ph], evs[it*5+ph], evs[it*5+ph+1]); acc[ph] += ph_ms[ph]; }
            h->it_ms[0][it] = ph_ms[2] + ph_ms[3]; h->it_ms[1][it] = ph_ms[1]; h->it_ms[2][it] = ph_ms[0]; h->it_ms[3][it] = ph_ms[4];
        }
        // reference order of the timing outputs: tTime, simTime, sweepTime, bpTime, nisTime, initTime (DDPWrappers.cuh:11)
        times_ms[1] = acc[2] + acc[3]; times_ms[2] = acc[1]; times_ms[3] = acc[0]; times_ms[4] = acc[4];
    }
    return 0;
}

extern "C" int pddp_solve_device(pddp_handle h, const float *d_x0, const float *d_u0, const float *d_xGoal, int ignoreFirstDefectFlag,
                                 float *d_x_out, float *d_u_out, float *d_Jout, int *d_alphaOut, int *d_iters_out, double *times_ms){
    if (!h){ return PDDP_E_INVALID; }
    DevState &S = h->S; const int B = S.B, N = S.N, n = S.n, m = S.m; int rc;
    CK(cudaSetDevice(h->cfg.device));
    h->launches = 0; h->graph_launches = 0;
    if (h->nccl.comm && (h->next_rollout || !h->next_clear)){ h->err = "step-size sharding: cold starts only (rollout = 0, clear = 1)"; return PDDP_E_INVALID; }
    h->xchg_ms = 0; h->xchg_calls = 0;
    CK(cudaEventRecord(h->ev[0], h->stream));
    if ((rc = launch_reset(h, ignoreFirstDefectFlag, h->next_clear))){ return rc; }
    CK(cudaMemcpyAsync(S.xp, d_x0, (size_t)B*N*n*4, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(S.up, d_u0, (size_t)B*N*m*4, cudaMemcpyDeviceToDevice, h->stream));
    if (d_xGoal != S.xGoal){ CK(cudaMemcpyAsync(S.xGoal, d_xGoal, (size_t)B*n*4, cudaMemcpyDeviceToDevice, h->stream)); }      // pddp_solve uploads straight into S.xGoal
    if ((rc = launch_init(h, h->next_rollout))){ return rc; }
    h->next_clear = 1; h->next_rollout = 0;          // the flags apply to one solve
    CK(cudaEventRecord(h->ev[1], h->stream));
    if ((rc = run_iterations(h, times_ms, h->groups))){ return rc; }
    CK(cudaEventRecord(h->ev[2], h->stream));
    store_kernel<<<B, 256, 0, h->stream>>>(S, d_x_out ? d_x_out : h->d_xout, d_u_out ? d_u_out : h->d_uout, d_iters_out ? d_iters_out : h->d_iters);
    h->launches += 1; CK(cudaGetLastError());
    if (d_Jout){ CK(cudaMemcpyAsync(d_Jout, S.Jout, (size_t)B*(S.max_iter+1)*4, cudaMemcpyDeviceToDevice, h->stream)); }
    if (d_alphaOut){ CK(cudaMemcpyAsync(d_alphaOut, S.alphaOut, (size_t)B*(S.max_iter+1)*4, cudaMemcpyDeviceToDevice, h->stream)); }
    CK(cudaEventRecord(h->ev[3], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (times_ms){
        float t_all = 0, t_init = 0, t_store = 0;
        cudaEventElapsedTime(&t_all, h->ev[0], h->ev[3]); cudaEventElapsedTime(&t_init, h->ev[0], h->ev[1]); cudaEventElapsedTime(&t_store, h->ev[2], h->ev[3]);
        times_ms[0] = t_all; times_ms[5] = t_init + t_store;
    }
    return 0;
}

extern "C" int pddp_solve(pddp_handle h, const float *x0, const float *u0, const float *xGoal, int forwardRolloutFlag, int clearVarsFlag,
                          int ignoreFirstDefectFlag, float *x_out, float *u_out, float *Jout, int *alphaOut, int *iters_out, double *times_ms){
    if (!h){ return PDDP_E_INVALID; }
    h->next_clear = clearVarsFlag ? 1 : 0; h->next_rollout = forwardRolloutFlag ? 1 : 0;
    if (!x0 || !u0 || !xGoal){ h->err = "null input"; return PDDP_E_INVALID; }
    DevState &S = h->S; const int B = S.B, N = S.N, n = S.n, m = S.m, L = S.max_iter + 1; int rc;
    CK(cudaSetDevice(h->cfg.device));
    // host -> pinned staging -> device (inputs), device -> pinned -> host (results): all inside the caller-visible time
    float *st = h->h_stage; float *sx = st, *su = sx + (size_t)B*N*n, *sg = su + (size_t)B*N*m;
    float *sJ = sg + (size_t)B*n; int *sA = reinterpret_cast<int*>(sJ + (size_t)B*L); int *sI = sA + (size_t)B*L;
    std::memcpy(sx, x0, (size_t)B*N*n*4); std::memcpy(su, u0, (size_t)B*N*m*4); std::memcpy(sg, xGoal, (size_t)B*n*4);
    // inputs travel to scratch device buffers (d_xout/d_uout double as input staging), then the device path runs
    CK(cudaMemcpyAsync(h->d_xout, sx, (size_t)B*N*n*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_uout, su, (size_t)B*N*m*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.xGoal, sg, (size_t)B*n*4, cudaMemcpyHostToDevice, h->stream));
    if ((rc = pddp_solve_device(h, h->d_xout, h->d_uout, S.xGoal, ignoreFirstDefectFlag, nullptr, nullptr, nullptr, nullptr, nullptr, times_ms))){ return rc; }
    CK(cudaMemcpyAsync(sx, h->d_xout, (size_t)B*N*n*4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(su, h->d_uout, (size_t)B*N*m*4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(sJ, S.Jout, (size_t)B*L*4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(sA, S.alphaOut, (size_t)B*L*4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(sI, h->d_iters, (size_t)B*4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (x_out){ std::memcpy(x_out, sx, (size_t)B*N*n*4); } if (u_out){ std::memcpy(u_out, su, (size_t)B*N*m*4); }
    if (Jout){ std::memcpy(Jout, sJ, (size_t)B*L*4); } if (alphaOut){ std::memcpy(alphaOut, sA, (size_t)B*L*4); }
    if (iters_out){ std::memcpy(iters_out, sI, (size_t)B*4); }
    return 0;
}

namespace pddp {
__global__ void selftest_rcp_kernel(unsigned long long *bad){
    unsigned long long local = 0;
    for (unsigned long long v = (unsigned long long)blockIdx.x*blockDim.x + threadIdx.x; v < (1ull << 32); v += (unsigned long long)gridDim.x*blockDim.x){
        const float x = __uint_as_float((unsigned)v);
        const float a = rcp_rn(x), b = __fdiv_rn(1.0f, x);
        const bool same = (__float_as_uint(a) == __float_as_uint(b)) || (a != a && b != b);
        local += same ? 0 : 1;
    }
    if (local){ atomicAdd(bad, local); }
}
__global__ void selftest_sincos_kernel(unsigned long long *bad){
    unsigned long long local = 0;
    for (unsigned long long v = (unsigned long long)blockIdx.x*blockDim.x + threadIdx.x; v < (1ull << 32); v += (unsigned long long)gridDim.x*blockDim.x){
        const float x = __uint_as_float((unsigned)v);
        float s, c; kuka::sincos_as_library(x, s, c);
        const float rs = sinf(x), rc = cosf(x);
        const bool same = ((__float_as_uint(s) == __float_as_uint(rs)) || (s != s && rs != rs)) && ((__float_as_uint(c) == __float_as_uint(rc)) || (c != c && rc != rc));
        local += same ? 0 : 1;
    }
    if (local){ atomicAdd(bad, local); }
}
}
// ---------------------------------------------------------------------------------------------------- receding horizon
// EE_COST: the reference's xTarget argument (costFunc / costGrad cost_arm.cuh:263-281; null for runiLQR_GPU, gv->d_xTarget for
// runiLQR_MPC_GPU, MPCHelpers.cuh:900).  HOST [batch][n], or NULL for "no target".
extern "C" int pddp_set_x_target(pddp_handle h, const float *xTarget){
    if (!h){ return PDDP_E_INVALID; }
    DevState &S = h->S; const size_t B = S.B, n = S.n;
    CK(cudaSetDevice(h->cfg.device));
    if (!xTarget){ S.xTarget = nullptr; return 0; }
    if (!h->d_xTarget){ void *q = nullptr; CK(cudaMalloc(&q, B*n*4)); h->d_xTarget = (float*)q; h->allocs.push_back(q); }
    CK(cudaMemcpy(h->d_xTarget, xTarget, B*n*4, cudaMemcpyHostToDevice));
    S.xTarget = h->d_xTarget;
    return 0;
}

// runiLQR_MPC_GPU's use_cost_shift argument (MPCHelpers.cuh:866,876): finalCostShift = shiftAmount of the step.  EE_COST only (the
// joint-space cost ignores it, as in the reference).
extern "C" int pddp_mpc_set_cost_shift(pddp_handle h, int use_cost_shift){
    if (!h){ return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    if (use_cost_shift && !h->d_cost_shift){ void *q = nullptr; CK(cudaMalloc(&q, (size_t)h->S.B*sizeof(int))); CK(cudaMemset(q, 0, (size_t)h->S.B*sizeof(int))); h->d_cost_shift = (int*)q; h->allocs.push_back(q); }
    h->use_cost_shift = use_cost_shift != 0;
    h->S.cost_shift = nullptr;                                        // set per step by pddp_mpc_step, cleared after it
    return 0;
}

extern "C" int pddp_mpc_init(pddp_handle h, const float *x_init, const float *u_init){
    if (!h){ return PDDP_E_INVALID; }
    if (!x_init || !u_init){ h->err = "null input"; return PDDP_E_INVALID; }
    DevState &S = h->S; const size_t B = S.B, N = S.N, n = S.n, m = S.m, A = S.A;
    CK(cudaSetDevice(h->cfg.device));
    if (!h->mpc.cx){
        auto da = [&](float **p, size_t cnt){ void *q = nullptr; if (cudaMalloc(&q, cnt*4) != cudaSuccess){ return false; } *p = (float*)q; h->allocs.push_back(q); return true; };
        float *xa = nullptr;
        if (!da(&h->mpc.cx, B*N*n) || !da(&h->mpc.cu, B*N*m) || !da(&h->mpc.cd, B*N*n) || !da(&h->mpc.x_old, B*N*n) || !da(&h->mpc.u_old, B*N*m) ||
            !da(&h->mpc.KT_old, B*N*n*m) || !da(&xa, B*n)){ h->err = "cudaMalloc failed (receding-horizon state)"; return PDDP_E_CUDA; }
        h->d_xActual = xa; h->mpc.xActual = xa;
        void *q = nullptr; CK(cudaMalloc(&q, 3*B*sizeof(int))); h->d_mpc_flags = (int*)q; h->allocs.push_back(q);
        h->mpc.shift = h->d_mpc_flags; h->mpc.clear = h->d_mpc_flags + B;
        h->mpc_pack_bytes = (B*N*(n + m + n*m) + 2*B*((size_t)S.max_iter + 1) + 2*B)*4;
        q = nullptr; CK(cudaMalloc(&q, h->mpc_pack_bytes)); h->d_mpc_pack = (float*)q; h->allocs.push_back(q);
        CK(cudaMallocHost((void**)&h->h_mpc_pack, h->mpc_pack_bytes));
        if (!h->ops){
            h->smem_mpc = 2*36*kuka::NB*sizeof(float) + sizeof(SimGroupSmem);
            CK(cudaFuncSetAttribute(mpc_load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_mpc));
        }
    }
    // the state runiLQR_MPC_GPU expects to find (LCMHelpers.cuh:224-233 + the caller's set-up): the plan in the current slot and in
    // xp/up, everything else zero
    CK(cudaMemcpyAsync(h->mpc.cx, x_init, B*N*n*4, cudaMemcpyHostToDevice, h->stream)); CK(cudaMemcpyAsync(h->mpc.cu, u_init, B*N*m*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.xp, x_init, B*N*n*4, cudaMemcpyHostToDevice, h->stream)); CK(cudaMemcpyAsync(S.up, u_init, B*N*m*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->mpc.cd, 0, B*N*n*4, h->stream)); CK(cudaMemsetAsync(S.dp, 0, B*N*n*4, h->stream)); CK(cudaMemsetAsync(S.d, 0, B*A*N*n*4, h->stream));
    CK(cudaMemsetAsync(S.xp2, 0, B*N*n*4, h->stream));
    CK(cudaMemsetAsync(S.KT, 0, B*N*n*m*4, h->stream)); CK(cudaMemsetAsync(S.du, 0, B*N*m*4, h->stream));
    CK(cudaMemsetAsync(S.Pbuf[0], 0, B*N*n*n*4, h->stream)); CK(cudaMemsetAsync(S.Pbuf[1], 0, B*N*n*n*4, h->stream));
    CK(cudaMemsetAsync(S.pbuf[0], 0, B*N*n*4, h->stream)); CK(cudaMemsetAsync(S.pbuf[1], 0, B*N*n*4, h->stream));
    CK(cudaMemsetAsync(h->mpc.x_old, 0, B*N*n*4, h->stream)); CK(cudaMemsetAsync(h->mpc.u_old, 0, B*N*m*4, h->stream)); CK(cudaMemsetAsync(h->mpc.KT_old, 0, B*N*n*m*4, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemset(S.alphaIndex, 0, B*sizeof(int)));                 // the plan starts in slot 0 (MPCHelpers.cuh:236)
    CK(cudaMemset(S.iter, 0, B*sizeof(int)));                       // both cost-to-go buffers are zero: no roles to keep yet
    h->mpc_lss.assign(B, 0); h->mpc_ready = true;
    return 0;
}

extern "C" int pddp_mpc_step(pddp_handle h, const float *xActual, const float *xGoal, const int *shiftAmount, int max_iter, int clear_vars,
                             int ignoreFirstDefectFlag, float *x, float *u, float *KT, float *Jout, int *alphaOut, int *iters_out, int *last_successful_solve){
    if (!h){ return PDDP_E_INVALID; }
    if (!h->mpc_ready){ h->err = "pddp_mpc_init first"; return PDDP_E_INVALID; }
    if (h->nccl.comm){ h->err = "step-size sharding: plain solves only"; return PDDP_E_INVALID; }
    if (!xActual || !xGoal || !shiftAmount || !x || !u || !KT){ h->err = "null input"; return PDDP_E_INVALID; }
    DevState &S = h->S; const int B = S.B, N = S.N, n = S.n, m = S.m, L = S.max_iter + 1; int rc;
    if (max_iter < 1 || max_iter > S.max_iter){ h->err = "max_iter of the step must be in [1, config max_iter]"; return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    h->launches = 0; h->graph_launches = 0;
    std::vector<int> flags(3*(size_t)B);
    for (int b = 0; b < B; b++){
        if (shiftAmount[b] < 0){ h->err = "negative shiftAmount"; return PDDP_E_INVALID; }
        flags[b] = shiftAmount[b];
        flags[B + b] = (h->mpc_lss[b] > 10 /* SOLVES_TO_RESET, MPCHelpers.cuh:34-36 */ || clear_vars) ? 1 : 0;
        flags[2*(size_t)B + b] = (h->mpc_lss[b] == 0) ? 1 : 0;     // a counter still at zero reaches one whatever the solve does: publish (:987-991)
    }
    CK(cudaMemcpyAsync(h->d_mpc_flags, flags.data(), 3*(size_t)B*sizeof(int), cudaMemcpyHostToDevice, h->stream));
    if (h->use_cost_shift && S.ee){ CK(cudaMemcpyAsync(h->d_cost_shift, flags.data(), (size_t)B*sizeof(int), cudaMemcpyHostToDevice, h->stream)); S.cost_shift = h->d_cost_shift; }
    CK(cudaMemcpyAsync(h->d_xActual, xActual, (size_t)B*n*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(S.xGoal, xGoal, (size_t)B*n*4, cudaMemcpyHostToDevice, h->stream));
    // loadVarsGPU_MPC + hand-over; the cost-to-go buffers keep their roles: the first backward pass must seed its blocks from the
    // shifted Pp (the older buffer) and overwrite P (the newer one) -- the load kernel puts them into the slots that the restarted
    // iteration counter selects, per problem (problems that stopped at different iterations have different parities)
    if (h->ops){ CKP(h->ops->launch_mpc_load(&S, &h->mpc, h->stream)); }
    else { mpc_load_kernel<<<B, 256, h->smem_mpc, h->stream>>>(S, h->mpc); }
    h->launches += 1; CK(cudaGetLastError());
    reset_kernel<<<B, 128, 0, h->stream>>>(S, h->cfg.rho_init, ignoreFirstDefectFlag);
    h->launches += 1; CK(cudaGetLastError());
    rc = launch_init(h, 0);
    if (!rc){ S.iter_cap = max_iter; rc = run_iterations(h, nullptr, 1); }
    S.iter_cap = S.max_iter; S.cost_shift = nullptr;
    if (rc){ return rc; }
    // the success scan (MPCHelpers.cuh:987-991, 757-758) runs inside mpc_store_kernel (the counter's value before the step went up with the flags), and everything the caller gets back returns
    // in one copy to pinned memory: one synchronisation per step.  Successful problems publish their plan and gains, the others keep
    // the caller's previous plan.
    mpc_store_kernel<<<B*MPC_STORE_SPLIT, 256, 0, h->stream>>>(S, h->mpc, h->d_mpc_flags + 2*(size_t)B, h->d_mpc_pack);
    h->launches += 1; CK(cudaGetLastError());
    CK(cudaMemcpyAsync(h->h_mpc_pack, h->d_mpc_pack, h->mpc_pack_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const float *sx = h->h_mpc_pack, *su = sx + (size_t)B*N*n, *sKT = su + (size_t)B*N*m;
    const int *aout = reinterpret_cast<const int*>(sKT + (size_t)B*N*n*m); const float *jout = reinterpret_cast<const float*>(aout + (size_t)B*L);
    const int *its = reinterpret_cast<const int*>(jout + (size_t)B*L), *succ = its + B;
    for (int b = 0; b < B; b++){
        h->mpc_lss[b] = succ[b] ? 1 : h->mpc_lss[b] + 1;
        if (succ[b]){
            std::memcpy(x + (size_t)b*N*n, sx + (size_t)b*N*n, (size_t)N*n*4); std::memcpy(u + (size_t)b*N*m, su + (size_t)b*N*m, (size_t)N*m*4);
            std::memcpy(KT + (size_t)b*N*n*m, sKT + (size_t)b*N*n*m, (size_t)N*n*m*4);
        }
        if (last_successful_solve){ last_successful_solve[b] = h->mpc_lss[b]; }
        if (iters_out){ iters_out[b] = its[b]; }
    }
    if (Jout){ std::memcpy(Jout, jout, (size_t)B*L*4); } if (alphaOut){ std::memcpy(alphaOut, aout, (size_t)B*L*4); }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- consumer side of the hand-off
// getHardwareControls (MPCHelpers.cuh:817-858): host arithmetic in the reference (float products and sums rounded one by one, the
// smoothing in double), restated operation by operation.
extern "C" int pddp_hardware_controls(int N, double time_step, const float *x, const float *u, const float *KT, double t0,
                                      const double *qActual, const double *qdActual, double tActual, int use_feedback, int pd_gains_on_state,
                                      double *u_prev, double alpha, double *q_out, double *u_out){
    constexpr int n = kuka::NX, m = kuka::NU, np = kuka::NB;
    if (N < 3 || !x || !u || !KT || !qActual || !qdActual || !q_out || !u_out){ return PDDP_E_INVALID; }
    const double step_us = (time_step*1000.0)*1000.0;                       // TIME_STEP_LENGTH_IN_us, :29-30
    const double dt = (tActual - t0)/step_us; const int k = static_cast<int>(dt); const double fraction = dt - static_cast<double>(k);
    if (k >= N - 2 || k < 0){ return 1; }                                   // beyond the plan (:827)
    const float *uk = u + (size_t)k*m;
    if (use_feedback){
        const float *KTk = KT + (size_t)k*n*m, *xd = x + (size_t)k*n, *xu = x + (size_t)(k+1)*n;
        volatile float dx[n];                                               // volatile: every float operation is rounded on its own (no contraction by the host compiler)
        const float w0 = static_cast<float>(1.0 - fraction), w1 = static_cast<float>(fraction);
        for (int i = 0; i < n; i++){
            volatile float a = w0*xd[i], b = w1*xu[i]; volatile float val = a + b;
            dx[i] = static_cast<float>(i < np ? qActual[i] : qdActual[i-np]) - val;
            if (pd_gains_on_state && i < np){ q_out[i] = static_cast<double>(val); }
        }
        for (int r = 0; r < m; r++){
            volatile float val = uk[r];
            for (int c = 0; c < n; c++){ volatile float p = KTk[c + r*n]*dx[c]; val = val - p; }
            u_out[r] = static_cast<double>(val);
        }
    } else {
        for (int i = 0; i < m; i++){ u_out[i] = static_cast<double>(uk[i]); }
    }
    if (!pd_gains_on_state){ for (int i = 0; i < np; i++){ q_out[i] = qActual[i]; } }
    if (u_prev && alpha > 0){ for (int i = 0; i < m; i++){ u_out[i] = (1 - alpha)*u_out[i] + alpha*u_prev[i]; u_prev[i] = u_out[i]; } }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- trajectory hand-off
// lcmt_trajectory_f (lcmtypes/drake/lcmt_trajectory_f.hpp): LCM wire format = 8-byte fingerprint, then the fields in declaration
// order, every scalar in network byte order.  Host-only code.
namespace {
const unsigned long long TRAJ_F_FINGERPRINT = ((0x8fb839bd5c6031eeULL << 1) + ((0x8fb839bd5c6031eeULL >> 63) & 1));   // :256-260
inline void put_be(unsigned char *&p, unsigned long long v, int nbytes){ for (int b = nbytes - 1; b >= 0; b--){ *p++ = (unsigned char)(v >> (8*b)); } }
inline unsigned long long get_be(const unsigned char *&p, int nbytes){ unsigned long long v = 0; for (int b = 0; b < nbytes; b++){ v = (v << 8) | *p++; } return v; }
inline void put_floats(unsigned char *&p, const float *src, long n_data, long n_total){
    for (long i = 0; i < n_total; i++){ unsigned int bits = 0; if (i < n_data){ std::memcpy(&bits, &src[i], 4); } put_be(p, bits, 4); }
}
}
extern "C" long pddp_traj_f_encoded_size(int x_size, int u_size, int KT_size){
    if (x_size < 0 || u_size < 0 || KT_size < 0){ return PDDP_E_INVALID; }
    return 8 + 8 + 3*4 + 4L*((long)x_size + u_size + KT_size);
}
extern "C" long pddp_traj_f_encode(long long utime, const float *x, int x_size, const float *u, int u_size, const float *KT, int KT_size,
                                   void *buf, long capacity){
    const long need = pddp_traj_f_encoded_size(x_size, u_size, KT_size);
    if (need < 0 || !buf || capacity < need || (x_size && !x) || (u_size && !u) || (KT_size && !KT)){ return PDDP_E_INVALID; }
    unsigned char *p = static_cast<unsigned char*>(buf);
    put_be(p, TRAJ_F_FINGERPRINT, 8); put_be(p, (unsigned long long)utime, 8);
    put_be(p, (unsigned int)x_size, 4); put_be(p, (unsigned int)u_size, 4); put_be(p, (unsigned int)KT_size, 4);
    put_floats(p, x, x_size, x_size); put_floats(p, u, u_size, u_size); put_floats(p, KT, KT_size, KT_size);
    return need;
}
extern "C" long pddp_traj_f_decode(const void *buf, long nbytes, long long *utime, int *x_size, int *u_size, int *KT_size,
                                   float *x, float *u, float *KT, long cap_x, long cap_u, long cap_KT){
    if (!buf || nbytes < 28){ return PDDP_E_INVALID; }
    const unsigned char *p = static_cast<const unsigned char*>(buf);
    if (get_be(p, 8) != TRAJ_F_FINGERPRINT){ return PDDP_E_INVALID; }
    const long long t = (long long)get_be(p, 8); const int xs = (int)get_be(p, 4), us = (int)get_be(p, 4), ks = (int)get_be(p, 4);
    if (xs < 0 || us < 0 || ks < 0 || nbytes < 28 + 4L*((long)xs + us + ks)){ return PDDP_E_INVALID; }
    if (utime){ *utime = t; } if (x_size){ *x_size = xs; } if (u_size){ *u_size = us; } if (KT_size){ *KT_size = ks; }
    auto take = [&](float *dst, long cap, int n){ for (int i = 0; i < n; i++){ unsigned int bits = (unsigned int)get_be(p, 4); if (dst && i < cap){ std::memcpy(&dst[i], &bits, 4); } } };
    take(x, cap_x, xs); take(u, cap_u, us); take(KT, cap_KT, ks);
    return 28 + 4L*((long)xs + us + ks);
}
// the message the reference's MPC loop publishes (LCMHelpers.cuh:245-252): the *_size fields carry BYTE counts
// (ld * TRAJ_RUNNER_TIME_STEPS * sizeof(float)), the arrays have that many floats with the data in their first quarter and zeros
// behind; without feedback (USE_FEEDBACK_IN_TRAJ_RUNNER 0) x and KT are empty.  buf == NULL returns the size needed.
extern "C" long pddp_traj_f_pack_reference(long long utime, const float *x, const float *u, const float *KT, int steps, int with_feedback,
                                           void *buf, long capacity){
    if (steps < 1 || !u || (with_feedback && (!x || !KT))){ return PDDP_E_INVALID; }
    const int n = kuka::NX, m = kuka::NU;
    const int u_size = m*steps*4, x_size = with_feedback ? n*steps*4 : 0, KT_size = with_feedback ? n*m*steps*4 : 0;
    const long need = pddp_traj_f_encoded_size(x_size, u_size, KT_size);
    if (!buf){ return need; }
    if (capacity < need){ return PDDP_E_INVALID; }
    unsigned char *p = static_cast<unsigned char*>(buf);
    put_be(p, TRAJ_F_FINGERPRINT, 8); put_be(p, (unsigned long long)utime, 8);
    put_be(p, (unsigned int)x_size, 4); put_be(p, (unsigned int)u_size, 4); put_be(p, (unsigned int)KT_size, 4);
    put_floats(p, x, with_feedback ? n*steps : 0, x_size); put_floats(p, u, m*steps, u_size); put_floats(p, KT, with_feedback ? n*m*steps : 0, KT_size);
    return need;
}

extern "C" int pddp_set_warm_start(pddp_handle h, const float *KT0, const float *P0, const float *p0, const float *d0){
    if (!h){ return PDDP_E_INVALID; }
    if (!KT0 || !P0 || !p0 || !d0){ h->err = "null warm-start array"; return PDDP_E_INVALID; }
    DevState &S = h->S; const size_t B = S.B, N = S.N, n = S.n, m = S.m;
    CK(cudaSetDevice(h->cfg.device));
    if (!h->w_KT){
        void *p = nullptr;
        CK(cudaMalloc(&p, B*N*n*m*4)); h->w_KT = (float*)p; h->allocs.push_back(p);
        CK(cudaMalloc(&p, B*N*n*n*4)); h->w_P = (float*)p; h->allocs.push_back(p);
        CK(cudaMalloc(&p, B*N*n*4)); h->w_p = (float*)p; h->allocs.push_back(p);
        CK(cudaMalloc(&p, B*N*n*4)); h->w_d = (float*)p; h->allocs.push_back(p);
    }
    CK(cudaMemcpyAsync(h->w_KT, KT0, B*N*n*m*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->w_P, P0, B*N*n*n*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->w_p, p0, B*N*n*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->w_d, d0, B*N*n*4, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return 0;
}
extern "C" int pddp_set_skip_unchanged(pddp_handle h, int on){
    if (!h){ return PDDP_E_INVALID; }
    h->S.skip_unchanged = on ? 1 : 0; return 0;
}
extern "C" int pddp_set_start_mode(pddp_handle h, int forwardRolloutFlag, int clearVarsFlag){
    if (!h){ return PDDP_E_INVALID; }
    h->next_rollout = forwardRolloutFlag ? 1 : 0; h->next_clear = clearVarsFlag ? 1 : 0; return 0;
}
extern "C" int pddp_selftest_rcp(unsigned long long *mismatches){
    if (!mismatches){ return PDDP_E_INVALID; }
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess || cudaMemset(d, 0, 8) != cudaSuccess){ return PDDP_E_CUDA; }
    pddp::selftest_rcp_kernel<<<148*8, 256>>>(d);
    const cudaError_t e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : PDDP_E_CUDA;
}
extern "C" int pddp_selftest_sincos(unsigned long long *mismatches){
    if (!mismatches){ return PDDP_E_INVALID; }
    unsigned long long *d = nullptr;
    if (cudaMalloc(&d, 8) != cudaSuccess || cudaMemset(d, 0, 8) != cudaSuccess){ return PDDP_E_CUDA; }
    pddp::selftest_sincos_kernel<<<148*8, 256>>>(d);
    const cudaError_t e = cudaMemcpy(mismatches, d, 8, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? 0 : PDDP_E_CUDA;
}
// ---------------------------------------------------------------------------------------------------- step-size sharding over GPUs
extern "C" int pddp_alpha_shard_unique_id(void *id128){
    if (!id128){ return PDDP_E_INVALID; }
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib){ g_plant_error = std::string("dlopen libnccl.so.2: ") + dlerror(); return PDDP_E_INVALID; }
    auto fn = reinterpret_cast<ncclResult_t (*)(ncclUniqueId*)>(dlsym(lib, "ncclGetUniqueId"));
    if (!fn){ return PDDP_E_INVALID; }
    static_assert(sizeof(ncclUniqueId) == 128, "NCCL unique id");
    return fn(static_cast<ncclUniqueId*>(id128)) == ncclSuccess ? 0 : PDDP_E_CUDA;
}
extern "C" int pddp_alpha_shard_init(pddp_handle h, int rank, int nranks, const void *id128){
    if (!h || !id128 || nranks < 1 || rank < 0 || rank >= nranks){ return PDDP_E_INVALID; }
    DevState &S = h->S;
    if (h->ops){ h->err = "step-size sharding is built for the Kuka kernels (PLANT 4)"; return PDDP_E_INVALID; }
    if (S.ee){ h->err = "step-size sharding: joint-space cost only"; return PDDP_E_INVALID; }
    if (S.A % nranks){ h->err = "n_alpha must be a multiple of the number of ranks"; return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    auto &nc = h->nccl;
    nc.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!nc.lib){ h->err = std::string("dlopen libnccl.so.2: ") + dlerror(); return PDDP_E_INVALID; }
    #define SYM(field, name) nc.field = reinterpret_cast<decltype(nc.field)>(dlsym(nc.lib, name)); if (!nc.field){ h->err = std::string("libnccl lacks ") + name; return PDDP_E_INVALID; }
    SYM(CommInitRank, "ncclCommInitRank") SYM(AllGather, "ncclAllGather") SYM(AllReduce, "ncclAllReduce") SYM(CommDestroy, "ncclCommDestroy") SYM(GetErrorString, "ncclGetErrorString")
    #undef SYM
    ncclUniqueId id; std::memcpy(&id, id128, sizeof(id));
    CKN(nc.CommInitRank(&nc.comm, nranks, id, rank));
    nc.rank = rank; nc.nranks = nranks;
    S.a_cnt = S.A / nranks; S.a_first = rank*S.a_cnt;
    const size_t per = (size_t)S.N*(2*S.n + S.m);
    void *p = nullptr;
    CK(cudaMalloc(&p, (size_t)S.B*S.a_cnt*2*4)); S.xchg_send = (float*)p; h->allocs.push_back(p);
    CK(cudaMalloc(&p, (size_t)S.B*S.A*2*4)); S.xchg_recv = (float*)p; h->allocs.push_back(p);
    CK(cudaMalloc(&p, (size_t)S.B*per*4)); S.acc_buf = (int*)p; h->allocs.push_back(p);
    for (auto &e : h->xev){ CK(cudaEventCreate(&e)); }
    return 0;
}
extern "C" int pddp_alpha_shard_stats(pddp_handle h, double *exchange_us_per_iteration, int *a_first, int *a_cnt){
    if (!h){ return PDDP_E_INVALID; }
    if (exchange_us_per_iteration){ *exchange_us_per_iteration = h->xchg_calls ? 1000.0*h->xchg_ms/h->xchg_calls : 0.0; }
    if (a_first){ *a_first = h->S.a_first; } if (a_cnt){ *a_cnt = h->S.a_cnt; }
    return 0;
}

extern "C" int pddp_set_bp_shape(pddp_handle h, int shape){ if (!h || shape < 0 || shape > 2){ return PDDP_E_INVALID; } h->bp_shape = shape; return 0; }
extern "C" long pddp_last_launch_count(pddp_handle h){ return h ? h->launches : 0; }
extern "C" long pddp_last_graph_launch_count(pddp_handle h){ return h ? h->graph_launches : 0; }
extern "C" int pddp_set_graphs(pddp_handle h, int on, int iterations_per_graph){
    if (!h || iterations_per_graph < 0 || iterations_per_graph > 1000){ return PDDP_E_INVALID; }
    h->use_graphs = on != 0; if (iterations_per_graph > 0){ h->graph_chunk = iterations_per_graph; }
    return 0;
}
// per-iteration device times (ms) of the last solve that was given a times_ms array and ran as one problem group: what the reference
// stores in its simTime / sweepTime / bpTime / nisTime arrays (DDPWrappers.cuh:60-107; there: host clock around each phase)
extern "C" int pddp_last_iteration_times(pddp_handle h, double *sim_ms, double *sweep_ms, double *bp_ms, double *nis_ms, int capacity){
    if (!h || capacity < 0){ return PDDP_E_INVALID; }
    const int cnt = (int)h->it_ms[0].size() < capacity ? (int)h->it_ms[0].size() : capacity;
    for (int i = 0; i < cnt; i++){
        if (sim_ms){ sim_ms[i] = h->it_ms[0][i]; } if (sweep_ms){ sweep_ms[i] = h->it_ms[1][i]; } if (bp_ms){ bp_ms[i] = h->it_ms[2][i]; } if (nis_ms){ nis_ms[i] = h->it_ms[3][i]; }
    }
    return cnt;
}
// largest L1 defect on the shooting-interval boundaries of every problem's final trajectory: what storeVarsGPU computes for the summary
// line of runiLQR_GPU (nisInitHelpers.cuh:739-750, DDPWrappers.cuh:134 `max_d`)
extern "C" int pddp_final_max_defect(pddp_handle h, float *max_d){
    if (!h || !max_d){ return PDDP_E_INVALID; }
    DevState &S = h->S; const size_t B = S.B, N = S.N, n = S.n, A = S.A; const int NBF = S.N / S.M;
    CK(cudaSetDevice(h->cfg.device)); for (auto st : h->gstreams){ CK(cudaStreamSynchronize(st)); }
    std::vector<int> src(B); std::vector<float> dT(B*A), dp(B*N*n);
    CK(cudaMemcpy(src.data(), S.final_src, B*sizeof(int), cudaMemcpyDeviceToHost)); CK(cudaMemcpy(dT.data(), S.dT, B*A*4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(dp.data(), S.dp, B*N*n*4, cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; b++){
        if (src[b] >= 0){ max_d[b] = dT[b*A + src[b]]; continue; }
        float best = 0.f;                                   // the accepted trajectory of an earlier iteration: defectKern's sum and maximum (fpHelpers.cuh:94-111)
        for (int k = NBF - 1; k < (int)N - 1; k += NBF){ volatile float acc = 0.f; for (size_t c = 0; c < n; c++){ acc = acc + std::fabs(dp[(b*N + k)*n + c]); } best = acc > best ? (float)acc : best; }
        max_d[b] = best;
    }
    return 0;
}
extern "C" int pddp_set_groups(pddp_handle h, int groups){
    if (!h || groups < 1 || groups > 8){ return PDDP_E_INVALID; }
    h->groups = (h->S.B >= 2*groups || groups == 1) ? groups : 1; return h->groups;
}

// WAFR_iLQR_examples.cu:67-121 with one std::default_random_engine(seed) per problem (the deterministic harness' inputs)
extern "C" int pddp_make_inputs_kuka(int N, int batch, unsigned seed0, float *x0, float *u0, float *xGoal){
    const double PI = 3.14159;
    const float q0[7] = {(float)(-0.5*PI), (float)(0.25*PI), (float)(0.167*PI), (float)(-0.167*PI), (float)(0.125*PI), (float)(0.167*PI), (float)(0.5*PI)};
    const float uu[7] = {(float)0.0, (float)-102.9832, (float)11.1968, (float)47.0724, (float)2.5993, (float)-7.0290, (float)-0.0907};
    const float xg[14] = {0, 0, 0, (float)(-0.25*PI), 0, (float)(0.25*PI), (float)(0.5*PI), 0, 0, 0, 0, 0, 0, 0};
    for (int b = 0; b < batch; b++){
        std::default_random_engine eng(seed0 + b);
        std::normal_distribution<double> dist(0.0, 0.001);
        for (int k = 0; k < N; k++){
            float *xk = x0 + ((size_t)b*N + k)*14;
            for (int i = 0; i < 7; i++){ xk[i] = q0[i]; }
            for (int i = 0; i < 7; i++){ xk[7+i] = static_cast<float>(dist(eng)); }
            float *uk = u0 + ((size_t)b*N + k)*7;
            for (int i = 0; i < 7; i++){ uk[i] = uu[i]; }
        }
        for (int i = 0; i < 14; i++){ xGoal[(size_t)b*14 + i] = xg[i]; }
    }
    return 0;
}

// initial guesses and goals of the reference's example for the other plants (WAFR_iLQR_examples.cu:19-33,72-78,87-90,110-115): one
// std::default_random_engine(seed) per problem, draws in knot order then state order -- the deterministic harness' inputs
extern "C" int pddp_make_inputs(int plant, int N, int batch, unsigned seed0, float *x0, float *u0, float *xGoal){
    if (plant == PDDP_PLANT_KUKA){ return pddp_make_inputs_kuka(N, batch, seed0, x0, u0, xGoal); }
    if (plant < PDDP_PLANT_PEND || plant > PDDP_PLANT_QUAD || !x0 || !u0 || !xGoal){ return PDDP_E_INVALID; }
    const int n = plant == PDDP_PLANT_PEND ? 2 : (plant == PDDP_PLANT_CART ? 4 : 12), m = plant == PDDP_PLANT_QUAD ? 4 : 1, np = n/2;
    for (int b = 0; b < batch; b++){
        std::default_random_engine eng(seed0 + b);
        std::normal_distribution<double> dist(0.0, 0.001);
        for (int k = 0; k < N; k++){
            float *xk = x0 + ((size_t)b*N + k)*n;
            if (plant == PDDP_PLANT_PEND){ xk[0] = 0.0f; xk[1] = static_cast<float>(dist(eng)); }
            else if (plant == PDDP_PLANT_CART){ xk[0] = 0.0f; xk[1] = 0.0f; xk[2] = static_cast<float>(dist(eng)); xk[3] = static_cast<float>(dist(eng)); }
            else { for (int i = 0; i < n; i++){ xk[i] = (i == 2) ? 0.5f : (i >= np ? static_cast<float>(dist(eng)) : 0.0f); } }
        }
        for (int k = 0; k < N; k++){ float *uk = u0 + ((size_t)b*N + k)*m; for (int i = 0; i < m; i++){ uk[i] = plant == PDDP_PLANT_QUAD ? (float)1.22625 : (float)0.01; } }
        float *g = xGoal + (size_t)b*n; for (int i = 0; i < n; i++){ g[i] = 0.0f; }
        if (plant == PDDP_PLANT_PEND){ g[0] = (float)3.1416; } else if (plant == PDDP_PLANT_CART){ g[1] = (float)3.1416; } else { g[0] = 7.0f; g[1] = 10.0f; g[2] = 0.5f; }
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------- plug-in unit calls
namespace {
// scratch device buffers of one unit call, released on every path out of it
struct Scratch {
    std::vector<void*> ptrs;
    ~Scratch(){ for (void *p : ptrs){ cudaFree(p); } }
    template <typename T> T *get(size_t count){ void *p = nullptr; if (cudaMalloc(&p, (count ? count : 1)*sizeof(T)) != cudaSuccess){ return nullptr; } ptrs.push_back(p); return static_cast<T*>(p); }
};
}
extern "C" int pddp_unit_dynamics(pddp_handle h, const float *x, const float *u, int nsamp, float *qdd){
    if (!h || nsamp < 1 || !x || !u || !qdd){ return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    const size_t n = h->S.n, m = h->S.m, np = h->S.npos; Scratch sc;
    float *dx = sc.get<float>(nsamp*n), *du = sc.get<float>(nsamp*m), *dq = sc.get<float>(nsamp*np);
    if (!dx || !du || !dq){ h->err = "cudaMalloc failed (unit call)"; return PDDP_E_CUDA; }
    CK(cudaMemcpy(dx, x, nsamp*n*4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(du, u, nsamp*m*4, cudaMemcpyHostToDevice));
    if (h->ops){ CKP(h->ops->unit_dynamics(&h->S, h->stream, dx, du, nsamp, dq)); }
    else { unit_dynamics_kernel<<<nsamp < 1184 ? nsamp : 1184, 32, h->smem_udyn, h->stream>>>(h->S.I, h->S.Tbody, h->S.grav, dx, du, nsamp, dq); }
    CK(cudaGetLastError()); CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(qdd, dq, nsamp*np*4, cudaMemcpyDeviceToHost));
    return 0;
}
static int unit_gradient_impl(pddp_handle h, const float *x, const float *u, int nsamp, float *AB, float *qdd, float *xnext){
    if (!h || nsamp < 1 || !x || !u){ return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    const size_t n = h->S.n, m = h->S.m, np = h->S.npos, nab = n*(n+m); Scratch sc;
    float *dx = sc.get<float>(nsamp*n), *du = sc.get<float>(nsamp*m), *dq = sc.get<float>(nsamp*np), *dab = sc.get<float>(nsamp*nab), *dxn = sc.get<float>(nsamp*n);
    if (!dx || !du || !dq || !dab || !dxn){ h->err = "cudaMalloc failed (unit call)"; return PDDP_E_CUDA; }
    CK(cudaMemcpy(dx, x, nsamp*n*4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(du, u, nsamp*m*4, cudaMemcpyHostToDevice));
    if (h->ops){ CKP(h->ops->unit_gradient(&h->S, h->stream, dx, du, nsamp, dab, dq, xnext ? dxn : nullptr)); }
    else {
        if (xnext){ h->err = "pddp_unit_integrator: plug-in plants only (the Kuka step is covered by the simulation traces)"; return PDDP_E_INVALID; }
        unit_gradient_kernel<<<nsamp < 888 ? nsamp : 888, 32, h->smem_ugrad, h->stream>>>(h->S.I, h->S.Tbody, h->S.grav, dx, du, nsamp, h->S.dt, dab, dq);
    }
    CK(cudaGetLastError()); CK(cudaStreamSynchronize(h->stream));
    if (AB){ CK(cudaMemcpy(AB, dab, nsamp*nab*4, cudaMemcpyDeviceToHost)); }
    if (qdd){ CK(cudaMemcpy(qdd, dq, nsamp*np*4, cudaMemcpyDeviceToHost)); }
    if (xnext){ CK(cudaMemcpy(xnext, dxn, nsamp*n*4, cudaMemcpyDeviceToHost)); }
    return 0;
}
extern "C" int pddp_unit_integrator_gradient(pddp_handle h, const float *x, const float *u, int nsamp, float *AB, float *qdd){
    if (!AB){ return PDDP_E_INVALID; }
    return unit_gradient_impl(h, x, u, nsamp, AB, qdd, nullptr);
}
extern "C" int pddp_unit_integrator(pddp_handle h, const float *x, const float *u, int nsamp, float *xnext){
    if (!xnext){ return PDDP_E_INVALID; }
    return unit_gradient_impl(h, x, u, nsamp, nullptr, nullptr, xnext);
}
extern "C" int pddp_unit_cost(pddp_handle h, const float *x, const float *u, const float *xGoal, const int *knot, int nsamp, float *J, float *H, float *g){
    if (!h || nsamp < 1 || !x || !u || !xGoal || !knot || !J || !H || !g){ return PDDP_E_INVALID; }
    if (!h->ops){ h->err = "pddp_unit_cost: plug-in plants only (the Kuka cost is fused into its kernels and covered by the phase traces)"; return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device));
    const size_t n = h->S.n, m = h->S.m, nm = n + m; Scratch sc;
    float *dx = sc.get<float>(nsamp*n), *du = sc.get<float>(nsamp*m), *dg = sc.get<float>(n), *dJ = sc.get<float>(nsamp), *dH = sc.get<float>(nsamp*nm*nm), *dgr = sc.get<float>(nsamp*nm);
    int *dk = sc.get<int>(nsamp);
    if (!dx || !du || !dg || !dJ || !dH || !dgr || !dk){ h->err = "cudaMalloc failed (unit call)"; return PDDP_E_CUDA; }
    CK(cudaMemcpy(dx, x, nsamp*n*4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(du, u, nsamp*m*4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dg, xGoal, n*4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dk, knot, nsamp*sizeof(int), cudaMemcpyHostToDevice));
    CKP(h->ops->unit_cost(&h->S, h->stream, dx, du, dg, dk, nsamp, dJ, dH, dgr));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(J, dJ, nsamp*4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(H, dH, nsamp*nm*nm*4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(g, dgr, nsamp*nm*4, cudaMemcpyDeviceToHost));
    return 0;
}

// ---------------------------------------------------------------------------------------------------- phase-level access
static bool resolve(pddp_handle h, const char *name, void **p, size_t *bytes){
    auto it = h->arrays.find(name);
    if (it == h->arrays.end()){ return false; }
    *p = it->second.first; *bytes = it->second.second; return true;
}
// "P","p" / "Pp","pp": the cost-to-go buffer the backward pass of the problem's current iteration writes / the one it seeds its blocks
// from (DevState::Pbuf).  The roles are per problem (parity of its iteration counter), so these arrays are gathered problem by problem.
static int pingpong(pddp_handle h, const char *name, void *host, long nbytes, bool to_device){
    const std::string s(name); DevState &S = h->S; const size_t B = S.B, N = S.N, n = S.n;
    const bool mat = (s == "P" || s == "Pp"), older = (s == "Pp" || s == "pp"); const size_t per = (mat ? N*n*n : N*n)*4;
    if ((size_t)nbytes != B*per){ h->err = std::string("bad array size: ") + name; return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device)); for (auto st : h->gstreams){ CK(cudaStreamSynchronize(st)); }
    std::vector<int> it(B); CK(cudaMemcpy(it.data(), S.iter, B*sizeof(int), cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < B; b++){
        const int idx = (it[b] & 1) ^ (older ? 1 : 0); char *dev = reinterpret_cast<char*>(mat ? S.Pbuf[idx] : S.pbuf[idx]) + b*per, *hp = static_cast<char*>(host) + b*per;
        if (to_device){ CK(cudaMemcpy(dev, hp, per, cudaMemcpyHostToDevice)); } else { CK(cudaMemcpy(hp, dev, per, cudaMemcpyDeviceToHost)); }
    }
    return 0;
}
static bool is_pingpong(const char *name){ const std::string s(name); return s == "P" || s == "Pp" || s == "p" || s == "pp"; }
// AB, H, g live in HBM with padded knot strides (kernels.cuh): the API speaks the reference's dense layout
static bool padded(pddp_handle h, const char *name, void **p, size_t *tile, size_t *stride){
    std::string s(name); DevState &S = h->S;
    if (s == "AB"){ *p = S.AB; *tile = (size_t)S.n*(S.n+S.m)*4; *stride = (size_t)S.ab_stride*4; return true; }
    if (s == "H"){ *p = S.H; *tile = (size_t)(S.n+S.m)*(S.n+S.m)*4; *stride = (size_t)S.h_stride*4; return true; }
    if (s == "g"){ *p = S.g; *tile = (size_t)(S.n+S.m)*4; *stride = (size_t)S.g_stride*4; return true; }
    return false;
}
extern "C" int pddp_set_array(pddp_handle h, const char *name, const void *src, long nbytes){
    if (!h){ return PDDP_E_INVALID; } void *p; size_t bytes;
    if (is_pingpong(name)){ return pingpong(h, name, const_cast<void*>(src), nbytes, true); }
    { size_t tile, stride; const size_t rows = (size_t)h->S.B*h->S.N;
      if (padded(h, name, &p, &tile, &stride)){
          if ((size_t)nbytes != tile*rows){ h->err = std::string("bad array size: ") + name; return PDDP_E_INVALID; }
          CK(cudaSetDevice(h->cfg.device)); CK(cudaStreamSynchronize(h->stream));
          CK(cudaMemcpy2D(p, stride, src, tile, tile, rows, cudaMemcpyHostToDevice)); return 0; } }
    if (!resolve(h, name, &p, &bytes) || (size_t)nbytes != bytes){ h->err = std::string("bad array name/size: ") + name; return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device)); CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice)); return 0;
}
extern "C" int pddp_get_array(pddp_handle h, const char *name, void *dst, long nbytes){
    if (!h){ return PDDP_E_INVALID; } void *p; size_t bytes;
    if (is_pingpong(name)){ return pingpong(h, name, dst, nbytes, false); }
    { size_t tile, stride; const size_t rows = (size_t)h->S.B*h->S.N;
      if (padded(h, name, &p, &tile, &stride)){
          if ((size_t)nbytes != tile*rows){ h->err = std::string("bad array size: ") + name; return PDDP_E_INVALID; }
          CK(cudaSetDevice(h->cfg.device)); for (auto st : h->gstreams){ CK(cudaStreamSynchronize(st)); }
          CK(cudaMemcpy2D(dst, tile, p, stride, tile, rows, cudaMemcpyDeviceToHost)); return 0; } }
    if (!resolve(h, name, &p, &bytes) || (size_t)nbytes != bytes){ h->err = std::string("bad array name/size: ") + name; return PDDP_E_INVALID; }
    CK(cudaSetDevice(h->cfg.device)); CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(dst, p, bytes, cudaMemcpyDeviceToHost)); return 0;
}
template <typename F>
static int timed_phase(pddp_handle h, F f){
    CK(cudaSetDevice(h->cfg.device));
    long l0 = h->launches;
    CK(cudaEventRecord(h->ev[4], h->stream));
    int rc = f(); if (rc){ return rc; }
    CK(cudaEventRecord(h->ev[5], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]); h->last_ms = ms; h->last_launches = (int)(h->launches - l0);
    return 0;
}
extern "C" int pddp_phase_load_init(pddp_handle h, const float *x0, const float *u0, const float *xGoal, int ignoreFirstDefectFlag){
    if (!h){ return PDDP_E_INVALID; }
    DevState &S = h->S; const size_t B = S.B, N = S.N, n = S.n, m = S.m;
    return timed_phase(h, [&](){
        const int clear = h->next_clear, rollout = h->next_rollout; h->next_clear = 1; h->next_rollout = 0;
        int rc = launch_reset(h, ignoreFirstDefectFlag, clear); if (rc){ return rc; }
        CK(cudaMemcpyAsync(S.xp, x0, B*N*n*4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(S.up, u0, B*N*m*4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(S.xGoal, xGoal, B*n*4, cudaMemcpyHostToDevice, h->stream));
        return launch_init(h, rollout);
    });
}
extern "C" int pddp_phase_backward_pass(pddp_handle h){ if (!h){ return PDDP_E_INVALID; } return timed_phase(h, [&](){ return launch_bp(h, h->stream, 0, h->S.B); }); }
extern "C" int pddp_phase_forward_sweep(pddp_handle h){ if (!h){ return PDDP_E_INVALID; } return timed_phase(h, [&](){ return launch_sweep(h, h->stream, 0, h->S.B); }); }
extern "C" int pddp_phase_forward_sim(pddp_handle h){ if (!h){ return PDDP_E_INVALID; } return timed_phase(h, [&](){ return launch_sim(h, h->stream, 0, h->S.B); }); }
extern "C" int pddp_phase_line_search(pddp_handle h){ if (!h){ return PDDP_E_INVALID; } return timed_phase(h, [&](){ return launch_select(h, h->stream, 0, h->S.B); }); }
extern "C" int pddp_phase_next_iteration(pddp_handle h){ if (!h){ return PDDP_E_INVALID; } return timed_phase(h, [&](){ return launch_nis(h, h->stream, 0, h->S.B); }); }
extern "C" int pddp_last_phase_stats(pddp_handle h, double *ms, int *launches){
    if (!h){ return PDDP_E_INVALID; } if (ms){ *ms = h->last_ms; } if (launches){ *launches = h->last_launches; } return 0;
}

#ifdef PDDP_SIM_TRACE
// trace build only: the clock64() stamps of the last forward-dynamics evaluation of block 0 (tools/sim_trace.py)
extern "C" int pddp_debug_simtrace(long long *out16){ return (int)cudaMemcpyFromSymbol(out16, pddp_simtrace, 16*sizeof(long long)); }
#endif
