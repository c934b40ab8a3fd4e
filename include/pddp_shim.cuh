// pddp_shim.cuh -- source-compatibility shim: the reference's solver-side templates on top of libpddp.so.
//
// examples/WAFR_iLQR_examples.cu includes ../config.cuh and calls, from its three test drivers,
//     allocateMemory_GPU<T>(...) / runiLQR_GPU<T>(...) / freeMemory_GPU<T>(...)      testGPU       (nisInitHelpers.cuh:766-772, DDPWrappers.cuh:8-21, nisInitHelpers.cuh:863-868)
//     allocateMemory_CPU[2]<T> / runiLQR_CPU[2]<T> / freeMemory_CPU[2]<T>            testCPU       (nisInitHelpers.cuh:884-965, DDPWrappers.cuh:140-365)
//     runSLQ_GPU<T>(...)                                                             testGPU_SLQ   (DDPWrappers.cuh:367)
// Including this header INSTEAD of ../config.cuh (one changed line: `#include "pddp_shim.cuh"`) makes the UNMODIFIED example
// compile and link against libpddp.so, for any PLANT (oracle/Makefile target `example_shim` does exactly that, and
// tests/test_gpu_shim.py runs the result):
//   * mode G runs through the library: same positional arguments, x0 / u0 in and out (nisInitHelpers.cuh:745-746), Jout / alphaOut slots
//     0..iters, the per-iteration timing arrays the example's statistics read, and the one-line summary of DDPWrappers.cuh:134 in the
//     same format (max_d included).  The ~45 work pointers the reference threads through every call are owned by the library, so the
//     shim hands back inert placeholders for them and keeps the solver handle in the slot of `d_x` (the first out-parameter).
//   * modes C / CS (the reference's CPU twins) and S (SLQ) are declared so that the example compiles, and stop with a message when
//     called: this library has no CPU path by design, and SLQ is out of scope (broken upstream, README.md:37).
//
// Compile-time macros of config.cuh are honoured as the front-end of the run-time pddp_config:
//   PLANT, NUM_TIME_STEPS, NUM_ALPHA, ALPHA_BASE, M_BLOCKS, MAX_ITER, TOL_COST, TOTAL_TIME, INTEGRATOR, RHO_INIT, RHO_MIN, RHO_MAX,
//   RHO_FACTOR, EXP_RED_MIN, EXP_RED_MAX, MAX_DEFECT_SIZE; for the arm (PLANT 4) also _Q1, _Q2, _R, _QF1, _QF2 and EE_COST with
//   _Q_EE1 ... _QF_xEE (with EE_COST 1 the caller's 6-float goal pose is what xGoal holds, as in the reference).  The other plants
//   carry their cost weights in their plant files (plants/cost_{pend,cart,quad}.cuh), here: pddp_default_config.
#pragma once
#include "pddp.h"
#include <cuda_runtime.h>
// what config.cuh brings in for its includers (utils/cudaUtils.h:37-44, utils/threadUtils.h, utils/exampleUtils.cuh, <sys/time.h>)
#include <sys/time.h>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#ifndef PLANT
#define PLANT 4
#endif
#ifndef EE_COST
#define EE_COST 0
#endif
#ifndef _Q_EE1                 // plants/cost_arm.cuh:106-117
#define _Q_EE1 0.1
#define _Q_EE2 0
#define _R_EE 0.0001
#define _QF_EE1 1000.0
#define _QF_EE2 0
#define _Q_xdEE 0.1
#define _QF_xdEE 1000.0
#define _Q_xEE 0.0
#define _QF_xEE 0.0
#endif
typedef float algType;
// config.cuh:21-61
#if PLANT == 1
    #define NUM_POS 1
    #define CONTROL_SIZE 1
    #ifndef RHO_INIT
    #define RHO_INIT 10.0
    #endif
#elif PLANT == 2
    #define NUM_POS 2
    #define CONTROL_SIZE 1
    #ifndef MAX_DEFECT_SIZE
    #define MAX_DEFECT_SIZE 0.75
    #endif
    #ifndef RHO_INIT
    #define RHO_INIT 10.0
    #endif
#elif PLANT == 3
    #define NUM_POS 6
    #define CONTROL_SIZE 4
    #ifndef ALPHA_BASE
    #define ALPHA_BASE 0.5
    #endif
    #ifndef NUM_ALPHA
    #define NUM_ALPHA 16
    #endif
    #ifndef RHO_INIT
    #define RHO_INIT 1.0
    #endif
#elif PLANT == 4
    #define NUM_POS 7
    #define CONTROL_SIZE 7
    #ifndef TOTAL_TIME
    #define TOTAL_TIME 0.5
    #endif
    #ifndef NUM_TIME_STEPS
    #define NUM_TIME_STEPS 64
    #endif
    #ifndef ALPHA_BASE
    #define ALPHA_BASE 0.5
    #endif
    #ifndef NUM_ALPHA
    #define NUM_ALPHA 16
    #endif
    #ifndef RHO_INIT
    #define RHO_INIT 12.5
    #endif
    #ifndef INTEGRATOR
    #define INTEGRATOR 1
    #endif
#else
    #error "PLANT: 1 pendulum, 2 cart-pole, 3 quadrotor, 4 Kuka iiwa14"
#endif
#define STATE_SIZE (2*NUM_POS)
// config.cuh:78-136
#ifndef INTEGRATOR
#define INTEGRATOR 3
#endif
#ifndef NUM_TIME_STEPS
#define NUM_TIME_STEPS 128
#endif
#ifndef TOTAL_TIME
#define TOTAL_TIME 4.0
#endif
#ifndef NUM_ALPHA
#define NUM_ALPHA 32
#endif
#ifndef ALPHA_BASE
#define ALPHA_BASE 0.75
#endif
#ifndef M_BLOCKS
#define M_BLOCKS 4
#endif
#define M_BLOCKS_B M_BLOCKS
#define M_BLOCKS_F M_BLOCKS
#ifndef MAX_ITER
#define MAX_ITER 100
#endif
#ifndef TOL_COST
#define TOL_COST 0.0001
#endif
#ifndef RHO_INIT
#define RHO_INIT 1.0
#endif
#ifndef RHO_MAX
#define RHO_MAX 10000000.0
#endif
#ifndef RHO_MIN
#define RHO_MIN 0.01
#endif
#ifndef RHO_FACTOR
#define RHO_FACTOR 1.25
#endif
#ifndef EXP_RED_MIN
#define EXP_RED_MIN 0.05
#endif
#ifndef EXP_RED_MAX
#define EXP_RED_MAX 1.25
#endif
#ifndef MAX_DEFECT_SIZE
#define MAX_DEFECT_SIZE 1.0
#endif
#ifndef _Q1
#define _Q1 0.1
#define _Q2 0.001
#define _R 0.0001
#define _QF1 1000.0
#define _QF2 1000.0
#endif
#ifndef USE_LIMITS_FLAG
#define USE_LIMITS_FLAG 0      // config.cuh:171-173
#endif
#ifndef USE_SMOOTH_ABS
#define USE_SMOOTH_ABS 0       // config.cuh:174-176
#endif
#ifndef SMOOTH_ABS_ALPHA
#define SMOOTH_ABS_ALPHA 0.2    // plants/cost_arm.cuh:119-121
#endif
#ifndef R_TL
#define R_TL 100.0             // plants/cost_arm.cuh:26-30
#define Q_PL 100.0
#define Q_VL 100.0
#endif
#define TIME_STEP (TOTAL_TIME/(NUM_TIME_STEPS-1))
#define NUM_STREAMS 18
#define DIM_x_r STATE_SIZE
#define DIM_u_r CONTROL_SIZE
#define DIM_x_c 1
#define DIM_u_c 1

inline void pddp_shim_die(pddp_handle h, const char *what){
    // the reference's gpuAssert prints and exits (utils/cudaUtils.cu:31-37); the shim keeps that behaviour at this level only
    std::fprintf(stderr, "GPUassert: %s: %s\n", what, pddp_last_error(h)); std::exit(1);
}

template <typename T>
void allocateMemory_GPU(T ***d_x, T ***h_d_x, T **d_xp, T **d_xp2, T ***d_u, T ***h_d_u, T **d_up, T **d_xGoal, T **xGoal, T **d_P, T **d_Pp, T **d_p, T **d_pp,
                        T **d_AB, T **d_H, T **d_g, T **d_KT, T **d_du, T ***d_d, T ***h_d_d, T **d_dp, T **d_dT, T **d_dM, T **d, T **d_ApBK, T **d_Bdu,
                        T **d_JT, T **J, T **d_dJexp, T **dJexp, T **alpha, T **d_alpha, int **alphaIndex, int **d_err, int **err,
                        int *ld_x, int *ld_u, int *ld_P, int *ld_p, int *ld_AB, int *ld_H, int *ld_g, int *ld_KT, int *ld_du, int *ld_d, int *ld_A,
                        cudaStream_t **streams, T **d_I = nullptr, T **d_Tbody = nullptr){
    static_assert(sizeof(T) == sizeof(float), "the library computes in float (algType, config.cuh:74)");
    pddp_config cfg; if (pddp_default_config(&cfg, PLANT, NUM_TIME_STEPS, 1) != 0){ pddp_shim_die(nullptr, "pddp_default_config"); }
    cfg.integrator = INTEGRATOR;
    cfg.n_alpha = NUM_ALPHA; cfg.alpha_base = (float)ALPHA_BASE; cfg.M = M_BLOCKS; cfg.max_iter = MAX_ITER; cfg.tol_cost = (float)TOL_COST;
    cfg.total_time = (float)TOTAL_TIME; cfg.rho_init = (float)RHO_INIT; cfg.rho_min = (float)RHO_MIN; cfg.rho_max = (float)RHO_MAX; cfg.rho_factor = (float)RHO_FACTOR;
    cfg.exp_red_min = (float)EXP_RED_MIN; cfg.exp_red_max = (float)EXP_RED_MAX; cfg.max_defect = (float)MAX_DEFECT_SIZE;
#if PLANT == 4
    cfg.Q1 = (float)_Q1; cfg.Q2 = (float)_Q2; cfg.R = (float)_R; cfg.QF1 = (float)_QF1; cfg.QF2 = (float)_QF2;
#endif
    cfg.ee_cost = EE_COST; cfg.Q_EE1 = (float)_Q_EE1; cfg.Q_EE2 = (float)_Q_EE2; cfg.QF_EE1 = (float)_QF_EE1; cfg.QF_EE2 = (float)_QF_EE2; cfg.R_EE = (float)_R_EE;
    cfg.Q_xdEE = (float)_Q_xdEE; cfg.QF_xdEE = (float)_QF_xdEE; cfg.Q_xEE = (float)_Q_xEE; cfg.QF_xEE = (float)_QF_xEE;
    cfg.use_limits = USE_LIMITS_FLAG; cfg.lim_Q_pos = (float)Q_PL; cfg.lim_Q_vel = (float)Q_VL; cfg.lim_R_tau = (float)R_TL;
    cfg.use_smooth_abs = USE_SMOOTH_ABS; cfg.smooth_abs_alpha = SMOOTH_ABS_ALPHA;
    pddp_handle h = nullptr;
    if (pddp_create(&cfg, &h) != 0){ pddp_shim_die(nullptr, "pddp_create"); }
    *d_x = reinterpret_cast<T**>(h);                                   // the handle travels in the d_x slot
    *h_d_x = nullptr; *d_xp = nullptr; *d_xp2 = nullptr; *d_u = nullptr; *h_d_u = nullptr; *d_up = nullptr; *d_xGoal = nullptr;
    *xGoal = (T*)std::calloc(STATE_SIZE, sizeof(T));                   // EE_COST: the first six floats are the goal pose
    *d_P = *d_Pp = *d_p = *d_pp = *d_AB = *d_H = *d_g = *d_KT = *d_du = nullptr; *d_d = nullptr; *h_d_d = nullptr;
    *d_dp = *d_dT = *d_dM = nullptr; *d = (T*)std::calloc(NUM_ALPHA, sizeof(T)); *d_ApBK = *d_Bdu = *d_JT = nullptr;
    *J = (T*)std::calloc(NUM_ALPHA, sizeof(T)); *d_dJexp = nullptr; *dJexp = (T*)std::calloc(2*M_BLOCKS, sizeof(T));
    *alpha = (T*)std::malloc(NUM_ALPHA*sizeof(T)); *d_alpha = nullptr; *alphaIndex = (int*)std::calloc(1, sizeof(int)); *d_err = nullptr; *err = (int*)std::calloc(M_BLOCKS, sizeof(int));
    *ld_x = STATE_SIZE; *ld_u = CONTROL_SIZE; *ld_P = STATE_SIZE; *ld_p = STATE_SIZE; *ld_AB = STATE_SIZE; *ld_H = STATE_SIZE + CONTROL_SIZE; *ld_g = STATE_SIZE + CONTROL_SIZE;
    *ld_KT = STATE_SIZE; *ld_du = CONTROL_SIZE; *ld_d = STATE_SIZE; *ld_A = STATE_SIZE;
    *streams = nullptr; if (d_I){ *d_I = nullptr; } if (d_Tbody){ *d_Tbody = nullptr; }
}

template <typename T>
void runiLQR_GPU(T *x0, T *u0, T *KT0, T *P0, T *p0, T *d0, T *xGoal, T *Jout, int *alphaOut, int forwardRolloutFlag, int clearVarsFlag, int ignoreFirstDefectFlag,
                 double *tTime, double *simTime, double *sweepTime, double *bpTime, double *nisTime, double *initTime, cudaStream_t *streams,
                 T **d_x, T **h_d_x, T *d_xp, T *d_xp2, T **d_u, T **h_d_u, T *d_up,
                 T *d_P, T *d_p, T *d_Pp, T *d_pp, T *d_AB, T *d_H, T *d_g, T *d_KT, T *d_du,
                 T **d_d, T **h_d_d, T *d_dp, T *d_dT, T *d, T *d_ApBK, T *d_Bdu, T *d_dM,
                 T *alpha, T *d_alpha, int *alphaIndex, T *d_JT, T *J, T *dJexp, T *d_dJexp, T *d_xGoal,
                 int *err, int *d_err, int ld_x, int ld_u, int ld_P, int ld_p, int ld_AB, int ld_H, int ld_g, int ld_KT, int ld_du, int ld_d, int ld_A,
                 T *d_I = nullptr, T *d_Tbody = nullptr){
    pddp_handle h = reinterpret_cast<pddp_handle>(d_x);
    std::vector<float> Jtmp(MAX_ITER+1); std::vector<int> atmp(MAX_ITER+1); int iters = 0; double times[6];
    // loadVarsGPU reads KT0, P0, p0, d0 only when clearVarsFlag = 0 (nisInitHelpers.cuh:622-631)
    if (!clearVarsFlag && pddp_set_warm_start(h, KT0, P0, p0, d0) != 0){ pddp_shim_die(h, "pddp_set_warm_start"); }
    if (pddp_solve(h, x0, u0, xGoal, forwardRolloutFlag, clearVarsFlag, ignoreFirstDefectFlag, x0, u0, Jtmp.data(), atmp.data(), &iters, times) != 0){ pddp_shim_die(h, "pddp_solve"); }
    // the reference only writes the slots it used (Jout[0..iters], alphaOut[0..iters])
    for (int i = 0; i <= iters; i++){ Jout[i] = Jtmp[i]; alphaOut[i] = atmp[i]; }
    *alphaIndex = atmp[iters] < 0 ? 0 : atmp[iters];
    // per-iteration times of each phase, as the reference's arrays hold them (device time here, host clock there): simTime / sweepTime /
    // bpTime of iteration i in slot i-1, nisTime of the setup that follows it in the same slot (DDPWrappers.cuh:60-107)
    *tTime = times[0]; *initTime = times[5];
    const int cnt = pddp_last_iteration_times(h, simTime, sweepTime, bpTime, nisTime, MAX_ITER);
    for (int i = cnt < 0 ? 0 : cnt; i < MAX_ITER; i++){ simTime[i] = sweepTime[i] = bpTime[i] = nisTime[i] = 0.0; }
    float max_d = 0.f; if (pddp_final_max_defect(h, &max_d) != 0){ pddp_shim_die(h, "pddp_final_max_defect"); }
    if (d){ d[*alphaIndex] = max_d; }
    std::printf("GPU Parallel blocks:[%d] t:[%f] with FP[%f], FS[%f], BP[%f], NIU[%f] Xf:[%.4f, %.4f] iters:[%d] cost:[%f] max_d[%f]\n",
                M_BLOCKS_B, *tTime, *simTime, *sweepTime, *bpTime, *nisTime, x0[ld_x*(NUM_TIME_STEPS-1)], x0[ld_x*(NUM_TIME_STEPS-1)+1], iters, (double)Jtmp[iters], (double)max_d);
}

template <typename T>
void freeMemory_GPU(T **d_x, T **h_d_x, T *d_xp, T *d_xp2, T **d_u, T **h_d_u, T *d_up, T *xGoal, T *d_xGoal, T *d_P, T *d_Pp, T *d_p, T *d_pp,
                    T *d_AB, T *d_H, T *d_g, T *d_KT, T *d_du, T **d_d, T **h_d_d, T *d_dp, T *d_dT, T *d_dM, T *d, T *d_ApBK, T *d_Bdu,
                    T *d_JT, T *J, T *d_dJexp, T *dJexp, T *alpha, T *d_alpha, int *alphaIndex, int *d_err, int *err,
                    cudaStream_t *streams, T *d_I = nullptr, T *d_Tbody = nullptr){
    pddp_destroy(reinterpret_cast<pddp_handle>(d_x));
    std::free(xGoal); std::free(d); std::free(J); std::free(dJexp); std::free(alpha); std::free(alphaIndex); std::free(err);
}

// ---- the reference's CPU twins and SLQ: declared so that the example compiles; calling them stops the program ------------------------
inline void pddp_shim_not_built(const char *what){
    std::fprintf(stderr, "%s: not built -- libpddp has no CPU path (by design: no CPU fallback) and no SLQ variant; run the example in mode G\n", what);
    std::exit(2);
}
template <typename T, typename... A> void allocateMemory_CPU(A...){ pddp_shim_not_built("allocateMemory_CPU"); }
template <typename T, typename... A> void allocateMemory_CPU2(A...){ pddp_shim_not_built("allocateMemory_CPU2"); }
template <typename T, typename... A> void runiLQR_CPU(A...){ pddp_shim_not_built("runiLQR_CPU"); }
template <typename T, typename... A> void runiLQR_CPU2(A...){ pddp_shim_not_built("runiLQR_CPU2"); }
template <typename T, typename... A> void freeMemory_CPU(A...){ pddp_shim_not_built("freeMemory_CPU"); }
template <typename T, typename... A> void freeMemory_CPU2(A...){ pddp_shim_not_built("freeMemory_CPU2"); }
template <typename T, typename... A> void runSLQ_GPU(A...){ pddp_shim_not_built("runSLQ_GPU"); }
