// pddp_shim.cuh -- source-compatibility shim: the reference's solver-side templates on top of libpddp.so.
//
// examples/WAFR_iLQR_examples.cu calls three templates with ~50 raw work pointers each:
//     allocateMemory_GPU<T>(...)   DDPHelpers/nisInitHelpers.cuh:766-772
//     runiLQR_GPU<T>(...)          DDPHelpers/DDPWrappers.cuh:8-21
//     freeMemory_GPU<T>(...)       DDPHelpers/nisInitHelpers.cuh:863-868
// Including this header INSTEAD of the reference's config.cuh keeps those call sites compiling unchanged (same names,
// same positional arguments).  The work pointers the reference threads through every call are owned by the library
// here, so the shim hands back inert placeholders for them and keeps the one thing that matters -- the solver handle --
// in the slot of `d_x` (the first out-parameter).  x0/u0 are in/out exactly as in the reference (the solution overwrites
// them, nisInitHelpers.cuh:745-746), Jout/alphaOut need MAX_ITER+1 slots, the six timing outputs are filled from the
// library's CUDA-event timings, and the one-line summary of DDPWrappers.cuh:134 is printed in the same format.
//
// Compile-time macros of config.cuh are honoured as the front-end of the run-time pddp_config:
//   NUM_TIME_STEPS, NUM_ALPHA, ALPHA_BASE, M_BLOCKS, MAX_ITER, TOL_COST, TOTAL_TIME, RHO_INIT, RHO_MIN, RHO_MAX,
//   RHO_FACTOR, EXP_RED_MIN, EXP_RED_MAX, MAX_DEFECT_SIZE, _Q1, _Q2, _R, _QF1, _QF2, EE_COST with _Q_EE1 ... _QF_xEE   (PLANT must be 4;
//   with EE_COST 1 the caller's 6-float goal pose is what xGoal holds, as in the reference).
#pragma once
#include "pddp.h"
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#ifndef PLANT
#define PLANT 4
#endif
#if PLANT != 4
#error "pddp_shim.cuh: only PLANT 4 (Kuka iiwa14) is built"
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
#define NUM_POS 7
#define STATE_SIZE 14
#define CONTROL_SIZE 7
#ifndef NUM_TIME_STEPS
#define NUM_TIME_STEPS 64
#endif
#ifndef TOTAL_TIME
#define TOTAL_TIME 0.5
#endif
#ifndef NUM_ALPHA
#define NUM_ALPHA 16
#endif
#ifndef ALPHA_BASE
#define ALPHA_BASE 0.5
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
#define RHO_INIT 12.5
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
#define TIME_STEP (TOTAL_TIME/(NUM_TIME_STEPS-1))
#define NUM_STREAMS 18
#define PI 3.14159
#define DIM_x_r STATE_SIZE
#define DIM_u_r CONTROL_SIZE

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
    pddp_config cfg; pddp_default_config_kuka(&cfg, NUM_TIME_STEPS, 1);
    cfg.n_alpha = NUM_ALPHA; cfg.alpha_base = (float)ALPHA_BASE; cfg.M = M_BLOCKS; cfg.max_iter = MAX_ITER; cfg.tol_cost = (float)TOL_COST;
    cfg.total_time = (float)TOTAL_TIME; cfg.rho_init = (float)RHO_INIT; cfg.rho_min = (float)RHO_MIN; cfg.rho_max = (float)RHO_MAX; cfg.rho_factor = (float)RHO_FACTOR;
    cfg.exp_red_min = (float)EXP_RED_MIN; cfg.exp_red_max = (float)EXP_RED_MAX; cfg.max_defect = (float)MAX_DEFECT_SIZE;
    cfg.Q1 = (float)_Q1; cfg.Q2 = (float)_Q2; cfg.R = (float)_R; cfg.QF1 = (float)_QF1; cfg.QF2 = (float)_QF2;
    cfg.ee_cost = EE_COST; cfg.Q_EE1 = (float)_Q_EE1; cfg.Q_EE2 = (float)_Q_EE2; cfg.QF_EE1 = (float)_QF_EE1; cfg.QF_EE2 = (float)_QF_EE2; cfg.R_EE = (float)_R_EE;
    cfg.Q_xdEE = (float)_Q_xdEE; cfg.QF_xdEE = (float)_QF_xdEE; cfg.Q_xEE = (float)_Q_xEE; cfg.QF_xEE = (float)_QF_xEE;
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
    // whole-solve device times of each phase; the reference stores per-iteration host times, so the totals go to slot 0
    *tTime = times[0]; simTime[0] = times[1]; sweepTime[0] = times[2]; bpTime[0] = times[3]; nisTime[0] = times[4]; *initTime = times[5];
    for (int i = 1; i < iters; i++){ simTime[i] = sweepTime[i] = bpTime[i] = nisTime[i] = 0.0; }
    std::printf("GPU Parallel blocks:[%d] t:[%f] with FP[%f], FS[%f], BP[%f], NIU[%f] Xf:[%.4f, %.4f] iters:[%d] cost:[%f] max_d[%f]\n",
                M_BLOCKS_B, *tTime, *simTime, *sweepTime, *bpTime, *nisTime, x0[ld_x*(NUM_TIME_STEPS-1)], x0[ld_x*(NUM_TIME_STEPS-1)+1], iters, (double)Jtmp[iters], 0.0);
}

template <typename T>
void freeMemory_GPU(T **d_x, T **h_d_x, T *d_xp, T *d_xp2, T **d_u, T **h_d_u, T *d_up, T *xGoal, T *d_xGoal, T *d_P, T *d_Pp, T *d_p, T *d_pp,
                    T *d_AB, T *d_H, T *d_g, T *d_KT, T *d_du, T **d_d, T **h_d_d, T *d_dp, T *d_dT, T *d_dM, T *d, T *d_ApBK, T *d_Bdu,
                    T *d_JT, T *J, T *d_dJexp, T *dJexp, T *alpha, T *d_alpha, int *alphaIndex, int *d_err, int *err,
                    cudaStream_t *streams, T *d_I = nullptr, T *d_Tbody = nullptr){
    pddp_destroy(reinterpret_cast<pddp_handle>(d_x));
    std::free(xGoal); std::free(d); std::free(J); std::free(dJexp); std::free(alpha); std::free(alphaIndex); std::free(err);
}
