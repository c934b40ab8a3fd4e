// Compile check of include/pddp_shim.cuh: the call sequence of examples/WAFR_iLQR_examples.cu::testGPU (lines 301-361)
// with the reference's argument lists, against libpddp.so.  Run on a GPU box: prints one summary line per solve.
#ifndef EE_COST
#define EE_COST 0        // -DEE_COST=1: the end-effector cost build of the example (WAFR_iLQR_examples.cu:37-43,101-104)
#endif
#define TOL_COST 0.0
#define NUM_TIME_STEPS 32
#define MAX_ITER 5
#include "../include/pddp_shim.cuh"
#ifndef PI
#define PI 3.14159        // the example defines it itself (WAFR_iLQR_examples.cu:36)
#endif
int main(){
    typedef algType T;
    int ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A; cudaStream_t *streams; T *alpha, *d_alpha; int *alphaIndex;
    T *d_P, *d_p, *d_Pp, *d_pp, *d_AB, *d_H, *d_g, *d_KT, *d_du; T **d_x, **d_u, **h_d_x, **h_d_u, *d_xp, *d_xp2, *d_up, *d_JT, *J;
    T **d_d, **h_d_d, *d_dp, *d_dT, *d, *d_ApBK, *d_Bdu, *d_dM; int *err, *d_err; T *dJexp, *d_dJexp; T *xGoal, *d_xGoal; T *d_I, *d_Tbody;
    allocateMemory_GPU<T>(&d_x, &h_d_x, &d_xp, &d_xp2, &d_u, &h_d_u, &d_up, &d_xGoal, &xGoal, &d_P, &d_Pp, &d_p, &d_pp, &d_AB, &d_H, &d_g, &d_KT, &d_du,
                          &d_d, &h_d_d, &d_dp, &d_dT, &d_dM, &d, &d_ApBK, &d_Bdu, &d_JT, &J, &d_dJexp, &dJexp, &alpha, &d_alpha, &alphaIndex, &d_err, &err,
                          &ld_x, &ld_u, &ld_P, &ld_p, &ld_AB, &ld_H, &ld_g, &ld_KT, &ld_du, &ld_d, &ld_A, &streams, &d_I, &d_Tbody);
    std::vector<T> x0(ld_x*NUM_TIME_STEPS), u0(ld_u*NUM_TIME_STEPS); T Jout[MAX_ITER+1]; int alphaOut[MAX_ITER+1];
    double tTime, initTime, fsim[MAX_ITER], fsw[MAX_ITER], bp[MAX_ITER], nis[MAX_ITER];
    pddp_make_inputs_kuka(NUM_TIME_STEPS, 1, 0, x0.data(), u0.data(), xGoal);
#if EE_COST
    { const T pose[6] = {(T)0.3638, (T)0.0, (T)1.0628, (T)(0.5*PI), (T)0.0, (T)(0.5*PI)}; for (int i = 0; i < STATE_SIZE; i++){ xGoal[i] = i < 6 ? pose[i] : (T)0; } }
#endif
    runiLQR_GPU<T>(x0.data(), u0.data(), nullptr, nullptr, nullptr, nullptr, xGoal, Jout, alphaOut, 0, 1, 1, &tTime, fsim, fsw, bp, nis, &initTime, streams,
                   d_x, h_d_x, d_xp, d_xp2, d_u, h_d_u, d_up, d_P, d_p, d_Pp, d_pp, d_AB, d_H, d_g, d_KT, d_du, d_d, h_d_d, d_dp, d_dT, d, d_ApBK, d_Bdu, d_dM,
                   alpha, d_alpha, alphaIndex, d_JT, J, dJexp, d_dJexp, d_xGoal, err, d_err, ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A, d_I, d_Tbody);
    freeMemory_GPU<T>(d_x, h_d_x, d_xp, d_xp2, d_u, h_d_u, d_up, xGoal, d_xGoal, d_P, d_Pp, d_p, d_pp, d_AB, d_H, d_g, d_KT, d_du, d_d, h_d_d, d_dp, d_dM, d_dT, d, d_ApBK, d_Bdu,
                      d_JT, J, d_dJexp, dJexp, alpha, d_alpha, alphaIndex, d_err, err, streams, d_I, d_Tbody);
    std::printf("J: %f -> %f, alpha trace:", Jout[0], Jout[MAX_ITER]); for (int i = 0; i <= MAX_ITER; i++){ std::printf(" %d", alphaOut[i]); } std::printf("\n");
    return 0;
}
