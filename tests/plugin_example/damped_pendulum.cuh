// damped_pendulum.cuh -- an example of a plant written by a user of the plug-in surface (parallel-ddp_b200/csrc/plugin/pddp_plugin.cuh):
// the reference's function names, argument order and calling convention (plants/dynamics_arm.cuh:2095-2097,2165-2167;
// plants/cost_arm.cuh:128-130,156-158), compiled around parallel-ddp_b200/csrc/plant_tu.cu into its own library and registered with
// pddp_load_plant_library (tests/test_gpu_plugin.py).  Pendulum with viscous friction:  theta_ddot = torque - 9.81 sin(theta) - 0.3 theta_dot.
#pragma once
#define NUM_POS 1
#define STATE_SIZE (2*NUM_POS)
#define CONTROL_SIZE 1

template <typename T> __host__ __device__ __forceinline__ void initI(T *s_I){ s_I[0] = static_cast<T>(0.3); }      // the model array carries the friction coefficient
template <typename T> __host__ __device__ __forceinline__ void initT(T *s_T){ return; }

template <typename T>
__host__ __device__ __forceinline__
void dynamics(T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody, T *s_eePos = nullptr, int reps = 1, T *s_eeVel = nullptr){
    int start, delta; singleLoopVals(&start, &delta);
    for (int r = start; r < reps; r += delta){
        s_qdd[r] = s_u[r] - static_cast<T>(9.81)*sin(s_x[STATE_SIZE*r]) - d_I[0]*s_x[STATE_SIZE*r + 1];
    }
}

template <typename T>
__host__ __device__ __forceinline__
void dynamicsGradient(T *s_dqdd, T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody){
#ifdef __CUDA_ARCH__
    if (threadIdx.x != 0 || threadIdx.y != 0){ return; }
#endif
    if (s_qdd != nullptr){ dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody); }
    s_dqdd[0] = -static_cast<T>(9.81)*cos(s_x[0]);
    s_dqdd[1] = -d_I[0];
    s_dqdd[2] = 1;
}

template <typename T>
__host__ __device__ __forceinline__
T costFunc(T *xk, T *uk, T *xgk, int k, T Q1, T Q2, T R, T QF1, T QF2){
    const bool last = (k == NUM_TIME_STEPS - 1);
    const T e0 = xk[0] - xgk[0], e1 = xk[1] - xgk[1];
    T cost = (last ? QF1 : Q1)*e0*e0 + (last ? QF2 : Q2)*e1*e1;
    if (!last){ cost += R*uk[0]*uk[0]; }
    return static_cast<T>(0.5)*cost;
}

template <typename T>
__host__ __device__ __forceinline__
void costGrad(T *Hk, T *gk, T *xk, T *uk, T *xgk, int k, int ld_H, T Q1, T Q2, T R, T QF1, T QF2){
    const bool last = (k == NUM_TIME_STEPS - 1);
    const T w[3] = {last ? QF1 : Q1, last ? QF2 : Q2, last ? static_cast<T>(0) : R};
    for (int i = 0; i < 3; i++){
        for (int j = 0; j < 3; j++){ Hk[i*ld_H + j] = (i == j) ? w[i] : static_cast<T>(0); }
        gk[i] = w[i]*(i < 2 ? xk[i] - xgk[i] : uk[0]);
    }
}
