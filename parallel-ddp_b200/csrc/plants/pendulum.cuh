// plants/pendulum.cuh -- PLANT 1: torque-driven simple pendulum, written against the plug-in surface (plugin/pddp_plugin.cuh).
// Model and cost of the reference's plants/dynamics_pend.cuh:30-51 and plants/cost_pend.cuh:20-54, with the solver-era (v0.2)
// signatures of plants/dynamics_arm.cuh:2095-2097,2165-2167 and plants/cost_arm.cuh:128-130,156-158.
//   state [theta, theta_dot], control [torque];   theta_ddot = torque + g sin(theta),  g = -9.81 (unit mass, unit length)
// Arithmetic: the literals are doubles as in the reference, so g sin(theta) and the squared errors are evaluated in double and
// rounded to T on assignment; pow(., 2) is the reference's way of squaring.
#pragma once
#define NUM_POS 1
#define STATE_SIZE (2*NUM_POS)
#define CONTROL_SIZE 1
#define PEND_G (-9.81)

template <typename T> __host__ __device__ __forceinline__ void initI(T *s_I){ return; }      // no rigid-body model data
template <typename T> __host__ __device__ __forceinline__ void initT(T *s_T){ return; }

// s_qdd[NUM_POS*r] for r < reps states stored back to back (the reference strides s_u by NUM_POS too, dynamics_pend.cuh:35)
template <typename T>
__host__ __device__ __forceinline__
void dynamics(T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody, T *s_eePos = nullptr, int reps = 1, T *s_eeVel = nullptr){
    int first, step; singleLoopVals(&first, &step);
    for (int r = first; r < reps; r += step){
        const T *x = s_x + STATE_SIZE*r, *u = s_u + NUM_POS*r;
        s_qdd[NUM_POS*r] = u[0] + PEND_G*sin(x[0]);
    }
}

// s_dqdd = [d/dtheta, d/dtheta_dot, d/dtorque] of theta_ddot; one thread of the group does the work (dynamics_pend.cuh:41-51)
template <typename T>
__host__ __device__ __forceinline__
void dynamicsGradient(T *s_dqdd, T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody){
#ifdef __CUDA_ARCH__
    if (threadIdx.x != 0 || threadIdx.y != 0){ return; }
#endif
    if (s_qdd != nullptr){ dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody); }
    s_dqdd[0] = PEND_G*cos(s_x[0]);
    s_dqdd[1] = 0.0;
    s_dqdd[2] = 1;
}

// 1/2 sum w_i (x_i - xg_i)^2 + 1/2 sum R u_i^2; the final knot takes QF1 (position) / QF2 (velocity) and no control term.
// Reference weights (cost_pend.cuh:20-24: the velocity falls through QR(i) to R): Q1 1.0, Q2 0.1, R 0.1, QF1 = QF2 1000.
template <typename T>
__host__ __device__ __forceinline__
T costFunc(T *xk, T *uk, T *xgk, int k, T Q1, T Q2, T R, T QF1, T QF2){
    const bool last = (k == NUM_TIME_STEPS - 1);
    T cost = 0.0;
    #pragma unroll
    for (int i = 0; i < STATE_SIZE; i++){ const T w = last ? (i < NUM_POS ? QF1 : QF2) : (i < NUM_POS ? Q1 : Q2); cost += w*pow(xk[i] - xgk[i], 2); }
    if (!last){
        #pragma unroll
        for (int i = 0; i < CONTROL_SIZE; i++){ cost += R*pow(uk[i], 2); }
    }
    return 0.5*cost;
}

// Hk: (n+m) x (n+m) with leading dimension ld_H (diagonal), gk: n+m
template <typename T>
__host__ __device__ __forceinline__
void costGrad(T *Hk, T *gk, T *xk, T *uk, T *xgk, int k, int ld_H, T Q1, T Q2, T R, T QF1, T QF2){
    const bool last = (k == NUM_TIME_STEPS - 1);
    #pragma unroll
    for (int i = 0; i < STATE_SIZE + CONTROL_SIZE; i++){
        const T w = i < STATE_SIZE ? (last ? (i < NUM_POS ? QF1 : QF2) : (i < NUM_POS ? Q1 : Q2)) : (last ? static_cast<T>(0.0) : R);
        #pragma unroll
        for (int j = 0; j < STATE_SIZE + CONTROL_SIZE; j++){ Hk[i*ld_H + j] = (i != j) ? static_cast<T>(0.0) : w; }
        gk[i] = w*(i < STATE_SIZE ? xk[i] - xgk[i] : uk[i - STATE_SIZE]);
    }
}
