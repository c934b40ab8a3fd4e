// plants/cartpole.cuh -- PLANT 2: cart with an unactuated pole, written against the plug-in surface (plugin/pddp_plugin.cuh).
// Model and cost of the reference's plants/dynamics_cart.cuh:30-76 and plants/cost_cart.cuh:19-68 with the solver-era (v0.2)
// signatures (plants/dynamics_arm.cuh:2095-2097,2165-2167; plants/cost_arm.cuh:128-130,156-158).
//   state [x, theta, x_dot, theta_dot], control [force on the cart]
//   [ mc+mp      mp l cos  ] [x_ddot    ]   [ mp l sin theta_dot^2 + f ]
//   [ mp l cos   mp l^2    ] [theta_ddot] = [ mp l sin g               ]        mc 10, mp 1, l 0.5, g -9.81
// solved with the closed-form 2x2 inverse.  The products with the double literals (mp l, g) are evaluated in double and rounded
// to T on assignment, everything else is T arithmetic -- as in the reference.
#pragma once
#define NUM_POS 2
#define STATE_SIZE (2*NUM_POS)
#define CONTROL_SIZE 1
#define CART_G (-9.81)
#define CART_MC 10
#define CART_MP 1
#define CART_L 0.5
#define CART_MPL (CART_MP * CART_L)
#define CART_MPLL (CART_MPL * CART_L)

template <typename T> __host__ __device__ __forceinline__ void initI(T *s_I){ return; }
template <typename T> __host__ __device__ __forceinline__ void initT(T *s_T){ return; }

template <typename T>
__host__ __device__ __forceinline__
void dynamics(T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody, T *s_eePos = nullptr, int reps = 1, T *s_eeVel = nullptr){
    int first, step; singleLoopVals(&first, &step);
    for (int r = first; r < reps; r += step){
        const T *x = s_x + STATE_SIZE*r, *u = s_u + NUM_POS*r; T *qdd = s_qdd + NUM_POS*r;
        const T c = cos(x[1]), s = sin(x[1]), w2 = x[3]*x[3];
        const T m00 = CART_MC + CART_MP, m11 = CART_MPLL, m01 = CART_MPL*c;
        const T ps = CART_MPL*s, f0 = ps*w2 + u[0], f1 = ps*CART_G;
        const T idet = 1/(m00*m11 - m01*m01);
        qdd[0] = idet*(m11*f0 - m01*f1);
        qdd[1] = idet*(m00*f1 - m01*f0);
    }
}

// s_dqdd: 2 x 5 column-major, columns [x, theta, x_dot, theta_dot, f]; one thread of the group (dynamics_cart.cuh:46-76)
template <typename T>
__host__ __device__ __forceinline__
void dynamicsGradient(T *s_dqdd, T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody){
#ifdef __CUDA_ARCH__
    if (threadIdx.x != 0 || threadIdx.y != 0){ return; }
#endif
    if (s_qdd != nullptr){ dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody); }
    const T c = cos(s_x[1]), s = sin(s_x[1]);
    const T w = s_x[3], w2 = w*w;
    const T m00 = CART_MC + CART_MP, m11 = CART_MPLL;
    const T m01 = CART_MPL*c, ps = CART_MPL*s;
    const T f0 = ps*w2 + s_u[0], f1 = ps*CART_G;
    const T det = m00*m11 - m01*m01, idet = 1/det;
    const T a0 = m11*f0 - m01*f1, a1 = m00*f1 - m01*f0;          // adjugate times right-hand side
    const T d1_du = idet*(-m01), d1_dw = idet*(-2*m01*ps*w);
    const T d0_du = idet*(m11),  d0_dw = idet*(2*m11*ps*w);
    const T m01_dth = -ps, f0_dth = m01*w2, f1_dth = m01*CART_G;
    const T a1_dth = m00*f1_dth - (m01_dth*f0 + m01*f0_dth);
    const T a0_dth = m11*f0_dth - (m01_dth*f1 + m01*f1_dth);
    const T idet_dth = -2*m01*ps*idet*idet;
    const T d1_dth = idet*a1_dth + idet_dth*a1;
    const T d0_dth = idet*a0_dth + idet_dth*a0;
    s_dqdd[0] = 0;       s_dqdd[1] = 0;
    s_dqdd[2] = d0_dth;  s_dqdd[3] = d1_dth;
    s_dqdd[4] = 0;       s_dqdd[5] = 0;
    s_dqdd[6] = d0_dw;   s_dqdd[7] = d1_dw;
    s_dqdd[8] = d0_du;   s_dqdd[9] = d1_du;
}

// Reference weights (cost_cart.cuh:31-37, N other than 256 / 512): Q1 0.01 (x and theta), Q2 0.001, R 0.0001, QF1 = QF2 1000.
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
