// plants/quadrotor.cuh -- PLANT 3: quadrotor with roll / pitch / yaw attitude, written against the plug-in surface
// (plugin/pddp_plugin.cuh).  Model and cost of the reference's plants/dynamics_quad.cuh:42-168 and plants/cost_quad.cuh:19-55 with
// the solver-era (v0.2) signatures (plants/dynamics_arm.cuh:2095-2097,2165-2167; plants/cost_arm.cuh:128-130,156-158).
//   state [x y z roll pitch yaw | their rates], control = the four rotor thrusts
//   mass 0.5, arm length 0.175, inertia diag(0.0023, 0.0023, 0.004), g -9.81
// The closed-form accelerations and their derivatives are the reference's expressions term by term and in its order (sums are
// evaluated left to right, double literals make a term double until it is rounded to T on assignment): the float results
// depend on that order, and parity with the reference is bit for bit.
#pragma once
#define NUM_POS 6
#define STATE_SIZE (2*NUM_POS)
#define CONTROL_SIZE 4
#define QUAD_G (-9.81)
#define QUAD_MASS 0.5
#define QUAD_INV_MASS 2
// weights of the rate terms of the cost (cost_quad.cuh:21-22 Q3 = Q4); Q1 weighs x y z, Q2 roll pitch yaw
#define QUAD_QRATE 2.0
// (Izz - Ixx)/Ixx, arm*thrust gains and torque ratio of the attitude equations as the reference's literals
#define QUAD_KI 0.73913043584
#define QUAD_KA 76.0869565217
#define QUAD_KB 6.125

template <typename T> __host__ __device__ __forceinline__ void initI(T *s_I){ return; }
template <typename T> __host__ __device__ __forceinline__ void initT(T *s_T){ return; }

template <typename T>
__host__ __device__ __forceinline__
void dynamics(T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody, T *s_eePos = nullptr, int reps = 1, T *s_eeVel = nullptr){
    int first, step; singleLoopVals(&first, &step);
    for (int r = first; r < reps; r += step){
        const T *x = s_x + STATE_SIZE*r, *u = s_u + NUM_POS*r; T *a = s_qdd + NUM_POS*r;
        const T c3 = cos(x[3]), c4 = cos(x[4]), c5 = cos(x[5]);
        const T s3 = sin(x[3]), s4 = sin(x[4]), s5 = sin(x[5]);
        const T pq = x[9]*x[10], pr = x[9]*x[11];
        const T qr = x[10]*x[11], rr = x[11]*x[11];
        // translation: total thrust along the body z axis, rotated into the world
        const T thrust = u[0] + u[1] + u[2] + u[3];
        a[0] = QUAD_INV_MASS*thrust*(s3*s5 + c3*c5*s4);
        a[1] = -QUAD_INV_MASS*thrust*(c5*s3 - c3*s4*s5);
        a[2] = QUAD_G + QUAD_INV_MASS*thrust*c3*c4;
        // attitude (the reference reads the un-strided s_x / s_u in four places, dynamics_quad.cuh:60-63: same values for reps = 1)
        const T yawU = u[0] - u[1] + u[2] - u[3], pitchU = u[2] - u[0];
        const T ic4 = 1/c4, c3c3 = c3*c3, c3c4 = c3*c4, sin2r = 2.0*s3*c3, cos2r = cos(2.0*s_x[3]);
        a[3] = ic4*(0.0005434782609*(32000.0*qr + 140000.0*(s_u[1] - s_u[3])*c4 - 28320.0*pq*s4 - 30160.0*qr*c3c3 + 1127.0*yawU*c3*s4 - 140000.0*pitchU*s3*s4 + 30160.0*pq*c3c3*s4 - 30160.0*qr*c3c4*c3c4 + 30160.0*s_x[10]*s_x[10]*c3*c4*s3 - 30160.0*rr*c3*c4*s3 + 30160.0*pr*c3*c4*s3*s4));
        a[4] = 76.08695652*pitchU*c3 - 0.6125*yawU*s3 - 1.0*pr*c4 - 8.195652174*pq*sin2r - 16.39130435*rr*c3c3*c4*s4 + 16.39130435*pr*c3c3*c4 + 16.39130435*qr*c3*s3*s4;
        a[5] = -ic4*(0.0005434782609*(13240.0*pq - 1127.0*yawU*c3 - 140000.0*pitchU*s3 - 16920.0*qr*s4 + 7540.0*rr*sin2r*2.0*s4*c4 - 15080.0*pq*cos2r - 15080.0*pr*sin2r*c4 + 15080.0*qr*cos2r*s4));
    }
}

// s_dqdd: 6 x 16 column-major, columns [x y z roll pitch yaw | rates | u0..u3]; one thread of the group (dynamics_quad.cuh:70-168)
template <typename T>
__host__ __device__ __forceinline__
void dynamicsGradient(T *s_dqdd, T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody){
#ifdef __CUDA_ARCH__
    if (threadIdx.x != 0 || threadIdx.y != 0){ return; }
#endif
    if (s_qdd != nullptr){ dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody); }
    for (int i = 0; i < NUM_POS*(STATE_SIZE + CONTROL_SIZE); i++){ s_dqdd[i] = 0; }
    #define D_(row, col) s_dqdd[(row) + (col)*NUM_POS]
    const T *x = s_x, *u = s_u;
    const T s3 = sin(x[3]), s4 = sin(x[4]), s5 = sin(x[5]);
    const T c3 = cos(x[3]), c4 = cos(x[4]), c5 = cos(x[5]);

    // linear accelerations: derivative of (thrust/m) R e_z with respect to the attitude and the four thrusts
    const T tm = (u[0] + u[1] + u[2] + u[3])*QUAD_INV_MASS;
    const T ax_du = (s3*s5 + c3*c5*s4)*QUAD_INV_MASS;
    D_(0,3) = (c3*s5 - c5*s3*s4)*tm;
    D_(0,4) = (c3*c4*c5)*tm;
    D_(0,5) = (c5*s3 - c3*s4*s5)*tm;
    D_(0,12) = ax_du; D_(0,13) = ax_du; D_(0,14) = ax_du; D_(0,15) = ax_du;

    const T ay_du = -(c5*s3 - c3*s4*s5)*QUAD_INV_MASS;
    D_(1,3) = -(c3*c5 + s3*s4*s5)*tm;
    D_(1,4) = (c3*c4*s5)*tm;
    D_(1,5) = (s3*s5 + c3*c5*s4)*tm;
    D_(1,12) = ay_du; D_(1,13) = ay_du; D_(1,14) = ay_du; D_(1,15) = ay_du;

    const T az_du = (c3*c4)/QUAD_MASS;
    D_(2,3) = -(c4*s3)*tm;
    D_(2,4) = -(c3*s4)*tm;
    D_(2,12) = az_du; D_(2,13) = az_du; D_(2,14) = az_du; D_(2,15) = az_du;

    // angular accelerations: shared products first
    const T rr = x[11]*x[11];
    const T pq = x[9]*x[10];
    const T pr = x[9]*x[11];
    const T qr = x[10]*x[11];
    const T c3c3 = c3*c3;
    const T c4c4 = c4*c4;
    const T c3c3c4 = c3c3*c4;
    const T c3c3c4c4 = c3c3*c4c4;
    const T c3c4s3 = c3*c4*s3;
    const T c3s3s4 = c3*s3*s4;
    const T c3c4s3s4 = c3c4s3*s4;
    const T ic4 = 1.0/c4;
    const T ic4ic4 = ic4/c4;
    const T du02 = -u[0] + u[2];
    const T du0123 = -u[0] + u[1] - u[2] + u[3];

    // roll
    const T rollA = QUAD_KB*c3*s4*ic4;
    const T rollB = QUAD_KA*s3;
    D_(3,3) = ic4*(QUAD_KA*c3*s4*du02 + QUAD_KB*s3*s4*du0123 + QUAD_KI*(x[10]*x[10]*(2*c3c3*c4 - c4) + rr*c4 - 2*rr*c3c3*c4 - pr*s4*c4 + 2.0*qr*s3*c3 + 2*pr*c3c3*c4*s4 - 2*x[11]*c3*c4c4*s3 - 2*x[10] - 2*pq*c3*s3*s4));
    D_(3,4) = ic4ic4*(pq + qr*s4 + QUAD_KA*s3*du02 - QUAD_KB*c3*du0123 + QUAD_KI*(pq*(1 + c3c3) + qr*(1 - c3c3*s4 + c3c3c4c4*s4) + pr*c3c4s3*c4c4));
    D_(3,9) = ic4*(x[10]*s4 + QUAD_KI*(x[10]*(c3c3 - 1)*s4 + x[11]*c3c4s3s4));
    D_(3,10) = ic4*(x[11] + x[9]*s4 - QUAD_KI*(x[11]*(1 + c3c3 + c3c3c4c4) - x[9]*(s4 + c3c3*s4) - 2*x[10]*c3c4s3));
    D_(3,11) = ic4*(x[10] - QUAD_KI*(x[10] + x[10]*c3c3 + x[9]*c3c4s3s4 - 2*x[11]*c3c4s3 - x[10]*c3c3c4c4));
    D_(3,12) = rollA - rollB;
    D_(3,13) = QUAD_KA - rollA;
    D_(3,14) = rollA + rollB;
    D_(3,15) = -QUAD_KA - rollA;

    // pitch
    const T pitchA = QUAD_KB*s3;
    const T pitchB = QUAD_KA*c3;
    D_(4,3) = QUAD_KB*c3*du0123 - QUAD_KA*s3*du02 + QUAD_KI*(pq*(1 - 2*c3c3) + qr*s4*(2*c3c3 - 1) + 2*rr*c3c4s3s4 - 2*pr*c3c4s3);
    D_(4,4) = QUAD_KI*(rr*c3c3 - 2.0*rr*c3c3c4c4 - c3c3*s4 + qr*c3c4s3) + pr*s4;
    D_(4,9) = QUAD_KI*(x[11]*c3c3c4 - x[10]*s3*c3) - x[11]*c4;
    D_(4,10) = QUAD_KI*(x[11]*c3s3s4 - x[9]*s3*c3);
    D_(4,11) = QUAD_KI*(x[9]*c3c3c4 + x[10]*c3s3s4 - 2.0*x[11]*c3c3c4*s4) - x[9]*c4;
    D_(4,12) = -pitchB - pitchA;
    D_(4,13) = pitchA;
    D_(4,14) = pitchB - pitchA;
    D_(4,15) = pitchA;

    // yaw
    const T yawA = -QUAD_KB*c3*ic4;
    const T yawB = QUAD_KA*s3*ic4;
    D_(5,3) = ic4*(QUAD_KA*c3*du02 + QUAD_KB*s3*du0123 + QUAD_KI*(rr*(s4*c4 - 2.0*c3c3c4*s4) + pr*(2.0*c3c3c4 - c4) - 2.0*pq*s3*c3 + 2.0*qr*c3s3s4));
    D_(5,4) = ic4ic4*(qr + s4*(pq + QUAD_KA*s3*du02 - QUAD_KB*c3*du0123) - QUAD_KI*(qr*(1 + c3c3) + pq*(s4 - c3c3*s4) + rr*c3c4s3*c4c4));
    D_(5,9) = ic4*(x[10] + QUAD_KI*(x[10]*(c3c3 - 1) + x[11]*c3c4s3));
    D_(5,10) = ic4*(x[9] + x[11]*s4 - QUAD_KI*(x[11]*(c3c3*s4 + 1) - x[9]*(1 + c3c3)));
    D_(5,11) = ic4*(x[10]*s4 + QUAD_KI*(x[9]*c3c4s3 - x[10]*s4*(1.0 + c3c3) - 0.25*x[11]*c3c4s3s4));
    D_(5,12) = -yawA - yawB;
    D_(5,13) = yawA;
    D_(5,14) = -yawA + yawB;
    D_(5,15) = yawA;
    #undef D_
}

// state weights: Q1 on x y z, Q2 on roll pitch yaw, QUAD_QRATE on all rates; R on the thrusts; the final knot takes QF1 on positions
// and QF2 on rates and has no control term.  Reference values (cost_quad.cuh:19-24): Q1 0.01, Q2 0.001, R 5.0, QF1 = QF2 1000.
template <typename T>
__host__ __device__ __forceinline__
T quadStateWeight(int i, bool last, T Q1, T Q2, T QF1, T QF2){
    return last ? (i < NUM_POS ? QF1 : QF2) : (i < 3 ? Q1 : (i < NUM_POS ? Q2 : static_cast<T>(QUAD_QRATE)));
}
template <typename T>
__host__ __device__ __forceinline__
T costFunc(T *xk, T *uk, T *xgk, int k, T Q1, T Q2, T R, T QF1, T QF2){
    const bool last = (k == NUM_TIME_STEPS - 1);
    T cost = 0.0;
    #pragma unroll
    for (int i = 0; i < STATE_SIZE; i++){ cost += quadStateWeight<T>(i, last, Q1, Q2, QF1, QF2)*pow(xk[i] - xgk[i], 2); }
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
        const T w = i < STATE_SIZE ? quadStateWeight<T>(i, last, Q1, Q2, QF1, QF2) : (last ? static_cast<T>(0.0) : R);
        #pragma unroll
        for (int j = 0; j < STATE_SIZE + CONTROL_SIZE; j++){ Hk[i*ld_H + j] = (i != j) ? static_cast<T>(0.0) : w; }
        gk[i] = w*(i < STATE_SIZE ? xk[i] - xgk[i] : uk[i - STATE_SIZE]);
    }
}
