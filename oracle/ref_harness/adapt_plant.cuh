/*
 * adapt_plant.cuh -- TEST INFRASTRUCTURE, not product code.
 *
 * The reference's pendulum / cart-pole / quadrotor plant files (plants/{cost,dynamics}_{pend,cart,quad}.cuh) kept their v0.1
 * signatures (costFunc(xk,uk,xgk,k), dynamics(...,s_eePos,reps)) and their weight macros (Q1, Q2, R, QF ...) collide with the
 * v0.2 parameter names of DDPHelpers/fpHelpers.cuh:134, so config.cuh does not compile for PLANT 1-3 at HEAD.
 * This adapter includes those SAME reference files, unmodified and where they lie, inside a namespace, removes the colliding
 * macros afterwards (they are already expanded inside the function bodies), and forwards the v0.2 call signatures
 * (plants/cost_arm.cuh:130,158; plants/dynamics_arm.cuh:2097,2167) to them.  oracle/Makefile points config.cuh's plant
 * #include lines here (sed into oracle/_ref/gen/); nothing of the reference is copied into the repo.
 */
#if EE_COST
#error "PLANT 1-3 have no end effector"
#endif
namespace v01 {
#if PLANT == 1
#include "plants/cost_pend.cuh"
#include "plants/dynamics_pend.cuh"
#elif PLANT == 2
#include "plants/cost_cart.cuh"
#include "plants/dynamics_cart.cuh"
#elif PLANT == 3
#include "plants/cost_quad.cuh"
#include "plants/dynamics_quad.cuh"
#else
#error "adapt_plant.cuh is for PLANT 1, 2, 3"
#endif
}
#undef Q1
#undef Q2
#undef Q3
#undef Q4
#undef QX
#undef QT
#undef R
#undef QF
#undef QR
#undef PI
#undef GRAVITY
// defaults of the v0.2 weight arguments (DDPWrappers.cuh:18-21); the v0.1 cost functions ignore them
#define _Q1 0
#define _Q2 0
#define _R 0
#define _QF1 0
#define _QF2 0
#define _Q_EE1 0
#define _Q_EE2 0
#define _R_EE 0
#define _QF_EE1 0
#define _QF_EE2 0
#define _Q_xdEE 0
#define _QF_xdEE 0
#define _Q_xEE 0
#define _QF_xEE 0
#define _Q_EEV1 0
#define _Q_EEV2 0
#define _QF_EEV1 0
#define _QF_EEV2 0
template <typename T> __host__ __device__ __forceinline__ void initI(T *s_I){ return; }
template <typename T> __host__ __device__ __forceinline__ void initT(T *s_T){ return; }
template <typename T> __host__ __device__ __forceinline__
void dynamics(T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody, T *s_eePos = nullptr, int reps = 1, T *s_eeVel = nullptr){ v01::dynamics<T>(s_qdd, s_x, s_u, d_I, d_Tbody, s_eePos, reps); }
template <typename T> __host__ __device__ __forceinline__
void dynamicsGradient(T *s_dqdd, T *s_qdd, T *s_x, T *s_u, T *d_I, T *d_Tbody){ v01::dynamicsGradient<T>(s_dqdd, s_qdd, s_x, s_u, d_I, d_Tbody); }
template <typename T> __host__ __device__ __forceinline__
T costFunc(T *xk, T *uk, T *xgk, int k, T Q1 = _Q1, T Q2 = _Q2, T R = _R, T QF1 = _QF1, T QF2 = _QF2){ return v01::costFunc<T>(xk, uk, xgk, k); }
template <typename T> __host__ __device__ __forceinline__
void costGrad(T *Hk, T *gk, T *xk, T *uk, T *xgk, int k, int ld_H, T Q1 = _Q1, T Q2 = _Q2, T R = _R, T QF1 = _QF1, T QF2 = _QF2){ v01::costGrad<T>(Hk, gk, xk, uk, xgk, k, ld_H); }
