/*
 * pddp_oracle_plants.c -- TEST INFRASTRUCTURE ONLY (see pddp_oracle.h).
 *
 * CPU restatement of the reference's pendulum / cart-pole / quadrotor plants (plants/{dynamics,cost}_{pend,cart,quad}.cuh) and of its
 * Midpoint and RK3 integrators (utils/integrators.cuh:56-233), part of liboracle*.so.
 *
 * Unlike pddp_oracle.c this file is written as ordinary C expressions with the reference's operand order: these plants mix float
 * variables with double literals (GRAVITY, the quadrotor's inertia ratios, pow(.,2)), so most of their arithmetic is double
 * arithmetic rounded to float on assignment, and the fused multiply-adds of the device build sit inside long double sums.  The two
 * arithmetic variants come from the compiler instead of from macros (oracle/Makefile):
 *     host arithmetic   -ffp-contract=off          every operation rounded on its own, like the reference's x86-64 host build
 *     GPU arithmetic    -ffp-contract=fast -mfma   a product feeding a sum is fused; for p1 + p2 the left product is fused -- the rule
 *                                                  nvcc applies to the reference's device code (pddp_oracle.c header)
 * and sinf / cosf are the library's of each side (ORC_SINF / ORC_COSF: glibc, or the CUDA 12.9 restatement in pddp_oracle.c).
 * Pins: tests/golden/p*_unit_H.npz, p*_trace_H*.npz (reference host build) and p*_unit_G.npz, p*_solve_G*.npz (reference GPU run).
 */
#include "pddp_oracle.h"
#include <math.h>
#include <string.h>

float orc_sinf_impl(float x); float orc_cosf_impl(float x);       /* pddp_oracle.c */
#define SINF(x) orc_sinf_impl(x)
#define COSF(x) orc_cosf_impl(x)
#ifdef ORACLE_F64
#define COS64(x) cos(x)
#define POW2(x) ((x)*(x))
#else
#define COS64(x) cos((double)(x))
#define POW2(x) pow((x), 2)
#endif

/* ------------------------------------------------------------------ pendulum: dynamics_pend.cuh:30-51 */
#define PEND_G (-9.81)
static void pend_dynamics(const float *x, const float *u, float *qdd){ qdd[0] = u[0] + PEND_G*SINF(x[0]); }
static void pend_gradient(const float *x, const float *u, float *qdd, float *dqdd){
    if (qdd){ pend_dynamics(x, u, qdd); }
    dqdd[0] = PEND_G*COSF(x[0]); dqdd[1] = 0.0; dqdd[2] = 1;
}

/* ------------------------------------------------------------------ cart-pole: dynamics_cart.cuh:30-76 */
#define CART_G (-9.81)
#define CART_MC 10
#define CART_MP 1
#define CART_L 0.5
#define CART_MPL (CART_MP * CART_L)
#define CART_MPLL (CART_MPL * CART_L)
static void cart_dynamics(const float *x, const float *u, float *qdd){
    const float c = COSF(x[1]), s = SINF(x[1]), w2 = x[3]*x[3];
    const float m00 = CART_MC + CART_MP, m11 = CART_MPLL, m01 = CART_MPL*c;
    const float ps = CART_MPL*s, f0 = ps*w2 + u[0], f1 = ps*CART_G;
    const float idet = 1/(m00*m11 - m01*m01);
    qdd[0] = idet*(m11*f0 - m01*f1);
    qdd[1] = idet*(m00*f1 - m01*f0);
}
static void cart_gradient(const float *x, const float *u, float *qdd, float *dqdd){
    if (qdd){ cart_dynamics(x, u, qdd); }
    const float c = COSF(x[1]), s = SINF(x[1]);
    const float w = x[3], w2 = w*w;
    const float m00 = CART_MC + CART_MP, m11 = CART_MPLL;
    const float m01 = CART_MPL*c, ps = CART_MPL*s;
    const float f0 = ps*w2 + u[0], f1 = ps*CART_G;
    const float det = m00*m11 - m01*m01, idet = 1/det;
    const float a0 = m11*f0 - m01*f1, a1 = m00*f1 - m01*f0;
    const float d1_du = idet*(-m01), d1_dw = idet*(-2*m01*ps*w);
    const float d0_du = idet*(m11),  d0_dw = idet*(2*m11*ps*w);
    const float m01_dth = -ps, f0_dth = m01*w2, f1_dth = m01*CART_G;
    const float a1_dth = m00*f1_dth - (m01_dth*f0 + m01*f0_dth);
    const float a0_dth = m11*f0_dth - (m01_dth*f1 + m01*f1_dth);
    const float idet_dth = -2*m01*ps*idet*idet;
    const float d1_dth = idet*a1_dth + idet_dth*a1;
    const float d0_dth = idet*a0_dth + idet_dth*a0;
    dqdd[0] = 0; dqdd[1] = 0; dqdd[2] = d0_dth; dqdd[3] = d1_dth; dqdd[4] = 0; dqdd[5] = 0;
    dqdd[6] = d0_dw; dqdd[7] = d1_dw; dqdd[8] = d0_du; dqdd[9] = d1_du;
}

/* ------------------------------------------------------------------ quadrotor: dynamics_quad.cuh:42-168 (reps = 1) */
#define QUAD_G (-9.81)
#define QUAD_MASS 0.5
#define QUAD_INV_MASS 2
#define QUAD_KI 0.73913043584
#define QUAD_KA 76.0869565217
#define QUAD_KB 6.125
static void quad_dynamics(const float *x, const float *u, float *a){
        const float c3 = COSF(x[3]), c4 = COSF(x[4]), c5 = COSF(x[5]);
        const float s3 = SINF(x[3]), s4 = SINF(x[4]), s5 = SINF(x[5]);
        const float pq = x[9]*x[10], pr = x[9]*x[11];
        const float qr = x[10]*x[11], rr = x[11]*x[11];
        const float thrust = u[0] + u[1] + u[2] + u[3];
        a[0] = QUAD_INV_MASS*thrust*(s3*s5 + c3*c5*s4);
        a[1] = -QUAD_INV_MASS*thrust*(c5*s3 - c3*s4*s5);
        a[2] = QUAD_G + QUAD_INV_MASS*thrust*c3*c4;
        const float yawU = u[0] - u[1] + u[2] - u[3], pitchU = u[2] - u[0];
        const float ic4 = 1/c4, c3c3 = c3*c3, c3c4 = c3*c4, sin2r = 2.0*s3*c3, cos2r = COS64(2.0*x[3]);
        a[3] = ic4*(0.0005434782609*(32000.0*qr + 140000.0*(u[1] - u[3])*c4 - 28320.0*pq*s4 - 30160.0*qr*c3c3 + 1127.0*yawU*c3*s4 - 140000.0*pitchU*s3*s4 + 30160.0*pq*c3c3*s4 - 30160.0*qr*c3c4*c3c4 + 30160.0*x[10]*x[10]*c3*c4*s3 - 30160.0*rr*c3*c4*s3 + 30160.0*pr*c3*c4*s3*s4));
        a[4] = 76.08695652*pitchU*c3 - 0.6125*yawU*s3 - 1.0*pr*c4 - 8.195652174*pq*sin2r - 16.39130435*rr*c3c3*c4*s4 + 16.39130435*pr*c3c3*c4 + 16.39130435*qr*c3*s3*s4;
        a[5] = -ic4*(0.0005434782609*(13240.0*pq - 1127.0*yawU*c3 - 140000.0*pitchU*s3 - 16920.0*qr*s4 + 7540.0*rr*sin2r*2.0*s4*c4 - 15080.0*pq*cos2r - 15080.0*pr*sin2r*c4 + 15080.0*qr*cos2r*s4));
}
static void quad_gradient(const float *x, const float *u, float *qdd, float *dqdd){
    if (qdd){ quad_dynamics(x, u, qdd); }
    memset(dqdd, 0, sizeof(float)*6*16);
    #define D_(row, col) dqdd[(row) + (col)*6]
    const float s3 = SINF(x[3]), s4 = SINF(x[4]), s5 = SINF(x[5]);
    const float c3 = COSF(x[3]), c4 = COSF(x[4]), c5 = COSF(x[5]);
    const float tm = (u[0] + u[1] + u[2] + u[3])*QUAD_INV_MASS;
    const float ax_du = (s3*s5 + c3*c5*s4)*QUAD_INV_MASS;
    D_(0,3) = (c3*s5 - c5*s3*s4)*tm;
    D_(0,4) = (c3*c4*c5)*tm;
    D_(0,5) = (c5*s3 - c3*s4*s5)*tm;
    D_(0,12) = ax_du; D_(0,13) = ax_du; D_(0,14) = ax_du; D_(0,15) = ax_du;
    const float ay_du = -(c5*s3 - c3*s4*s5)*QUAD_INV_MASS;
    D_(1,3) = -(c3*c5 + s3*s4*s5)*tm;
    D_(1,4) = (c3*c4*s5)*tm;
    D_(1,5) = (s3*s5 + c3*c5*s4)*tm;
    D_(1,12) = ay_du; D_(1,13) = ay_du; D_(1,14) = ay_du; D_(1,15) = ay_du;
    const float az_du = (c3*c4)/QUAD_MASS;
    D_(2,3) = -(c4*s3)*tm;
    D_(2,4) = -(c3*s4)*tm;
    D_(2,12) = az_du; D_(2,13) = az_du; D_(2,14) = az_du; D_(2,15) = az_du;
    const float rr = x[11]*x[11];
    const float pq = x[9]*x[10];
    const float pr = x[9]*x[11];
    const float qr = x[10]*x[11];
    const float c3c3 = c3*c3;
    const float c4c4 = c4*c4;
    const float c3c3c4 = c3c3*c4;
    const float c3c3c4c4 = c3c3*c4c4;
    const float c3c4s3 = c3*c4*s3;
    const float c3s3s4 = c3*s3*s4;
    const float c3c4s3s4 = c3c4s3*s4;
    const float ic4 = 1.0/c4;
    const float ic4ic4 = ic4/c4;
    const float du02 = -u[0] + u[2];
    const float du0123 = -u[0] + u[1] - u[2] + u[3];
    const float rollA = QUAD_KB*c3*s4*ic4;
    const float rollB = QUAD_KA*s3;
    D_(3,3) = ic4*(QUAD_KA*c3*s4*du02 + QUAD_KB*s3*s4*du0123 + QUAD_KI*(x[10]*x[10]*(2*c3c3*c4 - c4) + rr*c4 - 2*rr*c3c3*c4 - pr*s4*c4 + 2.0*qr*s3*c3 + 2*pr*c3c3*c4*s4 - 2*x[11]*c3*c4c4*s3 - 2*x[10] - 2*pq*c3*s3*s4));
    D_(3,4) = ic4ic4*(pq + qr*s4 + QUAD_KA*s3*du02 - QUAD_KB*c3*du0123 + QUAD_KI*(pq*(1 + c3c3) + qr*(1 - c3c3*s4 + c3c3c4c4*s4) + pr*c3c4s3*c4c4));
    D_(3,9) = ic4*(x[10]*s4 + QUAD_KI*(x[10]*(c3c3 - 1)*s4 + x[11]*c3c4s3s4));
    D_(3,10) = ic4*(x[11] + x[9]*s4 - QUAD_KI*(x[11]*(1 + c3c3 + c3c3c4c4) - x[9]*(s4 + c3c3*s4) - 2*x[10]*c3c4s3));
    D_(3,11) = ic4*(x[10] - QUAD_KI*(x[10] + x[10]*c3c3 + x[9]*c3c4s3s4 - 2*x[11]*c3c4s3 - x[10]*c3c3c4c4));
    D_(3,12) = rollA - rollB;
    D_(3,13) = QUAD_KA - rollA;
    D_(3,14) = rollA + rollB;
    D_(3,15) = -QUAD_KA - rollA;
    const float pitchA = QUAD_KB*s3;
    const float pitchB = QUAD_KA*c3;
    D_(4,3) = QUAD_KB*c3*du0123 - QUAD_KA*s3*du02 + QUAD_KI*(pq*(1 - 2*c3c3) + qr*s4*(2*c3c3 - 1) + 2*rr*c3c4s3s4 - 2*pr*c3c4s3);
    D_(4,4) = QUAD_KI*(rr*c3c3 - 2.0*rr*c3c3c4c4 - c3c3*s4 + qr*c3c4s3) + pr*s4;
    D_(4,9) = QUAD_KI*(x[11]*c3c3c4 - x[10]*s3*c3) - x[11]*c4;
    D_(4,10) = QUAD_KI*(x[11]*c3s3s4 - x[9]*s3*c3);
    D_(4,11) = QUAD_KI*(x[9]*c3c3c4 + x[10]*c3s3s4 - 2.0*x[11]*c3c3c4*s4) - x[9]*c4;
    D_(4,12) = -pitchB - pitchA;
    D_(4,13) = pitchA;
    D_(4,14) = pitchB - pitchA;
    D_(4,15) = pitchA;
    const float yawA = -QUAD_KB*c3*ic4;
    const float yawB = QUAD_KA*s3*ic4;
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

/* ------------------------------------------------------------------ GPU arithmetic of the three plants
 * Which multiply-adds the device build fuses inside these expressions follows no rule that could be restated by hand (nvcc fuses the
 * RIGHT product of s3*s5 + c3*c5*s4, the left one elsewhere; ptxas fuses some of what NVVM left) and a C compiler's own contraction
 * differs from it, so the ORACLE_FMA build takes the statements read off nvcc's PTX for the same closed forms (oracle/tools/ptx2c.py).
 * Both variants are pinned: the expressions above to the reference's host build, these to its GPU run. */
#if ORACLE_FMA && !defined(ORACLE_F64)
#include <stdint.h>
static float bitsf(uint32_t v){ float f; memcpy(&f, &v, 4); return f; }
static double bitsd(uint64_t v){ double f; memcpy(&f, &v, 8); return f; }
#define SIN64(x) sin(x)
#include "plants_gpu_arith.inc"
#define PLANT_DYN(nm, x, u, qdd) nm##_dynamics_gpu(x, u, qdd)
/* gradient = the acceleration (when asked for) + the block of the real kernel that fills s_dqdd, fed with the sines / cosines it
 * receives from the inlined library code ahead of it.  Live-in order of each block: see plants_gpu_arith.inc; which trigonometric
 * value sits in which register was resolved against the reference's GPU dumps (tools/quad_livein_search.py). */
int orc_quad_perm[6] = {0, 2, 4, 3, 5, 1};
static void pend_gradient_gpu(const float *x, const float *u, float *qdd, float *dqdd){
    if (qdd){ pend_dynamics_gpu(x, u, qdd); }
    dqdd[0] = PEND_G*COSF(x[0]); dqdd[1] = 0.0; dqdd[2] = 1;          /* one double product: nothing to fuse */
}
static void cart_gradient_gpu(const float *x, const float *u, float *qdd, float *dqdd){
    if (qdd){ cart_dynamics_gpu(x, u, qdd); }
    float li[2] = {COSF(x[1]), SINF(x[1])}; cart_gradient_tail_gpu(x, u, li, dqdd);
}
static void quad_gradient_gpu(const float *x, const float *u, float *qdd, float *dqdd){
    if (qdd){ quad_dynamics_gpu(x, u, qdd); }
    const float t[6] = {COSF(x[3]), COSF(x[4]), COSF(x[5]), SINF(x[3]), SINF(x[4]), SINF(x[5])};
    float li[6]; for (int i = 0; i < 6; i++){ li[i] = t[orc_quad_perm[i]]; }
    memset(dqdd, 0, sizeof(float)*6*16);
    quad_gradient_tail_gpu(x, u, li, dqdd);
}
#define PLANT_GRAD(nm, x, u, qdd, dqdd) nm##_gradient_gpu(x, u, qdd, dqdd)
#else
#define PLANT_DYN(nm, x, u, qdd) nm##_dynamics(x, u, qdd)
#define PLANT_GRAD(nm, x, u, qdd, dqdd) nm##_gradient(x, u, qdd, dqdd)
#endif

void orc_plant_dynamics(const orc_cfg *c, const float *x, const float *u, float *qdd){
    if (c->plant == ORC_PLANT_PEND){ PLANT_DYN(pend, x, u, qdd); } else if (c->plant == ORC_PLANT_CART){ PLANT_DYN(cart, x, u, qdd); } else { PLANT_DYN(quad, x, u, qdd); }
}
void orc_plant_gradient(const orc_cfg *c, const float *x, const float *u, float *qdd, float *dqdd){
    if (c->plant == ORC_PLANT_PEND){ PLANT_GRAD(pend, x, u, qdd, dqdd); } else if (c->plant == ORC_PLANT_CART){ PLANT_GRAD(cart, x, u, qdd, dqdd); } else { PLANT_GRAD(quad, x, u, qdd, dqdd); }
}

/* ------------------------------------------------------------------ costs: cost_pend.cuh:20-54, cost_cart.cuh:19-68, cost_quad.cuh:19-55
 * weight of state i: pendulum / cart-pole Q1 (positions), Q2 (rates); quadrotor Q1 (x y z), Q2 (roll pitch yaw), 2.0 (rates);
 * final knot QF1 / QF2; R on the controls (none at the final knot).  The squares are pow(., 2) in double. */
static float state_weight(const orc_cfg *c, int i, int last){
    if (last){ return i < c->npos ? c->QF1 : c->QF2; }
    if (c->plant == ORC_PLANT_QUAD){ return i < 3 ? c->Q1 : (i < 6 ? c->Q2 : (float)2.0); }
    return i < c->npos ? c->Q1 : c->Q2;
}
float orc_plant_cost(const orc_cfg *c, const float *x, const float *u, const float *xg, int k){
    const int last = (k == c->N - 1);
    float cost = 0.0;
    for (int i = 0; i < c->n; i++){ const float w = state_weight(c, i, last); cost += w*POW2(x[i] - xg[i]); }
    if (!last){ for (int i = 0; i < c->m; i++){ cost += c->R*POW2(u[i]); } }
    return 0.5*cost;
}
void orc_plant_cost_grad(const orc_cfg *c, float *H, float *g, const float *x, const float *u, const float *xg, int k){
    const int last = (k == c->N - 1), n = c->n, nm = c->n + c->m;
    for (int i = 0; i < nm; i++){
        const float w = i < n ? state_weight(c, i, last) : (last ? (float)0.0 : c->R);
        for (int j = 0; j < nm; j++){ H[i*nm + j] = (i != j) ? (float)0.0 : w; }
        g[i] = w*(i < n ? x[i] - xg[i] : u[i - n]);
    }
}

/* ------------------------------------------------------------------ integrators, utils/integrators.cuh (any plant through orc_dynamics) */
void orc_dynamics(const orc_cfg *c, const float *x, const float *u, float *qdd);
void orc_dynamics_gradient_any(const orc_cfg *c, const float *x, const float *u, float *qdd, float *dqdd);
static float dxd(const float *dqdd, int np, int r, int c){ return r < np ? (float)(r + np == c ? 1 : 0) : dqdd[(c-1)*np + r]; }   /* :15-17 */

/* The integrators are float arithmetic throughout, and here the device build's contraction is regular (read off the PTX of
 * plugin/integrators.cuh and of the reference's GPU dumps): a product feeding a sum is fused, of two products the left one. */
#if ORACLE_FMA && !defined(ORACLE_F64)
#define FMAF(a,b,c) fmaf((a),(b),(c))
#else
#define FMAF(a,b,c) ((float)((float)((a)*(b)) + (c)))
#endif
#define MULF(a,b) ((float)((a)*(b)))
#define ADDF(a,b) ((float)((a)+(b)))
#define SUBF(a,b) ((float)((a)-(b)))

void orc_integrator_generic(const orc_cfg *c, const float *x, const float *u, float *xn){
    const int np = c->npos; const float dt = c->dt, h = MULF((float)(0.5), dt);
    float a1[ORC_MAX_N], a2[ORC_MAX_N], a3[ORC_MAX_N], x2[ORC_MAX_N], x3[ORC_MAX_N];
    orc_dynamics(c, x, u, a1);
    if (c->integrator == ORC_INT_EULER){                                     /* :24-36 */
        for (int i = 0; i < np; i++){ xn[i] = FMAF(dt, x[i+np], x[i]); xn[i+np] = FMAF(dt, a1[i], x[i+np]); }
    } else if (c->integrator == ORC_INT_MIDPOINT){                           /* :56-83: q advances with the INITIAL velocity (:78) */
        for (int i = 0; i < np; i++){ x2[i] = FMAF(h, x[i+np], x[i]); x2[i+np] = FMAF(h, a1[i], x[i+np]); }
        orc_dynamics(c, x2, u, a2);
        for (int i = 0; i < np; i++){ xn[i] = FMAF(dt, x[i+np], x[i]); xn[i+np] = FMAF(dt, a2[i], x[i+np]); }
    } else {                                                                  /* RK3 :123-160 */
        const float dt6 = (float)(dt/(float)(6));
        for (int i = 0; i < np; i++){ x2[i] = FMAF(h, x[i+np], x[i]); x2[i+np] = FMAF(h, a1[i], x[i+np]); }
        orc_dynamics(c, x2, u, a2);
        for (int i = 0; i < np; i++){      /* 2*v is exact, so 2*v - w is one rounding with or without fusion */
            x3[i] = FMAF(dt, SUBF(MULF((float)(2), x2[i+np]), x[i+np]), x[i]); x3[i+np] = FMAF(dt, SUBF(MULF((float)(2), a2[i]), a1[i]), x[i+np]);
        }
        orc_dynamics(c, x3, u, a3);
        for (int i = 0; i < np; i++){
            xn[i] = FMAF(dt6, ADDF(FMAF((float)(4), x2[i+np], x[i+np]), x3[i+np]), x[i]);
            xn[i+np] = FMAF(dt6, ADDF(FMAF((float)(4), a2[i], a1[i]), a3[i]), x[i+np]);
        }
    }
}

#define ORC_ND (ORC_MAX_N*(ORC_MAX_N + ORC_MAX_M))
void orc_integrator_gradient_generic(const orc_cfg *c, const float *x, const float *u, float *AB, float *qdd_out){
    const int np = c->npos, n = c->n, nm = c->n + c->m; const float dt = c->dt, h = MULF((float)(0.5), dt), dt2 = MULF((float)(2), dt);
    float a1[ORC_MAX_N], a2[ORC_MAX_N], a3[ORC_MAX_N], x2[ORC_MAX_N], x3[ORC_MAX_N];
    float d1[ORC_ND], d2[ORC_ND], d3[ORC_ND], G1[ORC_ND], G2[ORC_ND];
    orc_dynamics_gradient_any(c, x, u, a1, d1);
    if (qdd_out){ for (int i = 0; i < np; i++){ qdd_out[i] = a1[i]; } }
    #define DLT(a, b) ((float)((a) == (b) ? 1 : 0))
    if (c->integrator == ORC_INT_EULER){                                     /* :38-53 */
        for (int ky = 0; ky < nm; ky++){ for (int kx = 0; kx < n; kx++){ AB[ky*n + kx] = FMAF(dt, dxd(d1, np, kx, ky), DLT(ky, kx)); } }
    } else if (c->integrator == ORC_INT_MIDPOINT){                           /* :86-121 */
        for (int i = 0; i < np; i++){ x2[i] = FMAF(h, x[i+np], x[i]); x2[i+np] = FMAF(h, a1[i], x[i+np]); }
        orc_dynamics_gradient_any(c, x2, u, a2, d2);
        for (int ky = 0; ky < nm; ky++){ for (int kx = 0; kx < n; kx++){
            float val = 0;
            for (int i = 0; i < n; i++){
                const float A2_val = FMAF(h, dxd(d2, np, kx, i), DLT(kx, i)), AB1_val = FMAF(h, dxd(d1, np, i, ky), DLT(ky, i));
                val = FMAF(A2_val, AB1_val, val);
            }
            /* the B-column term is formed in its own branch of the kernel and added behind it: a rounded product, then a sum */
            AB[ky*n + kx] = ADDF(val, (ky < n ? (float)(0) : MULF(h, dxd(d2, np, kx, ky))));
        }}
    } else {                                                                  /* RK3 :162-233, stage states as written (:181-182,190-191) */
        const float dt6 = (float)(dt/(float)(6)), dt23 = (float)(dt2/(float)(3));
        for (int i = 0; i < np; i++){ x2[i] = FMAF(h, x[i+np], x[i]); x2[i+np] = FMAF(h, a1[i], x[i]); }
        orc_dynamics_gradient_any(c, x2, u, a2, d2);
        for (int i = 0; i < np; i++){ x3[i] = FMAF(dt2, x2[i+np], FMAF(dt, x[i+np], x[i])); x3[i+np] = FMAF(dt2, a2[i], FMAF(dt, a1[i], x[i])); }
        orc_dynamics_gradient_any(c, x3, u, a3, d3);
        for (int ky = 0; ky < nm; ky++){ for (int kx = 0; kx < n; kx++){
            float val = 0;
            for (int i = 0; i < n; i++){ val = FMAF(dxd(d2, np, kx, i), FMAF(h, dxd(d1, np, i, ky), DLT(ky, i)), val); }
            G1[kx + n*ky] = ADDF(val, (ky < n ? (float)(0) : dxd(d2, np, kx, ky)));
        }}
        for (int ky = 0; ky < nm; ky++){ for (int kx = 0; kx < n; kx++){
            float val = 0;
            for (int i = 0; i < n; i++){
                const float inner = ADDF(FMAF(dt2, G1[ky*n + i], -MULF(dt, dxd(d1, np, i, ky))), DLT(ky, i));
                val = FMAF(dxd(d3, np, kx, i), inner, val);
            }
            G2[kx + n*ky] = ADDF(val, (ky < n ? (float)(0) : dxd(d3, np, kx, ky)));
        }}
        for (int ky = 0; ky < nm; ky++){ for (int kx = 0; kx < n; kx++){
            AB[kx + n*ky] = ADDF(FMAF(dt6, G2[kx + n*ky], FMAF(dt6, dxd(d1, np, kx, ky), MULF(dt23, G1[kx + n*ky]))), DLT(kx, ky));
        }}
    }
    #undef DLT
}
