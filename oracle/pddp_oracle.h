/*
 * pddp_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C, single-threaded CPU restatement of the reference's parallel-DDP/iLQR hot path
 * (plancherb1/parallel-DDP @ 665d2d4), following the *GPU* control flow of
 * DDPHelpers/DDPWrappers.cuh::runiLQR_GPU (best-alpha line search, real defect check,
 * tree-ordered cost reduction).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library; the product (parallel-ddp_b200/) never does.
 *
 * Parity pin: oracle/liboracle.so (ORACLE_FMA=0) is checked bit-for-bit against the
 * reference's own host instantiation of the same routines (oracle/_ref/ref_driver_N* `trace H`
 * and `unit H`, fixtures in tests/golden/), and oracle/liboracle_fma.so (ORACLE_FMA=1, the
 * fused-multiply-add pattern nvcc emits for the reference's device code) against the
 * reference's GPU run (`trace G`, `unit G`) -- see tests/test_oracle_golden.py.
 *
 * The end-effector cost path (EE_COST 1: orc_ee_*, orc_cfg.ee_cost / xTarget / final_cost_shift) is pinned the same way against
 * oracle/_ref/ref_ee_N* (`unit H`, `unit G`, `solve G`, `warm G`) and ref_mpc_ee_N32 (receding horizon).
 *
 * All matrices are column-major with leading dimension = row count, per-knot arrays are
 * contiguous [k][col][row] exactly as the reference allocates them (nisInitHelpers.cuh:776,797-798).
 */
#ifndef PDDP_ORACLE_H
#define PDDP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_ALPHA 64
#define ORC_MAX_N 16   /* state size bound */
#define ORC_MAX_M 8    /* control size bound */

enum { ORC_PLANT_PEND = 1, ORC_PLANT_CART = 2, ORC_PLANT_QUAD = 3, ORC_PLANT_KUKA = 4 };
enum { ORC_INT_EULER = 1, ORC_INT_MIDPOINT = 2, ORC_INT_RK3 = 3 };

typedef struct {
    int plant;            /* config.cuh:21-61 PLANT */
    int n, m, npos;       /* STATE_SIZE, CONTROL_SIZE, NUM_POS */
    int N;                /* NUM_TIME_STEPS (power of two, 32..1024: cudaUtils.h:187-207) */
    int n_alpha;          /* NUM_ALPHA */
    int M;                /* M_BLOCKS (= M_BLOCKS_B = M_BLOCKS_F, config.cuh:90-92) */
    int integrator;       /* INTEGRATOR */
    int max_iter;         /* MAX_ITER */
    int expred_host_order;/* 1: sum dJexp like the reference's HOST build (bpHelpers.cuh:330-331), 0: like its kernel (:327-328,416) */
    float dt;             /* (T)TIME_STEP */
    float alpha[ORC_MAX_ALPHA]; /* (T)pow(ALPHA_BASE,i) nisInitHelpers.cuh:829 */
    float rho_init, rho_min, rho_max, rho_factor;
    float exp_red_min, exp_red_max, max_defect, tol_cost;
    float Q1, Q2, R, QF1, QF2;   /* cost_arm.cuh:96-103 weights (other plants map theirs here) */
    float I[252], Tbody[252];    /* Kuka spatial inertias / fixed joint transforms (dynamics_arm.cuh:71-427) */
    float gravity;               /* GRAVITY dynamics_arm.cuh:42-46 (9.81; 0 under MPC_MODE) */
    int   ee_cost;               /* EE_COST config.cuh:165-167: 1 = end-effector pose cost (goal = 6 floats), 0 = joint-space cost */
    float Q_EE1, Q_EE2, QF_EE1, QF_EE2, R_EE, Q_xdEE, QF_xdEE, Q_xEE, QF_xEE;   /* cost_arm.cuh:106-117 */
    int   use_xtarget;           /* EE_COST: the nominal-state terms measure x from xTarget (non-null on the receding-horizon path, MPCHelpers.cuh:900) */
    float xTarget[ORC_MAX_N];
    int   final_cost_shift;      /* EE_COST: finalCostShift of runiLQR_MPC_GPU (MPCHelpers.cuh:876): the pose terms take their final weights from knot N-1-shift on */
    int   use_limits;            /* USE_LIMITS_FLAG config.cuh:171-173: quadratic penalties beyond the joint / velocity / torque limits (joint-space cost) */
    float Q_PL, Q_VL, R_TL;      /* cost_arm.cuh:26-30 */
    int   use_smooth_abs;        /* USE_SMOOTH_ABS config.cuh:174-176 (EE_COST): pose term sqrt(2 c + alpha^2) - alpha, cost_arm.cuh:218-220,242-252 */
    float sa_alpha, sa_alpha2;   /* (T)SMOOTH_ABS_ALPHA and (T)(SMOOTH_ABS_ALPHA*SMOOTH_ABS_ALPHA) (the product is formed in double), cost_arm.cuh:119-121 */
} orc_cfg;

/* work arrays of one problem, reference layouts */
typedef struct {
    float *x, *u, *d;          /* [n_alpha][N][n|m|n] */
    float *xp, *xp2, *up, *dp; /* [N][.] */
    float *AB, *H, *g;         /* [N][n*(n+m)], [N][(n+m)^2], [N][n+m] */
    float *P, *p, *Pp, *pp;    /* [N][n*n], [N][n] */
    float *KT, *du, *ApBK, *Bdu;
    float *xg;                 /* [n] */
    float J[ORC_MAX_ALPHA], dT[ORC_MAX_ALPHA];
    float dJexp[2*16];
    int   err[16];
    /* solver scalars */
    float prevJ, dJ, z, rho, drho;
    int   iter, alphaIndex, ignore_defect;
    float JTp[ORC_MAX_ALPHA*16];   /* EE_COST: per (alpha, shooting interval) cost partials d_JT[b + alpha*M] (fpHelpers.cuh:299) */
} orc_ws;

void  orc_default_cfg_kuka(orc_cfg *c, int N);                 /* headline constants, SURVEY A.1 (I/Tbody must be filled by caller) */
void  orc_default_cfg_plant(orc_cfg *c, int plant, int N, int n_alpha, int integrator);   /* PLANT 1-3: config.cuh:21-61 + the plants' cost weights */
orc_ws *orc_ws_alloc(const orc_cfg *c);
void  orc_ws_free(orc_ws *w);

/* plant plug-ins (Kuka) */
void  orc_kuka_dynamics(const orc_cfg *c, const float *x, const float *u, float *qdd);                  /* dynamics_arm.cuh:2095-2163 */
void  orc_kuka_dynamics_gradient(const orc_cfg *c, const float *x, const float *u, float *qdd, float *dqdd /*[147]*/); /* :2165-2289 */
int   orc_kuka_gradient_stages(const orc_cfg *c, const float *x, const float *u, float *stages);          /* debug dump, returns float count */
/* generic plant dispatch */
void  orc_dynamics(const orc_cfg *c, const float *x, const float *u, float *qdd);
void  orc_integrator(const orc_cfg *c, const float *x, const float *u, float *xnext);                 /* integrators.cuh */
void  orc_integrator_gradient(const orc_cfg *c, const float *x, const float *u, float *AB, float *qdd_out);
float orc_cost(const orc_cfg *c, const float *x, const float *u, const float *xg, int k);               /* cost_*.cuh costFunc */
void  orc_cost_grad(const orc_cfg *c, float *H, float *g, const float *x, const float *u, const float *xg, int k);
void  orc_dynamics_gradient_any(const orc_cfg *c, const float *x, const float *u, float *qdd, float *dqdd /*[npos*(n+m)]*/);   /* dynamicsGradient of any plant */

/* end-effector cost plug-ins (EE_COST 1) */
void  orc_ee_pos(const orc_cfg *c, const float *x, float *ee /*[6]*/, float *dee /*[7][6] or NULL*/);           /* compute_eePos dynamics_arm.cuh:1877-1923 */
float orc_ee_cost(const orc_cfg *c, const float *ee, const float *goal, const float *x, const float *u, int k);  /* costFunc cost_arm.cuh:306-325 */
void  orc_ee_cost_grad(const orc_cfg *c, float *H, float *g, const float *ee, const float *dee, const float *goal,
                       const float *x, const float *u, int k);                                                 /* costGrad cost_arm.cuh:328-388 */

/* phases, operating on a workspace */
void  orc_load(const orc_cfg *c, orc_ws *w, const float *x0, const float *u0, const float *xg);        /* loadVarsGPU, clear=1, rollout=0 */
void  orc_init(const orc_cfg *c, orc_ws *w, float *Jout, int *alphaOut);                               /* initAlgGPU */
void  orc_load_ex(const orc_cfg *c, orc_ws *w, const float *x0, const float *u0, const float *xg,
                  const float *KT0, const float *P0, const float *p0, const float *d0, int clear, int rollout, int ignore_first); /* loadVarsGPU, all flags */
void  orc_init_ex(const orc_cfg *c, orc_ws *w, float *Jout, int *alphaOut, int rollout);
int   orc_backward_pass(const orc_cfg *c, orc_ws *w);                                                   /* backwardPassGPU incl. rho retry */
void  orc_backward_pass_once(const orc_cfg *c, orc_ws *w, float rho);                                   /* one backPassKern launch */
void  orc_forward_sweep(const orc_cfg *c, orc_ws *w);                                                   /* forwardSweepKern, all alpha */
void  orc_forward_sim(const orc_cfg *c, orc_ws *w);                                                     /* forwardSimKern, all (interval, alpha) */
void  orc_cost_defect(const orc_cfg *c, orc_ws *w);                                                     /* costKern + defectKern */
void  orc_line_search(const orc_cfg *c, orc_ws *w);                                                     /* fpHelpers.cuh:374-376,395-408 */
int   orc_accept_reject(const orc_cfg *c, orc_ws *w, float *Jout, int *alphaOut);                       /* acceptRejectTrajGPU; returns 1 = stop */
void  orc_next_iteration_setup(const orc_cfg *c, orc_ws *w);                                            /* nextIterationSetupGPU */

/* whole solve: returns iterations used; x_out/u_out = accepted trajectory (storeVarsGPU) */
int   orc_solve(const orc_cfg *c, const float *x0, const float *u0, const float *xg,
                float *x_out, float *u_out, float *Jout /*[max_iter+1]*/, int *alphaOut /*[max_iter+1]*/);

/* same with runiLQR_GPU's warm-start inputs and flags (KT0/P0/p0/d0 may be NULL when clear = 1) */
int   orc_solve_ex(const orc_cfg *c, const float *x0, const float *u0, const float *xg, const float *KT0, const float *P0, const float *p0, const float *d0,
                   int rollout, int clear, int ignore_first, float *x_out, float *u_out, float *Jout, int *alphaOut);

/* receding horizon (runiLQR_MPC_GPU, MPCHelpers.cuh:862-1045, without its wall-clock budget): persistent state in orc_mpc */
typedef struct orc_mpc_s orc_mpc;
orc_mpc *orc_mpc_alloc(const orc_cfg *c, const float *x_init, const float *u_init, const float *xg);
void  orc_mpc_free(orc_mpc *mp);
int   orc_mpc_step(const orc_cfg *c, orc_mpc *mp, const float *xActual, const float *xg, int shift, int max_iter, int clear_vars, int ignore_first,
                   float *Jout, int *alphaOut);
const float *orc_mpc_x(const orc_mpc *mp); const float *orc_mpc_u(const orc_mpc *mp); const float *orc_mpc_KT(const orc_mpc *mp);
int   orc_mpc_last_successful_solve(const orc_mpc *mp);

int   orc_fma_mode(void); /* the ORACLE_FMA this library was built with */

#ifdef __cplusplus
}
#endif
#endif
