/*
 * pddp.h -- C-ABI of the B200-native batched parallel-DDP/iLQR solver (libpddp.so).
 *
 * Plain C, POD arguments only.  These entry points are what a binding of the reference's solver surface binds to;
 * each one names the reference interface it replaces (paths relative to plancherb1/parallel-DDP @ 665d2d4).
 * The reference has no ABI: its surface is a set of header-only C++ templates taking ~50 raw pointers
 * (DDPHelpers/DDPWrappers.cuh:8-21).  include/pddp_shim.cuh re-creates those templates on top of this ABI so that
 * examples/WAFR_iLQR_examples.cu compiles unchanged; INTEGRATION.md shows the ctypes / C++ bindings.
 *
 * Conventions: all matrices column-major, per-knot arrays contiguous [k][col][row] with leading dimension = row
 * count (nisInitHelpers.cuh:776,797-798); a leading `batch` dimension is added in front of every per-problem array
 * (the reference solves one problem per call).  Every function returns 0 on success or a negative PDDP_E_* code and
 * never calls exit(); pddp_last_error() gives the message (the reference prints and exit()s, utils/cudaUtils.cu:31-37).
 */
#ifndef PDDP_H
#define PDDP_H
#ifdef __cplusplus
extern "C" {
#endif

#define PDDP_E_INVALID   (-1)   /* bad argument / unsupported configuration */
#define PDDP_E_CUDA      (-2)   /* CUDA runtime error (message in pddp_last_error) */
#define PDDP_E_NODEVICE  (-3)   /* no usable CUDA device: there is NO CPU fallback */

#define PDDP_PLANT_PEND 1
#define PDDP_PLANT_CART 2
#define PDDP_PLANT_QUAD 3
#define PDDP_PLANT_KUKA 4

#define PDDP_MAX_ALPHA 32

/* Run-time form of the reference's compile-time flag system (config.cuh:21-136). */
typedef struct pddp_config {
    int   plant;          /* PLANT (config.cuh:21-61): 1 pendulum, 2 cart-pole, 3 quadrotor, 4 Kuka iiwa14, or a registered plug-in */
    int   N;              /* NUM_TIME_STEPS: power of two in [32,1024] (cudaUtils.h:187-207 reduction trees) */
    int   n_alpha;        /* NUM_ALPHA <= PDDP_MAX_ALPHA */
    int   M;              /* M_BLOCKS = M_BLOCKS_B = M_BLOCKS_F (config.cuh:90-92), must divide N */
    int   max_iter;       /* MAX_ITER (config.cuh:83) */
    int   batch;          /* independent problems solved per call (new; the reference has none) */
    int   device;         /* CUDA device ordinal */
    int   integrator;     /* INTEGRATOR 1 Euler | 2 Midpoint | 3 RK3 (config.cuh:78-80) */
    float alpha_base;     /* ALPHA_BASE: alpha_i = (float)pow(alpha_base, i) (nisInitHelpers.cuh:829) */
    float total_time;     /* TOTAL_TIME: dt = (float)(total_time/(N-1)) (config.cuh:136) */
    float rho_init, rho_min, rho_max, rho_factor;      /* config.cuh:98-104 */
    float exp_red_min, exp_red_max;                    /* config.cuh:116-122 */
    float max_defect;                                  /* MAX_DEFECT_SIZE (config.cuh:123-126) */
    float tol_cost;                                    /* TOL_COST (config.cuh:85-87) */
    float Q1, Q2, R, QF1, QF2;                         /* plants/cost_arm.cuh:96-103 */
    float gravity;                                     /* GRAVITY (dynamics_arm.cuh:42-46): 9.81, or 0 for the reference's MPC_MODE builds */
    int   ee_cost;                                     /* EE_COST (config.cuh:165-167): 1 = end-effector pose cost -- xGoal[b][0..5] is the goal
                                                          pose [x y z roll pitch yaw] (the other n-6 floats of a problem's slot are ignored) and
                                                          the weights below replace Q1..QF2; 0 = joint-space cost */
    float Q_EE1, Q_EE2, QF_EE1, QF_EE2, R_EE;          /* plants/cost_arm.cuh:106-111: running / final weights of the position (1) and
                                                          orientation (2) errors, control weight */
    float Q_xdEE, QF_xdEE, Q_xEE, QF_xEE;              /* cost_arm.cuh:112-115: running / final weights on joint velocities (xd) and angles (x) */
    int   use_limits;                                  /* USE_LIMITS_FLAG (config.cuh:171-173; PLANT 4, either cost): quadratic penalties beyond
                                                          0.8 x the iiwa14's joint, velocity and torque limits in the cost and its gradient
                                                          (plants/cost_arm.cuh:11-94,136-150,176-200: joint-space cost, Hessian untouched;
                                                          :289-291,310-312,341-343,374-376: end-effector cost, Hessian diagonal included) */
    float lim_Q_pos, lim_Q_vel, lim_R_tau;             /* Q_PL, Q_VL, R_TL of cost_arm.cuh:26-30: penalty weights of the position / velocity / torque
                                                          limits (100 each; named apart from the reference's macros so that the header shim can assign them) */
    int   use_smooth_abs;                              /* USE_SMOOTH_ABS (config.cuh:174-176; EE_COST 1): the pose term c of a knot's cost becomes
                                                          sqrt(2 c + alpha^2) - alpha and its gradient is divided by sqrt(2 c + alpha^2)
                                                          (plants/cost_arm.cuh:218-220,242-252; the Gauss-Newton Hessian is not touched) */
    double smooth_abs_alpha;                           /* SMOOTH_ABS_ALPHA (cost_arm.cuh:119-121, 0.2): a double, as the reference squares it in double */
} pddp_config;

typedef struct pddp_solver *pddp_handle;

/* Fill *cfg with the reference's Kuka defaults for the WAFR iLQR example (config.cuh:43-58, WAFR_iLQR_examples.cu:4-11). */
void pddp_default_config_kuka(pddp_config *cfg, int N, int batch);

/* The same for any built-in plant: PLANT 1 pendulum, 2 cart-pole, 3 quadrotor take config.cuh:21-136's defaults (RK3, M_BLOCKS 4,
 * NUM_ALPHA 32 / ALPHA_BASE 0.75 -- quadrotor 16 / 0.5 --, TOTAL_TIME 4, RHO_INIT 10 / 10 / 1, MAX_DEFECT_SIZE 1 / 0.75 / 1) and the
 * weights of plants/cost_{pend,cart,quad}.cuh mapped onto Q1 (positions; quadrotor: x y z), Q2 (rates; quadrotor: roll pitch yaw,
 * its rates are weighted 2.0 by the plant), R, QF1, QF2; PLANT 4 = pddp_default_config_kuka. */
int pddp_default_config(pddp_config *cfg, int plant, int N, int batch);
/* NUM_POS, STATE_SIZE, CONTROL_SIZE of a plant (config.cuh:21-61); any pointer may be NULL */
int pddp_plant_dims(int plant, int *num_pos, int *state_size, int *control_size);

/* ---- plant plug-ins (include/pddp_plant.h, parallel-ddp_b200/csrc/plugin/pddp_plugin.cuh): a plant written as a
 * reference-style header (dynamics, dynamicsGradient, costFunc, costGrad, initI, initT with the signatures of plants/dynamics_arm.cuh:
 * 2095-2097,2165-2167 and plants/cost_arm.cuh:128-130,156-158) and compiled around csrc/plant_tu.cu becomes available to pddp_create
 * under its own pddp_config.plant number.  pddp_load_plant_library dlopens such a library and returns the plant number (>= 1) or a
 * negative code; pddp_plant_error() has the reason. */
struct pddp_plant_ops;
int pddp_register_plant(const struct pddp_plant_ops *ops);
int pddp_load_plant_library(const char *path);
const char *pddp_plant_error(void);

/* replaces allocateMemory_GPU (nisInitHelpers.cuh:766-861): owns every device array, stream and the model constants */
int pddp_create(const pddp_config *cfg, pddp_handle *out);
/* replaces freeMemory_GPU (nisInitHelpers.cuh:863-882) */
void pddp_destroy(pddp_handle h);
/* message of the last error on this handle (or of the last failed pddp_create when h == NULL) */
const char *pddp_last_error(pddp_handle h);

/* replaces runiLQR_GPU (DDPWrappers.cuh:8-138) for `batch` problems at once.
 *   x0 [batch][N][n], u0 [batch][N][m], xGoal [batch][n]           HOST, reference layout (in)
 *   x_out/u_out (may alias x0/u0: the reference overwrites x0,u0)   HOST (out)
 *   Jout [batch][max_iter+1], alphaOut [batch][max_iter+1]         cost / chosen-alpha traces (-1 = rejected), unused
 *                                                                   slots are left as NaN / -99
 *   iters_out [batch]                                               iterations used per problem
 *   times_ms[6]: total, sim, sweep, bp, nis, init  -- device time of each phase summed over iterations (CUDA events);
 *                may be NULL.
 *   forwardRolloutFlag / clearVarsFlag are loadVarsGPU's (nisInitHelpers.cuh:594-652): clear = 0 starts from the P0, p0,
 *   KT0, d0 given to pddp_set_warm_start; rollout = 1 first simulates the given trajectory with alpha[0], du = 0 and the
 *   feedback gains around it, and iterates from the result (alphaOut[0] = 0 instead of -1). */
int pddp_solve(pddp_handle h, const float *x0, const float *u0, const float *xGoal,
               int forwardRolloutFlag, int clearVarsFlag, int ignoreFirstDefectFlag,
               float *x_out, float *u_out, float *Jout, int *alphaOut, int *iters_out, double *times_ms);

/* Same as pddp_solve but with inputs/outputs already resident on the device (used by bench.py's `value` leg). */
int pddp_solve_device(pddp_handle h, const float *d_x0, const float *d_u0, const float *d_xGoal,
                      int ignoreFirstDefectFlag, float *d_x_out, float *d_u_out, float *d_Jout, int *d_alphaOut,
                      int *d_iters_out, double *times_ms);

/* Synthetic inputs of the reference's benchmark driver (WAFR_iLQR_examples.cu:67-121) for problems seed0..seed0+batch-1,
 * each drawn from std::default_random_engine(seed) exactly as the deterministic harness does. HOST buffers. */
int pddp_make_inputs_kuka(int N, int batch, unsigned seed0, float *x0, float *u0, float *xGoal);

/* runiLQR_MPC_GPU's `use_cost_shift` argument (MPCHelpers.cuh:866,876): when on, every pddp_mpc_step evaluates the pose terms of the
 * end-effector cost with their final weights from knot N-1-shiftAmount on (finalCostShift = shiftAmount).  Ignored by the
 * joint-space cost, as in the reference. */
int pddp_mpc_set_cost_shift(pddp_handle h, int use_cost_shift);

/* EE_COST only: the reference's `xTarget` argument of costFunc / costGrad (plants/cost_arm.cuh:263-281) -- the nominal-state terms
 * (Q_xEE, Q_xdEE) then measure x from it.  runiLQR_GPU passes none (the default here); runiLQR_MPC_GPU always passes its
 * gv->d_xTarget (MPCHelpers.cuh:900), so a receding-horizon caller sets it.  HOST [batch][n]; NULL removes it. */
int pddp_set_x_target(pddp_handle h, const float *xTarget);

/* Inputs of the example for any built-in plant (WAFR_iLQR_examples.cu:19-33,72-78,87-90,110-115), seeded like pddp_make_inputs_kuka */
int pddp_make_inputs(int plant, int N, int batch, unsigned seed0, float *x0, float *u0, float *xGoal);

/* Consumer side of the hand-off: replaces getHardwareControls (DDPHelpers/MPCHelpers.cuh:817-858), the host function that turns
 * the published plan and a measured state into the joint command.  Zero-order hold on u and KT, first-order hold on x:
 *   steps = (tActual - t0) / (time_step * 1e6)   [times in microseconds, time_step = TOTAL_TIME/(N-1) in seconds]
 *   k = (int)steps, f = steps - k;  returns 1 (and writes nothing) when k >= N-2 or k < 0
 *   u_out = u[k] - K[k] (xActual - ((1-f) x[k] + f x[k+1]))          (float arithmetic, as the reference's T)
 *   q_out = qActual (PD_GAINS_ON_STATE 0, the reference's default) or the interpolated plan position (pd_gains_on_state != 0)
 *   use_feedback = 0 (USE_FEEDBACK_IN_TRAJ_RUNNER 0): u_out = u[k]
 *   alpha > 0 with u_prev != NULL: u_out = (1-alpha) u_out + alpha u_prev, u_prev = u_out          (double arithmetic)
 * x [N][14], u [N][7], KT [N][14*7] in the reference layouts; qActual, qdActual, q_out, u_out, u_prev: 7 doubles.  HOST only. */
int pddp_hardware_controls(int N, double time_step, const float *x, const float *u, const float *KT, double t0,
                           const double *qActual, const double *qdActual, double tActual, int use_feedback, int pd_gains_on_state,
                           double *u_prev, double alpha, double *q_out, double *u_out);

/* ---- plant plug-ins, evaluated on the device for n independent (x,u) samples (HOST buffers) -----------------------
 * dynamics (plants/dynamics_arm.cuh:2095), _integratorGradient (utils/integrators.cuh:38-53) */
int pddp_unit_dynamics(pddp_handle h, const float *x, const float *u, int n, float *qdd);
int pddp_unit_integrator_gradient(pddp_handle h, const float *x, const float *u, int n, float *AB, float *qdd);
/* plug-in plants: x_{k+1} of _integrator (utils/integrators.cuh:24-36,56-83,123-160) and costFunc / costGrad (plants/cost_*.cuh) for n
 * samples; knot[i] is the k argument (N-1 selects the final cost), xGoal one goal for all samples, H [n][(n+m)^2], g [n][n+m] */
int pddp_unit_integrator(pddp_handle h, const float *x, const float *u, int n, float *xnext);
int pddp_unit_cost(pddp_handle h, const float *x, const float *u, const float *xGoal, const int *knot, int n, float *J, float *H, float *g);

/* ---- phase-level entry points on the solver's device state (for knot-level parity tests and ncu captures) ----------
 * The state is addressed by array name: "x","u","d" ([batch][n_alpha][N][.]), "xp","xp2","up","dp","AB","H","g","P","p",
 * "Pp","pp","KT","du","ApBK","Bdu" ([batch][N][.]), "xGoal" ([batch][n]), "J","dT" ([batch][n_alpha]),
 * "dJexp" ([batch][2M]) and the scalars "rho","drho","prevJ","dJ","z" (float [batch]), "iter","alphaIndex",
 * "ignore_defect","done" (int [batch]).  All in the reference's layouts. */
int pddp_set_array(pddp_handle h, const char *name, const void *host_src, long nbytes);
int pddp_get_array(pddp_handle h, const char *name, void *host_dst, long nbytes);
int pddp_phase_load_init(pddp_handle h, const float *x0, const float *u0, const float *xGoal, int ignoreFirstDefectFlag); /* loadVarsGPU + initAlgGPU */
int pddp_phase_backward_pass(pddp_handle h);      /* backwardPassGPU      bpHelpers.cuh:484-517 */
int pddp_phase_forward_sweep(pddp_handle h);      /* forwardSweepKern     fpHelpers.cuh:55-63   */
int pddp_phase_forward_sim(pddp_handle h);        /* forwardSimKern + costKern + defectKern  fpHelpers.cuh:277-301,132-152,94-111 */
int pddp_phase_line_search(pddp_handle h);        /* fpHelpers.cuh:374-376,395-408 + acceptRejectTrajGPU nisInitHelpers.cuh:487-518 */
int pddp_phase_next_iteration(pddp_handle h);     /* nextIterationSetupGPU nisInitHelpers.cuh:245-279 */
/* device time (ms) of the last phase call and the kernels it launched */
int pddp_last_phase_stats(pddp_handle h, double *ms, int *launches);

/* Number of independent problem groups, each iterated on its own CUDA stream so that one group's latency-bound kernels
 * (backward pass, sweep, selection) overlap another group's throughput-bound ones (sim, next-iteration setup).  Results do
 * not depend on it.  Default 2 (env PDDP_GROUPS) for batches of 4 problems or more, else 1; the per-phase entries of times_ms and
 * pddp_last_iteration_times need 1 group (the phases of different groups overlap, there is no per-phase time then).
 * Returns the value in effect. */
int pddp_set_groups(pddp_handle h, int groups);

/* Per-iteration device times (ms) of the last pddp_solve* call that was given a times_ms array and ran as ONE problem group
 * (pddp_set_groups(h, 1); a batch below 4 problems always does): the entries of the reference's simTime / sweepTime / bpTime / nisTime
 * arrays (DDPWrappers.cuh:60-107).  sim includes the cost / defect reductions and the line search, as in the reference.  Returns the
 * number of iterations written (<= capacity); any array may be NULL. */
int pddp_last_iteration_times(pddp_handle h, double *sim_ms, double *sweep_ms, double *bp_ms, double *nis_ms, int capacity);
/* max_d of runiLQR_GPU's summary line (DDPWrappers.cuh:134): the largest L1 defect over the shooting-interval boundaries of each
 * problem's final trajectory, [batch] floats (HOST) */
int pddp_final_max_defect(pddp_handle h, float *max_d);

/* ---- line search sharded over GPUs (SURVEY 8e second paragraph; north_star "a single NCCL exchange at line-search selection").
 * One process per GPU, every rank owns a handle of the SAME configuration and calls pddp_solve* with the SAME inputs.  After
 * pddp_alpha_shard_init a solve runs like this on every rank: backward pass and next-iteration setup replicated, forward sweep and
 * simulation only for the rank's n_alpha / nranks step sizes, then ONE ncclAllGather of the (J, defect) pairs of all step sizes, the
 * reference's sequential scan (fpHelpers.cuh:395-408) on identical data on every rank -- so the chosen step size, the rho schedule and
 * the exit decisions are those of the unsharded solve bit for bit --, and the accepted candidate handed from its owner to the others
 * (ncclAllReduce of its bit patterns, 4 N (2n+m) bytes per problem).  NCCL is bound with dlopen("libnccl.so.2") when this is called.
 *   pddp_alpha_shard_unique_id: 128 bytes (ncclUniqueId) made by rank 0 and given to all ranks by the caller (any transport)
 *   pddp_alpha_shard_stats: mean device time of the two collectives per iteration (us, sampled every 8th iteration) and the rank's range
 * Cold starts of the joint-space cost on PLANT 4 only; n_alpha must be a multiple of nranks. */
int pddp_alpha_shard_unique_id(void *id128);
int pddp_alpha_shard_init(pddp_handle h, int rank, int nranks, const void *id128);
int pddp_alpha_shard_stats(pddp_handle h, double *exchange_us_per_iteration, int *a_first, int *a_cnt);

/* Device-resident iteration loop (DDPWrappers.cuh:52-114 is a host loop with >= 10 synchronisations per iteration):
 * the iterations of a solve are replayed from CUDA graphs of `iterations_per_graph` iterations each (default 10, env PDDP_GRAPH_CHUNK),
 * captured at the first solve of a shape; with TOL_COST > 0 the host polls the number of unfinished problems between graphs.  On by
 * default (env PDDP_GRAPHS=0 / on = 0: plain stream launches).  A solve that asks for per-phase times with one problem group is timed
 * launch by launch instead.  pddp_last_graph_launch_count: graphs launched by the last solve (pddp_last_launch_count counts the kernels
 * inside them). */
int pddp_set_graphs(pddp_handle h, int on, int iterations_per_graph);
long pddp_last_graph_launch_count(pddp_handle h);

/* Shape of the backward pass of PLANT 4: 0 = chosen by launch size (default, env PDDP_BP_SHAPE), 1 = warp chains (bp_warp.cuh), 2 = block-
 * cooperative (kernels.cuh).  Identical results. */
int pddp_set_bp_shape(pddp_handle h, int shape);

/* number of kernels launched by the last pddp_solve* call on this handle */
long pddp_last_launch_count(pddp_handle h);

/* Warm-start inputs of runiLQR_GPU (its KT0, P0, p0, d0 arguments, DDPWrappers.cuh:8): HOST arrays [batch][N][.] in the
 * reference layouts -- KT0 14x7 (98 floats per knot), P0 14x14 (196), p0 14, d0 14.  They are kept on the device and used
 * by every later solve with clearVarsFlag = 0. */
int pddp_set_warm_start(pddp_handle h, const float *KT0, const float *P0, const float *p0, const float *d0);
/* loadVarsGPU flags of the NEXT pddp_solve_device call (pddp_solve takes them as arguments); they fall back to
 * rollout = 0, clear = 1 afterwards. */
int pddp_set_start_mode(pddp_handle h, int forwardRolloutFlag, int clearVarsFlag);

/* ---- receding horizon: runiLQR_MPC_GPU (DDPHelpers/MPCHelpers.cuh:862-1045) for `batch` independent arms.
 * The solver keeps what the reference keeps on the device between solves (current plan and its defects, gains, both cost-to-go
 * buffers, the per-problem last_successful_solve counter).  The wall-clock budget of the reference (USE_MAX_SOLVER_TIME) is not
 * reproduced: max_iter is the limit.  Build the config with gravity = 0 to match the reference's MPC_MODE (dynamics_arm.cuh:42-43).
 *
 * pddp_mpc_init: x_init [batch][N][14], u_init [batch][N][7] (HOST) -- the plan the first step starts from (trajVars x,u and the
 *                current candidate slot of the reference); gains, cost-to-go and defects start at zero.
 * pddp_mpc_step: xActual [batch][14] measured state, xGoal [batch][14], shiftAmount [batch] = whole knots elapsed since the previous
 *                step (the reference derives it from its plant clock, MPCHelpers.cuh:875), max_iter <= config max_iter,
 *                clear_vars / ignoreFirstDefectFlag as in the reference.  Does loadVarsGPU_MPC (shift, zero-order hold, open-loop
 *                rollout from xActual over the whole horizon: FULL_ROLLOUT 1), initAlgGPU, the iteration loop and storeVarsGPU_MPC.
 *                x, u, KT (HOST, [batch][N][14|7|98]) play trajVars: they are overwritten only for the problems whose solve took a
 *                step (last_successful_solve == 1 afterwards); otherwise the device state falls back to the shifted previous plan.
 *                Jout / alphaOut [batch][config max_iter + 1], iters_out, last_successful_solve [batch] may be NULL. */
int pddp_mpc_init(pddp_handle h, const float *x_init, const float *u_init);
int pddp_mpc_step(pddp_handle h, const float *xActual, const float *xGoal, const int *shiftAmount, int max_iter, int clear_vars,
                  int ignoreFirstDefectFlag, float *x, float *u, float *KT, float *Jout, int *alphaOut, int *iters_out,
                  int *last_successful_solve);

/* ---- trajectory hand-off format: the LCM message drake::lcmt_trajectory_f (lcmtypes/drake/lcmt_trajectory_f.hpp) the reference's
 * MPC loop publishes to its trajectory runner (LCMHelpers.cuh:245-256).  Wire format: 8-byte fingerprint, utime (int64), x_size,
 * u_size, KT_size (int32), then x[x_size], u[u_size], KT[KT_size] floats, everything in network byte order.  Host-only helpers;
 * they return the number of bytes written / read or a negative PDDP_E_* code. */
long pddp_traj_f_encoded_size(int x_size, int u_size, int KT_size);
long pddp_traj_f_encode(long long utime, const float *x, int x_size, const float *u, int u_size, const float *KT, int KT_size,
                        void *buf, long capacity);
long pddp_traj_f_decode(const void *buf, long nbytes, long long *utime, int *x_size, int *u_size, int *KT_size,
                        float *x, float *u, float *KT, long cap_x, long cap_u, long cap_KT);
/* Exactly what LCMHelpers.cuh:245-252 builds from trajVars for `steps` = TRAJ_RUNNER_TIME_STEPS knots: the size fields are BYTE
 * counts and the arrays are that long, data in the first quarter, zeros behind (with_feedback = USE_FEEDBACK_IN_TRAJ_RUNNER).
 * buf == NULL returns the size needed. */
long pddp_traj_f_pack_reference(long long utime, const float *x, const float *u, const float *KT, int steps, int with_feedback,
                                void *buf, long capacity);

/* Opt-in (default off, env PDDP_SKIP_UNCHANGED=1): skip the gradient / cost-derivative refresh of a problem whose line search was
 * rejected -- its trajectory, hence AB, H, g, is unchanged, results are bit-identical.  The reference recomputes them
 * (nisInitHelpers.cuh:245-279) and the default does too, so that timed work matches the reference's iteration for iteration. */
int pddp_set_skip_unchanged(pddp_handle h, int on);

/* self-test: compares the library's reciprocal (pddp_math.cuh rcp_rn) with the IEEE division 1.0f/x the reference
 * compiles to (e.g. DDPHelpers/invHelpers.cuh pivot reciprocals) on all 2^32 float bit patterns; *mismatches = count. */
int pddp_selftest_rcp(unsigned long long *mismatches);
/* self-test: compares the branch-free restatement of sinf / cosf the Kuka dynamics use (plant_kuka.cuh sincos_as_library; the
 * reference calls sin() / cos() on floats, dynamics_arm.cuh:429-479) with the CUDA library's sinf and cosf on all 2^32 float bit
 * patterns, sign of zero and NaN-ness included; *mismatches = count. */
int pddp_selftest_sincos(unsigned long long *mismatches);

#ifdef __cplusplus
}
#endif
#endif
