/*
 * ref_mpc.cu -- TEST INFRASTRUCTURE, not product code.
 *
 * Deterministic driver around the UNMODIFIED reference's receding-horizon wrapper runiLQR_MPC_GPU
 * (DDPHelpers/MPCHelpers.cuh:862-1045), built with MPC_MODE 1 (config.cuh:185-186,272-273: gravity 0) and without the
 * wall-clock budget (USE_MAX_SOLVER_TIME 0), so that the iteration cap is the only exit besides convergence.
 * It plays what the LCM examples do around that call (LCMHelpers.cuh:224-243) with a synthetic plant clock and a synthetic
 * measured state: step s happens `shift` knots after step s-1, the measured state is the planned state at that knot plus a
 * fixed perturbation.  After every step it dumps the published trajectory (trajVars x, u, KT), the failure counter and the
 * cost / step-size trace of the solve.
 *
 *   mpc <seed> <nsteps> <shift> <max_iter> <out.bin> [use_cost_shift]
 *
 * Built twice (oracle/Makefile): ref_mpc_N* with the joint-space cost, ref_mpc_ee_N* with -DEE_COST=1 -- the configuration of
 * examples/WAFR_MPC_examples.cu: end-effector pose goal, nominal-state terms measured from xTarget (always non-null on this
 * path, MPCHelpers.cuh:889,900), roll/pitch/yaw weights set non-zero here so that the atan2 terms take part.
 */
#define PLANT 4
#ifndef EE_COST
#define EE_COST 0
#endif
#if EE_COST
#define _Q_EE1 0.1
#define _Q_EE2 0.01
#define _R_EE 0.0001
#define _QF_EE1 1000.0
#define _QF_EE2 10.0
#define _Q_xdEE 0.1
#define _QF_xdEE 1000.0
#define _Q_xEE 0.001
#define _QF_xEE 1.0
#endif
#define MPC_MODE 1
#define USE_MAX_SOLVER_TIME 0
#define USE_WAFR_URDF 1
#define _Q1 0.1
#define _Q2 0.001
#define _R  0.0001
#define _QF1 1000.0
#define _QF2 1000.0
#define TOL_COST 0.0001

#include "config.cuh"
#include <random>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdlib>

typedef algType T;
#define NT NUM_TIME_STEPS
static FILE *g_out = nullptr;
static void dumpraw(const char *name, const char *dtype, const void *p, size_t n, size_t sz){ fprintf(g_out, "%s %s %zu\n", name, dtype, n); fwrite(p, sz, n, g_out); }
static void dumpf(const std::string &name, const float *p, size_t n){ dumpraw(name.c_str(), "f32", p, n, 4); }
static void dumpi(const std::string &name, const int *p, size_t n){ dumpraw(name.c_str(), "i32", p, n, 4); }
static std::string nm(const char *base, int s){ char b[64]; snprintf(b, sizeof b, "s%d.%s", s, base); return b; }

int main(int argc, char **argv){
	if ((argc != 7 && argc != 8) || std::string(argv[1]) != "mpc"){fprintf(stderr, "usage: ref_mpc mpc <seed> <nsteps> <shift> <max_iter> <out.bin> [use_cost_shift]\n"); return 2;}
	const bool use_cost_shift = argc == 8 && atoi(argv[7]) != 0;      // runiLQR_MPC_GPU's last argument: final pose weights on the last shift+1 knots (MPCHelpers.cuh:876)
	const unsigned seed = (unsigned)atoi(argv[2]); const int nsteps = atoi(argv[3]), shift = atoi(argv[4]), max_iter = atoi(argv[5]);
	g_out = fopen(argv[6], "wb"); if (!g_out){perror("open"); return 1;}
	trajVars<T> tv; GPUVars<T> gv; matDimms md; algTrace<T> data; costParams<T> cst;
	allocateMemory_GPU_MPC<T>(&gv, &md, &tv);
	loadCost<T>(&cst);
	// initial plan: the WAFR start posture with small random velocities, gravity-compensation torques (WAFR_iLQR_examples.cu:67-121)
	std::default_random_engine eng(seed); std::normal_distribution<double> dist(0.0, 0.001);
	for (int k = 0; k < NT; k++){
		T *xk = tv.x + k*md.ld_x;
		xk[0] = -0.5*PI; xk[1] = 0.25*PI; xk[2] = 0.167*PI; xk[3] = -0.167*PI; xk[4] = 0.125*PI; xk[5] = 0.167*PI; xk[6] = 0.5*PI;
		for (int i = 0; i < NUM_POS; i++){xk[NUM_POS+i] = static_cast<T>(dist(eng));}
	}
	for (int k = 0; k < NT; k++){for (int i = 0; i < CONTROL_SIZE; i++){tv.u[k*md.ld_u+i] = static_cast<T>(0.01*(i+1));}}
	memset(tv.KT, 0, md.ld_KT*DIM_KT_c*NT*sizeof(T));
	const T goal[] = {0,0,0,-0.25*PI,0,0.25*PI,0.5*PI,0,0,0,0,0,0,0};
	#if EE_COST
		const T pose[] = {(T)0.3638, (T)0.0, (T)1.0628, (T)(0.5*PI), (T)0.0, (T)(0.5*PI)};
		for (int i = 0; i < 6; i++){gv.xGoal[i] = pose[i];}
		for (int i = 0; i < STATE_SIZE; i++){gv.xTarget[i] = goal[i];}
	#else
		for (int i = 0; i < STATE_SIZE; i++){gv.xGoal[i] = goal[i]; gv.xTarget[i] = goal[i];}
	#endif
	// device state the wrapper expects to find: the plan in candidate slot alphaIndex = 0, everything else zero
	gpuErrchk(cudaMemcpy(gv.h_d_x[0], tv.x, md.ld_x*NT*sizeof(T), cudaMemcpyHostToDevice));
	gpuErrchk(cudaMemcpy(gv.h_d_u[0], tv.u, md.ld_u*NT*sizeof(T), cudaMemcpyHostToDevice));
	gpuErrchk(cudaMemcpy(gv.d_xp, tv.x, md.ld_x*NT*sizeof(T), cudaMemcpyHostToDevice));
	gpuErrchk(cudaMemcpy(gv.d_up, tv.u, md.ld_u*NT*sizeof(T), cudaMemcpyHostToDevice));
	gpuErrchk(cudaMemset(gv.d_xp2, 0, md.ld_x*NT*sizeof(T)));
	for (int a = 0; a < NUM_ALPHA; a++){
		gpuErrchk(cudaMemset(gv.h_d_d[a], 0, md.ld_d*NT*sizeof(T)));
		if (a){gpuErrchk(cudaMemset(gv.h_d_x[a], 0, md.ld_x*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.h_d_u[a], 0, md.ld_u*NT*sizeof(T)));}
	}
	gpuErrchk(cudaMemset(gv.d_KT, 0, md.ld_KT*DIM_KT_c*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.d_du, 0, md.ld_du*NT*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_P, 0, md.ld_P*DIM_P_c*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.d_Pp, 0, md.ld_P*DIM_P_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_p, 0, md.ld_p*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.d_pp, 0, md.ld_p*NT*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_AB, 0, md.ld_AB*DIM_AB_c*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.d_H, 0, md.ld_H*DIM_H_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_g, 0, md.ld_g*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.d_dp, 0, md.ld_d*NT*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_ApBK, 0, md.ld_A*DIM_A_c*NT*sizeof(T))); gpuErrchk(cudaMemset(gv.d_Bdu, 0, md.ld_d*NT*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_dT, 0, NUM_ALPHA*sizeof(T))); gpuErrchk(cudaMemset(gv.d_JT, 0, NUM_ALPHA*sizeof(T)));
	gpuErrchk(cudaMemset(gv.d_dJexp, 0, 2*M_BLOCKS_B*sizeof(T)));
	gpuErrchk(cudaDeviceSynchronize());
	dumpf("x_init", tv.x, md.ld_x*NT); dumpf("u_init", tv.u, md.ld_u*NT); dumpf("xGoal", gv.xGoal, EE_COST ? 6 : STATE_SIZE);
	#if EE_COST
		dumpf("xTarget", gv.xTarget, STATE_SIZE);
		{const float wts[9] = {(float)cst.Q_EE1, (float)cst.Q_EE2, (float)cst.QF_EE1, (float)cst.QF_EE2, (float)cst.R_EE, (float)cst.Q_xdEE, (float)cst.QF_xdEE, (float)cst.Q_xEE, (float)cst.QF_xEE};
		 dumpf("weights", wts, 9);}
	#endif
	const double dt_us = TIME_STEP_LENGTH_IN_us;
	int64_t t_plant = 1000000;      // arbitrary clock origin
	std::vector<int> shifts, iters_per_step, lss;
	for (int s = 0; s < nsteps; s++){
		int expect_shift = 0;
		if (tv.first_pass){
			tv.t0_plant = t_plant; tv.t0_sys = t_plant; tv.first_pass = false;
			for (int i = 0; i < STATE_SIZE; i++){gv.xActual[i] = tv.x[i];}
		}
		else{
			// the plant clock advances by `shift` knots and a half (floor() of the reference then yields exactly `shift`)
			t_plant = tv.t0_plant + static_cast<int64_t>((shift + 0.5)*dt_us); expect_shift = shift;
			for (int i = 0; i < STATE_SIZE; i++){gv.xActual[i] = tv.x[shift*md.ld_x + i] + static_cast<T>(0.002*std::sin(1.0 + i + 3*s));}
		}
		dumpf(nm("xActual", s), gv.xActual, STATE_SIZE);
		const size_t j0 = data.J.size();
		runiLQR_MPC_GPU<T>(&tv, &gv, &md, &data, &cst, t_plant, t_plant, 0, max_iter, 1e12, s == 0 ? 1 : 0, use_cost_shift);
		std::vector<float> J(data.J.begin() + j0, data.J.end()); std::vector<int> al(data.alpha.begin() + j0, data.alpha.end());
		dumpf(nm("Jout", s), J.data(), J.size()); dumpi(nm("alphaOut", s), al.data(), al.size());
		dumpf(nm("x", s), tv.x, md.ld_x*NT); dumpf(nm("u", s), tv.u, md.ld_u*NT); dumpf(nm("KT", s), tv.KT, md.ld_KT*DIM_KT_c*NT);
		shifts.push_back(expect_shift); iters_per_step.push_back((int)J.size() - 1); lss.push_back(tv.last_successful_solve);
	}
	int meta[7] = {NT, NUM_ALPHA, M_BLOCKS, nsteps, shift, max_iter, use_cost_shift ? 1 : 0}; dumpi("meta", meta, 7);
	dumpi("shifts", shifts.data(), shifts.size()); dumpi("iters", iters_per_step.data(), iters_per_step.size()); dumpi("last_successful_solve", lss.data(), lss.size());
	fclose(g_out);
	return 0;
}
