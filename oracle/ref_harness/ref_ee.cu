/*
 * ref_ee.cu -- TEST INFRASTRUCTURE, not product code.
 *
 * The end-effector cost configuration (EE_COST 1) of the UNMODIFIED reference around a deterministic driver: the same
 * entry points as ref_driver.cu (allocateMemory_GPU / runiLQR_GPU, allocateMemory_CPU / runiLQR_CPU), compiled as their
 * EE_COST instantiation (plants/cost_arm.cuh:204-389, DDPHelpers/fpHelpers.cuh:259-265,298-300,
 * DDPHelpers/nisInitHelpers.cuh:51-84), with the goal of examples/WAFR_iLQR_examples.cu:37-43,101-104.
 * The cost weights are run-time arguments of the reference's functions; this driver sets the roll/pitch/yaw weights
 * (zero by default, cost_arm.cuh:107-117) to non-zero values so that the atan2 terms take part in the result.
 *
 * Modes
 *   solve <G|C> <seed0> <nseeds> <tol_cost> <out.bin>     (C = runiLQR_CPU, serial line search)
 *   unit  <G|H> <nsamples> <seed> <out.bin>               (per state: cost, g, H of costGradientHessianKern / ...Threaded)
 *   warm  G     <seed> <tol_cold> <tol_warm> <out.bin>    (cold solve, then warm starts with (rollout, clear) = (1,0), (0,0), (1,1))
 */
#define EE_COST 1
#define USE_WAFR_URDF 1
#define _Q_EE1 0.1
#define _Q_EE2 0.01
#define _R_EE 0.0001
#define _QF_EE1 1000.0
#define _QF_EE2 10.0
#define _Q_xdEE 0.1
#define _QF_xdEE 1000.0
#define _Q_xEE 0.001
#define _QF_xEE 1.0
double g_tol_cost = 0.0;
#define TOL_COST g_tol_cost

#include "config.cuh"
#include <random>
#include <string>
#include <vector>
#include <cmath>
#include <cstring>
#include <cstdlib>

typedef algType T;
#define NA NUM_ALPHA
#define NT NUM_TIME_STEPS
#define SENT_ALPHA (-99)

static FILE *g_out = nullptr;
static void dumpf(const char *name, const float *p, size_t n){ fprintf(g_out, "%s f32 %zu\n", name, n); fwrite(p, sizeof(float), n, g_out); }
static void dumpi(const char *name, const int *p, size_t n){ fprintf(g_out, "%s i32 %zu\n", name, n); fwrite(p, sizeof(int), n, g_out); }

// initial trajectory of WAFR_iLQR_examples.cu:67-95 with a fixed seed, goal of :37-43
static void loadXU_seeded(T *x, T *u, T *xGoal, int ld_x, int ld_u, unsigned seed){
	std::default_random_engine eng(seed);
	std::normal_distribution<double> dist(0.0, 0.001);
	const double PI_ = 3.14159;
	for (int k = 0; k < NT; k++){
		T *xk = x + k*ld_x;
		xk[0] = -0.5*PI_;	xk[1] = 0.25*PI_;	xk[2] = 0.167*PI_;
		xk[3] = -0.167*PI_;	xk[4] = 0.125*PI_;	xk[5] = 0.167*PI_;	xk[6] = 0.5*PI_;
		for (int k2 = 0; k2 < NUM_POS; k2++){xk[NUM_POS+k2] = static_cast<T>(dist(eng));}
	}
	for (int k = 0; k < NT; k++){
		T *uk = u + k*ld_u;
		uk[0] = 0.0;		uk[1] = -102.9832;	uk[2] = 11.1968;
		uk[3] = 47.0724;	uk[4] = 2.5993;		uk[5] = -7.0290;	uk[6] = -0.0907;
	}
	const T temp[] = {(T)0.3638, (T)0.0, (T)1.0628, (T)(0.5*PI_), (T)0.0, (T)(0.5*PI_)};
	for (int i = 0; i < 6; i++){xGoal[i] = temp[i];}
}

struct GpuVars {
	int ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A;
	cudaStream_t *streams;
	T *alpha, *d_alpha; int *alphaIndex;
	T *d_P, *d_p, *d_Pp, *d_pp, *d_AB, *d_H, *d_g, *d_KT, *d_du;
	T **d_x, **d_u, **h_d_x, **h_d_u, *d_xp, *d_xp2, *d_up, *d_JT, *J;
	T **d_d, **h_d_d, *d_dp, *d_dT, *d, *d_ApBK, *d_Bdu, *d_dM;
	int *err, *d_err;
	T *dJexp, *d_dJexp;
	T *xGoal, *d_xGoal;
	T *d_I, *d_Tbody;
};
static void gpu_alloc(GpuVars &v){
	allocateMemory_GPU<T>(&v.d_x, &v.h_d_x, &v.d_xp, &v.d_xp2, &v.d_u, &v.h_d_u, &v.d_up, &v.d_xGoal, &v.xGoal,
		&v.d_P, &v.d_Pp, &v.d_p, &v.d_pp, &v.d_AB, &v.d_H, &v.d_g, &v.d_KT, &v.d_du,
		&v.d_d, &v.h_d_d, &v.d_dp, &v.d_dT, &v.d_dM, &v.d, &v.d_ApBK, &v.d_Bdu,
		&v.d_JT, &v.J, &v.d_dJexp, &v.dJexp, &v.alpha, &v.d_alpha, &v.alphaIndex, &v.d_err, &v.err,
		&v.ld_x, &v.ld_u, &v.ld_P, &v.ld_p, &v.ld_AB, &v.ld_H, &v.ld_g, &v.ld_KT, &v.ld_du, &v.ld_d, &v.ld_A,
		&v.streams, &v.d_I, &v.d_Tbody);
	// arrays the reference never fully writes (cudaMalloc garbage otherwise)
	gpuErrchk(cudaMemset(v.d_AB, 0, v.ld_AB*DIM_AB_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_H, 0, v.ld_H*DIM_H_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_g, 0, v.ld_g*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_ApBK, 0, v.ld_A*DIM_A_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_Bdu, 0, v.ld_d*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_dp, 0, v.ld_d*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_xp2, 0, v.ld_x*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_dJexp, 0, 2*M_BLOCKS_B*sizeof(T)));
	gpuErrchk(cudaDeviceSynchronize());
}

static int count_iters(const int *alphaOut){int it = 0; for (int i = 1; i <= MAX_ITER; i++){if (alphaOut[i] != SENT_ALPHA){it = i;}} return it;}

static int run_solve(char hw, unsigned seed0, int nseeds){
	std::vector<T> Jout((MAX_ITER+1)*nseeds, NAN); std::vector<int> alphaOut((MAX_ITER+1)*nseeds, SENT_ALPHA);
	std::vector<double> tTime(nseeds), initTime(nseeds), simT(MAX_ITER), swT(MAX_ITER), bpT(MAX_ITER), nisT(MAX_ITER);
	std::vector<int> iters(nseeds);
	std::vector<T> xin, uin, xout, uout, goal(6);
	if (hw == 'G'){
		GpuVars v; gpu_alloc(v);
		std::vector<T> x0(v.ld_x*NT), u0(v.ld_u*NT);
		for (int i = 0; i < nseeds; i++){
			loadXU_seeded(x0.data(), u0.data(), v.xGoal, v.ld_x, v.ld_u, seed0+i);
			xin.insert(xin.end(), x0.begin(), x0.end()); uin.insert(uin.end(), u0.begin(), u0.end());
			runiLQR_GPU<T>(x0.data(), u0.data(), nullptr, nullptr, nullptr, nullptr, v.xGoal, &Jout[i*(MAX_ITER+1)], &alphaOut[i*(MAX_ITER+1)], 0, 1, 1,
				&tTime[i], simT.data(), swT.data(), bpT.data(), nisT.data(), &initTime[i], v.streams,
				v.d_x, v.h_d_x, v.d_xp, v.d_xp2, v.d_u, v.h_d_u, v.d_up, v.d_P, v.d_p, v.d_Pp, v.d_pp, v.d_AB, v.d_H, v.d_g, v.d_KT, v.d_du,
				v.d_d, v.h_d_d, v.d_dp, v.d_dT, v.d, v.d_ApBK, v.d_Bdu, v.d_dM, v.alpha, v.d_alpha, v.alphaIndex, v.d_JT, v.J, v.dJexp, v.d_dJexp, v.d_xGoal,
				v.err, v.d_err, v.ld_x, v.ld_u, v.ld_P, v.ld_p, v.ld_AB, v.ld_H, v.ld_g, v.ld_KT, v.ld_du, v.ld_d, v.ld_A, v.d_I, v.d_Tbody);
			iters[i] = count_iters(&alphaOut[i*(MAX_ITER+1)]);
			xout.insert(xout.end(), x0.begin(), x0.end()); uout.insert(uout.end(), u0.begin(), u0.end());
		}
		for (int i = 0; i < 6; i++){goal[i] = v.xGoal[i];}
	}
	else{
		int ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A;
		T *alpha, *P, *p, *Pp, *pp, *AB, *H, *g, *KT, *du, *x, *u, *xp, *xp2, *up, *JT, *d, *dp, *ApBK, *Bdu, *dJexp, *xGoal, *I, *Tbody;
		int *err;
		allocateMemory_CPU<T>(&x, &xp, &xp2, &u, &up, &xGoal, &P, &Pp, &p, &pp, &AB, &H, &g, &KT, &du, &d, &dp, &ApBK, &Bdu,
			&JT, &dJexp, &alpha, &err, &ld_x, &ld_u, &ld_P, &ld_p, &ld_AB, &ld_H, &ld_g, &ld_KT, &ld_du, &ld_d, &ld_A, &I, &Tbody);
		std::vector<T> x0(ld_x*NT), u0(ld_u*NT);
		for (int i = 0; i < nseeds; i++){
			loadXU_seeded(x0.data(), u0.data(), xGoal, ld_x, ld_u, seed0+i);
			xin.insert(xin.end(), x0.begin(), x0.end()); uin.insert(uin.end(), u0.begin(), u0.end());
			runiLQR_CPU<T>(x0.data(), u0.data(), nullptr, nullptr, nullptr, nullptr, xGoal, &Jout[i*(MAX_ITER+1)], &alphaOut[i*(MAX_ITER+1)], 0, 1, 1,
				&tTime[i], simT.data(), swT.data(), bpT.data(), nisT.data(), &initTime[i],
				x, xp, xp2, u, up, P, p, Pp, pp, AB, H, g, KT, du, d, dp, ApBK, Bdu, alpha, JT, dJexp, err,
				ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A, I, Tbody);
			iters[i] = count_iters(&alphaOut[i*(MAX_ITER+1)]);
			xout.insert(xout.end(), x0.begin(), x0.end()); uout.insert(uout.end(), u0.begin(), u0.end());
		}
		for (int i = 0; i < 6; i++){goal[i] = xGoal[i];}
	}
	int meta[4] = {NT, NA, M_BLOCKS, nseeds}; dumpi("meta", meta, 4);
	const float wts[9] = {(float)_Q_EE1, (float)_Q_EE2, (float)_QF_EE1, (float)_QF_EE2, (float)_R_EE, (float)_Q_xdEE, (float)_QF_xdEE, (float)_Q_xEE, (float)_QF_xEE};
	dumpf("weights", wts, 9); dumpf("xGoal", goal.data(), 6);
	dumpf("Jout", Jout.data(), Jout.size()); dumpi("alphaOut", alphaOut.data(), alphaOut.size()); dumpi("iters", iters.data(), nseeds);
	dumpf("x_in", xin.data(), xin.size()); dumpf("u_in", uin.data(), uin.size());
	dumpf("x_out", xout.data(), xout.size()); dumpf("u_out", uout.data(), uout.size());
	return 0;
}

// ---------------------------------------------------------------- unit: cost / g / H of single states, knot index 0 (running) and NT-1 (final)
static int run_unit(char hw, int n, unsigned seed){
	std::default_random_engine eng(seed);
	std::normal_distribution<double> dq(0.0, 1.0), dqd(0.0, 2.0), duu(0.0, 50.0);
	// the reference's kernels index knots 0..NT-1 of one trajectory: states are laid out as n/NT trajectories of NT knots
	const int ntraj = (n + NT - 1) / NT; n = ntraj * NT;
	std::vector<T> x(n*STATE_SIZE), u(n*CONTROL_SIZE), J(n), H(n*DIM_H_r*DIM_H_c, 0), g(n*DIM_g_r), xg(6);
	{std::vector<T> xx(STATE_SIZE*NT), uu(CONTROL_SIZE*NT); loadXU_seeded(xx.data(), uu.data(), xg.data(), STATE_SIZE, CONTROL_SIZE, seed);
	 for (int k = 0; k < n; k++){
		for (int i = 0; i < STATE_SIZE; i++){x[k*STATE_SIZE+i] = static_cast<T>(i < NUM_POS ? dq(eng) : dqd(eng));}
		for (int i = 0; i < CONTROL_SIZE; i++){u[k*CONTROL_SIZE+i] = static_cast<T>(duu(eng));}
	 }}
	std::vector<T> Tbody(36*NUM_POS); initT<T>(Tbody.data());
	if (hw == 'G'){
		T *d_x, *d_u, *d_g, *d_H, *d_xg, *d_Tb, *d_JT;
		gpuErrchk(cudaMalloc(&d_x, x.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_u, u.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_g, g.size()*sizeof(T)));
		gpuErrchk(cudaMalloc(&d_H, H.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_xg, 6*sizeof(T))); gpuErrchk(cudaMalloc(&d_Tb, Tbody.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_JT, n*sizeof(T)));
		gpuErrchk(cudaMemcpy(d_x, x.data(), x.size()*sizeof(T), cudaMemcpyHostToDevice)); gpuErrchk(cudaMemcpy(d_u, u.data(), u.size()*sizeof(T), cudaMemcpyHostToDevice));
		gpuErrchk(cudaMemcpy(d_xg, xg.data(), 6*sizeof(T), cudaMemcpyHostToDevice)); gpuErrchk(cudaMemcpy(d_Tb, Tbody.data(), Tbody.size()*sizeof(T), cudaMemcpyHostToDevice));
		gpuErrchk(cudaMemset(d_H, 0, H.size()*sizeof(T)));
		for (int t = 0; t < ntraj; t++){
			costGradientHessianKern<T><<<NT,dim3(8,7)>>>(d_x + t*NT*STATE_SIZE, d_u + t*NT*CONTROL_SIZE, d_g + t*NT*DIM_g_r, d_H + t*NT*DIM_H_r*DIM_H_c, d_xg,
				STATE_SIZE, CONTROL_SIZE, DIM_H_r, DIM_g_r, d_Tb, d_JT + t*NT);
			gpuErrchk(cudaPeekAtLastError());
		}
		gpuErrchk(cudaDeviceSynchronize());
		gpuErrchk(cudaMemcpy(g.data(), d_g, g.size()*sizeof(T), cudaMemcpyDeviceToHost)); gpuErrchk(cudaMemcpy(H.data(), d_H, H.size()*sizeof(T), cudaMemcpyDeviceToHost));
		gpuErrchk(cudaMemcpy(J.data(), d_JT, n*sizeof(T), cudaMemcpyDeviceToHost));
	}
	else{
		for (int t = 0; t < ntraj; t++){
			for (int k = 0; k < NT; k++){
				threadDesc_t one; one.dim = 1; one.tid = k; one.reps = 1;        // tid doubles as the knot index and the JT slot
				costGradientHessianThreaded<T>(one, &x[t*NT*STATE_SIZE], &u[t*NT*CONTROL_SIZE], &g[t*NT*DIM_g_r], &H[t*NT*DIM_H_r*DIM_H_c], xg.data(),
					STATE_SIZE, CONTROL_SIZE, DIM_H_r, DIM_g_r, Tbody.data(), &J[t*NT]);
			}
		}
	}
	int meta[2] = {NT, n}; dumpi("meta", meta, 2);
	const float wts[9] = {(float)_Q_EE1, (float)_Q_EE2, (float)_QF_EE1, (float)_QF_EE2, (float)_R_EE, (float)_Q_xdEE, (float)_QF_xdEE, (float)_Q_xEE, (float)_QF_xEE};
	dumpf("weights", wts, 9); dumpf("xGoal", xg.data(), 6);
	dumpf("x", x.data(), x.size()); dumpf("u", u.data(), u.size()); dumpf("J", J.data(), J.size()); dumpf("g", g.data(), g.size()); dumpf("H", H.data(), H.size());
	return 0;
}

// ---------------------------------------------------------------- warm starts of loadVarsGPU under the end-effector cost (the rollout
// takes its initial cost from forwardSimKern's partials, nisInitHelpers.cuh:384,646-651)
static void dumpd(const char *name, const double *p, size_t n){ fprintf(g_out, "%s f64 %zu\n", name, n); fwrite(p, sizeof(double), n, g_out); }
static int run_warm(unsigned seed, double tol1, double tol2){
	GpuVars v; gpu_alloc(v);
	std::vector<T> x0(v.ld_x*NT), u0(v.ld_u*NT);
	std::vector<T> Jout(MAX_ITER+1, NAN); std::vector<int> alphaOut(MAX_ITER+1, SENT_ALPHA);
	std::vector<double> simT(MAX_ITER), swT(MAX_ITER), bpT(MAX_ITER), nisT(MAX_ITER); double tTime, initTime;
	loadXU_seeded(x0.data(), u0.data(), v.xGoal, v.ld_x, v.ld_u, seed);
	g_tol_cost = tol1;
	#define RUN_(ROLL, CLEAR, KT0_, P0_, p0_, d0_) runiLQR_GPU<T>(x0.data(), u0.data(), KT0_, P0_, p0_, d0_, v.xGoal, Jout.data(), alphaOut.data(), ROLL, CLEAR, 1, \
		&tTime, simT.data(), swT.data(), bpT.data(), nisT.data(), &initTime, v.streams, \
		v.d_x, v.h_d_x, v.d_xp, v.d_xp2, v.d_u, v.h_d_u, v.d_up, v.d_P, v.d_p, v.d_Pp, v.d_pp, v.d_AB, v.d_H, v.d_g, v.d_KT, v.d_du, \
		v.d_d, v.h_d_d, v.d_dp, v.d_dT, v.d, v.d_ApBK, v.d_Bdu, v.d_dM, v.alpha, v.d_alpha, v.alphaIndex, v.d_JT, v.J, v.dJexp, v.d_dJexp, v.d_xGoal, \
		v.err, v.d_err, v.ld_x, v.ld_u, v.ld_P, v.ld_p, v.ld_AB, v.ld_H, v.ld_g, v.ld_KT, v.ld_du, v.ld_d, v.ld_A, v.d_I, v.d_Tbody)
	RUN_(0, 1, nullptr, nullptr, nullptr, nullptr);
	dumpf("J1", Jout.data(), Jout.size()); dumpi("alpha1", alphaOut.data(), alphaOut.size());
	std::vector<T> KT0(v.ld_KT*DIM_KT_c*NT), P0(v.ld_P*DIM_P_c*NT), p0(v.ld_p*NT), d0(v.ld_d*NT);
	gpuErrchk(cudaMemcpy(KT0.data(), v.d_KT, KT0.size()*sizeof(T), cudaMemcpyDeviceToHost));
	gpuErrchk(cudaMemcpy(P0.data(), v.d_P, P0.size()*sizeof(T), cudaMemcpyDeviceToHost));
	gpuErrchk(cudaMemcpy(p0.data(), v.d_p, p0.size()*sizeof(T), cudaMemcpyDeviceToHost));
	gpuErrchk(cudaMemcpy(d0.data(), v.h_d_d[*v.alphaIndex], d0.size()*sizeof(T), cudaMemcpyDeviceToHost));
	// the reference never writes the last knot of KT / P / p: zero what it left there so that the inputs are well defined
	for (int i = 0; i < v.ld_KT*DIM_KT_c; i++){KT0[(size_t)v.ld_KT*DIM_KT_c*(NT-1) + i] = 0;}
	for (int i = 0; i < v.ld_P*DIM_P_c; i++){P0[(size_t)v.ld_P*DIM_P_c*(NT-1) + i] = 0;}
	for (int i = 0; i < v.ld_p; i++){p0[(size_t)v.ld_p*(NT-1) + i] = 0;}
	// perturbed start: measured first knot off the plan, goal pose moved
	std::vector<T> xs = x0, us = u0;
	for (int i = 0; i < NUM_POS; i++){xs[i] += static_cast<T>(0.01*(i+1)/NUM_POS); xs[NUM_POS+i] += static_cast<T>(0.02*(NUM_POS-i)/NUM_POS);}
	v.xGoal[0] += static_cast<T>(0.05); v.xGoal[2] -= static_cast<T>(0.03); v.xGoal[4] += static_cast<T>(0.1);
	dumpf("x_in", xs.data(), xs.size()); dumpf("u_in", us.data(), us.size()); dumpf("xGoal", v.xGoal, 6);
	dumpf("KT0", KT0.data(), KT0.size()); dumpf("P0", P0.data(), P0.size()); dumpf("p0", p0.data(), p0.size()); dumpf("d0", d0.data(), d0.size());
	g_tol_cost = tol2;
	const int flags[3][2] = {{1,0},{0,0},{1,1}};
	for (int c = 0; c < 3; c++){
		x0 = xs; u0 = us; std::fill(Jout.begin(), Jout.end(), NAN); std::fill(alphaOut.begin(), alphaOut.end(), SENT_ALPHA);
		RUN_(flags[c][0], flags[c][1], KT0.data(), P0.data(), p0.data(), d0.data());
		char nmb[32];
		snprintf(nmb, sizeof nmb, "Jout_%d%d", flags[c][0], flags[c][1]); dumpf(nmb, Jout.data(), Jout.size());
		snprintf(nmb, sizeof nmb, "alphaOut_%d%d", flags[c][0], flags[c][1]); dumpi(nmb, alphaOut.data(), alphaOut.size());
		snprintf(nmb, sizeof nmb, "x_out_%d%d", flags[c][0], flags[c][1]); dumpf(nmb, x0.data(), x0.size());
		snprintf(nmb, sizeof nmb, "u_out_%d%d", flags[c][0], flags[c][1]); dumpf(nmb, u0.data(), u0.size());
	}
	#undef RUN_
	int meta[4] = {NT, NA, M_BLOCKS, 1}; dumpi("meta", meta, 4); dumpf("alpha", v.alpha, NA);
	const float wts[9] = {(float)_Q_EE1, (float)_Q_EE2, (float)_QF_EE1, (float)_QF_EE2, (float)_R_EE, (float)_Q_xdEE, (float)_QF_xdEE, (float)_Q_xEE, (float)_QF_xEE};
	dumpf("weights", wts, 9);
	double tols[2] = {tol1, tol2}; dumpd("tols", tols, 2);
	return 0;
}

int main(int argc, char **argv){
	if (argc < 2){fprintf(stderr, "usage: see header of ref_ee.cu\n"); return 2;}
	std::string mode(argv[1]);
	if (mode == "solve" && argc == 7){
		g_tol_cost = atof(argv[5]); g_out = fopen(argv[6], "wb"); if (!g_out){perror("open"); return 1;}
		int rc = run_solve(argv[2][0], (unsigned)atoi(argv[3]), atoi(argv[4])); fclose(g_out); return rc;
	}
	if (mode == "unit" && argc == 6){
		g_out = fopen(argv[5], "wb"); if (!g_out){perror("open"); return 1;}
		int rc = run_unit(argv[2][0], atoi(argv[3]), (unsigned)atoi(argv[4])); fclose(g_out); return rc;
	}
	if (mode == "warm" && argc == 7 && argv[2][0] == 'G'){
		g_out = fopen(argv[6], "wb"); if (!g_out){perror("open"); return 1;}
		int rc = run_warm((unsigned)atoi(argv[3]), atof(argv[4]), atof(argv[5])); fclose(g_out); return rc;
	}
	fprintf(stderr, "bad arguments\n"); return 2;
}
