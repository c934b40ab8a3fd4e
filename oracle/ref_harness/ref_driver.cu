/*
 * ref_driver.cu -- TEST INFRASTRUCTURE, not product code.
 *
 * A deterministic driver around the UNMODIFIED reference sources.  It includes
 * /root/reference/config.cuh at build time (never copied into this repo; see
 * oracle/Makefile) and calls the reference's own entry points:
 *   allocateMemory_GPU / runiLQR_GPU / freeMemory_GPU   (DDPHelpers/nisInitHelpers.cuh:766,
 *   DDPHelpers/DDPWrappers.cuh:8) and the CPU twins runiLQR_CPU / runiLQR_CPU2.
 * It replaces only what examples/WAFR_iLQR_examples.cu does around those calls:
 *   - time(0) seeding (WAFR_iLQR_examples.cu:60)  -> one fixed seed per problem
 *   - the interactive stdin prompt (:196-218)     -> none
 *   - and it writes Jout / alphaOut / x / u (and per-phase snapshots in `trace`
 *     mode) to a file so that tests/golden/ can pin the oracle and the CUDA path.
 *
 * Modes
 *   solve  <G|C|P> <seed0> <nseeds> <tol_cost> <out.bin>
 *   trace  <G|H>   <seed>  <tol_cost> <max_dump_iters> <out.bin>     (H = host math, GPU control flow)
 *   warm   G       <seed>  <tol_cold> <tol_warm> <out.bin>            (cold solve, then warm starts with (rollout, clear) = (1,0), (0,0), (1,1))
 *   unit   <G|H>   <nsamples> <seed> <out.bin>                       (dynamics / gradient / cost on random x,u)
 *   time   <G|C|P> <seed0> <nseeds> <tol_cost>                       (prints one summary line)
 */
#ifndef PLANT
#define PLANT 4
#endif
#define EE_COST 0
#if PLANT == 4
#define USE_WAFR_URDF 1
#define _Q1 0.1
#define _Q2 0.001
#define _R  0.0001
#define _QF1 1000.0
#define _QF2 1000.0
#endif
// PLANT 1-3 (pendulum, cart-pole, quadrotor): the plant files of the reference still have their v0.1 signatures and do not compile
// inside config.cuh at HEAD.  oracle/Makefile builds those configurations against a config.cuh whose three plant
// #include lines are redirected (sed, into oracle/_ref/gen/) to ref_harness/adapt_plant.cuh, which includes the SAME reference
// plant files inside a namespace and forwards the v0.2 call signatures to them.  Every other reference file is compiled as is.
// run-time cost tolerance: the macro is only used in host code (nisInitHelpers.cuh:393,395,456,512)
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

// ---------------------------------------------------------------- file output
static FILE *g_out = nullptr;
static void dumpf(const char *name, const float *p, size_t n){
	fprintf(g_out, "%s f32 %zu\n", name, n); fwrite(p, sizeof(float), n, g_out);
}
static void dumpi(const char *name, const int *p, size_t n){
	fprintf(g_out, "%s i32 %zu\n", name, n); fwrite(p, sizeof(int), n, g_out);
}
static void dumpd(const char *name, const double *p, size_t n){
	fprintf(g_out, "%s f64 %zu\n", name, n); fwrite(p, sizeof(double), n, g_out);
}
static void dump_dev(const char *name, const T *d_p, size_t n){
	std::vector<T> h(n); gpuErrchk(cudaMemcpy(h.data(), d_p, n*sizeof(T), cudaMemcpyDeviceToHost)); dumpf(name, h.data(), n);
}
static std::string nm(const char *base, int iter, const char *phase){
	char b[128]; snprintf(b, sizeof(b), "it%d.%s.%s", iter, phase, base); return std::string(b);
}

// ---------------------------------------------------------------- inputs (WAFR_iLQR_examples.cu:67-121, fixed seed)
static void loadXU_seeded(T *x, T *u, T *xGoal, int ld_x, int ld_u, unsigned seed){
	std::default_random_engine eng(seed);
	std::normal_distribution<double> dist(0.0, 0.001);
#if PLANT == 4
	for (int k = 0; k < NT; k++){
		T *xk = x + k*ld_x;
		xk[0] = -0.5*PI;	xk[1] = 0.25*PI;	xk[2] = 0.167*PI;
		xk[3] = -0.167*PI;	xk[4] = 0.125*PI;	xk[5] = 0.167*PI;	xk[6] = 0.5*PI;
		for (int k2 = 0; k2 < NUM_POS; k2++){xk[NUM_POS+k2] = static_cast<T>(dist(eng));}
	}
	for (int k = 0; k < NT; k++){
		T *uk = u + k*ld_u;
		uk[0] = 0.0;		uk[1] = -102.9832;	uk[2] = 11.1968;
		uk[3] = 47.0724;	uk[4] = 2.5993;		uk[5] = -7.0290;	uk[6] = -0.0907;
	}
	const T temp[] = {0,0,0,-0.25*PI,0,0.25*PI,0.5*PI,0,0,0,0,0,0,0};
#else
	// the other plants of the example (WAFR_iLQR_examples.cu:19-33,72-78,87-90,110-115): draws in knot order, state order
	for (int k = 0; k < NT; k++){
		T *xk = x + k*ld_x;
	#if PLANT == 1
		xk[0] = 0.0;	xk[1] = static_cast<T>(dist(eng));
	#elif PLANT == 2
		xk[0] = 0.0;	xk[1] = 0.0;	xk[2] = static_cast<T>(dist(eng));	xk[3] = static_cast<T>(dist(eng));
	#elif PLANT == 3
		for (int k2 = 0; k2 < STATE_SIZE; k2++){if (k2 == 2){xk[k2] = 0.5;} else if (k2 >= NUM_POS){xk[k2] = static_cast<T>(dist(eng));} else{xk[k2] = 0.0;}}
	#endif
	}
	for (int k = 0; k < NT; k++){
		T *uk = u + k*ld_u;
	#if PLANT == 3
		for (int k2 = 0; k2 < CONTROL_SIZE; k2++){uk[k2] = 1.22625;}
	#else
		uk[0] = 0.01;
	#endif
	}
	#if PLANT == 1
	const T temp[] = {3.1416, 0.0};
	#elif PLANT == 2
	const T temp[] = {0.0, 3.1416, 0.0, 0.0};
	#elif PLANT == 3
	const T temp[] = {7.0, 10.0, 0.5, 0,0,0,0,0,0,0,0,0};
	#endif
#endif
	for (int i = 0; i < STATE_SIZE; i++){xGoal[i] = temp[i];}
}

// ---------------------------------------------------------------- GPU state
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
	// the reference never writes AB[N-1], most of H[N-1], ApBK/Bdu[N-1] (cudaMalloc garbage): zero them so dumps are deterministic
	gpuErrchk(cudaMemset(v.d_AB, 0, v.ld_AB*DIM_AB_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_H, 0, v.ld_H*DIM_H_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_g, 0, v.ld_g*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_ApBK, 0, v.ld_A*DIM_A_c*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_Bdu, 0, v.ld_d*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_dp, 0, v.ld_d*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_xp2, 0, v.ld_x*NT*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_JT, 0, NA*sizeof(T)));
	gpuErrchk(cudaMemset(v.d_dJexp, 0, 2*M_BLOCKS_B*sizeof(T)));
	gpuErrchk(cudaDeviceSynchronize());
}
static void gpu_free(GpuVars &v){
	freeMemory_GPU<T>(v.d_x, v.h_d_x, v.d_xp, v.d_xp2, v.d_u, v.h_d_u, v.d_up, v.xGoal, v.d_xGoal, v.d_P, v.d_Pp, v.d_p, v.d_pp, v.d_AB, v.d_H, v.d_g, v.d_KT, v.d_du,
		v.d_d, v.h_d_d, v.d_dp, v.d_dM, v.d_dT, v.d, v.d_ApBK, v.d_Bdu, v.d_JT, v.J, v.d_dJexp, v.dJexp, v.alpha, v.d_alpha, v.alphaIndex, v.d_err, v.err,
		v.streams, v.d_I, v.d_Tbody);
}

static double now_ms(){struct timeval t; gettimeofday(&t, NULL); return get_time_ms(t);}

static int count_iters(const int *alphaOut){int it = 0; for (int i = 1; i <= MAX_ITER; i++){if (alphaOut[i] != SENT_ALPHA){it = i;}} return it;}

// ---------------------------------------------------------------- solve / time
static int run_solve(char hw, unsigned seed0, int nseeds, bool dump){
	std::vector<T> Jout((MAX_ITER+1)*nseeds, NAN); std::vector<int> alphaOut((MAX_ITER+1)*nseeds, SENT_ALPHA);
	std::vector<double> tTime(nseeds), initTime(nseeds), simT(MAX_ITER), swT(MAX_ITER), bpT(MAX_ITER), nisT(MAX_ITER);
	std::vector<int> iters(nseeds);
	std::vector<T> xin, uin, xout, uout;
	double t_all0 = 0, t_all1 = 0;
	if (hw == 'G'){
		GpuVars v; gpu_alloc(v);
		std::vector<T> x0(v.ld_x*NT), u0(v.ld_u*NT);
		t_all0 = now_ms();
		for (int i = 0; i < nseeds; i++){
			loadXU_seeded(x0.data(), u0.data(), v.xGoal, v.ld_x, v.ld_u, seed0+i);
			if (dump){xin.insert(xin.end(), x0.begin(), x0.end()); uin.insert(uin.end(), u0.begin(), u0.end());}
			runiLQR_GPU<T>(x0.data(), u0.data(), nullptr, nullptr, nullptr, nullptr, v.xGoal, &Jout[i*(MAX_ITER+1)], &alphaOut[i*(MAX_ITER+1)], 0, 1, 1,
				&tTime[i], simT.data(), swT.data(), bpT.data(), nisT.data(), &initTime[i], v.streams,
				v.d_x, v.h_d_x, v.d_xp, v.d_xp2, v.d_u, v.h_d_u, v.d_up, v.d_P, v.d_p, v.d_Pp, v.d_pp, v.d_AB, v.d_H, v.d_g, v.d_KT, v.d_du,
				v.d_d, v.h_d_d, v.d_dp, v.d_dT, v.d, v.d_ApBK, v.d_Bdu, v.d_dM, v.alpha, v.d_alpha, v.alphaIndex, v.d_JT, v.J, v.dJexp, v.d_dJexp, v.d_xGoal,
				v.err, v.d_err, v.ld_x, v.ld_u, v.ld_P, v.ld_p, v.ld_AB, v.ld_H, v.ld_g, v.ld_KT, v.ld_du, v.ld_d, v.ld_A, v.d_I, v.d_Tbody);
			iters[i] = count_iters(&alphaOut[i*(MAX_ITER+1)]);
			if (dump){xout.insert(xout.end(), x0.begin(), x0.end()); uout.insert(uout.end(), u0.begin(), u0.end());}
		}
		t_all1 = now_ms();
		if (dump){dumpf("xGoal", v.xGoal, STATE_SIZE); dumpf("alpha", v.alpha, NA);}
		gpu_free(v);
	}
	else{
		int serial = (hw == 'C');
		int ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A;
		T *alpha, *P, *p, *Pp, *pp, *AB, *H, *g, *KT, *du, *x, *u, *xp, *xp2, *up, *JT, *d, *dp, *ApBK, *Bdu, *dJexp, *xGoal, *I, *Tbody;
		int *err; T **xs, **us, **ds, **JTs;
		if (serial){allocateMemory_CPU<T>(&x, &xp, &xp2, &u, &up, &xGoal, &P, &Pp, &p, &pp, &AB, &H, &g, &KT, &du, &d, &dp, &ApBK, &Bdu,
			&JT, &dJexp, &alpha, &err, &ld_x, &ld_u, &ld_P, &ld_p, &ld_AB, &ld_H, &ld_g, &ld_KT, &ld_du, &ld_d, &ld_A, &I, &Tbody);}
		else{allocateMemory_CPU2<T>(&xs, &x, &xp, &xp2, &us, &u, &up, &xGoal, &P, &Pp, &p, &pp, &AB, &H, &g, &KT, &du, &ds, &d, &dp, &ApBK, &Bdu,
			&JTs, &dJexp, &alpha, &err, &ld_x, &ld_u, &ld_P, &ld_p, &ld_AB, &ld_H, &ld_g, &ld_KT, &ld_du, &ld_d, &ld_A, &I, &Tbody);}
		std::vector<T> x0(ld_x*NT), u0(ld_u*NT);
		t_all0 = now_ms();
		for (int i = 0; i < nseeds; i++){
			loadXU_seeded(x0.data(), u0.data(), xGoal, ld_x, ld_u, seed0+i);
			if (dump){xin.insert(xin.end(), x0.begin(), x0.end()); uin.insert(uin.end(), u0.begin(), u0.end());}
			if (serial){
				runiLQR_CPU<T>(x0.data(), u0.data(), nullptr, nullptr, nullptr, nullptr, xGoal, &Jout[i*(MAX_ITER+1)], &alphaOut[i*(MAX_ITER+1)], 0, 1, 1,
					&tTime[i], simT.data(), swT.data(), bpT.data(), nisT.data(), &initTime[i],
					x, xp, xp2, u, up, P, p, Pp, pp, AB, H, g, KT, du, d, dp, ApBK, Bdu, alpha, JT, dJexp, err,
					ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A, I, Tbody);
			}
			else{
				runiLQR_CPU2<T>(x0.data(), u0.data(), nullptr, nullptr, nullptr, nullptr, xGoal, &Jout[i*(MAX_ITER+1)], &alphaOut[i*(MAX_ITER+1)], 0, 1, 1,
					&tTime[i], simT.data(), swT.data(), bpT.data(), nisT.data(), &initTime[i],
					xs, x, xp, xp2, us, u, up, P, p, Pp, pp, AB, H, g, KT, du, ds, d, dp, ApBK, Bdu, alpha, JTs, dJexp, err,
					ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A, I, Tbody);
			}
			iters[i] = count_iters(&alphaOut[i*(MAX_ITER+1)]);
			if (dump){xout.insert(xout.end(), x0.begin(), x0.end()); uout.insert(uout.end(), u0.begin(), u0.end());}
		}
		t_all1 = now_ms();
		if (dump){dumpf("xGoal", xGoal, STATE_SIZE); dumpf("alpha", alpha, NA);}
		// (memory intentionally not freed: freeMemory_CPU2 double-frees JTs[0], nisInitHelpers.cuh:961,963)
	}
	long total_iters = 0; double total_ms = 0;
	for (int i = 0; i < nseeds; i++){total_iters += iters[i]; total_ms += tTime[i];}
	if (dump){
		int meta[4] = {NT, NA, M_BLOCKS, nseeds}; dumpi("meta", meta, 4);
		dumpf("Jout", Jout.data(), Jout.size()); dumpi("alphaOut", alphaOut.data(), alphaOut.size()); dumpi("iters", iters.data(), nseeds);
		dumpf("x_in", xin.data(), xin.size()); dumpf("u_in", uin.data(), uin.size());
		dumpf("x_out", xout.data(), xout.size()); dumpf("u_out", uout.data(), uout.size());
		dumpd("tTime_ms", tTime.data(), nseeds);
	}
	fprintf(stderr, "REFSUMMARY {\"hw\": \"%c\", \"N\": %d, \"alpha\": %d, \"M\": %d, \"nseeds\": %d, \"total_iters\": %ld, \"sum_solve_ms\": %.3f, \"wall_ms\": %.3f, \"iters_per_sec\": %.3f, \"cores\": %u}\n",
		hw, NT, NA, M_BLOCKS, nseeds, total_iters, total_ms, t_all1 - t_all0, total_iters/(total_ms/1000.0), std::thread::hardware_concurrency());
	return 0;
}

// ---------------------------------------------------------------- trace (GPU): same call sequence as runiLQR_GPU (DDPWrappers.cuh:24-131), with dumps in between
static void snap_traj(GpuVars &v, int iter, const char *ph, bool xs, bool us, bool ds){
	for (int a = 0; a < NA; a++){
		char b[32];
		if (xs){snprintf(b, 32, "x%d", a); dump_dev(nm(b,iter,ph).c_str(), v.h_d_x[a], v.ld_x*NT);}
		if (us){snprintf(b, 32, "u%d", a); dump_dev(nm(b,iter,ph).c_str(), v.h_d_u[a], v.ld_u*NT);}
		if (ds){snprintf(b, 32, "d%d", a); dump_dev(nm(b,iter,ph).c_str(), v.h_d_d[a], v.ld_d*NT);}
	}
}
static void snap_nis(GpuVars &v, int iter, const char *ph){
	dump_dev(nm("AB",iter,ph).c_str(), v.d_AB, v.ld_AB*DIM_AB_c*NT);
	dump_dev(nm("H",iter,ph).c_str(), v.d_H, v.ld_H*DIM_H_c*NT);
	dump_dev(nm("g",iter,ph).c_str(), v.d_g, v.ld_g*NT);
	dump_dev(nm("Pp",iter,ph).c_str(), v.d_Pp, v.ld_P*DIM_P_c*NT);
	dump_dev(nm("pp",iter,ph).c_str(), v.d_pp, v.ld_p*NT);
	dump_dev(nm("xp",iter,ph).c_str(), v.d_xp, v.ld_x*NT);
	dump_dev(nm("xp2",iter,ph).c_str(), v.d_xp2, v.ld_x*NT);
	dump_dev(nm("up",iter,ph).c_str(), v.d_up, v.ld_u*NT);
	dump_dev(nm("dp",iter,ph).c_str(), v.d_dp, v.ld_d*NT);
	snap_traj(v, iter, ph, true, true, true);
}
static int run_trace_gpu(unsigned seed, int maxdump){
	GpuVars v; gpu_alloc(v);
	std::vector<T> x0(v.ld_x*NT), u0(v.ld_u*NT);
	std::vector<T> Jout(MAX_ITER+1, NAN); std::vector<int> alphaOut(MAX_ITER+1, SENT_ALPHA);
	loadXU_seeded(x0.data(), u0.data(), v.xGoal, v.ld_x, v.ld_u, seed);
	dumpf("x_in", x0.data(), x0.size()); dumpf("u_in", u0.data(), u0.size()); dumpf("xGoal", v.xGoal, STATE_SIZE); dumpf("alpha", v.alpha, NA);
	int meta[4] = {NT, NA, M_BLOCKS, 1}; dumpi("meta", meta, 4);
	{
		std::vector<T> h(36*NUM_POS);
		gpuErrchk(cudaMemcpy(h.data(), v.d_I, 36*NUM_POS*sizeof(T), cudaMemcpyDeviceToHost)); dumpf("I", h.data(), h.size());
		gpuErrchk(cudaMemcpy(h.data(), v.d_Tbody, 36*NUM_POS*sizeof(T), cudaMemcpyDeviceToHost)); dumpf("Tbody", h.data(), h.size());
	}
	T prevJ, dJ, z; int iter = 1; T rho = RHO_INIT; T drho = 1.0; *v.alphaIndex = 0; int ignoreFirstDefectFlag = 1;
	dim3 ADimms(DIM_A_r,1); dim3 bpDimms(8,7); dim3 dynDimms(8,7); dim3 FPBlocks(M_BLOCKS_F,NUM_ALPHA); dim3 intDimms(NUM_TIME_STEPS-1,1);
	loadVarsGPU<T>(v.d_x,v.h_d_x,v.d_xp,x0.data(),v.d_u,v.h_d_u,v.d_up,u0.data(),v.d_P,v.d_Pp,nullptr,v.d_p,v.d_pp,nullptr,v.d_KT,nullptr,v.d_du,v.d_dT,v.d_d,v.h_d_d,nullptr,v.d_AB,v.d_err,v.xGoal,v.d_xGoal,v.d_alpha,
		v.d_Tbody,v.d_I,v.d_JT,1,0,v.streams,dynDimms,v.ld_x,v.ld_u,v.ld_P,v.ld_p,v.ld_KT,v.ld_du,v.ld_d,v.ld_AB);
	initAlgGPU<T>(v.d_x,v.h_d_x,v.d_xp,v.d_xp2,v.d_u,v.h_d_u,v.d_up,v.d_d,v.h_d_d,v.d_dp,v.d_dT,v.d_AB,v.d_H,v.d_g,v.d_KT,v.d_du,v.d_JT,&prevJ,v.d_xGoal,v.d_alpha,v.alphaIndex,
		alphaOut.data(),Jout.data(),v.streams,dynDimms,intDimms,0,v.ld_x,v.ld_u,v.ld_d,v.ld_AB,v.ld_H,v.ld_g,v.ld_KT,v.ld_du,v.d_I,v.d_Tbody);
	snap_nis(v, 0, "init");
	{T s[1] = {prevJ}; dumpf("it0.init.prevJ", s, 1);}
	while (1){
		bool dmp = iter <= maxdump;
		T rho_used;
		if (backwardPassGPU<T>(v.d_AB,v.d_P,v.d_p,v.d_Pp,v.d_pp,v.d_H,v.d_g,v.d_KT,v.d_du,v.h_d_d[*v.alphaIndex],v.d_ApBK,v.d_Bdu,
				v.h_d_x[*v.alphaIndex],v.d_xp2,v.d_dJexp,v.err,v.d_err,&rho,&drho,v.streams,bpDimms,
				v.ld_AB,v.ld_P,v.ld_p,v.ld_H,v.ld_g,v.ld_KT,v.ld_du,v.ld_A,v.ld_d,v.ld_x)){break;}
		gpuErrchk(cudaDeviceSynchronize());
		rho_used = rho;
		if (dmp){
			dump_dev(nm("P",iter,"bp").c_str(), v.d_P, v.ld_P*DIM_P_c*NT);	dump_dev(nm("p",iter,"bp").c_str(), v.d_p, v.ld_p*NT);
			dump_dev(nm("KT",iter,"bp").c_str(), v.d_KT, v.ld_KT*DIM_KT_c*NT);	dump_dev(nm("du",iter,"bp").c_str(), v.d_du, v.ld_du*NT);
			dump_dev(nm("ApBK",iter,"bp").c_str(), v.d_ApBK, v.ld_A*DIM_A_c*NT);	dump_dev(nm("Bdu",iter,"bp").c_str(), v.d_Bdu, v.ld_d*NT);
			dump_dev(nm("dJexp",iter,"bp").c_str(), v.d_dJexp, 2*M_BLOCKS_B);
			T s[1] = {rho_used}; dumpf(nm("rho",iter,"bp").c_str(), s, 1);
		}
		forwardSweepKern<T><<<NUM_ALPHA,ADimms,0,v.streams[0]>>>(v.d_x,v.d_ApBK,v.d_Bdu,v.h_d_d[*v.alphaIndex],v.d_xp,v.d_alpha,v.ld_x,v.ld_d,v.ld_A);
		gpuErrchk(cudaPeekAtLastError());	gpuErrchk(cudaDeviceSynchronize());
		if (dmp){snap_traj(v, iter, "sweep", true, false, false);}
		forwardSimGPU<T>(v.d_x,v.d_xp,v.d_xp2,v.d_u,v.d_KT,v.d_du,v.alpha,v.d_alpha,v.d,v.d_d,v.d_dT,v.dJexp,v.d_dJexp,v.J,v.d_JT,v.d_xGoal,&dJ,&z,prevJ,
			v.streams,dynDimms,FPBlocks,v.alphaIndex,&ignoreFirstDefectFlag,v.ld_x,v.ld_u,v.ld_KT,v.ld_du,v.ld_d,v.d_I,v.d_Tbody);
		gpuErrchk(cudaDeviceSynchronize());
		if (dmp){
			snap_traj(v, iter, "sim", true, true, true);
			dumpf(nm("J",iter,"sim").c_str(), v.J, NA); dumpf(nm("dT",iter,"sim").c_str(), v.d, NA); dumpf(nm("dJexpSum",iter,"sim").c_str(), v.dJexp, 2);
			T s[3] = {dJ, z, prevJ}; dumpf(nm("dJ_z_prevJ",iter,"sim").c_str(), s, 3);
			int si[2] = {*v.alphaIndex, ignoreFirstDefectFlag}; dumpi(nm("alphaIndex_ignore",iter,"sim").c_str(), si, 2);
		}
		int it_before = iter;
		if (acceptRejectTrajGPU<T>(v.h_d_x,v.d_xp,v.h_d_u,v.d_up,v.h_d_d,v.d_dp,v.J,&prevJ,&dJ,&rho,&drho,v.alphaIndex,alphaOut.data(),Jout.data(),&iter,v.streams,v.ld_x,v.ld_u,v.ld_d)){break;}
		nextIterationSetupGPU<T>(v.d_x,v.h_d_x,v.d_xp,v.d_u,v.h_d_u,v.d_up,v.d_d,v.h_d_d,v.d_dp,v.d_AB,v.d_H,v.d_g,v.d_P,v.d_p,v.d_Pp,v.d_pp,v.d_xGoal,v.alphaIndex,
			v.streams,dynDimms,intDimms,v.ld_x,v.ld_u,v.ld_d,v.ld_AB,v.ld_H,v.ld_g,v.ld_P,v.ld_p,v.d_I,v.d_Tbody);
		gpuErrchk(cudaDeviceSynchronize());
		if (dmp){
			snap_nis(v, it_before, "nis");
			T s[4] = {rho, drho, prevJ, dJ}; dumpf(nm("rho_drho_prevJ_dJ",it_before,"nis").c_str(), s, 4);
			int si[1] = {*v.alphaIndex}; dumpi(nm("alphaIndex",it_before,"nis").c_str(), si, 1);
		}
	}
	gpuErrchk(cudaDeviceSynchronize());
	storeVarsGPU(v.h_d_x,x0.data(),v.h_d_u,u0.data(),v.alphaIndex,v.streams,v.ld_x,v.ld_u,v.d_d,v.d_dT,v.d,v.ld_d);
	int iters = count_iters(alphaOut.data());
	dumpf("Jout", Jout.data(), Jout.size()); dumpi("alphaOut", alphaOut.data(), alphaOut.size()); dumpi("iters", &iters, 1);
	dumpf("x_out", x0.data(), x0.size()); dumpf("u_out", u0.data(), u0.size());
	fprintf(stderr, "trace G seed %u: iters %d cost %f\n", seed, iters, (double)prevJ);
	gpu_free(v);
	return 0;
}

// ---------------------------------------------------------------- trace (H): the GPU control flow (best-alpha, real defect, tree-ordered J)
// evaluated with the reference's HOST instantiations of the same math routines.  This pins every arithmetic routine of the
// oracle bit-for-bit without a GPU (x86-64 host code: no FMA contraction, glibc sinf/cosf).
static T tree_sum(std::vector<T> v){ // reduceSum order, cudaUtils.h:187-207 (N a power of two)
	for (int s = NT/2; s >= 2; s /= 2){for (int t = 0; t < s; t++){v[t] += v[t+s];}}
	v[0] += v[1]; return v[0];
}
static T tree_max(std::vector<T> v){
	for (int s = NT/2; s >= 2; s /= 2){for (int t = 0; t < s; t++){v[t] = max(v[t], v[t+s]);}}
	v[0] = max(v[0], v[1]); return v[0];
}
static int run_trace_host(unsigned seed, int maxdump){
	int ld_x, ld_u, ld_P, ld_p, ld_AB, ld_H, ld_g, ld_KT, ld_du, ld_d, ld_A;
	T *alpha, *P, *p, *Pp, *pp, *AB, *H, *g, *KT, *du, *x, *u, *xp, *xp2, *up, *d, *dp, *ApBK, *Bdu, *dJexp, *xGoal, *I, *Tbody;
	int *err; T **xs, **us, **ds, **JTs;
	allocateMemory_CPU2<T>(&xs, &x, &xp, &xp2, &us, &u, &up, &xGoal, &P, &Pp, &p, &pp, &AB, &H, &g, &KT, &du, &ds, &d, &dp, &ApBK, &Bdu,
		&JTs, &dJexp, &alpha, &err, &ld_x, &ld_u, &ld_P, &ld_p, &ld_AB, &ld_H, &ld_g, &ld_KT, &ld_du, &ld_d, &ld_A, &I, &Tbody);
	memset(AB, 0, ld_AB*DIM_AB_c*NT*sizeof(T)); memset(H, 0, ld_H*DIM_H_c*NT*sizeof(T)); memset(g, 0, ld_g*NT*sizeof(T));
	memset(ApBK, 0, ld_A*DIM_A_c*NT*sizeof(T)); memset(Bdu, 0, ld_d*NT*sizeof(T)); memset(dp, 0, ld_d*NT*sizeof(T));
	std::vector<T> x0(ld_x*NT), u0(ld_u*NT);
	std::vector<T> Jout(MAX_ITER+1, NAN); std::vector<int> alphaOut(MAX_ITER+1, SENT_ALPHA);
	loadXU_seeded(x0.data(), u0.data(), xGoal, ld_x, ld_u, seed);
	dumpf("x_in", x0.data(), x0.size()); dumpf("u_in", u0.data(), u0.size()); dumpf("xGoal", xGoal, STATE_SIZE); dumpf("alpha", alpha, NA);
	int meta[4] = {NT, NA, M_BLOCKS, 1}; dumpi("meta", meta, 4);
	dumpf("I", I, 36*NUM_POS); dumpf("Tbody", Tbody, 36*NUM_POS);
	size_t nx = ld_x*NT, nu = ld_u*NT, nd = ld_d*NT;
	// loadVarsGPU semantics (nisInitHelpers.cuh:605-634)
	memcpy(xs[0], x0.data(), nx*sizeof(T)); memcpy(us[0], u0.data(), nu*sizeof(T)); memcpy(xp, x0.data(), nx*sizeof(T)); memcpy(up, u0.data(), nu*sizeof(T));
	memset(P, 0, ld_P*DIM_P_c*NT*sizeof(T)); memset(Pp, 0, ld_P*DIM_P_c*NT*sizeof(T)); memset(p, 0, ld_p*NT*sizeof(T)); memset(pp, 0, ld_p*NT*sizeof(T));
	memset(KT, 0, ld_KT*DIM_KT_c*NT*sizeof(T)); for (int a = 0; a < NA; a++){memset(ds[a], 0, nd*sizeof(T));} memset(du, 0, ld_du*NT*sizeof(T));
	T prevJ, dJ, z; int iter = 1; T rho = RHO_INIT; T drho = 1.0; int alphaIndex = 0; int ignore_defect = 1;
	threadDesc_t one; one.tid = 0; one.dim = 1;
	auto nis_math = [&](int a){
		one.reps = NT-1; integratorGradientThreaded<T>(one, xs[a], us[a], AB, ld_x, ld_u, ld_AB, I, Tbody);
		one.reps = NT;   costGradientHessianThreaded<T>(one, xs[a], us[a], g, H, xGoal, ld_x, ld_u, ld_H, ld_g);
	};
	auto bcast = [&](int a){
		for (int b = 0; b < NA; b++){if (b == a){continue;} memcpy(xs[b], xs[a], nx*sizeof(T)); memcpy(us[b], us[a], nu*sizeof(T)); memcpy(ds[b], ds[a], nd*sizeof(T));}
	};
	auto costJ = [&](int a){
		std::vector<T> v(NT); for (int k = 0; k < NT; k++){v[k] = costFunc<T>(&xs[a][k*ld_x], &us[a][k*ld_u], xGoal, k);} return tree_sum(v);
	};
	auto snap = [&](int it, const char *ph){
		dumpf(nm("AB",it,ph).c_str(), AB, ld_AB*DIM_AB_c*NT); dumpf(nm("H",it,ph).c_str(), H, ld_H*DIM_H_c*NT); dumpf(nm("g",it,ph).c_str(), g, ld_g*NT);
		dumpf(nm("Pp",it,ph).c_str(), Pp, ld_P*DIM_P_c*NT); dumpf(nm("pp",it,ph).c_str(), pp, ld_p*NT);
		dumpf(nm("xp",it,ph).c_str(), xp, nx); dumpf(nm("xp2",it,ph).c_str(), xp2, nx); dumpf(nm("up",it,ph).c_str(), up, nu); dumpf(nm("dp",it,ph).c_str(), dp, nd);
		for (int a = 0; a < NA; a++){char b[32];
			snprintf(b,32,"x%d",a); dumpf(nm(b,it,ph).c_str(), xs[a], nx); snprintf(b,32,"u%d",a); dumpf(nm(b,it,ph).c_str(), us[a], nu); snprintf(b,32,"d%d",a); dumpf(nm(b,it,ph).c_str(), ds[a], nd);}
	};
	// initAlgGPU semantics (nisInitHelpers.cuh:363-395)
	alphaOut[0] = -1; nis_math(0); bcast(0);
	memcpy(xp, xs[0], nx*sizeof(T)); memcpy(xp2, xs[0], nx*sizeof(T)); memcpy(up, us[0], nu*sizeof(T)); memcpy(dp, ds[0], nd*sizeof(T));
	prevJ = costJ(0); prevJ += static_cast<T>(2*TOL_COST); Jout[0] = prevJ - static_cast<T>(2*TOL_COST);
	snap(0, "init"); {T s[1] = {prevJ}; dumpf("it0.init.prevJ", s, 1);}
	std::vector<T> J(NA), dT(NA);
	while (1){
		bool dmp = iter <= maxdump;
		// backward pass: M_BLOCKS_B independent blocks (bpHelpers.cuh:422-481, FORCE_PARALLEL reads Pp/pp)
		// a block whose Huu is not positive definite reports err (only the 1-D and 4-D inverses can): rho up, P,p <- Pp,pp, again
		// (backwardPassGPU bpHelpers.cuh:497-511)
		int n_retry = 0;
		while (1){
			int fail = 0;
			for (int b = 0; b < M_BLOCKS_B; b++){
				threadDesc_t desc; desc.tid = b; desc.dim = M_BLOCKS_B; desc.reps = 1;
				backPassThreaded<T>(desc, AB, P, p, Pp, pp, H, g, KT, du, ds[alphaIndex], ApBK, Bdu, xs[alphaIndex], xp2, dJexp, err, ld_AB, ld_P, ld_p, ld_H, ld_g, ld_KT, ld_du, ld_A, ld_d, ld_x, rho);
			}
			for (int b = 0; b < M_BLOCKS_B; b++){fail |= err[b];}
			if (!fail){break;}
			drho = max(drho*static_cast<T>(RHO_FACTOR),static_cast<T>(RHO_FACTOR)); rho = min(rho*drho, static_cast<T>(RHO_MAX));
			memcpy(P, Pp, ld_P*DIM_P_c*NT*sizeof(T)); memcpy(p, pp, ld_p*NT*sizeof(T)); n_retry++;
		}
		if (dmp){int si[1] = {n_retry}; dumpi(nm("retries",iter,"bp").c_str(), si, 1);}
		if (dmp){
			dumpf(nm("P",iter,"bp").c_str(), P, ld_P*DIM_P_c*NT); dumpf(nm("p",iter,"bp").c_str(), p, ld_p*NT);
			dumpf(nm("KT",iter,"bp").c_str(), KT, ld_KT*DIM_KT_c*NT); dumpf(nm("du",iter,"bp").c_str(), du, ld_du*NT);
			dumpf(nm("ApBK",iter,"bp").c_str(), ApBK, ld_A*DIM_A_c*NT); dumpf(nm("Bdu",iter,"bp").c_str(), Bdu, ld_d*NT);
			dumpf(nm("dJexp",iter,"bp").c_str(), dJexp, 2*M_BLOCKS_B); T s[1] = {rho}; dumpf(nm("rho",iter,"bp").c_str(), s, 1);
		}
		// forward sweep for every alpha, d of the current trajectory (DDPWrappers.cuh:73)
		{
			std::vector<T> dcur(ds[alphaIndex], ds[alphaIndex]+nd); // the kernel reads d[alphaIndex] while sweeping x[a]; d is not written by the sweep
			for (int a = 0; a < NA; a++){forwardSweep<T>(xs[a], ApBK, Bdu, dcur.data(), xp, alpha[a], ld_x, ld_d, ld_A);}
		}
		if (dmp){for (int a = 0; a < NA; a++){char b[32]; snprintf(b,32,"x%d",a); dumpf(nm(b,iter,"sweep").c_str(), xs[a], nx);}}
		// forward sim for every (interval, alpha) (fpHelpers.cuh:277-301)
		for (int a = 0; a < NA; a++){for (int b = 0; b < M_BLOCKS_F; b++){
			threadDesc_t desc; desc.tid = b; desc.dim = M_BLOCKS_F; desc.reps = 1;
			forwardSim<T>(desc, xs[a], us[a], KT, du, ds[a], alpha[a], xp, ld_x, ld_u, ld_KT, ld_du, ld_d, I, Tbody);
		}}
		memcpy(xp2, xp, nx*sizeof(T));
		for (int i = 1; i < M_BLOCKS_B; i++){dJexp[0] += dJexp[2*i]; dJexp[1] += dJexp[2*i+1];}
		for (int a = 0; a < NA; a++){
			J[a] = costJ(a);
			std::vector<T> v(NT, 0); for (int k = 0; k < NT; k++){if (onDefectBoundary(k)){for (int c = 0; c < DIM_d_r; c++){v[k] += abs(ds[a][k*ld_d+c]);}}}
			dT[a] = tree_max(v);
		}
		// line search (fpHelpers.cuh:395-408)
		dJ = -1; z = 0; T cdJ = -1; T cz = 0; bool JFlag, zFlag, dFlag;
		for (int i = 0; i < NA; i++){
			cdJ = prevJ - J[i]; JFlag = cdJ >= static_cast<T>(0) && cdJ > dJ;
			cz = cdJ / (alpha[i]*dJexp[0] + static_cast<T>(0.5)*alpha[i]*alpha[i]*dJexp[1]); zFlag = (static_cast<T>(EXP_RED_MIN) < cz && cz < static_cast<T>(EXP_RED_MAX));
			dFlag = ignore_defect ? 1 : dT[i] < static_cast<T>(MAX_DEFECT_SIZE);
			if (JFlag && zFlag && dFlag){if (dT[i] < static_cast<T>(MAX_DEFECT_SIZE)){ignore_defect = 0;} alphaIndex = i; dJ = cdJ; z = cz;}
		}
		if (dmp){
			for (int a = 0; a < NA; a++){char b[32];
				snprintf(b,32,"x%d",a); dumpf(nm(b,iter,"sim").c_str(), xs[a], nx); snprintf(b,32,"u%d",a); dumpf(nm(b,iter,"sim").c_str(), us[a], nu); snprintf(b,32,"d%d",a); dumpf(nm(b,iter,"sim").c_str(), ds[a], nd);}
			dumpf(nm("J",iter,"sim").c_str(), J.data(), NA); dumpf(nm("dT",iter,"sim").c_str(), dT.data(), NA); dumpf(nm("dJexpSum",iter,"sim").c_str(), dJexp, 2);
			T s[3] = {dJ, z, prevJ}; dumpf(nm("dJ_z_prevJ",iter,"sim").c_str(), s, 3);
			int si[2] = {alphaIndex, ignore_defect}; dumpi(nm("alphaIndex_ignore",iter,"sim").c_str(), si, 2);
		}
		// accept / reject (nisInitHelpers.cuh:493-516)
		int it_before = iter; int done = 0;
		if (dJ < static_cast<T>(0)){
			drho = max(drho*static_cast<T>(RHO_FACTOR),static_cast<T>(RHO_FACTOR)); rho = min(rho*drho, static_cast<T>(RHO_MAX));
			alphaIndex = 0; alphaOut[iter] = -1; Jout[iter] = prevJ;
			memcpy(xs[0], xp, nx*sizeof(T)); memcpy(us[0], up, nu*sizeof(T)); memcpy(ds[0], dp, nd*sizeof(T));
		}
		else{
			drho = min(drho/static_cast<T>(RHO_FACTOR), static_cast<T>(1.0/RHO_FACTOR)); rho = max(rho*drho, static_cast<T>(RHO_MIN));
			dJ = dJ/prevJ; prevJ = J[alphaIndex]; alphaOut[iter] = alphaIndex; Jout[iter] = J[alphaIndex];
			if (dJ < static_cast<T>(TOL_COST)){done = 1;}
		}
		if (!done){if (iter == MAX_ITER){done = 1;} else{iter += 1;}}
		if (done){break;}
		// next iteration setup (nisInitHelpers.cuh:259-276)
		nis_math(alphaIndex);
		memcpy(Pp, P, ld_P*DIM_P_c*NT*sizeof(T)); memcpy(pp, p, ld_p*NT*sizeof(T));
		bcast(alphaIndex);
		memcpy(xp, xs[alphaIndex], nx*sizeof(T)); memcpy(up, us[alphaIndex], nu*sizeof(T)); memcpy(dp, ds[alphaIndex], nd*sizeof(T));
		if (dmp){
			snap(it_before, "nis");
			T s[4] = {rho, drho, prevJ, dJ}; dumpf(nm("rho_drho_prevJ_dJ",it_before,"nis").c_str(), s, 4);
			int si[1] = {alphaIndex}; dumpi(nm("alphaIndex",it_before,"nis").c_str(), si, 1);
		}
	}
	int iters = count_iters(alphaOut.data());
	dumpf("Jout", Jout.data(), Jout.size()); dumpi("alphaOut", alphaOut.data(), alphaOut.size()); dumpi("iters", &iters, 1);
	dumpf("x_out", xs[alphaIndex], nx); dumpf("u_out", us[alphaIndex], nu);
	fprintf(stderr, "trace H seed %u: iters %d cost %f\n", seed, iters, (double)prevJ);
	return 0;
}

// ---------------------------------------------------------------- unit: plant functions on random (x,u) (test/testDynGrad.cu:13-19 distributions, gravity on)
template <typename TT>
__global__ void unitDynKern(TT *d_qdd, TT *d_x, TT *d_u, TT *d_I, TT *d_Tbody, int n){
	__shared__ TT s_x[STATE_SIZE]; __shared__ TT s_u[CONTROL_SIZE]; __shared__ TT s_qdd[NUM_POS];
	for (int k = blockIdx.x; k < n; k += gridDim.x){
		int tid = threadIdx.x + threadIdx.y*blockDim.x;
		if (tid < STATE_SIZE){s_x[tid] = d_x[k*STATE_SIZE+tid];} if (tid < CONTROL_SIZE){s_u[tid] = d_u[k*CONTROL_SIZE+tid];}
		__syncthreads();
		dynamics<TT>(s_qdd, s_x, s_u, d_I, d_Tbody);
		__syncthreads();
		if (tid < NUM_POS){d_qdd[k*NUM_POS+tid] = s_qdd[tid];}
		__syncthreads();
	}
}
template <typename TT>
__global__ void unitGradKern(TT *d_AB, TT *d_qdd, TT *d_x, TT *d_u, TT *d_I, TT *d_Tbody, int n){
	__shared__ TT s_x[STATE_SIZE]; __shared__ TT s_u[CONTROL_SIZE]; __shared__ TT s_qdd[NUM_POS]; __shared__ TT s_dqdd[3*NUM_POS*NUM_POS];
	int k = blockIdx.x; if (k >= n){return;}
	int tid = threadIdx.x + threadIdx.y*blockDim.x;
	if (tid < STATE_SIZE){s_x[tid] = d_x[k*STATE_SIZE+tid];} if (tid < CONTROL_SIZE){s_u[tid] = d_u[k*CONTROL_SIZE+tid];}
	__syncthreads();
	_integratorGradient<TT>(&d_AB[k*DIM_AB_r*DIM_AB_c], s_x, s_u, s_qdd, s_dqdd, d_I, d_Tbody, (TT)TIME_STEP, DIM_AB_r);
	__syncthreads();
	if (tid < NUM_POS){d_qdd[k*NUM_POS+tid] = s_qdd[tid];}
}
static int run_unit(char hw, int n, unsigned seed){
	std::default_random_engine eng(seed);
	std::normal_distribution<double> dq(0.0, 2.0), dqd(0.0, 5.0), duu(0.0, 50.0), dsm(0.0, 0.001);
	std::vector<T> x(n*STATE_SIZE), u(n*CONTROL_SIZE), qdd(n*NUM_POS), qdd2(n*NUM_POS), AB(n*DIM_AB_r*DIM_AB_c), J(n), Jf(n), H(n*DIM_H_r*DIM_H_c, 0), g(n*DIM_g_r), Hf(n*DIM_H_r*DIM_H_c, 0), gf(n*DIM_g_r);
	std::vector<T> xg(STATE_SIZE);
	{std::vector<T> xx(STATE_SIZE*NT), uu(CONTROL_SIZE*NT); loadXU_seeded(xx.data(), uu.data(), xg.data(), STATE_SIZE, CONTROL_SIZE, seed);
	 // first half of the samples: near the benchmark's nominal point (small velocities); second half: wide random
	 for (int k = 0; k < n; k++){
		if (k < n/2){for (int i = 0; i < STATE_SIZE; i++){x[k*STATE_SIZE+i] = xx[i] + (i < NUM_POS ? static_cast<T>(0.3*dq(eng)) : static_cast<T>(dqd(eng)*0.2));}
		             for (int i = 0; i < CONTROL_SIZE; i++){u[k*CONTROL_SIZE+i] = uu[i] + static_cast<T>(0.2*duu(eng));}}
		else{for (int i = 0; i < STATE_SIZE; i++){x[k*STATE_SIZE+i] = static_cast<T>(i < NUM_POS ? dq(eng) : dqd(eng));}
		     for (int i = 0; i < CONTROL_SIZE; i++){u[k*CONTROL_SIZE+i] = static_cast<T>(duu(eng));}}
	 }}
	std::vector<T> I(36*NUM_POS), Tbody(36*NUM_POS); initI<T>(I.data()); initT<T>(Tbody.data());
	if (hw == 'G'){
		T *d_x, *d_u, *d_qdd, *d_qdd2, *d_AB, *d_I, *d_Tb;
		gpuErrchk(cudaMalloc(&d_x, x.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_u, u.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_qdd, qdd.size()*sizeof(T)));
		gpuErrchk(cudaMalloc(&d_qdd2, qdd.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_AB, AB.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_I, I.size()*sizeof(T))); gpuErrchk(cudaMalloc(&d_Tb, I.size()*sizeof(T)));
		gpuErrchk(cudaMemcpy(d_x, x.data(), x.size()*sizeof(T), cudaMemcpyHostToDevice)); gpuErrchk(cudaMemcpy(d_u, u.data(), u.size()*sizeof(T), cudaMemcpyHostToDevice));
		gpuErrchk(cudaMemcpy(d_I, I.data(), I.size()*sizeof(T), cudaMemcpyHostToDevice)); gpuErrchk(cudaMemcpy(d_Tb, Tbody.data(), I.size()*sizeof(T), cudaMemcpyHostToDevice));
		unitDynKern<T><<<min(n,1024),dim3(8,7)>>>(d_qdd, d_x, d_u, d_I, d_Tb, n); gpuErrchk(cudaPeekAtLastError());
		unitGradKern<T><<<n,dim3(8,7)>>>(d_AB, d_qdd2, d_x, d_u, d_I, d_Tb, n); gpuErrchk(cudaPeekAtLastError());
		gpuErrchk(cudaDeviceSynchronize());
		gpuErrchk(cudaMemcpy(qdd.data(), d_qdd, qdd.size()*sizeof(T), cudaMemcpyDeviceToHost)); gpuErrchk(cudaMemcpy(qdd2.data(), d_qdd2, qdd.size()*sizeof(T), cudaMemcpyDeviceToHost));
		gpuErrchk(cudaMemcpy(AB.data(), d_AB, AB.size()*sizeof(T), cudaMemcpyDeviceToHost));
	}
	else{
		for (int k = 0; k < n; k++){
			T s_dqdd[3*NUM_POS*NUM_POS];
			dynamics<T>(&qdd[k*NUM_POS], &x[k*STATE_SIZE], &u[k*CONTROL_SIZE], I.data(), Tbody.data());
			_integratorGradient<T>(&AB[k*DIM_AB_r*DIM_AB_c], &x[k*STATE_SIZE], &u[k*CONTROL_SIZE], &qdd2[k*NUM_POS], s_dqdd, I.data(), Tbody.data(), (T)TIME_STEP, DIM_AB_r);
		}
	}
	// cost plug-ins are evaluated on the host in both modes (they are 30 flops; GPU values come from the trace goldens)
	for (int k = 0; k < n; k++){
		J[k] = costFunc<T>(&x[k*STATE_SIZE], &u[k*CONTROL_SIZE], xg.data(), 0); Jf[k] = costFunc<T>(&x[k*STATE_SIZE], &u[k*CONTROL_SIZE], xg.data(), NT-1);
		costGrad<T>(&H[k*DIM_H_r*DIM_H_c], &g[k*DIM_g_r], &x[k*STATE_SIZE], &u[k*CONTROL_SIZE], xg.data(), 0, DIM_H_r);
		costGrad<T>(&Hf[k*DIM_H_r*DIM_H_c], &gf[k*DIM_g_r], &x[k*STATE_SIZE], &u[k*CONTROL_SIZE], xg.data(), NT-1, DIM_H_r);
	}
	int meta[4] = {NT, NA, M_BLOCKS, n}; dumpi("meta", meta, 4);
	T dt[1] = {(T)TIME_STEP}; dumpf("dt", dt, 1);
	dumpf("I", I.data(), I.size()); dumpf("Tbody", Tbody.data(), Tbody.size()); dumpf("xGoal", xg.data(), xg.size());
	dumpf("x", x.data(), x.size()); dumpf("u", u.data(), u.size()); dumpf("qdd", qdd.data(), qdd.size()); dumpf("qdd_from_grad", qdd2.data(), qdd2.size());
	dumpf("AB", AB.data(), AB.size()); dumpf("J_run", J.data(), n); dumpf("J_final", Jf.data(), n);
	dumpf("H_run", H.data(), H.size()); dumpf("g_run", g.data(), g.size()); dumpf("H_final", Hf.data(), Hf.size()); dumpf("g_final", gf.data(), gf.size());
	return 0;
}

// ---------------------------------------------------------------- warm start (loadVarsGPU's clearVarsFlag = 0 / forwardRolloutFlag = 1, nisInitHelpers.cuh:594-652)
// solve 1: cold start (tol1).  Its solution, feedback gains, cost-to-go and defects become the warm-start inputs of three
// second solves (tol2) from a perturbed first knot and a moved goal: (rollout, clear) = (1,0), (0,0), (1,1).
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
	// perturbed start: measured first knot off the plan, goal moved
	std::vector<T> xs = x0, us = u0;
	for (int i = 0; i < NUM_POS; i++){xs[i] += static_cast<T>(0.01*(i+1)/NUM_POS); xs[NUM_POS+i] += static_cast<T>(0.02*(NUM_POS-i)/NUM_POS);}
	v.xGoal[3] += static_cast<T>(0.1); v.xGoal[5] -= static_cast<T>(0.05);
	dumpf("x_in", xs.data(), xs.size()); dumpf("u_in", us.data(), us.size()); dumpf("xGoal", v.xGoal, STATE_SIZE);
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
	double tols[2] = {tol1, tol2}; dumpd("tols", tols, 2);
	gpu_free(v);
	return 0;
}

int main(int argc, char **argv){
	if (argc < 2){fprintf(stderr, "usage: see header of ref_driver.cu\n"); return 2;}
	std::string mode(argv[1]);
	if (mode == "solve" && argc == 7){
		g_tol_cost = atof(argv[5]); g_out = fopen(argv[6], "wb"); if (!g_out){perror("open"); return 1;}
		int rc = run_solve(argv[2][0], (unsigned)atoi(argv[3]), atoi(argv[4]), true); fclose(g_out); return rc;
	}
	if (mode == "time" && argc == 6){
		g_tol_cost = atof(argv[5]); return run_solve(argv[2][0], (unsigned)atoi(argv[3]), atoi(argv[4]), false);
	}
	if (mode == "trace" && argc == 7){
		g_tol_cost = atof(argv[4]); g_out = fopen(argv[6], "wb"); if (!g_out){perror("open"); return 1;}
		int rc = argv[2][0] == 'G' ? run_trace_gpu((unsigned)atoi(argv[3]), atoi(argv[5])) : run_trace_host((unsigned)atoi(argv[3]), atoi(argv[5]));
		fclose(g_out); return rc;
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
