/*
 * ref_hwc.cu -- TEST INFRASTRUCTURE, not product code.
 *
 * The consumer side of the trajectory hand-off: the UNMODIFIED reference's getHardwareControls (DDPHelpers/MPCHelpers.cuh:817-858),
 * which turns the published plan (x, u, KT) and a measured state into the joint command -- zero-order hold on u and KT, first-order
 * hold on x, optional exponential smoothing.  Host code only.  Writes the inputs and, for a sweep of measurement times (inside the
 * plan, on knot boundaries, before its start and past its end), the outputs.
 *
 *   ref_hwc <seed> <out.bin>
 */
#define PLANT 4
#define EE_COST 0
#define MPC_MODE 1
#define USE_WAFR_URDF 1
#include "config.cuh"
#include <random>
#include <vector>
#include <cstdio>

typedef algType T;
#define NT NUM_TIME_STEPS
static FILE *g_out = nullptr;
static void dumpraw(const char *name, const char *dtype, const void *p, size_t n, size_t sz){ fprintf(g_out, "%s %s %zu\n", name, dtype, n); fwrite(p, sz, n, g_out); }

int main(int argc, char **argv){
	if (argc != 3){fprintf(stderr, "usage: ref_hwc <seed> <out.bin>\n"); return 2;}
	g_out = fopen(argv[2], "wb"); if (!g_out){perror("open"); return 1;}
	std::default_random_engine eng((unsigned)atoi(argv[1])); std::normal_distribution<double> nd(0.0, 1.0);
	const int ld_x = STATE_SIZE, ld_u = CONTROL_SIZE, ld_KT = DIM_KT_r;
	std::vector<T> x(ld_x*NT), u(ld_u*NT), KT(ld_KT*DIM_KT_c*NT);
	for (auto &v : x){v = static_cast<T>(nd(eng));} for (auto &v : u){v = static_cast<T>(20.0*nd(eng));} for (auto &v : KT){v = static_cast<T>(5.0*nd(eng));}
	const double t0 = 1234567.0, step_us = TIME_STEP_LENGTH_IN_us;
	std::vector<double> times, q_out, u_out, qa, qda, alphas; std::vector<int> err;
	double q_prev[NUM_POS] = {0}, u_prev[CONTROL_SIZE] = {0};
	const double fracs[] = {-0.5, 0.0, 0.25, 1.0, 1.5, 7.999, 13.37, NT - 3.0, NT - 2.5, NT - 2.0, NT - 1.0, NT + 4.0};
	for (int pass = 0; pass < 2; pass++){           // pass 1: with smoothing (alpha 0.3, carried q_prev / u_prev)
		for (double f : fracs){
			double qA[NUM_POS], qdA[NUM_POS], qo[NUM_POS] = {0}, uo[CONTROL_SIZE] = {0};
			for (int i = 0; i < NUM_POS; i++){qA[i] = nd(eng); qdA[i] = nd(eng);}
			const double tA = t0 + f*step_us, alpha = pass ? 0.3 : 0.0;
			int e = getHardwareControls<T>(qo, uo, x.data(), u.data(), KT.data(), t0, qA, qdA, tA, ld_x, ld_u, ld_KT, pass ? q_prev : nullptr, pass ? u_prev : nullptr, alpha);
			times.push_back(tA); err.push_back(e); alphas.push_back(alpha);
			qa.insert(qa.end(), qA, qA + NUM_POS); qda.insert(qda.end(), qdA, qdA + NUM_POS);
			q_out.insert(q_out.end(), qo, qo + NUM_POS); u_out.insert(u_out.end(), uo, uo + CONTROL_SIZE);
		}
	}
	int meta[2] = {NT, (int)times.size()}; double consts[2] = {t0, (double)TIME_STEP};
	dumpraw("meta", "i32", meta, 2, 4); dumpraw("consts", "f64", consts, 2, 8);
	dumpraw("x", "f32", x.data(), x.size(), 4); dumpraw("u", "f32", u.data(), u.size(), 4); dumpraw("KT", "f32", KT.data(), KT.size(), 4);
	dumpraw("tActual", "f64", times.data(), times.size(), 8); dumpraw("alpha", "f64", alphas.data(), alphas.size(), 8);
	dumpraw("qActual", "f64", qa.data(), qa.size(), 8); dumpraw("qdActual", "f64", qda.data(), qda.size(), 8);
	dumpraw("q_out", "f64", q_out.data(), q_out.size(), 8); dumpraw("u_out", "f64", u_out.data(), u_out.size(), 8); dumpraw("err", "i32", err.data(), err.size(), 4);
	fclose(g_out);
	return 0;
}
