/*
 * ref_stages.cu -- TEST INFRASTRUCTURE.  Runs the reference's dynamicsGradient pipeline (plants/dynamics_arm.cuh:2165-2289)
 * on the GPU by calling the reference's own stage functions in the reference's order, and copies the intermediate
 * arrays out after each stage, so that contraction (FMA) differences of a restatement can be localised.
 * Inputs: the (x,u) samples of a `unit` dump.   usage: ref_stages <unit.bin> <out.bin>
 */
#define EE_COST 0
#define USE_WAFR_URDF 1
#define _Q1 0.1
#define _Q2 0.001
#define _R  0.0001
#define _QF1 1000.0
#define _QF2 1000.0
#define TOL_COST 0.0
#include "config.cuh"
#include <vector>
#include <string>
#include <cstring>
typedef algType T;
#define NP NUM_POS
// per-sample output layout (floats)
#define O_dTA 0
#define O_dJ (O_dTA + 36*NP*NP)
#define O_dIw (O_dJ + 6*NP*NP)
#define O_Iw (O_dIw + 36*NP*NP)
#define O_Icrbs (O_Iw + 36*NP)
#define O_Minv (O_Icrbs + 36*NP)
#define O_qdd (O_Minv + NP*NP)
#define O_dM (O_qdd + NP)
#define O_dqddM (O_dM + NP*NP*NP)
#define O_dTwist (O_dqddM + NP*NP)
#define O_dJdotV (O_dTwist + 12*NP*NP)
#define O_dWb (O_dJdotV + 12*NP*NP)
#define O_dTau (O_dWb + 12*NP*NP)
#define O_dqdd (O_dTau + 2*NP*NP)
#define O_TOTAL (O_dqdd + 3*NP*NP)

__device__ void cpy(T *dst, const T *src, int n){ for (int i = threadIdx.x + threadIdx.y*blockDim.x; i < n; i += blockDim.x*blockDim.y){ dst[i] = src[i]; } __syncthreads(); }

__global__ void stagesKern(T *out, T *d_x, T *d_u, T *d_I, T *d_Tbody, int n){
	__shared__ T s_x[STATE_SIZE]; __shared__ T s_u[CONTROL_SIZE]; __shared__ T s_qdd[NP]; __shared__ T s_dqdd[3*NP*NP];
	__shared__ T s_I[36*NP]; __shared__ T s_Icrbs[36*NP]; __shared__ T s_TA[42*NP]; __shared__ T s_dTA[36*NP*NP];
	__shared__ T s_J[6*NP]; __shared__ T s_dJ[6*NP*NP]; __shared__ T s_JdotV[6*NP]; __shared__ T s_twist[6*NP];
	__shared__ T s_W[6*NP]; __shared__ T s_F[6*NP]; __shared__ T s_temp[36*NP]; __shared__ T s_temp2[36*NP]; __shared__ T s_temp3[36*NP*NP];
	int k = blockIdx.x; if (k >= n){return;}
	T *o = out + (size_t)k*O_TOTAL;
	int tid = threadIdx.x + threadIdx.y*blockDim.x;
	if (tid < STATE_SIZE){s_x[tid] = d_x[k*STATE_SIZE+tid];} if (tid < CONTROL_SIZE){s_u[tid] = d_u[k*CONTROL_SIZE+tid];}
	__syncthreads();
	// --- the reference's own sequence (dynamics_arm.cuh:2197-2288), stage dumps in between
	T *s_Tb = s_temp; T *s_dTb = s_temp2;
	load_Tb(s_x,s_Tb,d_Tbody,s_W,s_F,s_dTb);
	load_I(s_I,d_I);
	__syncthreads();
	T *s_T = s_Icrbs;
	compute_T_TA_J(s_Tb,s_T,s_TA,s_J);
	__syncthreads();
	T *s_dT = s_temp3; T *s_dTp = &s_temp2[16*NP];
	compute_dT_dTA_dJ(s_Tb,s_dTb,s_T,s_dT,s_dTp,s_TA,s_dTA,s_dJ);
	__syncthreads();
	cpy(o + O_dTA, s_dTA, 36*NP*NP); cpy(o + O_dJ, s_dJ, 6*NP*NP);
	compute_Iw_Icrbs_twist(s_I,s_Icrbs,s_twist,s_TA,s_J,s_x,s_temp,s_dTA,s_temp2);
	T *s_dIw = s_dTA;
	__syncthreads();
	cpy(o + O_dIw, s_dIw, 36*NP*NP); cpy(o + O_Iw, s_I, 36*NP); cpy(o + O_Icrbs, s_Icrbs, 36*NP);
	compute_JdotV(s_JdotV,s_twist,s_J,s_x,s_temp);
	__syncthreads();
	T *s_M = s_temp2; T *s_Tau = &s_temp2[2*NP*NP];
	compute_M_Tau(s_M, s_Tau, s_W, s_JdotV, s_F, s_Icrbs, s_twist, s_J, s_I, s_x, s_u, s_temp, s_temp2, s_TA);
	__syncthreads();
	invertMatrix<T,NP,1>(s_M,s_F);
	T *s_Minv = &s_temp2[NP*NP];
	__syncthreads();
	compute_qdd(s_qdd,s_Minv,s_Tau);
	__syncthreads();
	cpy(o + O_Minv, s_Minv, NP*NP); cpy(o + O_qdd, s_qdd, NP);
	T *s_dM = s_temp3;
	compute_dM(s_dM,s_Icrbs,s_dIw,s_J,s_dJ,s_F,s_TA);
	__syncthreads();
	cpy(o + O_dM, s_dM, NP*NP*NP);
	compute_dqdd_dM(s_dqdd,s_dM,s_Minv,s_qdd,s_temp);
	__syncthreads();
	cpy(o + O_dqddM, s_dqdd, NP*NP);
	T *s_dTwist = s_temp3; T *s_dJdotV = &s_temp3[6*(2*NP)*NP]; T *s_dWb = &s_temp3[6*(4*NP)*NP];
	compute_dtwist(s_dTwist,s_J,s_dJ,s_x);
	__syncthreads();
	compute_dJdotV(s_dJdotV,s_twist,s_dTwist,s_J,s_dJ,s_x,s_temp,s_TA);
	__syncthreads();
	compute_dWb(s_dWb,s_JdotV,s_dJdotV,s_twist,s_dTwist,s_I,s_dIw,s_temp,s_TA,s_F);
	__syncthreads();
	cpy(o + O_dTwist, s_dTwist, 12*NP*NP); cpy(o + O_dJdotV, s_dJdotV, 12*NP*NP); cpy(o + O_dWb, s_dWb, 12*NP*NP);
	T *s_dTau = s_temp;
	compute_dTau(s_dTau,s_dWb,s_W,s_J,s_dJ);
	__syncthreads();
	cpy(o + O_dTau, s_dTau, 2*NP*NP);
	finish_dqdd(s_dqdd,s_dTau,s_Minv);
	__syncthreads();
	cpy(o + O_dqdd, s_dqdd, 3*NP*NP);
}

static bool read_arr(const std::vector<char> &buf, const char *want, std::vector<float> &out){
	size_t pos = 0;
	while (pos < buf.size()){
		size_t nl = pos; while (buf[nl] != '\n'){nl++;}
		std::string line(&buf[pos], nl-pos); char name[64], dt[8]; size_t cnt; sscanf(line.c_str(), "%63s %7s %zu", name, dt, &cnt);
		size_t esz = (strcmp(dt,"f64") == 0) ? 8 : 4;
		if (strcmp(name, want) == 0){out.resize(cnt); memcpy(out.data(), &buf[nl+1], cnt*4); return true;}
		pos = nl + 1 + cnt*esz;
	}
	return false;
}
int main(int argc, char **argv){
	if (argc != 3){fprintf(stderr,"usage: ref_stages unit.bin out.bin\n"); return 2;}
	FILE *f = fopen(argv[1],"rb"); fseek(f,0,SEEK_END); long sz = ftell(f); fseek(f,0,SEEK_SET); std::vector<char> buf(sz); if (fread(buf.data(),1,sz,f) != (size_t)sz){return 1;} fclose(f);
	std::vector<float> x, u; read_arr(buf,"x",x); read_arr(buf,"u",u); int n = x.size()/STATE_SIZE;
	std::vector<T> I(36*NP), Tb(36*NP); initI<T>(I.data()); initT<T>(Tb.data());
	T *d_x,*d_u,*d_I,*d_Tb,*d_o;
	gpuErrchk(cudaMalloc(&d_x,x.size()*4)); gpuErrchk(cudaMalloc(&d_u,u.size()*4)); gpuErrchk(cudaMalloc(&d_I,36*NP*4)); gpuErrchk(cudaMalloc(&d_Tb,36*NP*4)); gpuErrchk(cudaMalloc(&d_o,(size_t)n*O_TOTAL*4));
	gpuErrchk(cudaMemcpy(d_x,x.data(),x.size()*4,cudaMemcpyHostToDevice)); gpuErrchk(cudaMemcpy(d_u,u.data(),u.size()*4,cudaMemcpyHostToDevice));
	gpuErrchk(cudaMemcpy(d_I,I.data(),36*NP*4,cudaMemcpyHostToDevice)); gpuErrchk(cudaMemcpy(d_Tb,Tb.data(),36*NP*4,cudaMemcpyHostToDevice));
	stagesKern<<<n,dim3(8,7)>>>(d_o,d_x,d_u,d_I,d_Tb,n); gpuErrchk(cudaPeekAtLastError()); gpuErrchk(cudaDeviceSynchronize());
	std::vector<T> o((size_t)n*O_TOTAL); gpuErrchk(cudaMemcpy(o.data(),d_o,o.size()*4,cudaMemcpyDeviceToHost));
	FILE *g = fopen(argv[2],"wb");
	int offs[16] = {O_dTA,O_dJ,O_dIw,O_Iw,O_Icrbs,O_Minv,O_qdd,O_dM,O_dqddM,O_dTwist,O_dJdotV,O_dWb,O_dTau,O_dqdd,O_TOTAL,n};
	fprintf(g,"offsets i32 16\n"); fwrite(offs,4,16,g);
	fprintf(g,"stages f32 %zu\n", o.size()); fwrite(o.data(),4,o.size(),g); fclose(g);
	return 0;
}
