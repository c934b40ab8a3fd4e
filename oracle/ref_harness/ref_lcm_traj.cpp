/*
 * ref_lcm_traj.cpp -- TEST INFRASTRUCTURE.  Builds the trajectory hand-off message exactly as the reference's MPC loop does
 * (DDPHelpers/LCMHelpers.cuh:245-252: the *_size fields carry BYTE counts, the vectors are resized to that many floats and
 * the data sits in their first quarter) with the reference's own generated type lcmtypes/drake/lcmt_trajectory_f.hpp, and
 * writes the encoded bytes.   usage: ref_lcm_traj <N> <with_feedback 0|1> <out.bin>
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "lcmtypes/drake/lcmt_trajectory_f.hpp"
int main(int argc, char **argv){
    if (argc != 4){ return 2; }
    const int N = atoi(argv[1]), fb = atoi(argv[2]); const int ld_x = 14, ld_u = 7, ld_KT = 14, DIM_KT_c = 7;
    std::vector<float> x(ld_x*N), u(ld_u*N), KT(ld_KT*DIM_KT_c*N);
    for (size_t i = 0; i < x.size(); i++){ x[i] = 0.001f*(float)i - 0.5f; }
    for (size_t i = 0; i < u.size(); i++){ u[i] = 0.25f*(float)(i % 17) - 1.f; }
    for (size_t i = 0; i < KT.size(); i++){ KT[i] = 1.0f/(float)(i + 3); }
    drake::lcmt_trajectory_f dataOut; dataOut.utime = 1234567890123LL; int stepsSize = N*sizeof(float);
    int uSize = ld_u*stepsSize; dataOut.u_size = uSize; dataOut.u.resize(dataOut.u_size); memcpy(&(dataOut.u[0]), &u[0], uSize);
    if (fb){
        int xSize = ld_x*stepsSize; dataOut.x_size = xSize; dataOut.x.resize(dataOut.x_size); memcpy(&(dataOut.x[0]), &x[0], xSize);
        int KTSize = ld_KT*DIM_KT_c*stepsSize; dataOut.KT_size = KTSize; dataOut.KT.resize(dataOut.KT_size); memcpy(&(dataOut.KT[0]), &KT[0], KTSize);
    }
    else{ dataOut.x_size = 0; dataOut.KT_size = 0; dataOut.x.resize(dataOut.x_size); dataOut.KT.resize(dataOut.KT_size); }
    std::vector<unsigned char> buf(dataOut.getEncodedSize());
    int nb = dataOut.encode(buf.data(), 0, (int)buf.size());
    FILE *f = fopen(argv[3], "wb"); fwrite(buf.data(), 1, nb, f); fclose(f);
    printf("%d bytes\n", nb);
    return 0;
}
