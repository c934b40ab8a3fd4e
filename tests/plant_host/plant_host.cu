// plant_host.cu -- TEST INFRASTRUCTURE.  Host instantiation of a plug-in plant header (parallel-ddp_b200/csrc/plants/*.cuh) and of
// plugin/integrators.cuh behind a C interface, so that the CPU test-suite can compare the plant code the CUDA kernels are built
// from with the reference's own host build (tests/golden/p*_unit_H.npz) without a GPU.  Built by tests/test_plants_cpu.py:
//   nvcc -O3 -std=c++17 -Xcompiler -fPIC -shared -DPDDP_PLANT_HEADER='"plants/pendulum.cuh"' -I parallel-ddp_b200/csrc plant_host.cu
#include "plugin/pddp_plugin.cuh"
#include PDDP_PLANT_HEADER
#include "plugin/integrators.cuh"
typedef float T;
extern "C" {
void ph_dims(int *d){ d[0] = NUM_POS; d[1] = STATE_SIZE; d[2] = CONTROL_SIZE; }
void ph_dynamics(const T *x, const T *u, T *qdd){ dynamics<T>(qdd, (T*)x, (T*)u, nullptr, nullptr); }
void ph_integrator(int integ, const T *x, const T *u, T dt, T *xn){
    T qdd[NUM_POS];
    if (integ == 1){ _integrator<T,1>(xn, (T*)x, (T*)u, qdd, nullptr, nullptr, dt); }
    else if (integ == 2){ _integrator<T,2>(xn, (T*)x, (T*)u, qdd, nullptr, nullptr, dt); }
    else { _integrator<T,3>(xn, (T*)x, (T*)u, qdd, nullptr, nullptr, dt); }
}
void ph_integrator_gradient(int integ, const T *x, const T *u, T dt, T *AB, T *qdd){
    T dqdd[NUM_POS*(STATE_SIZE+CONTROL_SIZE)];
    if (integ == 1){ _integratorGradient<T,1>(AB, (T*)x, (T*)u, qdd, dqdd, nullptr, nullptr, dt, STATE_SIZE); }
    else if (integ == 2){ _integratorGradient<T,2>(AB, (T*)x, (T*)u, qdd, dqdd, nullptr, nullptr, dt, STATE_SIZE); }
    else { _integratorGradient<T,3>(AB, (T*)x, (T*)u, qdd, dqdd, nullptr, nullptr, dt, STATE_SIZE); }
}
T ph_cost(int N, const T *x, const T *u, const T *xg, int k, const T *w){ pddp_plugin::host_num_time_steps = N; return costFunc<T>((T*)x, (T*)u, (T*)xg, k, w[0], w[1], w[2], w[3], w[4]); }
void ph_cost_grad(int N, T *H, T *g, const T *x, const T *u, const T *xg, int k, const T *w){
    pddp_plugin::host_num_time_steps = N; costGrad<T>(H, g, (T*)x, (T*)u, (T*)xg, k, STATE_SIZE+CONTROL_SIZE, w[0], w[1], w[2], w[3], w[4]);
}
}
