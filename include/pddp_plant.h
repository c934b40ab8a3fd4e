/*
 * pddp_plant.h -- binary interface between libpddp.so and a plant translation unit (plant plug-in).
 *
 * The reference has no such interface: a plant is a pair of headers (plants/dynamics_*.cuh, plants/cost_*.cuh) that config.cuh
 * includes by PLANT number (config.cuh:239-262), and the solver templates are recompiled around it.  Here a plant author writes the
 * same kind of header -- same function names, argument order and calling convention, parallel-ddp_b200/csrc/plugin/pddp_plugin.cuh --
 * and compiles parallel-ddp_b200/csrc/plant_tu.cu around it:
 *
 *     nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -shared \
 *          -DPDDP_PLANT_ID=<id> -DPDDP_PLANT_HEADER='"my_plant.cuh"' -DPDDP_PLANT_NAME='"my plant"' \
 *          -I<dir of my_plant.cuh> parallel-ddp_b200/csrc/plant_tu.cu -o libmyplant.so
 *
 * The result exports `pddp_plant_entry`, which returns the table below (the plant's instances of the solver kernels behind plain
 * function pointers).  pddp_load_plant_library(path) or pddp_register_plant(ops) makes the plant available to pddp_create under
 * pddp_config.plant = <id>.  Pendulum, cart-pole and quadrotor (ids 1-3) are built into libpddp.so through exactly this route.
 * Every launcher returns 0 or a cudaError_t value; `state` is the library's device-state struct (checked through state_size).
 */
#ifndef PDDP_PLANT_H
#define PDDP_PLANT_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define PDDP_PLANT_ABI 2

typedef struct pddp_plant_ops {
    int abi;                     /* PDDP_PLANT_ABI */
    int plant_id;                /* pddp_config.plant value this table serves */
    int num_pos, state_size, control_size;      /* NUM_POS, STATE_SIZE, CONTROL_SIZE of the plant header */
    size_t state_size_bytes, mpc_size_bytes;    /* sizeof the device-state structs the table was compiled against */
    const char *name;
    void (*init_model)(float *I, float *Tbody);                                   /* initI / initT on the host, 36*num_pos floats each */
    int (*prepare)(int max_M);                                                    /* once per handle: kernel attributes */
    int (*launch_bp)(const void *state, void *stream, int b0, int nb);            /* backPassKern + rho retry, problems [b0, b0+nb) */
    int (*launch_sweep)(const void *state, void *stream, int b0, int nb, int num_sms);
    int (*launch_sim)(const void *state, void *stream, int b0, int nb, int n_cand);
    int (*launch_init_cost)(const void *state, void *stream);
    int (*launch_nis)(const void *state, void *stream, int mode, int b0, int nb);
    int (*launch_mpc_load)(const void *state, const void *mpc, void *stream);
    int (*unit_dynamics)(const void *state, void *stream, const float *d_x, const float *d_u, int nsamp, float *d_qdd);
    int (*unit_gradient)(const void *state, void *stream, const float *d_x, const float *d_u, int nsamp, float *d_AB, float *d_qdd, float *d_xnext);
    int (*unit_cost)(const void *state, void *stream, const float *d_x, const float *d_u, const float *d_xg, const int *d_knot, int nsamp,
                     float *d_J, float *d_H, float *d_g);
} pddp_plant_ops;

/* symbol a plant library exports */
typedef const pddp_plant_ops *(*pddp_plant_entry_fn)(void);

#ifdef __cplusplus
}
#endif
#endif
