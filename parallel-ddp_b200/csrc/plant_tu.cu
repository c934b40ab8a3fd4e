// plant_tu.cu -- one plant translation unit: the plant header (a reference-style plug-in, plugin/pddp_plugin.cuh) + the solver
// kernels that call it (plugin/plugin_kernels.cuh) + the operations table libpddp dispatches through (include/pddp_plant.h).
//
//   -DPDDP_PLANT_ID=<n>  -DPDDP_PLANT_HEADER='"plants/pendulum.cuh"'  [-DPDDP_PLANT_NAME='"pendulum"']
//   -DPDDP_PLANT_BUILTIN : export pddp_plant_entry_<n> (linked into libpddp.so) instead of pddp_plant_entry (stand-alone library)
//
// Compiled WITHOUT -fmad=false: plant code is ordinary CUDA C++ and gets the compiler's default contraction, exactly as the
// reference's plant headers do inside its kernels; the solver arithmetic around it is written with explicit intrinsics.
#ifndef PDDP_PLANT_ID
#error "define PDDP_PLANT_ID and PDDP_PLANT_HEADER"
#endif
#include "plugin/pddp_plugin.cuh"
#include "dev_state.cuh"
#include "../../include/pddp_plant.h"
#define PDDP_CAT2(a, b) a##b
#define PDDP_CAT(a, b) PDDP_CAT2(a, b)
#define PDDP_PLANT_NS PDDP_CAT(pddp_plant_, PDDP_PLANT_ID)

// the plant, the integrators on top of it and the kernels around both live in the plant's own namespace: plant headers all define
// the same names (dynamics, costFunc, ...), and several plants are linked into one library
namespace PDDP_PLANT_NS {
using namespace pddp;
#include PDDP_PLANT_HEADER
#if !defined(NUM_POS) || !defined(STATE_SIZE) || !defined(CONTROL_SIZE)
#error "the plant header must define NUM_POS, STATE_SIZE and CONTROL_SIZE"
#endif
#include "plugin/integrators.cuh"
#include "plugin/plugin_kernels.cuh"
}

#ifndef PDDP_PLANT_NAME
#define PDDP_PLANT_NAME PDDP_PLANT_HEADER
#endif
#ifdef PDDP_PLANT_BUILTIN
#define PDDP_ENTRY PDDP_CAT(pddp_plant_entry_, PDDP_PLANT_ID)
#else
#define PDDP_ENTRY pddp_plant_entry
#endif

extern "C" const pddp_plant_ops *PDDP_ENTRY(void){
    static const pddp_plant_ops ops = {
        PDDP_PLANT_ABI, PDDP_PLANT_ID, NUM_POS, STATE_SIZE, CONTROL_SIZE, sizeof(pddp::DevState), sizeof(pddp::MpcState), PDDP_PLANT_NAME,
        PDDP_PLANT_NS::plug_init_model, PDDP_PLANT_NS::plug_prepare, PDDP_PLANT_NS::plug_launch_bp, PDDP_PLANT_NS::plug_launch_sweep, PDDP_PLANT_NS::plug_launch_sim,
        PDDP_PLANT_NS::plug_launch_init_cost, PDDP_PLANT_NS::plug_launch_nis, PDDP_PLANT_NS::plug_launch_mpc_load,
        PDDP_PLANT_NS::plug_unit_dynamics, PDDP_PLANT_NS::plug_unit_gradient, PDDP_PLANT_NS::plug_unit_cost };
    return &ops;
}
