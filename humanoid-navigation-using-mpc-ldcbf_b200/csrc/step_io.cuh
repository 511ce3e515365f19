// Argument block of one batched MPC step, shared by the short-horizon (mpc_step.cu) and long-horizon (mpc_long.cu) kernels.
#pragma once
#include "ldcbf_common.cuh"

namespace ldcbf {

struct StepIO {
    const double* x0; const double* theta0; const double* goal; const int8_t* foot;
    const double* c_eta; const int32_t* nobs; const double* delta; const double* limits;
    double* U; double* X; double* theta; double* omega; double* obj; int32_t* status; int32_t* iters;
    // loop-shaped variant (ldcbf_mpc_step_packed_f64): state rows [B,6] in (x0/theta0/foot unused), next rows [B,10] out
    const double* state6; double* next10;
};

__device__ __forceinline__ void load_state(const StepIO& io, int b, double4& x, double& th) {
    if (io.state6) {
        const double2* s = reinterpret_cast<const double2*>(io.state6) + 3 * (size_t)b;
        const double2 a = s[0], c = s[1], e = s[2];
        x = make_double4(a.x, a.y, c.x, c.y);
        th = e.x;
    } else {
        x = reinterpret_cast<const double4*>(io.x0)[b];          // (p_x, v_x, p_y, v_y)
        th = io.theta0[b];
    }
}

// Long horizons (N > LDCBF_MAX_HORIZON, one thread block per scenario), mpc_long.cu.
int launch_long_horizon(const StepConst& C, int B, int N, int max_obs, const StepIO& io, cudaStream_t st);

}  // namespace ldcbf
