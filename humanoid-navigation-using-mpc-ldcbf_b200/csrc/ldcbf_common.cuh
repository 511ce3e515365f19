// Shared device/host helpers of libldcbf_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "ldcbf_mpc.h"

namespace ldcbf {

// Constants derived from ldcbf_params on the host once per call and passed by value to the kernels
// (kernel parameters live in the constant bank: uniform operands cost no register or LDS traffic).
struct StepConst {
    // LIP, reference HumanoidMpc.py:34-48: per-axis A_d = [[ch, sh/beta],[beta*sh, ch]], B_d = [1-ch, -beta*sh]
    double ch, sh_over_beta, beta_sh;
    double gtil;          // beta*sh/(ch-1): v_{k+1} = -v_k + gtil (p_{k+1} - p_k)
    double inv_one_m_ch;  // 1/(1-ch):       u_k = (p_{k+1} - ch p_k - (sh/beta) v_k)/(1-ch)
    double alpha_over_pi;
    double l_max_x, l_max_y, l_min_x, l_min_y;
    double v_min0, v_min1, v_max0, v_max1;
    double legx_mid, legx_half, legy_mid, legy_half, vlat_mid, vlat_half;   // two-sided rows as mid +- half
    double omega_max, omega_min;
    double foot_offset, stop_objective, sampling_time;
    double eps_active, eps_const_row, eps_infeasible;
    int max_iter;
    bool cold_start;      // LDCBF_FLAG_COLD_START: no initial active-set guess / no warm start
    bool coop_lanes;      // LDCBF_FLAG_COOP_LANES: warp-per-scenario solver for small batches (N <= 3, <= 8 obstacles)
};

inline StepConst make_const(const ldcbf_params& p) {
    StepConst c;
    const double beta = sqrt(p.gravity / p.com_height);   // HumanoidMpc.py:20
    const double ch = cosh(beta * p.delta_t), sh = sinh(beta * p.delta_t);
    c.ch = ch;
    c.sh_over_beta = sh / beta;
    c.beta_sh = beta * sh;
    c.gtil = beta * sh / (ch - 1.0);
    c.inv_one_m_ch = 1.0 / (1.0 - ch);
    c.alpha_over_pi = p.alpha / 3.141592653589793;
    c.l_max_x = p.l_max_x; c.l_max_y = p.l_max_y; c.l_min_x = p.l_min_x; c.l_min_y = p.l_min_y;
    c.v_min0 = p.v_min[0]; c.v_min1 = p.v_min[1]; c.v_max0 = p.v_max[0]; c.v_max1 = p.v_max[1];
    c.legx_mid = 0.5 * (p.l_max_x + p.l_min_x); c.legx_half = 0.5 * (p.l_max_x - p.l_min_x);
    c.legy_mid = 0.5 * (p.l_max_y + p.l_min_y); c.legy_half = 0.5 * (p.l_max_y - p.l_min_y);
    c.vlat_mid = 0.5 * (p.v_max[1] + p.v_min[1]); c.vlat_half = 0.5 * (p.v_max[1] - p.v_min[1]);
    c.omega_max = p.omega_max; c.omega_min = p.omega_min;
    c.foot_offset = p.foot_offset; c.stop_objective = p.stop_objective; c.sampling_time = p.sampling_time;
    c.eps_active = p.eps_active; c.eps_const_row = p.eps_const_row; c.eps_infeasible = p.eps_infeasible;
    c.max_iter = p.max_iter;
    c.cold_start = (p.flags & LDCBF_FLAG_COLD_START) != 0;
    c.coop_lanes = (p.flags & LDCBF_FLAG_COOP_LANES) != 0;
    return c;
}

void set_last_error(cudaError_t e);
// Library-owned stream-ordered memory pool of the current device (mpc_step.cu): record hand-off of the large-batch
// solve, half-plane scratch of rollouts with more than LDCBF_MAX_OBSTACLES obstacles.  nullptr on failure.
cudaMemPool_t workspace_pool();
inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_error(e); return LDCBF_E_LAUNCH; }
    return LDCBF_OK;
}

// K1 launcher with an arbitrary position layout: p_x = pos[b*stride], p_y = pos[b*stride + y_off]
// ([B,2] positions: (2,1); [B,4] states: (4,2); [B,5] states with heading: (5,2)).
int launch_halfplanes(int B, int max_obs, int max_verts, const double* pos, int pos_stride, int y_off,
                      const double* verts, const int32_t* nverts, const int32_t* nobs, double* c_eta,
                      bool fast_geometry, void* cuda_stream);

}  // namespace ldcbf
