// bam_sediment_fast.cuh -- per-INPUT invariants of Sediment_Apply, shared by the prediction fast path (bam_predict.cu) and
// the log-posterior fast path (bam_logpost.cu).
//
// When the pseudo-parameters d, s, nu, g are inputs or fixed values and the critical-shear-stress coefficients are
// fixed (the shipped BaM_Sediment* option files), everything in Sediment_Apply except three numbers per parameter set
// depends on the input row only:
//     MPM      dim t1 (tau-css)^t2                  Camenen  dim t1 tau^t2 exp(-t3 css/tau)
//     Nielsen  dim t1 tau^t2 (tau-css)              Smart    dim t1 S0^t2 (s-1)^(1/6) d^(1/6) g^(-1/2) tau^t3 (tau-css)
// all have the form  t1 * B(input) * exp(t2 * L1(input) + t3 * L2(input)).  `sed_obs_prep` evaluates (B, L1, L2) and the row
// status (infeasible inputs / tau <= css -> 0) once per input row with the reference's libm-class functions
// (SedimentTransport_model.f90:155-187,290-304).
#pragma once
#include "bam_device.cuh"

struct SedObs { double B, L1, L2, st; };  // st: 0 normal, 1 -> value 0 (tau <= css), 2 infeasible row

template <int NX>
static __global__ void sed_obs_prep(DevProblem p, int nobsPred, int t0, int nT, double* const* __restrict__ xspag,
                             SedObs* __restrict__ obs) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nT) return;
    double in[NX];
#pragma unroll
    for (int j = 0; j < NX; ++j) in[j] = xspag[j][t0 + t];
    SedObs o{0.0, 0.0, 0.0, 2.0};
    bool ok = true;
#pragma unroll
    for (int j = 0; j < NX; ++j) ok &= in[j] > 0.0;
    double pseudo[4];
    int kin = 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (p.sedPpo[i] == BAMGPU_PPO_FIX) pseudo[i] = p.sedPseudo[i];
        else { pseudo[i] = (kin < NX) ? in[kin < NX ? kin : 0] : 0.0; ++kin; }
    }
    const double d = pseudo[0], s = pseudo[1], v = pseudo[2], g = pseudo[3];
    ok = ok && d > 0.0 && s > 0.0 && v > 0.0 && g > 0.0 && s > 1.0;
    if (ok) {
        const double SD = pow(g * (s - 1.0) / (v * v), 1.0 / 3.0) * d;
        const double css = (p.sedCss == BAMGPU_CSS_SOULSBY) ? (0.3 / (1.0 + 1.2 * SD)) * 0.055 * (1.0 - exp(-1.0 * 0.02 * SD))
                                                           : (0.24 / SD) * 0.055 * (1.0 - exp(-1.0 * 0.02 * SD));
        const double tau = (in[0] * in[1]) / ((s - 1.0) * d);
        if (tau > 0.0) {
            if (tau <= css) o.st = 1.0;
            else {
                const double dim = sqrt(g * (s - 1.0) * (d * d * d));
                o.st = 0.0;
                switch (p.sedFormula) {
                    case BAMGPU_SED_MPM: o.B = dim; o.L1 = log(tau - css); break;
                    case BAMGPU_SED_CAMENEN: o.B = dim; o.L1 = log(tau); o.L2 = -1.0 * (css / tau); break;
                    case BAMGPU_SED_NIELSEN: o.B = dim * (tau - css); o.L1 = log(tau); break;
                    default:
                        o.B = dim * pow(s - 1.0, 1.0 / 6.0) * pow(d, 1.0 / 6.0) * pow(g, -0.5) * (tau - css);
                        o.L1 = log(in[1]); o.L2 = log(tau);
                        break;
                }
            }
        }
    }
    obs[t] = o;
}

using SedPrepKernel = void (*)(DevProblem, int, int, int, double* const*, SedObs*);
static inline SedPrepKernel pick_sed_prep(int nX) {
    switch (nX) {
        case 2: return sed_obs_prep<2>;
        case 3: return sed_obs_prep<3>;
        case 4: return sed_obs_prep<4>;
        case 5: return sed_obs_prep<5>;
        case 6: return sed_obs_prep<6>;
    }
    return nullptr;
}

// the per-input form applies (no pseudo-parameter is a fitted parameter, fixed css coefficients, css not a parameter)
static inline bool sed_fast_form(const DevProblem& p) {
    if (p.model != BAMGPU_MDL_SEDIMENT || p.sedCss == BAMGPU_CSS_CONSTANT || p.sedCo != BAMGPU_CO_FIX) return false;
    for (int i = 0; i < 4; ++i)
        if (p.sedPpo[i] == BAMGPU_PPO_PAR) return false;
    return true;
}
