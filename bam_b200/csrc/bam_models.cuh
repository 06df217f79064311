// bam_models.cuh -- device implementations of the closed-form models on the hot path.
//
// Contract (mirrors the plugin seam XXX_Apply of ModelLibrary_tools.f90:10-22, evaluated for ONE
// observation): given the inputs of one observation, the full parameter vector `th` (FIX filled
// in, VAR resolved) and the per-parameter-set derived values `der` computed once by
// model_prepare(), write the nY outputs and return row feasibility.
#pragma once
#include "bam_device.cuh"

// -------------------------------------------------------------------------------------------
// BaRatin (BaRatin_model.f90).  theta = (k_i, a_i, c_i) per control; der = offsets b_i.
// -------------------------------------------------------------------------------------------

// ApplyContinuity, BaRatin_model.f90:330-396.  Same operation order as the reference so the
// offsets b (which select H-b<0 branches) see identical operands.
__device__ inline bool baratin_prepare(const DevProblem& p, const double* th, double* b) {
    const int nc = p.ncontrol;
    if (th[1] <= 0.0 || th[2] == 0.0) return false;
    b[0] = th[0];
    for (int j = 1; j < nc; ++j) {
        double kj = th[3 * j], aj = th[3 * j + 1], cj = th[3 * j + 2];
        if (kj <= th[3 * (j - 1)]) return false;
        for (int i = 0; i < j; ++i)
            if (kj < b[i]) return false;
        if (aj <= 0.0 || cj == 0.0) return false;
        double som = 0.0;
        for (int i = 0; i < j; ++i) {
            int d = (int)((p.ctrlMask >> ((j - 1) * 8 + i)) & 1ull) - (int)((p.ctrlMask >> (j * 8 + i)) & 1ull);
            som = som + (double)d * th[3 * i + 1] * pow(kj - b[i], th[3 * i + 2]);
        }
        if (som < 0.0) return false;
        b[j] = kj - pow((1.0 / aj) * som, 1.0 / cj);
    }
    // der[nc + i] = log a_i (every a_i > 0 here): lets the per-observation kernel evaluate
    // a_i (H-b_i)^c_i as exp(c_i log(H-b_i) + log a_i), ~2x fewer FP64 instructions than pow()
    for (int i = 0; i < nc; ++i) b[nc + i] = log(th[3 * i + 1]);
    return true;
}

// BaRatin_computeQ + GetHRange, BaRatin_model.f90:194-252,398-436
__device__ __forceinline__ bool baratin_apply(const DevProblem& p, double H, const double* __restrict__ th,
                                              const double* __restrict__ b, double& Q) {
    const int nc = p.ncontrol;
    int range = nc;
    for (int i = 0; i < nc; ++i)
        if (H < th[3 * i]) { range = i; break; }
    double q = 0.0;
    bool ok = true;
    if (range > 0) {
        unsigned row = (unsigned)((p.ctrlMask >> ((range - 1) * 8)) & 0xffull);
        for (int i = 0; i < range; ++i) {
            double d = H - b[i];
            if ((row >> i) & 1u) {
                if (d < 0.0) { q = 0.0; break; }
                // a_i d^c_i; d == 0 gives log = -inf -> exp(-inf) = 0 for c > 0, +inf for c < 0, as pow() does
                q = q + exp(fma(th[3 * i + 2], log(d), b[nc + i]));
            } else if (d < 0.0) { ok = false; break; }
        }
    }
    Q = q;
    return ok;
}

// -------------------------------------------------------------------------------------------
// Linear (Linear_model.f90:62-101): Y = X . reshape(theta,(nX,nY))
// -------------------------------------------------------------------------------------------
template <int NX, int NY>
__device__ __forceinline__ void linear_apply(const double* x, const double* __restrict__ th, double* y) {
#pragma unroll
    for (int j = 0; j < NY; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < NX; ++i) s = s + x[i] * th[j * NX + i];
        y[j] = s;
    }
}

// -------------------------------------------------------------------------------------------
// Sediment (SedimentTransport_model.f90:100-192,251-310,525-607)
// -------------------------------------------------------------------------------------------
template <int NX>
__device__ inline bool sediment_apply(const DevProblem& p, const double* in, const double* __restrict__ th,
                                      double& out) {
#pragma unroll
    for (int i = 0; i < NX; ++i)
        if (in[i] <= 0.0) return false;
    const int f = p.sedFormula, cssId = p.sedCss;
    const int nbase = (f == BAMGPU_SED_MPM || f == BAMGPU_SED_NIELSEN) ? 2 : 3;
    int kt = nbase, kin = 2;
    double cte = 0.0;
    if (cssId == BAMGPU_CSS_CONSTANT) cte = th[kt++];
    double pseudo[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (p.sedPpo[i] == BAMGPU_PPO_FIX) pseudo[i] = p.sedPseudo[i];
        else if (p.sedPpo[i] == BAMGPU_PPO_PAR) pseudo[i] = th[kt++];
        else { pseudo[i] = (kin < NX) ? in[kin] : 0.0; ++kin; }
    }
    double c0, c1, c2, c3 = 0.0;
    if (cssId != BAMGPU_CSS_CONSTANT && p.sedCo == BAMGPU_CO_PAR) {
        c0 = th[kt]; c1 = th[kt + 1]; c2 = th[kt + 2];
        if (cssId == BAMGPU_CSS_SOULSBY) c3 = th[kt + 3];
    } else if (cssId == BAMGPU_CSS_SOULSBY) { c0 = 0.3; c1 = 1.2; c2 = 0.055; c3 = 0.02; }
    else { c0 = 0.24; c1 = 0.055; c2 = 0.02; }
    const double d = pseudo[0], s = pseudo[1], v = pseudo[2], g = pseudo[3];
    if (d <= 0.0 || s <= 0.0 || v <= 0.0 || g <= 0.0) return false;
    double css;
    if (cssId == BAMGPU_CSS_CONSTANT) css = cte;
    else {
        double SD = pow(g * (s - 1.0) / (v * v), 1.0 / 3.0) * d;
        if (cssId == BAMGPU_CSS_SOULSBY) css = (c0 / (1.0 + c1 * SD)) * c2 * (1.0 - exp(-1.0 * c3 * SD));
        else css = (c0 / SD) * c1 * (1.0 - exp(-1.0 * c2 * SD));
    }
    if (s <= 1.0) return false;
    const double tau = (in[0] * in[1]) / ((s - 1.0) * d);
    if (tau <= 0.0) return false;
    if (tau <= css) { out = 0.0; return true; }
    const double dim = sqrt(g * (s - 1.0) * (d * d * d));
    switch (f) {
        case BAMGPU_SED_MPM: out = dim * th[0] * pow(tau - css, th[1]); break;
        case BAMGPU_SED_CAMENEN: out = dim * th[0] * pow(tau, th[1]) * exp(-1.0 * th[2] * (css / tau)); break;
        case BAMGPU_SED_NIELSEN: out = dim * th[0] * pow(tau, th[1]) * (tau - css); break;
        default:
            out = dim * th[0] * pow(in[1], th[1]) * pow(s - 1.0, 1.0 / 6.0) * pow(d, 1.0 / 6.0) * pow(g, -0.5) *
                  pow(tau, th[2]) * (tau - css);
            break;
    }
    return true;
}

// -------------------------------------------------------------------------------------------
// TextFile (TextFile_model.f90:231-309): stack VM over host-compiled bytecode.  All lanes run the
// same program, so the interpreter loop is divergence-free except for the error exits.
// -------------------------------------------------------------------------------------------
enum { VM_IMMED = 1, VM_NEG, VM_ADD, VM_SUB, VM_MUL, VM_DIV, VM_POW, VM_ABS, VM_EXP, VM_LOG10, VM_LOG, VM_SQRT,
       VM_SINH, VM_COSH, VM_TANH, VM_SIN, VM_COS, VM_TAN, VM_ASIN, VM_ACOS, VM_ATAN, VM_IS0, VM_ISPOS, VM_ISNEG,
       VM_ISSPOS, VM_ISSNEG, VM_VARBEGIN };

template <int NX>
__device__ inline bool textfile_eval(const DevProblem& p, int k, const double* in, const double* __restrict__ th,
                                     double& res) {
    double st[BAM_VMSTACK];
    int sp = -1;
    const int* __restrict__ code = p.vmCode + p.vmCodeOff[k];
    const double* __restrict__ imm = p.vmImmed + p.vmImmedOff[k];
    int dp = 0;
    const int ncode = p.vmNcode[k];
    for (int ip = 0; ip < ncode; ++ip) {
        const int op = __ldg(code + ip);
        switch (op) {
            case VM_IMMED: st[++sp] = __ldg(imm + dp); ++dp; break;
            case VM_NEG: st[sp] = -st[sp]; break;
            case VM_ADD: st[sp - 1] = st[sp - 1] + st[sp]; --sp; break;
            case VM_SUB: st[sp - 1] = st[sp - 1] - st[sp]; --sp; break;
            case VM_MUL: st[sp - 1] = st[sp - 1] * st[sp]; --sp; break;
            case VM_DIV:
                if (st[sp] == 0.0) return false;
                st[sp - 1] = st[sp - 1] / st[sp]; --sp; break;
            case VM_POW: st[sp - 1] = pow(st[sp - 1], st[sp]); --sp; break;
            case VM_ABS: st[sp] = fabs(st[sp]); break;
            case VM_EXP: st[sp] = exp(st[sp]); break;
            case VM_LOG10:
                if (st[sp] <= 0.0) return false;
                st[sp] = log10(st[sp]); break;
            case VM_LOG:
                if (st[sp] <= 0.0) return false;
                st[sp] = log(st[sp]); break;
            case VM_SQRT:
                if (st[sp] < 0.0) return false;
                st[sp] = sqrt(st[sp]); break;
            case VM_SINH: st[sp] = sinh(st[sp]); break;
            case VM_COSH: st[sp] = cosh(st[sp]); break;
            case VM_TANH: st[sp] = tanh(st[sp]); break;
            case VM_SIN: st[sp] = sin(st[sp]); break;
            case VM_COS: st[sp] = cos(st[sp]); break;
            case VM_TAN: st[sp] = tan(st[sp]); break;
            case VM_ASIN:
                if (st[sp] < -1.0 || st[sp] > 1.0) return false;
                st[sp] = asin(st[sp]); break;
            case VM_ACOS:
                if (st[sp] < -1.0 || st[sp] > 1.0) return false;
                st[sp] = acos(st[sp]); break;
            case VM_ATAN: st[sp] = atan(st[sp]); break;
            case VM_IS0: st[sp] = (st[sp] == 0.0) ? 1.0 : 0.0; break;
            case VM_ISPOS: st[sp] = (st[sp] >= 0.0) ? 1.0 : 0.0; break;
            case VM_ISNEG: st[sp] = (st[sp] <= 0.0) ? 1.0 : 0.0; break;
            case VM_ISSPOS: st[sp] = (st[sp] > 0.0) ? 1.0 : 0.0; break;
            case VM_ISSNEG: st[sp] = (st[sp] < 0.0) ? 1.0 : 0.0; break;
            default: {
                const int v = op - VM_VARBEGIN;
                double val = 0.0;
                if (v < NX) {
#pragma unroll
                    for (int i = 0; i < NX; ++i)
                        if (i == v) val = in[i];
                } else val = th[v - NX];
                st[++sp] = val;
                break;
            }
        }
    }
    res = st[0];
    return !(res != res);
}

// -------------------------------------------------------------------------------------------
// dispatch
// -------------------------------------------------------------------------------------------

// once per (parameter set, VAR combination): derived values + feasibility
__device__ inline bool model_prepare(const DevProblem& p, const double* th, double* der) {
    if (p.model == BAMGPU_MDL_BARATIN) return baratin_prepare(p, th, der);
    return true;
}

// one observation.  y[NY]; returns row feasibility (ApplyModel's vfeas, ModelLibrary_tools.f90:419-432)
template <int MODEL, int NX, int NY>
__device__ __forceinline__ bool model_apply(const DevProblem& p, const double* x, const double* __restrict__ th,
                                            const double* __restrict__ der, double* y) {
    if (MODEL == BAMGPU_MDL_BARATIN) {
        return baratin_apply(p, x[0], th, der, y[0]);
    } else if (MODEL == BAMGPU_MDL_LINEAR) {
        linear_apply<NX, NY>(x, th, y);
        return true;
    } else if (MODEL == BAMGPU_MDL_SEDIMENT) {
        return sediment_apply<NX>(p, x, th, y[0]);
    } else {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NY; ++k) ok = textfile_eval<NX>(p, k, x, th, y[k]) && ok;
        return ok;
    }
}
