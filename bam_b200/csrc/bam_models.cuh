// bam_models.cuh -- device implementations of the closed-form models on the hot path.
//
// Contract (mirrors the plugin seam XXX_Apply of ModelLibrary_tools.f90:10-22, evaluated for ONE
// observation): given the inputs of one observation, the full parameter vector `th` (FIX filled
// in, VAR resolved) and the per-parameter-set derived values `der` computed once by
// model_prepare(), write the nY outputs and return row feasibility.
#pragma once
#include "bam_device.cuh"
#include "bam_fastmath.cuh"

// ---- table-driven log / exp for the generic kernels ------------------------------------------------------------
// Shared-memory addresses of the 256-entry log and exp tables of bam_fastmath.cuh (0, 0: not staged -> libdevice).  The
// models whose per-observation cost is dominated by libdevice pow / log / exp (Vegetation: ~10 evaluations of x^chi inside
// the bracketed Newton, three powers of the water depth; the TextFile VM's ^, exp, log opcodes) take them through these
// wrappers: tlog is within 5.4e-16 (absolute), texp within 2.5e-16 (relative) of libm (tests/test_fastmath.py).
struct FmTab {
    unsigned lt, et;
};
__device__ __forceinline__ double fm_log(double x, FmTab t) {  // x > 0 expected by the caller's own domain test
    return (t.lt != 0u && (unsigned)__double2hiint(x) - 0x00100000u < 0x7fe00000u) ? tlog_pos(x, t.lt) : log(x);
}
__device__ __forceinline__ double fm_exp(double x, FmTab t) { return t.et != 0u ? texp(x, t.et) : exp(x); }
__device__ __forceinline__ double fm_pow(double b, double e, FmTab t) {  // pow() itself outside b > 0 (signs, zeros, infinities)
    return (t.lt != 0u && t.et != 0u && (unsigned)__double2hiint(b) - 0x00100000u < 0x7fe00000u) ? texp(e * tlog_pos(b, t.lt), t.et) : pow(b, e);
}


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
// BaRatinBAC (BaRatinBAC_model.f90:96-666): theta = (b_i, a_i, c_i); the activation stages k_i follow from the
// continuity of the rating curve: k_i = b_i when control i is ADDED, the root of
//     f(x) = c2 log(x-b2) - c1 log(x-b1) - log(a1/a2)        (fContinuity :521-526)
// when control i REPLACES control i-1 (SolveContinuity :396-518, cases I-VI).  der = [k_i | ktype_i].
// -------------------------------------------------------------------------------------------
#define BAM_BAC_EPS 1.1920928955078125e-07 /* epsilon(1.): the reference uses the SINGLE-precision epsilon (:400) */

__device__ __forceinline__ double bac_f(double x, double a1, double b1, double c1, double a2, double b2, double c2) {
    return c2 * log(x - b2) - c1 * log(x - b1) - log(a1 / a2);
}
__device__ __forceinline__ double bac_df(double x, double b1, double c1, double b2, double c2) {
    return (c2 / (x - b2)) - (c1 / (x - b1));
}

// findRoot (:603-636) -> znewton with tolX = 1e-4, tolF = 1e-6, unit scales, itmax = 10000; same declared convention
// as the Vegetation solver (rtsafe + polish; miniDMSL is un-vendored), identical to oracle/bam_oracle.c:bac_find_root
__device__ inline bool bac_find_root(double a1, double b1, double c1, double a2, double b2, double c2, double x1, double x2,
                                     double fl, double fh, double& root) {
    if (fl == 0.0) { root = x1; return true; }
    if (fh == 0.0) { root = x2; return true; }
    if ((fl > 0.0) == (fh > 0.0) || fl != fl || fh != fh) return false;
    double xl, xh;
    if (fl < 0.0) { xl = x1; xh = x2; } else { xl = x2; xh = x1; }
    double rts = 0.5 * (x1 + x2), dxold = fabs(x2 - x1), dx = dxold;
    double f = bac_f(rts, a1, b1, c1, a2, b2, c2), df = bac_df(rts, b1, c1, b2, c2);
    for (int j = 0; j < 10000; ++j) {
        if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
            dxold = dx; dx = 0.5 * (xh - xl); rts = xl + dx;
        } else {
            dxold = dx; dx = f / df; rts = rts - dx;
        }
        const bool convx = fabs(dx) <= 1e-4;
        f = bac_f(rts, a1, b1, c1, a2, b2, c2);
        df = bac_df(rts, b1, c1, b2, c2);
        if (f < 0.0) xl = rts; else xh = rts;
        if (convx || fabs(f) <= 1e-6) break;
    }
    for (int j = 0; j < 4 && f != 0.0 && df != 0.0; ++j) {  // polish to the root itself
        const double lo = fmin(xl, xh), hi = fmax(xl, xh);
        double xn = rts - f / df;
        xn = fmin(fmax(xn, lo), hi);
        if (xn == rts) break;
        rts = xn;
        f = bac_f(rts, a1, b1, c1, a2, b2, c2);
        df = bac_df(rts, b1, c1, b2, c2);
        if (f < 0.0) xl = rts; else xh = rts;
    }
    root = rts;
    return root == root;
}

// SolveContinuity (:396-518); returns the number of roots (0 | 1)
__device__ inline int bac_solve(double a1, double b1, double c1, double a2, double b2, double c2, double lbound, double hbound,
                                double& root, int& troot) {
    const double eps = BAM_BAC_EPS;
    troot = 0;
    if (fabs(c1 - c2) <= eps) {  // case I
        if (fabs(a1 - a2) <= eps) return 0;
        const double foo1 = pow(a1, 1.0 / c1), foo2 = pow(a2, 1.0 / c1);
        root = (foo2 * b2 - foo1 * b1) / (foo2 - foo1);
        if (root > lbound && root < hbound) { troot = 1; return 1; }
        return 0;
    }
    if (fabs(b1 - b2) <= eps) {  // case II
        root = b1 + pow(a1 / a2, 1.0 / (c2 - c1));
        if (root > lbound && root < hbound) { troot = 2; return 1; }
        return 0;
    }
    const double x0 = (c2 * b1 - c1 * b2) / (c2 - c1);
    if (x0 <= lbound) {  // case III: monotonic within the bounds
        const double f1 = bac_f(lbound + eps, a1, b1, c1, a2, b2, c2), f2 = bac_f(hbound, a1, b1, c1, a2, b2, c2);
        if (f1 * f2 > 0.0) return 0;
        if (bac_find_root(a1, b1, c1, a2, b2, c2, lbound + eps, hbound, f1, f2, root)) { troot = 3; return 1; }
        return 0;
    }
    const double fx0 = bac_f(x0, a1, b1, c1, a2, b2, c2);
    if (fabs(fx0) <= eps) { root = x0; troot = 5; return 1; }  // case V
    if ((c2 > c1 && fx0 > 0.0) || (c2 < c1 && fx0 < 0.0)) { troot = 6; return 0; }  // case VI
    troot = 4;  // case IV: keep the largest of the two roots
    if (x0 >= hbound) return 0;
    const double f2 = bac_f(hbound, a1, b1, c1, a2, b2, c2);
    if (f2 * fx0 > 0.0) return 0;
    return bac_find_root(a1, b1, c1, a2, b2, c2, x0, hbound, fx0, f2, root) ? 1 : 0;
}

// ApplyContinuity (:237-394): der[0..nc) = k, der[nc..2nc) = ktype
__device__ inline bool baratinbac_prepare(const DevProblem& p, const double* th, double* der) {
    const int nc = p.ncontrol;
    for (int i = 0; i < nc; ++i)
        if (th[3 * i + 1] <= 0.0 || th[3 * i + 2] == 0.0) return false;
    der[0] = th[0];
    der[nc] = 0.0;
    for (int j = nc - 1; j >= 1; --j) {  // from the highest control down
        const double bj = th[3 * j], bjm = th[3 * (j - 1)];
        const double lbound = fmax(bj, bjm);
        const double hbound = (j == nc - 1) ? p.modelOptD : der[j + 1];
        bool addition = true;
        for (int i = 0; i < j; ++i)
            addition &= ((p.ctrlMask >> (j * 8 + i)) & 1ull) == ((p.ctrlMask >> ((j - 1) * 8 + i)) & 1ull);
        if (addition) { der[j] = bj; der[nc + j] = 0.0; }
        else {
            double root = 0.0;
            int troot = 0;
            if (bac_solve(th[3 * (j - 1) + 1], bjm, th[3 * (j - 1) + 2], th[3 * j + 1], bj, th[3 * j + 2], lbound, hbound, root, troot) == 0)
                return false;
            der[j] = root;
            der[nc + j] = (double)troot;
        }
        if (der[j] >= hbound) return false;
        for (int i = 0; i < j; ++i)
            if (der[j] < th[3 * i]) return false;
    }
    return true;
}

// BaRatin_computeQ (BaRatin_model.f90:194-252) with b from theta and k from der
__device__ __forceinline__ bool baratinbac_apply(const DevProblem& p, double H, const double* __restrict__ th,
                                                 const double* __restrict__ k, double& Q) {
    const int nc = p.ncontrol;
    int range = nc;
    for (int i = 0; i < nc; ++i)
        if (H < k[i]) { range = i; break; }
    double q = 0.0;
    bool ok = true;
    if (range > 0) {
        const unsigned row = (unsigned)((p.ctrlMask >> ((range - 1) * 8)) & 0xffull);
        for (int i = 0; i < range; ++i) {
            const double d = H - th[3 * i];
            if ((row >> i) & 1u) {
                if (d < 0.0) { q = 0.0; break; }
                q = q + th[3 * i + 1] * pow(d, th[3 * i + 2]);
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
                                     double& res, FmTab fm = FmTab{0u, 0u}) {
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
            case VM_POW: st[sp - 1] = fm_pow(st[sp - 1], st[sp], fm); --sp; break;
            case VM_ABS: st[sp] = fabs(st[sp]); break;
            case VM_EXP: st[sp] = fm_exp(st[sp], fm); break;
            case VM_LOG10:
                if (st[sp] <= 0.0) return false;
                st[sp] = log10(st[sp]); break;
            case VM_LOG:
                if (st[sp] <= 0.0) return false;
                st[sp] = fm_log(st[sp], fm); break;
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

// The same program for V parameter sets at once (the chains of a logpost_tile thread): one opcode fetch and one dispatch per
// V evaluations instead of per evaluation -- with a 5-opcode formula the interpreter's own bookkeeping (fetch, jump table,
// stack pointer) is 2/3 of the instructions.  Each set has its own inputs (latent cells differ per chain) and parameters.
// Returns the mask of sets whose evaluation FAILED (the error exits of textfile_eval, or a NaN result); the arithmetic per
// set is that of textfile_eval, operation for operation.
template <int NX, int V>
__device__ inline unsigned textfile_eval_multi(const DevProblem& p, int k, const double (*in)[NX], const double* const* th,
                                               double* res, FmTab fm = FmTab{0u, 0u}) {
    double st[BAM_VMSTACK][V];
    int sp = -1;
    unsigned fail = 0u;
    const int* __restrict__ code = p.vmCode + p.vmCodeOff[k];
    const double* __restrict__ imm = p.vmImmed + p.vmImmedOff[k];
    int dp = 0;
    const int ncode = p.vmNcode[k];
#define VM_EACH(expr) _Pragma("unroll") for (int v = 0; v < V; ++v) { expr; }
    for (int ip = 0; ip < ncode; ++ip) {
        const int op = __ldg(code + ip);
        switch (op) {
            case VM_IMMED: { const double c = __ldg(imm + dp); ++dp; ++sp; VM_EACH(st[sp][v] = c) break; }
            case VM_NEG: VM_EACH(st[sp][v] = -st[sp][v]) break;
            case VM_ADD: VM_EACH(st[sp - 1][v] = st[sp - 1][v] + st[sp][v]) --sp; break;
            case VM_SUB: VM_EACH(st[sp - 1][v] = st[sp - 1][v] - st[sp][v]) --sp; break;
            case VM_MUL: VM_EACH(st[sp - 1][v] = st[sp - 1][v] * st[sp][v]) --sp; break;
            case VM_DIV: VM_EACH(if (st[sp][v] == 0.0) fail |= 1u << v; st[sp - 1][v] = st[sp - 1][v] / st[sp][v]) --sp; break;
            case VM_POW: VM_EACH(st[sp - 1][v] = fm_pow(st[sp - 1][v], st[sp][v], fm)) --sp; break;
            case VM_ABS: VM_EACH(st[sp][v] = fabs(st[sp][v])) break;
            case VM_EXP: VM_EACH(st[sp][v] = fm_exp(st[sp][v], fm)) break;
            case VM_LOG10: VM_EACH(if (st[sp][v] <= 0.0) fail |= 1u << v; st[sp][v] = log10(st[sp][v])) break;
            case VM_LOG: VM_EACH(if (st[sp][v] <= 0.0) { fail |= 1u << v; st[sp][v] = 1.0; } st[sp][v] = fm_log(st[sp][v], fm)) break;
            case VM_SQRT: VM_EACH(if (st[sp][v] < 0.0) fail |= 1u << v; st[sp][v] = sqrt(st[sp][v])) break;
            case VM_SINH: VM_EACH(st[sp][v] = sinh(st[sp][v])) break;
            case VM_COSH: VM_EACH(st[sp][v] = cosh(st[sp][v])) break;
            case VM_TANH: VM_EACH(st[sp][v] = tanh(st[sp][v])) break;
            case VM_SIN: VM_EACH(st[sp][v] = sin(st[sp][v])) break;
            case VM_COS: VM_EACH(st[sp][v] = cos(st[sp][v])) break;
            case VM_TAN: VM_EACH(st[sp][v] = tan(st[sp][v])) break;
            case VM_ASIN: VM_EACH(if (st[sp][v] < -1.0 || st[sp][v] > 1.0) fail |= 1u << v; st[sp][v] = asin(st[sp][v])) break;
            case VM_ACOS: VM_EACH(if (st[sp][v] < -1.0 || st[sp][v] > 1.0) fail |= 1u << v; st[sp][v] = acos(st[sp][v])) break;
            case VM_ATAN: VM_EACH(st[sp][v] = atan(st[sp][v])) break;
            case VM_IS0: VM_EACH(st[sp][v] = (st[sp][v] == 0.0) ? 1.0 : 0.0) break;
            case VM_ISPOS: VM_EACH(st[sp][v] = (st[sp][v] >= 0.0) ? 1.0 : 0.0) break;
            case VM_ISNEG: VM_EACH(st[sp][v] = (st[sp][v] <= 0.0) ? 1.0 : 0.0) break;
            case VM_ISSPOS: VM_EACH(st[sp][v] = (st[sp][v] > 0.0) ? 1.0 : 0.0) break;
            case VM_ISSNEG: VM_EACH(st[sp][v] = (st[sp][v] < 0.0) ? 1.0 : 0.0) break;
            default: {
                const int q = op - VM_VARBEGIN;
                ++sp;
                if (q < NX) {
                    VM_EACH(double val = 0.0; _Pragma("unroll") for (int i = 0; i < NX; ++i) if (i == q) val = in[v][i]; st[sp][v] = val)
                } else {
                    VM_EACH(st[sp][v] = th[v][q - NX])
                }
                break;
            }
        }
    }
#undef VM_EACH
#pragma unroll
    for (int v = 0; v < V; ++v) {
        res[v] = st[0][v];
        if (res[v] != res[v]) fail |= 1u << v;
    }
    return fail;
}

// -------------------------------------------------------------------------------------------
// Vegetation (Vegetation_model.f90:87-569).  theta = (B,S0,n1,b1,c1 | chi,U_chi | growth pars | gmax).
// Per parameter set (vegetation_prepare): feasibility + growth-curve constants; for the temperature-driven
// formulas this is the per-year degree-day scan of Apply_growth (:420-532) over the WHOLE input series.
// der layout: [0] B sqrt(S0)/n1 (Dpar 1); temporal: [1] alpha, [2] t_f;
//             else [1+i] T_m of year i (Dpar 2..), then per year 5 values at [1+nYear+5i]: t_b, t_e, t_f, alpha, g==0 flag.
// -------------------------------------------------------------------------------------------
#define BAM_VEG_PI 3.14159265358979323846
#define BAM_VEG_GRAVITY 9.81

__device__ inline bool vegetation_prepare(const DevProblem& p, const double* th, double* der, const double* const* xc,
                                          int nser) {
    const bool temporal = p.vegGrowth == BAMGPU_VEG_YIN1_TEMPORAL;
    const int nG = temporal ? 3 : 4, nGmax = temporal ? 1 : p.vegNYear;
    const double B = th[0], S0 = th[1], n1 = th[2], c1 = th[4], chi = th[5], U_chi = th[6];
    const double* par = th + 7;
    const double* gmax = th + 7 + nG;
    der[0] = (S0 >= 0.0) ? B * sqrt(S0) / n1 : BAMGPU_UNDEF_RN;
    bool badg = false;
    for (int i = 0; i < nGmax; ++i) badg |= gmax[i] < 0.0;
    if (B <= 0.0 || S0 <= 0.0 || n1 <= 0.0 || c1 <= 0.0 || chi < -2.0 || chi > 0.0 || U_chi <= 0.0 || badg) return false;
    der[temporal ? 3 : 1 + 6 * p.vegNYear] = pow(B, chi) * pow(U_chi, chi);  // gamma2 = B^chi y^chi U_chi^chi (:216-218)
    if (temporal) {  // :395-419
        const double t0 = par[0], ti = par[1], tmax = par[2];
        if (t0 <= 0.0 || ti <= 0.0 || tmax <= 0.0 || t0 >= 1.0 || ti >= 1.0 || tmax >= 1.0 || t0 >= ti || ti >= tmax) return false;
        der[1] = (tmax - t0) / (tmax - ti);
        der[2] = 2.0 * tmax - ti;
        return true;
    }
    const double Tmin = par[0], Tb = par[1], Te = par[2];
    const int nYear = p.vegNYear;
    for (int i = 0; i < nYear; ++i) der[1 + i] = BAMGPU_UNDEF_RN;
    if (Te <= Tb) return false;
    const double *time = xc[0], *Traw = xc[2], *Tsmooth = xc[3];
    // The reference scans the whole series once per year (Apply_growth :420-532).  Here ONE sweep serves all years: the
    // running state of the year being visited (degree-day sum, previous time, first crossings) lives in registers and is
    // parked in that year's der slots when the series moves to another year, so every year still sees its own rows in
    // order of appearance with the same arithmetic -- nser iterations instead of nYear x nser.
    // parked state of year i: yr[0] t_b, yr[1] t_e, yr[2] degree-day sum, yr[3] previous time (= last time once the year has
    // a row), yr[4] flags (1 crossed T_b, 2 crossed T_e, 4 crossed T_m, 8 has rows), der[1+i] t_m
    for (int i = 0; i < nYear; ++i) {
        double* yr = der + 1 + nYear + 5 * i;
        yr[0] = 0.0; yr[1] = 0.0; yr[2] = 0.0; yr[3] = (double)i; yr[4] = 0.0;
        der[1 + i] = 0.0;
    }
    const bool yin1 = p.vegGrowth == BAMGPU_VEG_YIN1;
    {
        int cy = -1;
        double cumul = 0.0, prev = 0.0, time_b = 0.0, time_e = 0.0, time_m = 0.0;
        unsigned fl = 0u;
        const double Tm1 = yin1 ? par[3] : 0.0;
        // one row of the year in registers: the loop-carried chain is the degree-day sum alone (one fused add per row)
        auto step = [&](double tj, double tr, double ts) {
            if (tr >= Tmin) cumul = cumul + 365.0 * (tj - prev) * (tr - Tmin);
            prev = tj;
            fl |= 8u;
            if (!(fl & 1u) && cumul >= Tb) { fl |= 1u; time_b = tj; }
            if (!(fl & 2u) && cumul >= Te) { fl |= 2u; time_e = tj; }
            if (yin1 && !(fl & 4u) && ts >= Tm1) { fl |= 4u; time_m = tj; }
        };
        auto row = [&](double tj, double tr, double ts) {  // any row: switch the register state to its year first
            const double fy = floor(tj);
            if (!(fy >= 0.0 && fy < (double)nYear)) return;
            const int iy = (int)fy;
            if (iy != cy) {
                if (cy >= 0) {
                    double* yr = der + 1 + nYear + 5 * cy;
                    yr[0] = time_b; yr[1] = time_e; yr[2] = cumul; yr[3] = prev; yr[4] = (double)fl;
                    der[1 + cy] = time_m;
                }
                const double* yr = der + 1 + nYear + 5 * iy;
                time_b = yr[0]; time_e = yr[1]; cumul = yr[2]; prev = yr[3]; fl = (unsigned)yr[4];
                time_m = der[1 + iy];
                cy = iy;
            }
            step(tj, tr, ts);
        };
        int j = 0;
        // four rows per trip: the loads and the year tests of the four are independent, so their latencies overlap; rows
        // of the year already in registers (the usual case: a series in time order) skip the state switch
        for (; j + 4 <= nser; j += 4) {
            const double t0 = time[j], t1 = time[j + 1], t2 = time[j + 2], t3 = time[j + 3];
            const double r0 = Traw[j], r1 = Traw[j + 1], r2 = Traw[j + 2], r3 = Traw[j + 3];
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            if (yin1) { s0 = Tsmooth[j]; s1 = Tsmooth[j + 1]; s2 = Tsmooth[j + 2]; s3 = Tsmooth[j + 3]; }
            const double cyd = (double)cy;
            if (cy >= 0 && floor(t0) == cyd && floor(t1) == cyd && floor(t2) == cyd && floor(t3) == cyd) {
                // once every crossing of the year is found, its later rows only move "the last time of the year"
                const unsigned all = 1u | 2u | 8u | (yin1 ? 4u : 0u);
                if ((fl & all) == all) prev = t3;
                else { step(t0, r0, s0); step(t1, r1, s1); step(t2, r2, s2); step(t3, r3, s3); }
            } else {
                row(t0, r0, s0); row(t1, r1, s1); row(t2, r2, s2); row(t3, r3, s3);
            }
        }
        for (; j < nser; ++j) row(time[j], Traw[j], yin1 ? Tsmooth[j] : 0.0);
        if (cy >= 0) {
            double* yr = der + 1 + nYear + 5 * cy;
            yr[0] = time_b; yr[1] = time_e; yr[2] = cumul; yr[3] = prev; yr[4] = (double)fl;
            der[1 + cy] = time_m;
        }
    }
    // per year, in year order with the reference's exits: growth-curve constants; t_m parked in der[1+i] for the second sweep
    unsigned need = 0u;
    bool ok = true;
    for (int i = 0; i < nYear && ok; ++i) {
        double* yr = der + 1 + nYear + 5 * i;
        const double time_b = yr[0], time_e = yr[1], last = yr[3];
        const unsigned fl = (unsigned)yr[4];
        double time_m = der[1 + i];
        der[1 + i] = BAMGPU_UNDEF_RN;
        if (!(fl & 8u)) { ok = false; break; }  // "empty year": an error in the reference (:432-436); checked on the host where possible
        if (!(fl & 2u)) { ok = false; break; }
        const bool gotB = (fl & 1u) != 0u;
        yr[2] = 0.0; yr[3] = 0.0; yr[4] = gotB ? 0.0 : 1.0;
        if (!gotB) continue;  // no growth this year
        double alpha, time_f;
        if (yin1) {
            if (!(fl & 4u)) { ok = false; break; }
            if (time_m <= time_b) time_m = time_b;
            if (time_m >= time_e) { ok = false; break; }
            alpha = (time_e - time_b) / (time_e - time_m);
            time_f = 2.0 * time_e - time_m;
        } else if (p.vegGrowth == BAMGPU_VEG_YIN2) {
            if (par[3] <= 0.0) { ok = false; break; }
            alpha = par[3];
            time_f = time_e + (time_e - time_b) / alpha;
            time_m = ((alpha - 1.0) * time_f + 2.0 * time_b) / (alpha + 1.0);
        } else {
            if (par[3] <= 0.0) { ok = false; break; }
            time_f = time_e + par[3];
            alpha = (time_e - time_b) / (time_f - time_e);
            time_m = ((alpha - 1.0) * time_f + 2.0 * time_b) / (alpha + 1.0);
        }
        if (time_m < time_b) time_m = time_b;
        if (alpha <= 0.0 || time_m >= time_e || time_f > last) { ok = false; break; }
        yr[2] = time_f; yr[3] = alpha;
        der[1 + i] = time_m;
        need |= 1u << i;
    }
    if (!ok) {
        for (int i = 0; i < nYear; ++i)
            if ((need >> i) & 1u) der[1 + i] = BAMGPU_UNDEF_RN;
        return false;
    }
    // second sweep: optimal temperature = smoothed temperature at the first point of the year with time >= t_m
    {
        auto look = [&](double tj, int j) {
            const double fy = floor(tj);
            if (!(fy >= 0.0 && fy < (double)nYear)) return;
            const int iy = (int)fy;
            if (((need >> iy) & 1u) && tj >= der[1 + iy]) { der[1 + iy] = Tsmooth[j]; need &= ~(1u << iy); }
        };
        int j = 0;
        for (; j + 4 <= nser && need != 0u; j += 4) {
            const double t0 = time[j], t1 = time[j + 1], t2 = time[j + 2], t3 = time[j + 3];
            look(t0, j); look(t1, j + 1); look(t2, j + 2); look(t3, j + 3);
        }
        for (; j < nser && need != 0u; ++j) look(time[j], j);
    }
    for (int i = 0; i < nYear; ++i)
        if ((need >> i) & 1u) der[1 + i] = BAMGPU_UNDEF_RN;
    return true;
}

// Vegetation_fNewton (:536-569).  x^chi once per evaluation, as exp(chi log x) (x^(1+chi) = x x^chi): the reference's two
// pow() calls per evaluation dominate the per-observation cost.
// The loop invariants of a solve are computed once (veg_root): Q0^2, 1 / gamma2 and gamma / gamma2 (2 + chi) -- the reference
// divides by gamma2 twice in EVERY evaluation, and an FP64 division is ~30 instructions here; x^chi / gamma2 becomes a product
// (<= 1 ulp from the quotient: the iterates move by that much, the polished root does not).
struct VegNewton { double Q0sq, gam, rg2, k3, chi; };
// Newton step f / df: reciprocal (MUFU seed + one cubic correction, <= 1 ulp) times f for a derivative in the normal range,
// the IEEE quotient otherwise
__device__ __forceinline__ double veg_quot(double f, double df) {
    const unsigned hi = (unsigned)__double2hiint(df);
    return (hi - 0x00200000u < 0x7fb00000u) ? f * trcp(df) : f / df;
}
__device__ __forceinline__ void veg_fnewton(double x, const VegNewton& w, double& f, double& df, FmTab fm = FmTab{0u, 0u}) {
    // x > 0 (the bracket's lower end x = 0 is evaluated in closed form by veg_root)
    const double px = fm_exp(w.chi * fm_log(x, fm), fm);
    double m = px * w.rg2;
    if (1.0 < m) m = 1.0;
    f = x * x + (w.gam * (x * x) * m) - w.Q0sq;
    if (m == 1.0) df = 2.0 * x * (1.0 + w.gam);
    else df = 2.0 * x + (w.k3 * (x * px));
}

// znewton on [0, Q0]: declared convention (miniDMSL is un-vendored), the algorithm of oracle/bam_oracle.c:veg_znewton
__device__ inline bool veg_root(const DevProblem& p, double Q0, double gam, double gam2, double chi, double& root,
                                FmTab fm = FmTab{0u, 0u}) {
    double fl, fh, df, f, xl, xh;
    const double x1 = 0.0, x2 = Q0;
    VegNewton w;
    w.Q0sq = Q0 * Q0; w.gam = gam; w.rg2 = 1.0 / gam2; w.k3 = gam / gam2 * (2.0 + chi); w.chi = chi;
    // f(0) = 0 + gamma 0 m - Q0^2 with m = min(0^chi / gamma2, 1) in (0, 1]: no power needed (a NaN gamma still gives NaN)
    fl = 0.0 + (gam * 0.0) - w.Q0sq;
    int fcalls = 2;
    if (fl == 0.0) { root = x1; return fcalls <= p.vegItmax; }  // Q0 = 0
    if (!(Q0 > 0.0)) return false;                              // negative or NaN Q0: no bracket (f(Q0) was NaN)
    // f(Q0) = gamma Q0^2 m(Q0) with m in (0, 1]: positive for a positive finite gamma and gamma2 -- only its sign is needed
    // (a value that rounds to zero made the reference return Q0; the iteration below arrives there to rounding)
    if (gam > 0.0 && gam < CUDART_INF && gam2 > 0.0 && gam2 < CUDART_INF && w.Q0sq < CUDART_INF) fh = 1.0;
    else veg_fnewton(x2, w, fh, df, fm);
    if (fh == 0.0) { root = x2; return fcalls <= p.vegItmax; }
    if ((fl > 0.0) == (fh > 0.0) || fl != fl || fh != fh) return false;
    if (fl < 0.0) { xl = x1; xh = x2; } else { xl = x2; xh = x1; }
    // first iterate: Q0 / sqrt(1 + gamma), the root itself wherever the bending factor is saturated (m = 1) and a lower bound
    // of it otherwise (f is increasing on (0, Q0]), instead of the bracket's midpoint: 2-3 safeguarded Newton steps instead
    // of 6-7.  The path is ours to choose (the reference's is un-vendored); the polished root is the bracket's only root.
    double rts = Q0 * rsqrt(1.0 + gam), dxold = fabs(x2 - x1), dx = dxold;
    if (!(rts > 0.0 && rts < Q0)) rts = 0.5 * (x1 + x2);
    veg_fnewton(rts, w, f, df, fm);
    ++fcalls;
    const double tolx = p.vegTolX * p.vegXscale, tolf = p.vegTolF * p.vegFscale;
    for (int j = 0; j < p.vegItmax; ++j) {
        if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
            dxold = dx; dx = 0.5 * (xh - xl); rts = xl + dx;
        } else {
            dxold = dx; dx = veg_quot(f, df); rts = rts - dx;
        }
        const bool convx = fabs(dx) <= tolx;
        if (!(rts > 0.0)) return false;  // cannot happen inside a bracket (0, Q0] with f(0) < 0; guards the closed-form f(0)
        veg_fnewton(rts, w, f, df, fm);
        ++fcalls;
        if (f < 0.0) xl = rts; else xh = rts;
        if (convx || fabs(f) <= tolf) break;
    }
    for (int j = 0; j < 3 && f != 0.0 && df > 0.0; ++j) {  // polish to the root itself
        const double lo = fmin(xl, xh), hi = fmax(xl, xh);
        double xn = rts - veg_quot(f, df);
        xn = fmin(fmax(xn, lo), hi);
        if (xn == rts || !(xn > 0.0)) break;
        rts = xn;
        veg_fnewton(rts, w, f, df, fm);
        if (f < 0.0) xl = rts; else xh = rts;
    }
    root = rts;
    return fcalls <= p.vegItmax;
}

template <int NX, int NY>
__device__ inline bool vegetation_apply(const DevProblem& p, const double* in, const double* __restrict__ th,
                                        const double* __restrict__ der, double* out, FmTab fm = FmTab{0u, 0u}) {
    const double time = in[0], h = in[1];
    double g, g0, cosg, sing;
    if (p.vegGrowth == BAMGPU_VEG_YIN1_TEMPORAL) {
        const double t0 = th[7], tmax = th[9], alf = der[1], tf = der[2];
        const double tt = time - floor(time);
        if (tt >= t0 && tt <= tf) g0 = (fm_pow((tt - t0) / (tmax - t0), alf, fm)) * (tf - tt) / (tf - tmax);
        else g0 = 0.0;
        double sn;
        sincos(g0 * BAM_VEG_PI, &sn, &cosg);
        sing = (tt <= tmax) ? sn : -1.0 * sn;
        g = th[10] * g0;
    } else {
        const int nYear = p.vegNYear;
        const double fy = floor(time);
        if (fy >= 0.0 && fy < (double)nYear) {
            const int i = (int)fy;
            const double* yr = der + 1 + nYear + 5 * i;
            const double time_b = yr[0], time_e = yr[1], time_f = yr[2], alpha = yr[3];
            if (yr[4] != 0.0 || time <= time_b || time >= time_f) g0 = 0.0;
            else g0 = fm_pow((time - time_b) / (time_e - time_b), alpha, fm) * (time_f - time) / (time_f - time_e);
            g = g0 * th[11 + i];
            double sn;
            sincos(g0 * BAM_VEG_PI, &sn, &cosg);
            sing = (time <= time_e) ? sn : -1.0 * sn;
        } else g = g0 = cosg = sing = BAMGPU_UNDEF_RN;  // a point outside every year keeps the initial undefRN (:380)
    }
    if (NY > 1) out[1] = g;
    if (NY > 2) out[2] = cosg;
    if (NY > 3) out[3] = sing;
    if (NY > 4) out[4] = g0;
    const double y = h - th[3];
    if (y <= 0.0) { out[0] = 0.0; return true; }
    const double ly = fm_log(y, fm);  // y > 0: the three powers of y share one logarithm
    const double Q0 = der[0] * fm_exp(th[4] * ly, fm);
    if (g <= 0.0) { out[0] = Q0; return true; }
    const double chi = th[5];
    const double gam = (g * (fm.et != 0u ? texp(ly * (1.0 / 3.0), fm.et) : cbrt(y))) / (8.0 * BAM_VEG_GRAVITY * (th[2] * th[2]));
    const double gam2 = der[(p.vegGrowth == BAMGPU_VEG_YIN1_TEMPORAL) ? 3 : 1 + 6 * p.vegNYear] * fm_exp(chi * ly, fm);
    double q;
    if (!veg_root(p, Q0, gam, gam2, chi, q, fm)) return false;
    out[0] = q;
    return true;
}

// -------------------------------------------------------------------------------------------
// Second-wave pointwise models
// -------------------------------------------------------------------------------------------

// SuspendedLoad_Apply (SuspendedLoad_model.f90:38-116): Qs = t1 (h-t2)^t3 / (h-t4)^t5
__device__ __forceinline__ bool suspendedload_apply(double h, const double* __restrict__ th, double& out) {
    if (h < th[1] || h < th[3]) return false;
    const double num = pow(h - th[1], th[2]), den = pow(h - th[3], th[4]);
    out = th[0] * (num / den);
    return true;
}

// SWOT_Apply (SWOT_model.f90:73-148): Q = K (B-B0) sqrt(S-S0) (h-h0)^M ; variant hS: B = theta2, B0 = 0
template <int NX>
__device__ __forceinline__ bool swot_apply(const DevProblem& p, const double* in, const double* __restrict__ th, double& out) {
    const double K = th[0], S0 = th[2], h0 = th[3], M = th[4];
    double Bt, B0;
    if (p.modelOpt[0] == BAMGPU_SWOT_HS) { Bt = th[1]; B0 = 0.0; }
    else { Bt = in[NX > 2 ? 2 : 0]; B0 = th[1]; }
    if (Bt - B0 < 0.0 || in[1] - S0 < 0.0) return false;
    if (in[0] - h0 <= 0.0) { out = 0.0; return true; }
    out = K * (Bt - B0) * sqrt(in[1] - S0) * pow(in[0] - h0, M);
    return true;
}

// SGD_Apply (StageGradientDischarge_model.f90:57-113)
__device__ __forceinline__ bool sgd_apply(const double* in, const double* __restrict__ th, double& out) {
    const double K = th[0], B = th[1], S0 = th[2], h0 = th[3], M = th[4];
    const double h = in[0], dhdt = in[1];
    if (h - h0 <= 0.0) return false;
    const double Q0 = K * B * sqrt(S0) * pow(h - h0, M);
    const double c = M * K * sqrt(S0) * pow(h - h0, M - 1.0);
    const double corr = 1.0 + (1.0 / (c * S0)) * dhdt;
    if (corr <= 0.0) return false;
    out = Q0 * sqrt(corr);
    return true;
}

// Recession_h_Apply (Recession_model.f90:36-94): normalised two-exponential recession
__device__ __forceinline__ bool recession_apply(double t, const double* __restrict__ th, double& out) {
    out = th[0] * exp(-th[1] * t) + (1.0 - th[0] - th[2]) * exp(-th[3] * t) + th[2];
    return true;
}

// Ortho_Apply (Orthorectification_model.f90:66-190).  der[0..19] = Dpar (r11..r33, a1..a4, b1..b4, c1..c3),
// der[20] = f / pixel width, der[21] = f / pixel height.
__device__ inline bool ortho_prepare(const double* th, double* der) {
    const double x0 = th[0], y0 = th[1], z0 = th[2], azim = th[3], roll = th[4], tilt = th[5], f = th[6], c0 = th[7],
                 l0 = th[8], lambda = th[9], k = th[10];
    const double pc = lambda, pl = k * pc;
    if (pc <= 0.0 || pl <= 0.0 || k <= 0.0) return false;
    const double ac = f / pc, al = f / pl;
    const double r11 = cos(azim) * cos(roll) + sin(azim) * cos(tilt) * sin(roll);
    const double r12 = -1.0 * sin(azim) * cos(roll) + cos(azim) * cos(tilt) * sin(roll);
    const double r13 = sin(tilt) * sin(roll);
    const double r21 = -1.0 * cos(azim) * sin(roll) + sin(azim) * cos(tilt) * cos(roll);
    const double r22 = sin(azim) * sin(roll) + cos(azim) * cos(tilt) * cos(roll);
    const double r23 = sin(tilt) * cos(roll);
    const double r31 = sin(azim) * sin(tilt), r32 = cos(azim) * sin(tilt), r33 = -1.0 * cos(tilt);
    const double alpha = r31 * x0 + r32 * y0 + r33 * z0;
    if (alpha == 0.0) return false;
    const double a1 = (ac * r21 - r31 * c0) / alpha, a2 = (ac * r22 - r32 * c0) / alpha, a3 = (ac * r23 - r33 * c0) / alpha;
    const double a4 = -1.0 * (a1 * x0 + a2 * y0 + a3 * z0);
    const double b1 = (al * r11 - r31 * l0) / alpha, b2 = (al * r12 - r32 * l0) / alpha, b3 = (al * r13 - r33 * l0) / alpha;
    const double b4 = -1.0 * (b1 * x0 + b2 * y0 + b3 * z0);
    const double vals[22] = {r11, r12, r13, r21, r22, r23, r31, r32, r33, a1, a2, a3, a4, b1, b2, b3, b4,
                             -1.0 * r31 / alpha, -1.0 * r32 / alpha, -1.0 * r33 / alpha, ac, al};
    for (int i = 0; i < 22; ++i) der[i] = vals[i];
    return true;
}

__device__ inline bool ortho_apply(const DevProblem& p, const double* in, const double* __restrict__ th,
                                   const double* __restrict__ der, double* out) {
    const double dx = in[0] - th[0], dy = in[1] - th[1], dz = in[2] - th[2];
    const double denom = der[6] * dx + der[7] * dy + der[8] * dz;
    if (denom == 0.0) return false;
    const double c0 = th[7], l0 = th[8];
    const double c = c0 - der[20] * (der[3] * dx + der[4] * dy + der[5] * dz) / denom;
    const double l = l0 - der[21] * (der[0] * dx + der[1] * dy + der[2] * dz) / denom;
    if (p.modelOpt[0] == BAMGPU_DISTO_NONE) { out[0] = c; out[1] = l; return true; }
    // Disto_Apply, radial (:271-280): d = 1 + sum_i theta(11+i) r^(2i)
    const double r = sqrt((c - c0) * (c - c0) + (l - l0) * (l - l0));
    const double r2 = r * r;
    double d = 1.0, rp = 1.0;
    for (int i = 0; i < p.modelOpt[1]; ++i) { rp = rp * r2; d = d + th[11 + i] * rp; }
    out[0] = c0 + d * (c - c0);
    out[1] = l0 + d * (l - l0);
    return true;
}

// HydraulicControl_section_Apply (HydraulicControl_section_model.f90:40-112): critical-section control from a
// bathymetry table (h, A, w); the critical stage solves h + A/(2w) = H by linear interpolation of the table.
__device__ inline bool hcs_apply(const DevProblem& p, double H, const double* __restrict__ th, double& out) {
    const int np = p.hcsN;
    const double* __restrict__ hh = p.hcsTab;
    const double* __restrict__ AA = p.hcsTab + np;
    const double* __restrict__ lhs = p.hcsTab + 3 * np;
    if (H <= hh[0]) { out = 0.0; return true; }
    if (H >= hh[np - 1]) return false;
    int k1 = -1;  // maxloc(fx, mask = fx <= 0): first position of the largest non-positive fx
    double best = 0.0;
    for (int i = 0; i < np; ++i) {
        const double fx = lhs[i] - H;
        if (fx <= 0.0 && (k1 < 0 || fx > best)) { k1 = i; best = fx; }
    }
    const int k2 = k1 + 1;
    if (k1 < 0 || k2 >= np) return false;
    const double x1 = hh[k1], x2 = hh[k2];
    double y1 = lhs[k1] - H, y2 = lhs[k2] - H;
    const double hc = (x1 * y2 - x2 * y1) / (y2 - y1);
    y1 = AA[k1]; y2 = AA[k2];
    const double Ahc = (y1 * (x2 - hc) + y2 * (hc - x1)) / (x2 - x1);
    out = th[0] * sqrt(2.0 * th[1] * (H - hc)) * Ahc;
    return true;
}

// Segmentation_Apply (Segmentation_model.f90:52-140): theta = (mu_1..mu_nS, nS-1 shift times | durations).
// Per parameter set: shift times tau (-> der), feasibility incl. the >= nmin points per segment test, which scans the
// whole input series; per observation: Y = mu(segment), segment = first i with t < tau_i (GettRange :182-215).
__device__ inline bool segmentation_prepare(const DevProblem& p, const double* th, double* der, const double* time, int nser) {
    const int nS = p.modelOpt[0], nmin = p.modelOpt[1], option = p.modelOpt[2];
    if (nS == 1) return true;
    double prev = 0.0, acc = 0.0;
    for (int i = 0; i < nS - 1; ++i) {
        double dur, tau;
        if (option == 1) { tau = th[nS + i]; dur = (i == 0) ? 1.0 : tau - prev; }
        else { dur = th[nS + i]; acc = acc + dur; tau = acc; }
        if (dur <= 0.0 || dur <= p.modelOptD) return false;
        der[i] = tau;
        prev = tau;
    }
    double tmax = -INFINITY;
    for (int j = 0; j < nser; ++j) tmax = fmax(tmax, time[j]);
    for (int i = 0; i < nS - 1; ++i)
        if (der[i] >= tmax) return false;
    int cnt[BAM_MAXTHETA];
    for (int i = 0; i < nS; ++i) cnt[i] = 0;
    for (int j = 0; j < nser; ++j) {
        int seg = nS - 1;
        for (int i = 0; i < nS - 1; ++i)
            if (time[j] < der[i]) { seg = i; break; }
        ++cnt[seg];
    }
    for (int i = 0; i < nS; ++i)
        if (cnt[i] < nmin) return false;
    return true;
}

__device__ __forceinline__ bool segmentation_apply(const DevProblem& p, double t, const double* __restrict__ th,
                                                   const double* __restrict__ der, double& out) {
    const int nS = p.modelOpt[0];
    int seg = nS - 1;
    for (int i = 0; i < nS - 1; ++i)
        if (t < der[i]) { seg = i; break; }
    out = th[seg];
    return true;
}

// -------------------------------------------------------------------------------------------
// SFD (StageFallDischarge_model.f90:80-283, SFD_Val_General): twin-gauge rating curve.  theta = (KB, h0, M, L, h0',
// KBS0, delta, M').  Per observation the transition stage kappa between the variable-slope and the channel-control
// regime solves  KB (x-h0)^M sqrt((x-h2-delta)/L) = KBS0 (x-h0')^M'  (fNewton_Val_General :253-283) on
// [max(h0,h0',h2+delta)+0.001, upperbound] by znewton (declared convention as for Vegetation: rtsafe + polish).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void sfd_fnewton(double x, const double* __restrict__ th, double h2, double& f, double& df) {
    const double s = sqrt((x - h2 - th[6]) / th[3]);
    f = th[0] * pow(x - th[1], th[2]) * s - th[5] * pow(x - th[4], th[7]);
    df = th[0] * pow(x - th[1], th[2] - 1.0) / (2.0 * th[3] * s) *
             ((2.0 * th[2] + 1.0) * x - (2.0 * th[2] * (h2 + th[6]) + th[1])) -
         th[5] * th[7] * pow(x - th[4], th[7] - 1.0);
}

// kappaOut: the state variable (transition stage); stays undefRN on the rows that never solve for it (:126,143-144)
__device__ inline bool sfd_apply(const DevProblem& p, const double* in, const double* __restrict__ th, double& out,
                                 double* kappaOut = nullptr) {
    const double h1 = in[0], h2 = in[1];
    const double KB = th[0], h0 = th[1], M = th[2], L = th[3], h0p = th[4], KBS0 = th[5], delta = th[6], Mp = th[7];
    if (kappaOut) *kappaOut = BAMGPU_UNDEF_RN;
    if (h1 - h2 - delta <= 0.0) return false;
    if (h1 - h0 <= 0.0 || h2 - h0p <= 0.0) { out = 0.0; return true; }
    const double x1 = fmax(fmax(h0, h0p), h2 + delta) + 0.001, x2 = p.modelOptD;
    double fl, fh, df, f, xl, xh;
    sfd_fnewton(x1, th, h2, fl, df);
    sfd_fnewton(x2, th, h2, fh, df);
    int fcalls = 2;
    double kappa;
    if (fl == 0.0) kappa = x1;
    else if (fh == 0.0) kappa = x2;
    else {
        if ((fl > 0.0) == (fh > 0.0) || fl != fl || fh != fh) return false;  // znewton err > 0: root not bracketed
        if (fl < 0.0) { xl = x1; xh = x2; } else { xl = x2; xh = x1; }
        double rts = 0.5 * (x1 + x2), dxold = fabs(x2 - x1), dx = dxold;
        sfd_fnewton(rts, th, h2, f, df);
        ++fcalls;
        for (int j = 0; j < p.vegItmax; ++j) {
            if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
                dxold = dx; dx = 0.5 * (xh - xl); rts = xl + dx;
            } else {
                dxold = dx; dx = f / df; rts = rts - dx;
            }
            const bool convx = fabs(dx) <= p.vegTolX * p.vegXscale;
            sfd_fnewton(rts, th, h2, f, df);
            ++fcalls;
            if (f < 0.0) xl = rts; else xh = rts;
            if (convx || fabs(f) <= p.vegTolF * p.vegFscale) break;
        }
        for (int j = 0; j < 3 && f != 0.0 && df != 0.0; ++j) {  // polish
            const double lo = fmin(xl, xh), hi = fmax(xl, xh);
            double xn = rts - f / df;
            xn = fmin(fmax(xn, lo), hi);
            if (xn == rts) break;
            rts = xn;
            sfd_fnewton(rts, th, h2, f, df);
            if (f < 0.0) xl = rts; else xh = rts;
        }
        kappa = rts;
    }
    if (fcalls > p.vegItmax || !(kappa == kappa)) return false;
    if (kappaOut) *kappaOut = kappa;
    if (h1 <= kappa) out = (KB * pow(h1 - h0, M)) * sqrt((h1 - h2 - delta) / L);  // variable-slope model
    else out = KBS0 * pow(h1 - h0p, Mp);                                       // standard channel model
    return true;
}

// -------------------------------------------------------------------------------------------
// Mixture (Mixture_model.f90:74-135): Y = sum_i w_{j,i} X_i per event j, the weights of an event in [0,1] and summing to
// one (the last one is derived); with a missing source, + w_last * theta(nEvt nS + e) for element e.  The reference
// builds output row (j-1) nElt + e from the e-th INPUT row of event j (pack order); the host pairs the rows once at
// creation (bam::mixture_rows), so that here the model is pointwise: x[0] = event, x[1] = element position within the
// event (both 1-based), x[2..] = source concentrations.  der[j] = last weight of event j (the derived parameters).
// -------------------------------------------------------------------------------------------
__device__ inline bool mixture_prepare(const DevProblem& p, const double* th, double* der) {
    const int nS = p.modelOpt[0], nEvt = p.modelOpt[1], missing = p.modelOpt[3];
    for (int j = 0; j < nEvt; ++j) {
        double s = 0.0, last;
        bool bad = false;
        if (missing) {
            for (int i = 0; i < nS; ++i) { const double w = th[j * nS + i]; s = s + w; bad |= (w < 0.0 || w > 1.0); }
            if (s > 1.0) return false;
            last = 1.0 - s;
        } else {
            for (int i = 0; i < nS - 1; ++i) { const double w = th[j * (nS - 1) + i]; s = s + w; bad |= (w < 0.0 || w > 1.0); }
            last = 1.0 - s;
            bad |= (last < 0.0 || last > 1.0);
        }
        if (bad) return false;
        der[j] = last;
    }
    return true;
}

template <int NX>
__device__ __forceinline__ bool mixture_apply(const DevProblem& p, const double* x, const double* __restrict__ th,
                                              const double* __restrict__ der, double& out) {
    constexpr int nS = NX - 2;
    const int nEvt = p.modelOpt[1], missing = p.modelOpt[3];
    const int j = (int)x[0] - 1, e = (int)x[1] - 1;
    const double last = der[j];
    const double* __restrict__ w = th + (missing ? j * nS : j * (nS - 1));
    double y = 0.0;
#pragma unroll
    for (int i = 0; i < nS; ++i) {
        const double wi = (missing || i < nS - 1) ? w[i] : last;
        y = y + x[2 + i] * wi;
    }
    if (missing) y = y + last * th[nEvt * nS + e];
    out = y;
    return true;
}

// -------------------------------------------------------------------------------------------
// SFDTidal_Qmec (SFDTidal_Qmec_model.f90:61-197), time-recursive: Q(m) = Q(m-1) + dt (pg + bf + ad), first-order step
// for the second point, second-order Adams-Bashforth afterwards.  One thread walks the whole series of one parameter
// vector; values of the last three points live in registers.  Explicitly rounded operations (no FMA contraction): the
// recursion feeds every rounding back into itself, so the operation sequence is the reference's (and the oracle's).
// out[(y * n + t) * K]: y = 0 discharge, 1..3 = dt pg, dt bf, dt ad (the three state variables).
// -------------------------------------------------------------------------------------------
// variant 2 = SFDTidal_Qmec2 (SFDTidal_Qmec2_model.f90:61-187): theta = (Be, bevs, delta, ne, d1, d2, c, g, Q0, dx, dt), the
// effective width and the Manning coefficient are parameters instead of being derived from (y0, A0, be, phi).
__device__ inline bool sfdtidal_qmec_series(int variant, int n, const double* __restrict__ h1, const double* __restrict__ h2,
                                            const double* __restrict__ th, double* __restrict__ out, size_t K, int nOut) {
    // variant 3 = SFDTidal_Qmec0 (SFDTidal_Qmec0_model.f90:61-178): theta = (Be, he, dzeta, ne, d1, d2, c, g, Q0, dx, dt), stages
    //   MINUS datums and Ae = Be (he + hm): variant 2 with be = -he and the datums negated (x - (-y) and x + (-y) are the exact
    //   sum / difference), its own feasibility conditions (:96-98), no state variables;
    // variant 4 = SFDTidal_QmecMS (SFDTidal_QmecMS_model.f90:62-154): theta = (y0, A0, be, delta, phi, d1, d2, c, g, dx); the
    //   Manning-Strickler discharge of every point on its own, sign from the water-surface slope (:149-152)
    double be, delta, d1, d2, c, g, Q0, dx, dt, width, ne;
    if (variant == 2 || variant == 3) {
        width = th[0]; be = th[1]; delta = th[2]; ne = th[3]; d1 = th[4]; d2 = th[5]; c = th[6]; g = th[7]; Q0 = th[8];
        dx = th[9]; dt = th[10];
        if (width <= 0.0 || ne <= 0.0 || c <= 0.0 || g <= 0.0 || dx <= 0.0 || dt <= 0.0) return false;
        if (variant == 3) {
            if (be <= 0.0 || delta >= 0.0) return false;  // he <= 0, dzeta >= 0
            be = -be; d1 = -d1; d2 = -d2;
        }
    } else {
        const double y0 = th[0], A0 = th[1], phi = th[4];
        be = th[2]; delta = th[3]; d1 = th[5]; d2 = th[6]; c = th[7]; g = th[8];
        if (variant == 4) { Q0 = 0.0; dx = th[9]; dt = 1.0; }
        else { Q0 = th[9]; dx = th[10]; dt = th[11]; }
        if (A0 <= 0.0 || (y0 - be) <= 0.0 || phi <= 0.0 || c <= 0.0 || g <= 0.0 || dx <= 0.0 || dt <= 0.0) return false;
        width = __ddiv_rn(A0, __dsub_rn(y0, be));
        const double R0 = __ddiv_rn(A0, __dadd_rn(width, __dmul_rn(2.0, __dsub_rn(y0, be))));
        ne = sqrt(__dmul_rn(__dmul_rn(__dmul_rn(phi, __ddiv_rn(1.0, __dmul_rn(8.0, g))), A0), pow(R0, c)));
    }
    const double ne2 = __dmul_rn(ne, ne);
    auto geom = [&](int m, double& dy, double& ym, double& Ae, double& Rhe) -> bool {
        const double y1 = __dadd_rn(h1[m], d1), y2 = __dadd_rn(h2[m], d2);
        dy = __dsub_rn(y2, y1);
        ym = __ddiv_rn(__dadd_rn(y1, y2), 2.0);
        Ae = __dmul_rn(width, __dsub_rn(ym, be));
        const double Pe = __dadd_rn(width, __dmul_rn(2.0, __dsub_rn(ym, be)));
        Rhe = __ddiv_rn(Ae, Pe);
        return !(Ae < 0.0 || Pe < 0.0 || Rhe < 0.0);
    };
    // The reference checks the geometry of the whole series before it runs anything (:131-133).  Here the check rides
    // along with the recursion: a vector with a bad point returns false at the end, and what it wrote is never read
    // (the likelihood, prediction and residual kernels skip flagged vectors).
    bool ok = true;
    auto put = [&](int m, double q, double s1, double s2, double s3) {
        out[(size_t)m * K] = q;
        if (nOut > 1) {
            out[((size_t)n + m) * K] = s1;
            out[((size_t)2 * n + m) * K] = s2;
            out[((size_t)3 * n + m) * K] = s3;
        }
    };
    if (variant == 4) {  // Q(m) = (1/ne) Rhe^(c/2) Ae sqrt(|(dy + delta) / dx|), negative when the surface rises downstream
        const double rne = __ddiv_rn(1.0, ne), hc = __dmul_rn(0.5, c);
        for (int m = 0; m < n; ++m) {
            double dy0, ym0, Ae0, Rh0;
            ok &= geom(m, dy0, ym0, Ae0, Rh0);
            const double sl = __dadd_rn(dy0, delta);
            double q = __dmul_rn(__dmul_rn(__dmul_rn(rne, pow(Rh0, hc)), Ae0), sqrt(fabs(__ddiv_rn(sl, dx))));
            if (sl > 0.0) q = -q;
            put(m, q, 0.0, 0.0, 0.0);
        }
        return ok;
    }
    // point m-1 (suffix 1); of point m-2 only ym is still needed (see below)
    double dy1 = 0.0, ym1 = 0.0, Ae1 = 0.0, Rh1 = 0.0, ym2 = 0.0, Q1 = Q0;
    if (n > 0) { ok &= geom(0, dy1, ym1, Ae1, Rh1); put(0, Q0, 0.0, 0.0, 0.0); }
    auto pgf = [&](double Ae, double dy) { return __ddiv_rn(__dmul_rn(__dmul_rn(-g, Ae), __dadd_rn(dy, delta)), dx); };
    auto bff = [&](double Rh, double Q, double Ae) {
        return __ddiv_rn(__dmul_rn(__dmul_rn(__dmul_rn(-g, __ddiv_rn(ne2, pow(Rh, c))), Q), fabs(Q)), Ae);
    };
    auto adf = [&](double Q, double Ae, double dydt) { return __dmul_rn(__dmul_rn(__dmul_rn(2.0, __ddiv_rn(Q, Ae)), width), dydt); };
    // The "m-2" terms of step m are the "m-1" terms of step m-1, operand for operand: pgf(Ae(m-2), dy(m-2)),
    // bff(Rhe(m-2), Q(m-2), Ae(m-2)) and adf(Q(m-2), Ae(m-2), dydt2) with dydt2(m) = dydt1(m-1) -- also at m = 2, 3, where
    // the reference switches between the one-sided and the centred difference (:160-172).  They are carried over
    // instead of being recomputed (one pow and ~10 divisions per step), bit for bit the same values.
    double pgP = 0.0, bfP = 0.0, adP = 0.0;  // pgf / bff / adf of the previous step's point m-1
    for (int m = 1; m < n; ++m) {
        double dy0, ym0, Ae0, Rh0;
        ok &= geom(m, dy0, ym0, Ae0, Rh0);
        double pg, bf, ad;
        const double pgC = pgf(Ae1, dy1), bfC = bff(Rh1, Q1, Ae1);
        double adC;
        if (m == 1) {
            const double dydt1 = __ddiv_rn(__dsub_rn(ym0, ym1), dt);
            adC = adf(Q1, Ae1, dydt1);
            pg = -__dmul_rn(g, __ddiv_rn(__dmul_rn(Ae1, __dadd_rn(dy1, delta)), dx));  // -g*(Ae*(dy+delta)/dx) (:165)
            bf = bfC;
            ad = adC;
        } else {
            const double dydt1 = __ddiv_rn(__dsub_rn(ym0, ym2), __dmul_rn(2.0, dt));
            adC = adf(Q1, Ae1, dydt1);
            pg = __dsub_rn(__dmul_rn(1.5, pgC), __dmul_rn(0.5, pgP));
            bf = __dsub_rn(__dmul_rn(1.5, bfC), __dmul_rn(0.5, bfP));
            ad = __dsub_rn(__dmul_rn(1.5, adC), __dmul_rn(0.5, adP));
        }
        const double Qm = __dadd_rn(Q1, __dmul_rn(dt, __dadd_rn(__dadd_rn(pg, bf), ad)));
        put(m, Qm, __dmul_rn(dt, pg), __dmul_rn(dt, bf), __dmul_rn(dt, ad));
        pgP = pgC; bfP = bfC; adP = adC;
        ym2 = ym1;
        dy1 = dy0; ym1 = ym0; Ae1 = Ae0; Rh1 = Rh0; Q1 = Qm;
    }
    return ok;
}

// -------------------------------------------------------------------------------------------
// AlgaeBiomass (AlgaeBiomass_model.f90:55-197), time-recursive: b(t) = b(t-1) + mu0 Fb Ft Fi Fn b dt - (Rsen + removal + Rhyd)
// (b - bmin) dt, clamped to [bmin, bmax].  Inputs (series order): temperature, irradiance, depth, turbidity, removal.
// Explicitly rounded operations in the reference's order (the recursion feeds every rounding back into itself); pow / exp
// are libdevice's (1-2 ulp from libm: the logistic recursion is contractive, the difference does not grow).
// out[(y * n + t) * K]: y = 0 biomass, 1 biomass / bmax, then the states Fb, Ft, Fi, Fn, Rhyd.
// -------------------------------------------------------------------------------------------
#define BAM_EPSRE 2.2204460492503131e-16 /* DMSL epsRe (un-vendored): declared = epsilon(1._mrk) */
// DYN: DynamicVegetation_Apply (DynamicVegetation_model.f90:74-191) = the same recursion on depth = max(stage - b1, 0)
// (:137-138) with the discharge of the vegetation-affected rating curve in front of the outputs (:147-189): rows Q,
// biomass, biomass / bmax, then the states.  A time step whose Newton search fails is an infeasible ROW (feas(t), :171-182),
// stored as NaN and turned back into a row flag by model_apply.
template <bool DYN>
__device__ inline bool algae_series(const DevProblem& p, int n, const double* __restrict__ X, const double* __restrict__ th,
                                    double* __restrict__ out, size_t K, int nOut) {
    const double dt = th[0], b0 = th[1], bmin = th[2], bmax = th[3], mu0 = th[4], Tmin = th[5], Topt = th[6], Tmax = th[7],
                 Iopt = th[8], katt = th[9], kcw = th[10], Rsen = th[11], g = th[12], rho = th[13], J = th[14], tau0 = th[15],
                 bHyd = th[16], cHyd = th[17];
    if (dt <= 0.0 || bmin < 0.0 || b0 < bmin || b0 > bmax || bmin >= bmax || mu0 < 0.0 || Tmin >= Tmax || Topt >= Tmax ||
        Topt <= Tmin || Iopt <= 0.0 || katt < 0.0 || kcw < 0.0 || Rsen < 0.0 || g < 0.0 || rho < 0.0 || J < 0.0 || tau0 <= 0.0 ||
        bHyd < 0.0 || cHyd < 0.0)
        return false;  // :140-146
    const double Texp = __ddiv_rn(__dsub_rn(Topt, Tmin), __dsub_rn(Tmax, Topt));
    const double *temp = X, *irr = X + n, *depthIn = X + 2 * (size_t)n, *turb = X + 3 * (size_t)n, *rem = X + 4 * (size_t)n;
    constexpr int Y0 = DYN ? 1 : 0, NYS = DYN ? 3 : 2;  // first biomass row, first state row
    const double B = DYN ? th[18] : 0.0, n1 = DYN ? th[19] : 0.0, b1 = DYN ? th[20] : 0.0, c1 = DYN ? th[21] : 0.0,
                 chi = DYN ? th[22] : 0.0, U_chi = DYN ? th[23] : 0.0;
    const double kQ0 = DYN ? (B * sqrt(J)) / n1 : 0.0;                       // S0 = theta(iS0) = the slope J of the biomass budget (:30,:131)
    const double kG2 = DYN ? pow(B, chi) * pow(U_chi, chi) : 0.0;           // parameter-only factor of gamma2
    double bt = b0;
    for (int t = 0; t < n; ++t) {
        double dep = depthIn[t];
        if (DYN) {
            dep = __dsub_rn(dep, b1);
            if (dep <= 0.0) dep = 0.0;
        }
        const double Fb = __dsub_rn(1.0, __ddiv_rn(bt, bmax));
        double Ft = 0.0;
        const double T = temp[t];
        if (!(T <= Tmin || T >= Tmax)) {
            Ft = __dmul_rn(__ddiv_rn(__dsub_rn(Tmax, T), __dsub_rn(Tmax, Topt)), pow(__ddiv_rn(__dsub_rn(T, Tmin), __dsub_rn(Topt, Tmin)), Texp));
            if (Ft <= BAM_EPSRE) Ft = 0.0;
        }
        const double iz = __dmul_rn(irr[t], exp(__dmul_rn(-__dadd_rn(__dmul_rn(katt, turb[t]), kcw), dep)));
        const double Ir = __ddiv_rn(iz, Iopt);
        const double Fi = __dmul_rn(Ir, exp(__dsub_rn(1.0, Ir)));
        const double Fn = 1.0;
        const double tau = __dmul_rn(__dmul_rn(__dmul_rn(dep, g), J), rho);
        const double Rh = (tau <= tau0) ? 0.0 : __dmul_rn(cHyd, pow(__ddiv_rn(__dsub_rn(tau, tau0), tau0), bHyd));
        const double grow = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(mu0, Fb), Ft), Fi), Fn), bt), dt);
        const double loss = __dmul_rn(__dmul_rn(__dadd_rn(__dadd_rn(Rsen, rem[t]), Rh), __dsub_rn(bt, bmin)), dt);
        double b = __dsub_rn(__dadd_rn(bt, grow), loss);
        b = fmin(b, bmax);
        b = fmax(b, bmin);
        bt = b;
        if (DYN) {  // discharge of this time step (:147-189)
            double Q;
            if (dep <= 0.0) Q = 0.0;
            else {
                const double Q0 = kQ0 * pow(dep, c1);
                if (b <= 0.0) Q = Q0;
                else {
                    const double gam = (b * cbrt(dep)) / (8.0 * g * (n1 * n1));
                    const double gam2 = kG2 * pow(dep, chi);
                    if (!veg_root(p, Q0, gam, gam2, chi, Q)) Q = CUDART_NAN;
                }
            }
            out[(size_t)t * K] = Q;
        }
        out[((size_t)Y0 * n + t) * K] = b;
        if (nOut > Y0 + 1) out[((size_t)(Y0 + 1) * n + t) * K] = __ddiv_rn(b, bmax);
        if (nOut > NYS) {
            out[((size_t)NYS * n + t) * K] = Fb;
            out[((size_t)(NYS + 1) * n + t) * K] = Ft;
            out[((size_t)(NYS + 2) * n + t) * K] = Fi;
            out[((size_t)(NYS + 3) * n + t) * K] = Fn;
            out[((size_t)(NYS + 4) * n + t) * K] = Rh;
        }
    }
    return true;
}

// SFDTidal / SFDTidal2 / SFDTidal4 / SFDTidalJones (variant 1..4): stage-fall-discharge curves for tidal rivers whose slope
// term reads the stages at the lagged step tlag = clamp(t - Dt, 1, n), Dt = int(theta1) (SFDTidal_model.f90:110-118 and the
// same lines of the other three).  Not recursive, but an output needs OTHER rows of the input series, so they run on
// series_run with the time-recursive models.  Products are rounded operation by operation in the reference's order.
__device__ inline bool sfdtidal_lag_series(int variant, int n, const double* __restrict__ X, const double* __restrict__ th,
                                           double* __restrict__ out, size_t K) {
    const double* h1 = X;
    const double* h2 = X + n;
    const double* dhdt = X + 2 * (size_t)n;  // Jones only
    const long long Dt = (long long)th[0];   // Fortran real -> integer assignment truncates
    auto lag = [&](int t) -> int {           // 0-based
        const long long q = (long long)t - Dt;
        return (q < 0) ? 0 : ((q > n - 1) ? n - 1 : (int)q);
    };
    auto sgn = [](double x) { return (x < 0.0 || (x == 0.0 && signbit(x))) ? -1.0 : 1.0; };  // sign(1._mrk, x)
    if (variant == 1) {  // Dt, K, B1, B2, h0, M, L, deltaz
        const double Kc = th[1], B1 = th[2], B2 = th[3], h0 = th[4], M = th[5], L = th[6], dz = th[7];
        if (Kc <= 0.0 || B1 <= 0.0 || B2 <= 0.0 || M <= 0.0 || L <= 0.0) return false;
        for (int t = 0; t < n; ++t) {
            const int tl = lag(t);
            const double fall = __dsub_rn(__dsub_rn(h1[tl], h2[tl]), dz);
            double q = __ddiv_rn(__dmul_rn(__dmul_rn(sgn(fall), Kc), __dadd_rn(B2, B1)), 2.0);
            q = __dmul_rn(q, sqrt(__ddiv_rn(fabs(fall), L)));
            q = __dmul_rn(q, pow(__dsub_rn(__ddiv_rn(__dadd_rn(h1[tl], h2[tl]), 2.0), h0), M));
            if (__dsub_rn(h1[t], h0) <= 0.0) q = 0.0;
            out[(size_t)t * K] = q;
        }
        return true;
    }
    const double Kc = th[1], B = th[2], h0 = th[3], M = th[4], L = th[5], delta = th[6];
    if (variant == 2) {  // ..., alpha in (0, 1)
        const double alpha = th[7];
        if (Kc <= 0.0 || B <= 0.0 || M <= 0.0 || L <= 0.0 || alpha <= 0.0 || alpha >= 1.0) return false;
        for (int t = 0; t < n; ++t) {
            double q = 0.0;
            if (__dsub_rn(h1[t], h0) > 0.0) {
                const double fall = __dsub_rn(__dsub_rn(h1[t], __dmul_rn(alpha, h2[lag(t)])), delta);
                q = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(sgn(fall), Kc), B), sqrt(__ddiv_rn(fabs(fall), L))), pow(__dsub_rn(h1[t], h0), M));
            }
            out[(size_t)t * K] = q;
        }
        return true;
    }
    if (variant == 3) {  // ..., P = int(theta8): half-width of the window the slope is averaged over
        const long long P = (long long)th[7];
        if (Kc <= 0.0 || B <= 0.0 || M <= 0.0 || L <= 0.0 || P < 0) return false;
        auto Sf = [&](int t) { const int tl = lag(t); return __ddiv_rn(__dsub_rn(__dsub_rn(h1[tl], h2[tl]), delta), L); };
        for (int t = 0; t < n; ++t) {
            long long t1 = t, t2 = t;
            if (!((long long)t - P < 0 || (long long)t - P > n - 1 || (long long)t + P < 0 || (long long)t + P > n - 1)) { t1 = t - P; t2 = t + P; }
            double sum = 0.0;
            for (long long u = t1; u <= t2; ++u) sum = __dadd_rn(sum, Sf((int)u));
            const double Sm = __ddiv_rn(sum, (double)(t2 - t1 + 1));
            double q = 0.0;
            if (__dsub_rn(h1[t], h0) > 0.0)
                q = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(sgn(Sm), Kc), B), sqrt(fabs(Sm))), pow(__dsub_rn(h1[t], h0), M));
            out[(size_t)t * K] = q;
        }
        return true;
    }
    {   // Jones: ..., alpha > 0; third input dh/dt
        const double alpha = th[7];
        if (Kc <= 0.0 || B <= 0.0 || M <= 0.0 || L <= 0.0 || alpha <= 0.0) return false;
        for (int t = 0; t < n; ++t) {
            double q = 0.0;
            if (__dsub_rn(h1[t], h0) > 0.0) {
                const int tl = lag(t);
                const double Sf = __ddiv_rn(__dsub_rn(__dsub_rn(h1[tl], h2[tl]), delta), L);
                const double rad = __dmul_rn(fabs(Sf), __dsub_rn(1.0, __dmul_rn(alpha, dhdt[t])));
                q = __dmul_rn(__dmul_rn(__dmul_rn(__dmul_rn(sgn(Sf), Kc), B), sqrt(rad)), pow(__dsub_rn(h1[t], h0), M));
            }
            out[(size_t)t * K] = q;
        }
        return true;
    }
}

// one sequential pass of a time-recursive model for one parameter vector; false = infeasible vector
__device__ inline bool model_series(const DevProblem& p, const double* __restrict__ th, double* __restrict__ out, size_t K,
                                    int nOut) {
    if (p.model == BAMGPU_MDL_SFDTIDAL_QMEC || p.model == BAMGPU_MDL_SFDTIDAL_QMEC2 || p.model == BAMGPU_MDL_SFDTIDAL_QMEC0 ||
        p.model == BAMGPU_MDL_SFDTIDAL_QMECMS)
        return sfdtidal_qmec_series(p.model == BAMGPU_MDL_SFDTIDAL_QMEC2 ? 2 : (p.model == BAMGPU_MDL_SFDTIDAL_QMEC0 ? 3 : (p.model == BAMGPU_MDL_SFDTIDAL_QMECMS ? 4 : 1)),
                                    p.seriesN, p.seriesX, p.seriesX + p.seriesN, th, out, K, nOut);
    if (p.model == BAMGPU_MDL_SFDTIDAL_LAG) return sfdtidal_lag_series(p.modelOpt[0], p.seriesN, p.seriesX, th, out, K);
    if (p.model == BAMGPU_MDL_ALGAE) return algae_series<false>(p, p.seriesN, p.seriesX, th, out, K, nOut);
    if (p.model == BAMGPU_MDL_DYNVEG) return algae_series<true>(p, p.seriesN, p.seriesX, th, out, K, nOut);
    return false;
}

// -------------------------------------------------------------------------------------------
// dispatch
// -------------------------------------------------------------------------------------------

// once per (parameter set, VAR combination): derived values + feasibility.  xc[j] = column j of the input series
// the parameter set will be applied to (nser rows): only whole-series models (Vegetation) look at it.
__device__ inline bool model_prepare(const DevProblem& p, const double* th, double* der, const double* const* xc,
                                     int nser) {
    if (p.model == BAMGPU_MDL_BARATIN) return baratin_prepare(p, th, der);
    if (p.model == BAMGPU_MDL_BARATINBAC) return baratinbac_prepare(p, th, der);
    if (p.model == BAMGPU_MDL_VEGETATION) return vegetation_prepare(p, th, der, xc, nser);
    if (p.model == BAMGPU_MDL_ORTHO) return ortho_prepare(th, der);
    if (p.model == BAMGPU_MDL_SEGMENTATION) return segmentation_prepare(p, th, der, xc[0], nser);
    if (p.model == BAMGPU_MDL_MIXTURE) return mixture_prepare(p, th, der);
    if (p.model == BAMGPU_MDL_SFD)  // StageFallDischarge_model.f90:141-143
        return !(th[0] <= 0.0 || th[2] <= 0.0 || th[3] <= 0.0 || th[5] <= 0.0 || th[7] <= 0.0);
    if (p.model == BAMGPU_MDL_HCS) return !(th[0] <= 0.0 || th[1] <= 0.0);  // HydraulicControl_section_model.f90:70-72
    // parameter-set feasibility of the pointwise second-wave models
    if (p.model == BAMGPU_MDL_SUSPENDEDLOAD) return !(th[0] <= 0.0 || th[2] <= 0.0 || th[4] <= 0.0);  // SuspendedLoad_model.f90:96-99
    if (p.model == BAMGPU_MDL_SWOT) return !(th[0] <= 0.0 || th[4] <= 0.0);                             // SWOT_model.f90:128-130
    if (p.model == BAMGPU_MDL_SGD) return !(th[0] <= 0.0 || th[4] <= 0.0 || th[1] <= 0.0 || th[2] <= 0.0);  // SGD :94-96
    return true;
}

// one observation.  y[NY]; returns row feasibility (ApplyModel's vfeas, ModelLibrary_tools.f90:419-432)
template <int MODEL, int NX, int NY>
__device__ __forceinline__ bool model_apply(const DevProblem& p, const double* x, const double* __restrict__ th,
                                            const double* __restrict__ der, double* y, FmTab fm = FmTab{0u, 0u}) {
    if (MODEL == BAMGPU_MDL_BARATIN) {
        return baratin_apply(p, x[0], th, der, y[0]);
    } else if (MODEL == BAMGPU_MDL_LINEAR) {
        linear_apply<NX, NY>(x, th, y);
        return true;
    } else if (MODEL == BAMGPU_MDL_SEDIMENT) {
        return sediment_apply<NX>(p, x, th, y[0]);
    } else if (MODEL == BAMGPU_MDL_VEGETATION) {
        return vegetation_apply<NX, NY>(p, x, th, der, y, fm);
    } else if (MODEL == BAMGPU_MDL_SUSPENDEDLOAD) {
        return suspendedload_apply(x[0], th, y[0]);
    } else if (MODEL == BAMGPU_MDL_SWOT) {
        return swot_apply<NX>(p, x, th, y[0]);
    } else if (MODEL == BAMGPU_MDL_SGD) {
        return sgd_apply(x, th, y[0]);
    } else if (MODEL == BAMGPU_MDL_RECESSION_H) {
        return recession_apply(x[0], th, y[0]);
    } else if (MODEL == BAMGPU_MDL_ORTHO) {
        return ortho_apply(p, x, th, der, y);
    } else if (MODEL == BAMGPU_MDL_BARATINBAC) {
        return baratinbac_apply(p, x[0], th, der, y[0]);
    } else if (MODEL == BAMGPU_MDL_SFD) {
        return sfd_apply(p, x, th, y[0]);
    } else if (MODEL == BAMGPU_MDL_HCS) {
        return hcs_apply(p, x[0], th, y[0]);
    } else if (MODEL == BAMGPU_MDL_SEGMENTATION) {
        return segmentation_apply(p, x[0], th, der, y[0]);
    } else if (MODEL == BAMGPU_MDL_MIXTURE) {
        return mixture_apply<NX>(p, x, th, der, y[0]);
    } else if (MODEL == BAMGPU_MDL_SFDTIDAL_QMEC || MODEL == BAMGPU_MDL_SFDTIDAL_QMEC2 || MODEL == BAMGPU_MDL_ALGAE ||
               MODEL == BAMGPU_MDL_DYNVEG || MODEL == BAMGPU_MDL_SFDTIDAL_QMEC0 || MODEL == BAMGPU_MDL_SFDTIDAL_QMECMS ||
               MODEL == BAMGPU_MDL_SFDTIDAL_LAG) {  // time-recursive: produced by series_run, looked up here
#pragma unroll
        for (int k = 0; k < NY; ++k)
            y[k] = p.seriesY[((size_t)k * p.seriesN + (size_t)x[0]) * p.seriesK + (size_t)der[0]];
        return MODEL != BAMGPU_MDL_DYNVEG || y[0] == y[0];  // DynamicVegetation: a failed Newton search is an infeasible row
    } else {
        bool ok = true;
#pragma unroll
        for (int k = 0; k < NY; ++k) ok = textfile_eval<NX>(p, k, x, th, y[k], fm) && ok;
        return ok;
    }
}

// state variables (ModelLibrary_tools.f90 XtraSetup: model%nState): only SFD has one on this path (kappa)
__host__ __device__ constexpr int model_nstate(int model) {
    return model == BAMGPU_MDL_SFD ? 1 : ((model == BAMGPU_MDL_SFDTIDAL_QMEC || model == BAMGPU_MDL_SFDTIDAL_QMEC2) ? 3 : ((model == BAMGPU_MDL_ALGAE || model == BAMGPU_MDL_DYNVEG) ? 5 : 0));
}

// model_apply + the state variables of the observation (st[model_nstate(MODEL)])
template <int MODEL, int NX, int NY>
__device__ __forceinline__ bool model_apply_s(const DevProblem& p, const double* x, const double* __restrict__ th,
                                              const double* __restrict__ der, double* y, double* st, FmTab fm = FmTab{0u, 0u}) {
    if (MODEL == BAMGPU_MDL_SFD) return sfd_apply(p, x, th, y[0], st);
    if (MODEL == BAMGPU_MDL_SFDTIDAL_QMEC || MODEL == BAMGPU_MDL_SFDTIDAL_QMEC2 || MODEL == BAMGPU_MDL_ALGAE || MODEL == BAMGPU_MDL_DYNVEG) {
#pragma unroll
        for (int k = 0; k < model_nstate(MODEL); ++k)
            st[k] = p.seriesY[((size_t)(NY + k) * p.seriesN + (size_t)x[0]) * p.seriesK + (size_t)der[0]];
    }
    return model_apply<MODEL, NX, NY>(p, x, th, der, y, fm);
}
