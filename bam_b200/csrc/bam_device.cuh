// bam_device.cuh -- device-side problem description and math shared by all kernels.
//
// DevProblem is the HBM-resident, immutable image of the reference's INFER + MODEL globals
// (src/BaM_tools.f90:73-101; ModelLibrary_tools.f90:114-126).  It is passed BY VALUE to every
// kernel (a few hundred bytes of kernel-parameter space); the big tables it points to are
// column-major exactly as Fortran holds them, so a warp reading 32 consecutive observations
// of one variable issues one 256-byte coalesced request.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/bam_gpu.h"

#define BAM_MAXX 8
#define BAM_MAXY 8
#define BAM_MAXTHETA 160  /* Mixture: nS*nEvt+nElt weights (tests/BaM_Mixture: 148) */
#define BAM_MAXCTRL 8
#define BAM_MAXDER 64  /* per-parameter-set derived values kept in the chain table (BaRatin: b_i and log a_i; Vegetation: 1+6 nYear) */
#define BAM_VEG_MAXYEAR 10
#define BAM_MAXGAMMA 24
#define BAM_MAXCOPULA 96
#define BAM_VMSTACK 24
#define BAM_LOGTAB_N 256  /* table-driven log: reciprocal table entries (bam_fastmath.cuh) */
#define BAM_EXPTAB_N 256  /* table-driven exp: 2^(j/N) entries */
#define BAM_LOG2_NB 16    /* joint (exponent, mantissa) log table of the BaRatin fast path: binades in the window */
#define BAM_LOG2_N (BAM_LOG2_NB * 256)
#define BAM_SCTAB_N 256   /* table-driven sin / cos of 2 pi u: (cos, sin) at the bin centres 2 pi (j + 0.5) / N */

// chain status bits (priority order of the reference's early exits, GetPostLogPdf :3310-3382)
#define ST_PRIOR_INFEAS 1
#define ST_PRIOR_NULL 2
#define ST_LKH_INFEAS 4
#define ST_HYPER_INFEAS 8

struct DevProblem {
    int model;
    int n, nX, nY, ntheta;
    int nFit, nGamma, nLatent, P, nCombo, nDparModel, nDpar;
    // per-chain table row: [gamma | xbias | ybias | nCombo x (ntheta full theta + nDer derived)]
    int tabStride, tabGamma, tabXb, tabYb, tabCombo, comboStride, nDer;
    const double *X, *Xu, *Xb, *Y, *Yu, *Yb;
    const int *Xbindx, *Ybindx;
    const int* comboOf;   // n : VAR-combination id of each observation
    const int* comboStart;  // nCombo+1 : observations are stored sorted by combination; combo v = [start[v], start[v+1])
    const double2* logTab;  // BAM_LOGTAB_N x (1/c_i, -log(1/c_i)), c_i = 1 + (i+0.5)/N : table-driven log (bam_fastmath.cuh)
    const double* expTab;   // BAM_EXPTAB_N x 2^(j/N)                                    : table-driven exp
    const double2* logTab2; // BAM_LOG2_N x (1/c, log c), c = 2^e (1 + (i+0.5)/256), slot (e mod 16, i): window [logBaseHi, +16 binades)
    const double* expTab2;  // BAM_EXPTAB_N x 2^(j/N) with (j << 12) subtracted from the high word (bam_baratin_cb2.cuh)
    int logBaseHi;          // high word of the window's lower bound 2^(E0-1023)
    const double2* scTab;   // BAM_SCTAB_N x (cos, sin)(2 pi (j + 0.5) / N)              : Box-Muller angle of the noise
    const int* comboSrc;  // nCombo x ntheta : position in the parameter vector, or -1 for FIX
    const int* latIdx;    // n x nX : position of the latent "true input" in the vector, or -1
    const double* XuInv;  // n x nX : 1 / Xu (0 where Xu = 0)                    } problems with latent inputs and no negative Xu,
    const double* hyC;    // n : sum over the cells with Xu > 0 of -log sqrt(2 pi) - log Xu  } else nullptr
    double fixv[BAM_MAXTHETA];
    int hasLatent, hasXbias, hasYbias, hyperInfeasible, hasMissing;
    int Xerror[BAM_MAXX], nXb[BAM_MAXX], offXb[BAM_MAXX], tabOffXb[BAM_MAXX];
    int nYb[BAM_MAXY], offYb[BAM_MAXY], tabOffYb[BAM_MAXY];
    int sigfunc[BAM_MAXY], nrem[BAM_MAXY], gamOff[BAM_MAXY];
    const int *ptDist, *pgDist;
    const double *ptPar, *pgPar;
    const double* priorM;
    int kM;
    double mv, unfeas;
    // BaRatin
    int ncontrol;
    unsigned long long ctrlMask;  // bit (r*8+c) = ControlMatrix(r+1,c+1)
    // Sediment
    int sedFormula, sedCss, sedCo, sedPpo[4];
    double sedPseudo[4];
    // TextFile
    const int* vmCode;
    const double* vmImmed;
    int vmNcode[BAM_MAXY], vmCodeOff[BAM_MAXY], vmImmedOff[BAM_MAXY], vmNin, vmNpar;
    // SWOT {variant}, Orthorectification {distortion id, npar}, Segmentation {nS, nmin, option}, Mixture {nS, nEvt, nElt, missing}
    int partPacked;  // BaRatin fast paths: 1 = one double per (tile, chain) partial (logpost_final_combo reads nothing else), 0 = {lkh, hyper} pairs
    int modelOpt[4];
    double modelOptD;
    const double* hcsTab;  // HydraulicControl_section: 4 x hcsN (h | A | w | h + A/(2w))
    int hcsN;
    // time-recursive models (row f4): the model output of every vector is produced by one sequential pass over the series
    // (series_run) into seriesY[(y * seriesN + t) * seriesK + vector]; for those models column 0 of X holds the row index t
    // and der[0] of a vector its own index, so the pointwise kernels just look the value up (model_apply)
    int isSeries, seriesN, seriesK;
    const double* seriesX;  // seriesN x nX, the real inputs in series order
    double* seriesY;
    // Vegetation
    int vegGrowth, vegNYear, vegItmax;
    double vegXscale, vegFscale, vegTolX, vegTolF;
};

#define BAM_LOG_SQRT_2PI 0.91893853320467274178

// ---------------------------------------------------------------------------------------------
// distributions (BMSL Distribution_tools; conventions declared in SURVEY.md App. E / DESIGN.md)
// ---------------------------------------------------------------------------------------------

// 1/v for a positive, normal v: MUFU.RCP64H seed + two Newton steps (relative error ~1e-16); replaces the
// IEEE division (~35 instructions with its slow path) in the likelihood tail
__device__ __forceinline__ double fast_rcp_pos(double v) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v));
    double e = fma(-v, r, 1.0);
    r = fma(r, e, r);
    e = fma(-v, r, 1.0);
    return fma(r, e, r);
}

// log N(y; mu, sqrt(v)) given the VARIANCE v > 0: -log sqrt(2 pi) - 0.5 log v - 0.5 (y-mu)^2 / v.
// Same quantity as gauss_logpdf(y, mu, sqrt(v)) without the square root and the division.
__device__ __forceinline__ double gauss_logpdf_var(double y, double mu, double v) {
    const double r = y - mu;
    const double t = fma(r * r, fast_rcp_pos(v), log(v));
    return fma(-0.5, t, -BAM_LOG_SQRT_2PI);
}

// log N(x; mu, s) ; the caller guarantees s > 0
__device__ __forceinline__ double gauss_logpdf(double x, double mu, double s) {
    double z = (x - mu) / s;
    return -BAM_LOG_SQRT_2PI - log(s) - 0.5 * z * z;
}

// returns 0 OK; sets feas/isnull like BMSL GetPdf(loga=.true.)
__device__ inline void dist_logpdf(int dist, const double* __restrict__ par, double x, double& lp, bool& feas,
                                   bool& isnull) {
    feas = true;
    isnull = false;
    lp = BAMGPU_UNDEF_RN;
    switch (dist) {
        case BAMGPU_DIST_GAUSSIAN: {
            double s = par[1];
            if (!(s > 0.0)) { feas = false; return; }
            lp = gauss_logpdf(x, par[0], s);
            return;
        }
        case BAMGPU_DIST_UNIFORM: {
            double a = par[0], b = par[1];
            if (!(b > a)) { feas = false; return; }
            if (x < a || x > b) { isnull = true; return; }
            lp = -log(b - a);
            return;
        }
        case BAMGPU_DIST_TRIANGLE: {
            double pk = par[0], a = par[1], b = par[2];
            if (!(b > a) || pk < a || pk > b) { feas = false; return; }
            if (x < a || x > b) { isnull = true; return; }
            double d;
            if (x <= pk) d = (pk > a) ? 2.0 * (x - a) / ((b - a) * (pk - a)) : 2.0 / (b - a);
            else d = (b > pk) ? 2.0 * (b - x) / ((b - a) * (b - pk)) : 2.0 / (b - a);
            if (!(d > 0.0)) { isnull = true; return; }
            lp = log(d);
            return;
        }
        case BAMGPU_DIST_LOGNORMAL: {
            double s = par[1];
            if (!(s > 0.0)) { feas = false; return; }
            if (!(x > 0.0)) { isnull = true; return; }
            double lx = log(x);
            double z = (lx - par[0]) / s;
            lp = -BAM_LOG_SQRT_2PI - log(s) - lx - 0.5 * z * z;
            return;
        }
        case BAMGPU_DIST_EXPONENTIAL: {
            double sc = par[1];
            if (!(sc > 0.0)) { feas = false; return; }
            if (x < par[0]) { isnull = true; return; }
            lp = -log(sc) - (x - par[0]) / sc;
            return;
        }
        case BAMGPU_DIST_FLATPRIOR: lp = 0.0; return;
        case BAMGPU_DIST_FLATPRIOR_POS:
            if (x > 0.0) lp = 0.0; else isnull = true;
            return;
        case BAMGPU_DIST_FLATPRIOR_NEG:
            if (x < 0.0) lp = 0.0; else isnull = true;
            return;
        default: feas = false; return;
    }
}

__device__ inline void dist_cdf(int dist, const double* __restrict__ par, double x, double& u, bool& feas) {
    feas = true;
    u = 0.0;
    switch (dist) {
        case BAMGPU_DIST_GAUSSIAN:
            if (!(par[1] > 0.0)) { feas = false; return; }
            u = normcdf((x - par[0]) / par[1]);
            return;
        case BAMGPU_DIST_UNIFORM:
            if (!(par[1] > par[0])) { feas = false; return; }
            if (x <= par[0]) u = 0.0;
            else if (x >= par[1]) u = 1.0;
            else u = (x - par[0]) / (par[1] - par[0]);
            return;
        case BAMGPU_DIST_TRIANGLE: {
            double pk = par[0], a = par[1], b = par[2];
            if (!(b > a) || pk < a || pk > b) { feas = false; return; }
            if (x <= a) u = 0.0;
            else if (x >= b) u = 1.0;
            else if (x <= pk) u = (x - a) * (x - a) / ((b - a) * (pk - a));
            else u = 1.0 - (b - x) * (b - x) / ((b - a) * (b - pk));
            return;
        }
        case BAMGPU_DIST_LOGNORMAL:
            if (!(par[1] > 0.0)) { feas = false; return; }
            if (!(x > 0.0)) u = 0.0;
            else u = normcdf((log(x) - par[0]) / par[1]);
            return;
        case BAMGPU_DIST_EXPONENTIAL:
            if (!(par[1] > 0.0)) { feas = false; return; }
            if (x <= par[0]) u = 0.0;
            else u = 1.0 - exp(-(x - par[0]) / par[1]);
            return;
        default: feas = false; return;
    }
}

// Sigmafunk_Apply, src/BaM_tools.f90:3521-3571
__device__ __forceinline__ double sigmafunk(int funk, const double* par, double Y) {
    double ay = fabs(Y);
    switch (funk) {
        case BAMGPU_SIG_CONSTANT: return par[0];
        case BAMGPU_SIG_LINEAR: return fma(par[1], ay, par[0]);
        case BAMGPU_SIG_PROPORTIONAL: return par[0] * ay;
        case BAMGPU_SIG_POWER: return par[0] + par[1] * pow(ay, par[2]);
        case BAMGPU_SIG_EXPONENTIAL: return par[0] + (par[2] - par[0]) * (1.0 - exp(-(ay / par[1])));
        default: {  // Gaussian
            double r = ay / par[1];
            return par[0] + (par[2] - par[0]) * (1.0 - exp(-(r * r)));
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller.  Same specification as the oracle (oracle/bam_oracle.c), written
// independently: counter = (a_lo, a_hi, b, c), key = (seed_lo, seed_hi).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

__device__ __forceinline__ double philox_uniform(uint64_t seed, uint64_t a, uint32_t b, uint32_t c) {
    uint32_t c0 = (uint32_t)a, c1 = (uint32_t)(a >> 32), c2 = b, c3 = c;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    return u53(c0, c1);
}

__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t a, uint32_t b, uint32_t c) {
    uint32_t c0 = (uint32_t)a, c1 = (uint32_t)(a >> 32), c2 = b, c3 = c;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    double u1 = u53(c0, c1), u2 = u53(c2, c3);
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

// Structural-error noise of the prediction path (specification in bam_fastmath.cuh): inputs 4q..4q+3 share one
// Philox block; words (0,1) serve inputs 4q, 4q+1 and words (2,3) inputs 4q+2, 4q+3; the even input takes the
// cosine, the odd one the sine branch of Box-Muller; uniforms are (w + 0.5) 2^-32.
__device__ __forceinline__ double philox_normal_pred(uint64_t seed, uint64_t rep, uint32_t t, uint32_t y) {
    uint32_t c0 = (uint32_t)rep, c1 = (uint32_t)(rep >> 32), c2 = t >> 2, c3 = y;
    philox4x32_10(c0, c1, c2, c3, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t wa = (t & 2u) ? c2 : c0, wb = (t & 2u) ? c3 : c1;
    const double u1 = ((double)wa + 0.5) * (1.0 / 4294967296.0), u2 = ((double)wb + 0.5) * (1.0 / 4294967296.0);
    const double r = sqrt(-2.0 * log(u1));
    return r * ((t & 1u) ? sinpi(2.0 * u2) : cospi(2.0 * u2));
}
