/*
 * bam_oracle.c -- CPU restatement (ORACLE) of BaM's hot path.  TEST INFRASTRUCTURE ONLY
 * (see bam_oracle.h for who may use it and for the parity status).
 *
 * Plain C99, IEEE double, libm.  The loop order, the early exits and the order in which terms
 * are accumulated follow the reference (i-major / j-minor likelihood sum, prior -> likelihood
 * -> hyper, j-major hyper sum) so that this file is a line-by-line restatement of
 *   src/BaM_tools.f90:3006-3571 (posterior), :4115-4177 (theta expansion / VAR), :4202-4242 and
 *   :3386-3463 (vector layout), :882-1135 and :1292-1432 (prediction / envelopes),
 *   src/ModelLibrary_tools.f90:237-434 (dispatch + infeasible-row flagging),
 *   src/Models/{BaRatin,Linear,SedimentTransport,TextFile,Vegetation}_model.f90.
 * Compile with -ffp-contract=off: the reference is built by gfortran without FMA contraction.
 */
#include "bam_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define MAXCONTROL 16
#define MAXVMCODE 512
#define MAXVMIMMED 128
#define MAXVMOUT 8
#define MAXVMSTACK 64

static const double LOG_SQRT_2PI = 0.91893853320467274178; /* 0.5*log(2*pi) */
static const double SQRT_2PI = 2.50662827463100050242;

typedef struct {
    int ncode, nimmed, stacksize;
    int32_t code[MAXVMCODE];
    double immed[MAXVMIMMED];
} vm_prog;

struct oracle_handle {
    int model;
    int n, nX, nY, ntheta;
    double *X, *Xu, *Xb, *Y, *Yu, *Yb;
    int32_t *Xbindx, *Ybindx;
    int32_t *partype, *nval, *theta_off; /* theta_off: 0-based offset in the fitted vector, -1 for FIX */
    double* fix_value;
    int32_t* var_indx; /* n x ntheta, 1-based; all ones if no VAR */
    int nVar;
    int32_t *pt_dist, *pg_dist;
    double *pt_par, *pg_par;
    int32_t *sigma_func, *nremnant, *gamma_off; /* gamma_off: offset inside the gamma block */
    double* priorM;
    int kM;
    int32_t* xtra_i;
    int n_xtra_i;
    double* xtra_d;
    int n_xtra_d;
    double mv, unfeas;
    /* derived */
    int nFit, nGamma, nLatent, nXbTot, nYbTot, P, nDparModel, nDpar, nCombo, nState;
    int32_t *nXb, *nYb, *off_xb, *off_yb; /* offsets in the full vector */
    int32_t* Xerror;
    int32_t* latIdx;   /* n x nX: position in the full vector or -1 */
    int32_t* comboOf;  /* n: combo id of each observation */
    int32_t* comboRow; /* nCombo: first observation (0-based) of each combo = INFER%DparIndx - 1 */
    /* model specifics */
    int ncontrol;
    int32_t M[MAXCONTROL * MAXCONTROL]; /* row-major M[r*ncontrol+c] */
    int sed_formula, sed_css, sed_co, sed_ppo[4];
    double sed_pseudo[4];
    int vm_nin, vm_npar, vm_nout;
    vm_prog vm[MAXVMOUT];
    int model_opt[4];  /* SWOT {variant}; Orthorectification {distortion id, npar}; Segmentation {nS, nmin, option} */
    double model_optd; /* Segmentation tmin */
    int hcs_n;         /* HydraulicControl_section: rows of the h-A-w table (xtra_d, p x 3) */
    int veg_growth, veg_nyear, veg_itmax;          /* Vegetation xtra (Vegetation_XtraRead :244-289) */
    double veg_xscale, veg_fscale, veg_tolx, veg_tolf;
};

static void setmess(char* mess, const char* s) {
    if (!mess) return;
    strncpy(mess, s, BAMGPU_MESS_LEN - 1);
    mess[BAMGPU_MESS_LEN - 1] = 0;
}

static void* dup_mem(const void* src, size_t bytes) {
    void* p = malloc(bytes ? bytes : 1);
    if (src && bytes) memcpy(p, src, bytes);
    return p;
}

/* ------------------------------------------------------------------------------------------
 * Distributions (BMSL Distribution_tools -- un-vendored; conventions as declared in SURVEY App. E)
 * ------------------------------------------------------------------------------------------ */

static double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

/* Phi^-1: Abramowitz-Stegun 26.2.23 start, Halley iterations on erfc to full double precision. */
double oracle_normal_quantile(double p) {
    if (!(p > 0.0 && p < 1.0)) return NAN;
    int upper = p > 0.5;
    double q = upper ? 1.0 - p : p; /* exact for p in [0.5,1] */
    if (q == 0.5) return 0.0;
    double t = sqrt(-2.0 * log(q));
    double z = -(t - (2.515517 + t * (0.802853 + t * 0.010328)) /
                         (1.0 + t * (1.432788 + t * (0.189269 + t * 0.001308))));
    for (int it = 0; it < 12; ++it) {
        double e = norm_cdf(z) - q;
        double u = e * SQRT_2PI * exp(0.5 * z * z);
        double dz = u / (1.0 + 0.5 * z * u);
        z -= dz;
        if (fabs(dz) <= 1e-16 * fabs(z)) break;
    }
    return upper ? -z : z;
}

int oracle_logpdf(int dist, const double* par, double x, double* lp, int* feas, int* isnull) {
    *feas = 1;
    *isnull = 0;
    *lp = BAMGPU_UNDEF_RN;
    switch (dist) {
        case BAMGPU_DIST_GAUSSIAN: {
            double mu = par[0], s = par[1];
            if (!(s > 0.0)) { *feas = 0; return 0; }
            double z = (x - mu) / s;
            *lp = -LOG_SQRT_2PI - log(s) - 0.5 * z * z;
            return 0;
        }
        case BAMGPU_DIST_UNIFORM: {
            double a = par[0], b = par[1];
            if (!(b > a)) { *feas = 0; return 0; }
            if (x < a || x > b) { *isnull = 1; return 0; }
            *lp = -log(b - a);
            return 0;
        }
        case BAMGPU_DIST_TRIANGLE: {
            double pk = par[0], a = par[1], b = par[2];
            if (!(b > a) || pk < a || pk > b) { *feas = 0; return 0; }
            if (x < a || x > b) { *isnull = 1; return 0; }
            double d;
            if (x <= pk) d = (pk > a) ? 2.0 * (x - a) / ((b - a) * (pk - a)) : 2.0 / (b - a);
            else d = (b > pk) ? 2.0 * (b - x) / ((b - a) * (b - pk)) : 2.0 / (b - a);
            if (!(d > 0.0)) { *isnull = 1; return 0; }
            *lp = log(d);
            return 0;
        }
        case BAMGPU_DIST_LOGNORMAL: {
            double mu = par[0], s = par[1];
            if (!(s > 0.0)) { *feas = 0; return 0; }
            if (!(x > 0.0)) { *isnull = 1; return 0; }
            double lx = log(x);
            double z = (lx - mu) / s;
            *lp = -LOG_SQRT_2PI - log(s) - lx - 0.5 * z * z;
            return 0;
        }
        case BAMGPU_DIST_EXPONENTIAL: {
            double x0 = par[0], sc = par[1];
            if (!(sc > 0.0)) { *feas = 0; return 0; }
            if (x < x0) { *isnull = 1; return 0; }
            *lp = -log(sc) - (x - x0) / sc;
            return 0;
        }
        case BAMGPU_DIST_FLATPRIOR: *lp = 0.0; return 0;
        case BAMGPU_DIST_FLATPRIOR_POS:
            if (x > 0.0) *lp = 0.0; else *isnull = 1;
            return 0;
        case BAMGPU_DIST_FLATPRIOR_NEG:
            if (x < 0.0) *lp = 0.0; else *isnull = 1;
            return 0;
        default: *feas = 0; return 1;
    }
}

int oracle_cdf(int dist, const double* par, double x, double* u, int* feas) {
    *feas = 1;
    *u = BAMGPU_UNDEF_RN;
    switch (dist) {
        case BAMGPU_DIST_GAUSSIAN:
            if (!(par[1] > 0.0)) { *feas = 0; return 0; }
            *u = norm_cdf((x - par[0]) / par[1]);
            return 0;
        case BAMGPU_DIST_UNIFORM:
            if (!(par[1] > par[0])) { *feas = 0; return 0; }
            if (x <= par[0]) *u = 0.0;
            else if (x >= par[1]) *u = 1.0;
            else *u = (x - par[0]) / (par[1] - par[0]);
            return 0;
        case BAMGPU_DIST_TRIANGLE: {
            double pk = par[0], a = par[1], b = par[2];
            if (!(b > a) || pk < a || pk > b) { *feas = 0; return 0; }
            if (x <= a) *u = 0.0;
            else if (x >= b) *u = 1.0;
            else if (x <= pk) *u = (x - a) * (x - a) / ((b - a) * (pk - a));
            else *u = 1.0 - (b - x) * (b - x) / ((b - a) * (b - pk));
            return 0;
        }
        case BAMGPU_DIST_LOGNORMAL:
            if (!(par[1] > 0.0)) { *feas = 0; return 0; }
            if (!(x > 0.0)) *u = 0.0;
            else *u = norm_cdf((log(x) - par[0]) / par[1]);
            return 0;
        case BAMGPU_DIST_EXPONENTIAL:
            if (!(par[1] > 0.0)) { *feas = 0; return 0; }
            if (x <= par[0]) *u = 0.0;
            else *u = 1.0 - exp(-(x - par[0]) / par[1]);
            return 0;
        default: /* flat priors have no cdf: infeasible in a copula */
            *feas = 0;
            return 0;
    }
}

/* Gcop_getV, src/BaM_tools.f90:4246-4269: v = qnorm(cdf(x)); quantile failure -> infeasible */
static int gcop_getv(int dist, const double* par, double x, double* v, int* feas) {
    double u;
    int err = oracle_cdf(dist, par, x, &u, feas);
    if (err > 0) { *feas = 0; return err; }
    if (!*feas) return 0;
    if (!(u > 0.0 && u < 1.0)) { *feas = 0; return 0; }
    *v = oracle_normal_quantile(u);
    return 0;
}

/* Sigmafunk_Apply, src/BaM_tools.f90:3521-3571 */
static double sigmafunk(int funk, const double* par, double Y) {
    double ay = fabs(Y);
    switch (funk) {
        case BAMGPU_SIG_CONSTANT: return par[0];
        case BAMGPU_SIG_LINEAR: return par[0] + par[1] * ay;
        case BAMGPU_SIG_PROPORTIONAL: return par[0] * ay;
        case BAMGPU_SIG_POWER: return par[0] + par[1] * pow(ay, par[2]);
        case BAMGPU_SIG_EXPONENTIAL: return par[0] + (par[2] - par[0]) * (1.0 - exp(-(ay / par[1])));
        case BAMGPU_SIG_GAUSSIAN: {
            double r = ay / par[1];
            return par[0] + (par[2] - par[0]) * (1.0 - exp(-(r * r)));
        }
    }
    return NAN;
}

static int sigmafunk_npar(int funk) {
    switch (funk) {
        case BAMGPU_SIG_CONSTANT: return 1;
        case BAMGPU_SIG_LINEAR: return 2;
        case BAMGPU_SIG_PROPORTIONAL: return 1;
        case BAMGPU_SIG_POWER: return 3;
        case BAMGPU_SIG_EXPONENTIAL: return 3;
        case BAMGPU_SIG_GAUSSIAN: return 3;
    }
    return -1;
}

/* ------------------------------------------------------------------------------------------
 * TextFile model: formula compiler + stack VM (TextFile_model.f90:208-309, 551-857)
 * ------------------------------------------------------------------------------------------ */

enum { cImmed = 1, cNeg, cAdd, cSub, cMul, cDiv, cPow, cAbs, cExp, cLog10, cLog, cSqrt, cSinh, cCosh, cTanh,
       cSin, cCos, cTan, cAsin, cAcos, cAtan, cIs0, cIsPos, cIsNeg, cIsSPos, cIsSNeg, VarBegin };
static const char* const FUNCS[] = {"abs", "exp", "log10", "log", "sqrt", "sinh", "cosh", "tanh", "sin", "cos",
                                    "tan", "asin", "acos", "atan", "is0", "ispos", "isneg", "isspos", "issneg"};
#define NFUNCS 19
static const char OPS[] = "+-*/^";

typedef struct {
    const char* F; /* condensed formula */
    const char* const* var;
    int nvar;
    int32_t* code;
    int maxcode, ncode;
    double* immed;
    int maximmed, nimmed;
    int sp, stacksize;
    int err;
} vm_comp;

/* MathFunctionIndex (:463-482): first catalogue entry that is a case-insensitive prefix of s[0..len) */
static int math_function_index(const char* s, int len) {
    for (int j = 0; j < NFUNCS; ++j) {
        int k = (int)strlen(FUNCS[j]);
        if (k > len) continue; /* the reference compares a blank-padded, shorter string: never equal */
        int ok = 1;
        for (int q = 0; q < k; ++q)
            if (tolower((unsigned char)s[q]) != FUNCS[j][q]) { ok = 0; break; }
        if (ok) return cAbs + j;
    }
    return 0;
}

/* VariableIndex (:484-515): name runs until one of "+-*^/) " */
static int variable_index(const vm_comp* c, const char* s, int len) {
    int in = 0;
    while (in < len && !strchr("+-*/^) ", s[in])) ++in;
    for (int j = 0; j < c->nvar; ++j)
        if ((int)strlen(c->var[j]) == in && strncmp(s, c->var[j], in) == 0) return j + 1;
    return 0;
}

static int completely_enclosed(const char* F, int b, int e) { /* :621-642, 0-based inclusive */
    if (b > e || F[b] != '(' || F[e] != ')') return 0;
    int k = 0;
    for (int j = b + 1; j <= e - 1; ++j) {
        if (F[j] == '(') ++k;
        else if (F[j] == ')') --k;
        if (k < 0) break;
    }
    return k == 0;
}

static int is_binary_op(const char* F, int j) { /* :740-778 */
    if (F[j] == '+' || F[j] == '-') {
        if (j == 0) return 0;
        if (strchr("+-*/^(", F[j - 1])) return 0;
        if (isdigit((unsigned char)F[j + 1]) && strchr("eEdD", F[j - 1])) {
            int Dflag = 0, Pflag = 0;
            int k = j - 1;
            while (k > 0) {
                --k;
                if (isdigit((unsigned char)F[k])) Dflag = 1;
                else if (F[k] == '.') {
                    if (Pflag) break;
                    Pflag = 1;
                } else break;
            }
            if (Dflag && (k == 0 || strchr("+-*/^(", F[k]))) return 0;
        }
    }
    return 1;
}

static void add_byte(vm_comp* c, int32_t b) {
    if (c->ncode >= c->maxcode) { c->err = 1; return; }
    c->code[c->ncode++] = b;
}

/* RealNum (:780-853): [+|-][nnn][.nnn][e|E|d|D[+|-]nnn] at the start of s */
static double real_num(const char* s, int len, int* ok) {
    char buf[128];
    int in = 0, Bflag = 1, InMan = 0, Pflag = 0, Eflag = 0, InExp = 0, DInMan = 0, DInExp = 0;
    while (in < len) {
        char ch = s[in];
        if (ch == '+' || ch == '-') {
            if (Bflag) { InMan = 1; Bflag = 0; }
            else if (Eflag) { InExp = 1; Eflag = 0; }
            else break;
        } else if (isdigit((unsigned char)ch)) {
            if (Bflag) { InMan = 1; Bflag = 0; }
            else if (Eflag) { InExp = 1; Eflag = 0; }
            if (InMan) DInMan = 1;
            if (InExp) DInExp = 1;
        } else if (ch == '.') {
            if (Bflag) { Pflag = 1; InMan = 1; Bflag = 0; }
            else if (InMan && !Pflag) Pflag = 1;
            else break;
        } else if (ch == 'e' || ch == 'E' || ch == 'd' || ch == 'D') {
            if (InMan) { Eflag = 1; InMan = 0; }
            else break;
        } else break;
        ++in;
    }
    int bad = (in < 1) || !DInMan || ((Eflag || InExp) && !DInExp) || in >= (int)sizeof(buf);
    if (bad) { *ok = 0; return 0.0; }
    for (int q = 0; q < in; ++q) buf[q] = (s[q] == 'd' || s[q] == 'D') ? 'e' : s[q];
    buf[in] = 0;
    *ok = 1;
    return strtod(buf, NULL);
}

static void compile_substr(vm_comp* c, int b, int e) { /* CompileSubstr :644-738 (0-based inclusive) */
    const char* F = c->F;
    if (c->err || b > e) { c->err = 1; return; }
    if (F[b] == '+') { compile_substr(c, b + 1, e); return; }              /* case 1 */
    if (completely_enclosed(F, b, e)) { compile_substr(c, b + 1, e - 1); return; } /* case 2 */
    if (isalpha((unsigned char)F[b])) {
        int n = math_function_index(F + b, e - b + 1);
        if (n > 0) {
            const char* par = memchr(F + b, '(', e - b + 1);
            if (par) {
                int b2 = (int)(par - F);
                if (completely_enclosed(F, b2, e)) { /* case 3 */
                    compile_substr(c, b2 + 1, e - 1);
                    add_byte(c, n);
                    return;
                }
            }
        }
    } else if (F[b] == '-') {
        if (completely_enclosed(F, b + 1, e)) { /* case 4 */
            compile_substr(c, b + 2, e - 1);
            add_byte(c, cNeg);
            return;
        } else if (b + 1 <= e && isalpha((unsigned char)F[b + 1])) {
            int n = math_function_index(F + b + 1, e - b);
            if (n > 0) {
                const char* par = memchr(F + b + 1, '(', e - b);
                if (par) {
                    int b2 = (int)(par - F);
                    if (completely_enclosed(F, b2, e)) { /* case 5 */
                        compile_substr(c, b2 + 1, e - 1);
                        add_byte(c, n);
                        add_byte(c, cNeg);
                        return;
                    }
                }
            }
        }
    }
    for (int io = 0; io < 5; ++io) { /* increasing priority + - * / ^ ; right-most occurrence first */
        int k = 0;
        for (int j = e; j >= b; --j) {
            if (F[j] == ')') ++k;
            else if (F[j] == '(') --k;
            if (k == 0 && F[j] == OPS[io] && is_binary_op(F, j)) {
                if (io >= 2 && F[b] == '-') { /* case 6 */
                    compile_substr(c, b + 1, e);
                    add_byte(c, cNeg);
                    return;
                }
                compile_substr(c, b, j - 1); /* case 7 */
                compile_substr(c, j + 1, e);
                add_byte(c, cAdd + io);
                c->sp -= 1;
                return;
            }
        }
    }
    int b2 = b;
    if (F[b] == '-') ++b2;
    int n = 0;
    if (b2 <= e && (isdigit((unsigned char)F[b2]) || F[b2] == '.')) {
        int ok;
        double r = real_num(F + b2, e - b2 + 1, &ok);
        if (!ok || c->nimmed >= c->maximmed) { c->err = 1; return; }
        c->immed[c->nimmed++] = r;
        n = cImmed;
    } else if (b2 <= e) {
        n = variable_index(c, F + b2, e - b2 + 1);
        if (n > 0) n = VarBegin + n - 1;
    }
    if (n == 0) { c->err = 1; return; }
    add_byte(c, n);
    c->sp += 1;
    if (c->sp > c->stacksize) c->stacksize += 1;
    if (b2 > b) add_byte(c, cNeg);
}

int oracle_textfile_compile(const char* formula, const char* const* varnames, int nvar, int32_t* code,
                            int maxcode, double* immed, int maximmed, int* nimmed, int* stacksize) {
    /* parsef (:208-229): '**' -> '^', strip blanks */
    size_t L = strlen(formula);
    char* F = (char*)malloc(L + 1);
    size_t m = 0;
    for (size_t i = 0; i < L; ++i) {
        if (formula[i] == '*' && i + 1 < L && formula[i + 1] == '*') { F[m++] = '^'; ++i; continue; }
        if (formula[i] == ' ' || formula[i] == '\t' || formula[i] == '\r' || formula[i] == '\n') continue;
        F[m++] = formula[i];
    }
    F[m] = 0;
    if (m == 0) { free(F); return -1; }
    int par = 0;
    for (size_t i = 0; i < m; ++i) { /* minimal syntax check: parentheses */
        if (F[i] == '(') ++par;
        if (F[i] == ')') --par;
        if (par < 0) { free(F); return -1; }
    }
    if (par != 0) { free(F); return -1; }
    vm_comp c;
    memset(&c, 0, sizeof c);
    c.F = F; c.var = varnames; c.nvar = nvar; c.code = code; c.maxcode = maxcode;
    c.immed = immed; c.maximmed = maximmed;
    compile_substr(&c, 0, (int)m - 1);
    free(F);
    if (c.err) return -1;
    *nimmed = c.nimmed;
    *stacksize = c.stacksize;
    return c.ncode;
}

/* evalf (:231-309).  Returns error type (0 OK). */
static int vm_eval(const vm_prog* p, const double* val, double* res) {
    double st[MAXVMSTACK];
    int sp = -1, dp = 0;
    for (int ip = 0; ip < p->ncode; ++ip) {
        int op = p->code[ip];
        switch (op) {
            case cImmed: st[++sp] = p->immed[dp++]; break;
            case cNeg: st[sp] = -st[sp]; break;
            case cAdd: st[sp - 1] = st[sp - 1] + st[sp]; --sp; break;
            case cSub: st[sp - 1] = st[sp - 1] - st[sp]; --sp; break;
            case cMul: st[sp - 1] = st[sp - 1] * st[sp]; --sp; break;
            case cDiv:
                if (st[sp] == 0.0) { *res = 0.0; return 1; }
                st[sp - 1] = st[sp - 1] / st[sp]; --sp; break;
            case cPow: st[sp - 1] = pow(st[sp - 1], st[sp]); --sp; break;
            case cAbs: st[sp] = fabs(st[sp]); break;
            case cExp: st[sp] = exp(st[sp]); break;
            case cLog10:
                if (st[sp] <= 0.0) { *res = 0.0; return 3; }
                st[sp] = log10(st[sp]); break;
            case cLog:
                if (st[sp] <= 0.0) { *res = 0.0; return 3; }
                st[sp] = log(st[sp]); break;
            case cSqrt:
                if (st[sp] < 0.0) { *res = 0.0; return 3; }
                st[sp] = sqrt(st[sp]); break;
            case cSinh: st[sp] = sinh(st[sp]); break;
            case cCosh: st[sp] = cosh(st[sp]); break;
            case cTanh: st[sp] = tanh(st[sp]); break;
            case cSin: st[sp] = sin(st[sp]); break;
            case cCos: st[sp] = cos(st[sp]); break;
            case cTan: st[sp] = tan(st[sp]); break;
            case cAsin:
                if (st[sp] < -1.0 || st[sp] > 1.0) { *res = 0.0; return 4; }
                st[sp] = asin(st[sp]); break;
            case cAcos:
                if (st[sp] < -1.0 || st[sp] > 1.0) { *res = 0.0; return 4; }
                st[sp] = acos(st[sp]); break;
            case cAtan: st[sp] = atan(st[sp]); break;
            case cIs0: st[sp] = (st[sp] == 0.0) ? 1.0 : 0.0; break;
            case cIsPos: st[sp] = (st[sp] >= 0.0) ? 1.0 : 0.0; break;
            case cIsNeg: st[sp] = (st[sp] <= 0.0) ? 1.0 : 0.0; break;
            case cIsSPos: st[sp] = (st[sp] > 0.0) ? 1.0 : 0.0; break;
            case cIsSNeg: st[sp] = (st[sp] < 0.0) ? 1.0 : 0.0; break;
            default: st[++sp] = val[op - VarBegin]; break;
        }
    }
    *res = st[0];
    return 0;
}

/* split one list-directed record into items (comma / blank separated, quotes stripped, '!' ends it) */
static int split_items(const char* line, char items[][64], int maxitems) {
    int n = 0;
    const char* s = line;
    while (*s && n < maxitems) {
        while (*s == ' ' || *s == '\t' || *s == ',' || *s == '\r') ++s;
        if (!*s || *s == '\n' || *s == '!') break;
        int m = 0;
        if (*s == '"' || *s == '\'') {
            char q = *s++;
            while (*s && *s != q && m < 63) items[n][m++] = *s++;
            if (*s == q) ++s;
        } else {
            while (*s && !strchr(" \t,\r\n!", *s) && m < 63) items[n][m++] = *s++;
        }
        items[n][m] = 0;
        ++n;
    }
    return n;
}

static const char* next_line(const char* s, char* buf, int bufsz) {
    if (!s || !*s) return NULL;
    int m = 0;
    while (*s && *s != '\n') {
        if (m < bufsz - 1) buf[m++] = *s;
        ++s;
    }
    buf[m] = 0;
    if (*s == '\n') ++s;
    return s;
}

/* TxtMdl_load + TxtMdl_define (TextFile_model.f90:897-1033) from the text of the model file */
static int textfile_load(oracle_t* o, const char* text, char* mess) {
    char line[4096];
    char items[64][64];
    char names[64][64];
    const char* vp[64];
    const char* s = text;
    if (!s) { setmess(mess, "oracle: TextFile model needs xtra_text"); return 1; }
    s = next_line(s, line, sizeof line);
    if (!s) goto bad;
    int nIN = atoi(line);
    s = next_line(s, line, sizeof line);
    if (!s) goto bad;
    int nv = 0;
    if (nIN > 0) {
        int k = split_items(line, items, 64);
        if (k < nIN) goto bad;
        for (int i = 0; i < nIN; ++i) strcpy(names[nv++], items[i]);
    }
    s = next_line(s, line, sizeof line);
    if (!s) goto bad;
    int nPAR = atoi(line);
    s = next_line(s, line, sizeof line);
    if (!s) goto bad;
    if (nPAR > 0) {
        int k = split_items(line, items, 64);
        if (k < nPAR) goto bad;
        for (int i = 0; i < nPAR; ++i) strcpy(names[nv++], items[i]);
    }
    s = next_line(s, line, sizeof line);
    if (!s) goto bad;
    int nOUT = atoi(line);
    if (nOUT <= 0 || nOUT > MAXVMOUT) goto bad;
    for (int i = 0; i < nv; ++i) vp[i] = names[i];
    o->vm_nin = nIN; o->vm_npar = nPAR; o->vm_nout = nOUT;
    for (int k = 0; k < nOUT; ++k) {
        s = next_line(s, line, sizeof line);
        if (!s && k < nOUT - 1) goto bad;
        char* excl = strchr(line, '!');
        if (excl) *excl = 0;
        vm_prog* p = &o->vm[k];
        p->ncode = oracle_textfile_compile(line, vp, nv, p->code, MAXVMCODE, p->immed, MAXVMIMMED, &p->nimmed,
                                           &p->stacksize);
        if (p->ncode < 0 || p->stacksize > MAXVMSTACK) {
            setmess(mess, "oracle: TxtMdl_define: Syntax error");
            return 5;
        }
        if (!s) break;
    }
    return 0;
bad:
    setmess(mess, "oracle: TxtMdl_load: problem reading model text");
    return 1;
}

/* ------------------------------------------------------------------------------------------
 * Models.  X is n x nX column-major with leading dimension ldx.  Y n x nY, leading dim ldy.
 * vfeas[n] row feasibility.  Returns err.
 * ------------------------------------------------------------------------------------------ */

/* ApplyContinuity, BaRatin_model.f90:330-396 */
static int baratin_continuity(const oracle_t* o, const double* a, const double* c, const double* k, double* b) {
    int nc = o->ncontrol;
    if (a[0] <= 0.0 || c[0] == 0.0) return 0;
    b[0] = k[0];
    for (int j = 1; j < nc; ++j) {
        if (k[j] <= k[j - 1]) return 0;
        for (int i = 0; i < j; ++i)
            if (k[j] < b[i]) return 0;
        if (a[j] <= 0.0 || c[j] == 0.0) return 0;
        double som = 0.0;
        for (int i = 0; i < j; ++i)
            som = som + (double)(o->M[(j - 1) * nc + i] - o->M[j * nc + i]) * a[i] * pow(k[j] - b[i], c[i]);
        if (som < 0.0) return 0;
        b[j] = k[j] - pow((1.0 / a[j]) * som, 1.0 / c[j]);
    }
    return 1;
}

/* BaRatin_Apply + BaRatin_computeQ + GetHRange, BaRatin_model.f90:67-137,194-252,398-436 */
static void baratin_apply(const oracle_t* o, int n, const double* H, const double* theta, double* Q, double* b,
                          int32_t* vfeas) {
    int nc = o->ncontrol;
    double k[MAXCONTROL], a[MAXCONTROL], c[MAXCONTROL];
    for (int i = 0; i < nc; ++i) { k[i] = theta[3 * i]; a[i] = theta[3 * i + 1]; c[i] = theta[3 * i + 2]; }
    if (!baratin_continuity(o, a, c, k, b)) {
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return;
    }
    for (int t = 0; t < n; ++t) {
        double h = H[t];
        int range = nc;
        for (int i = 0; i < nc; ++i)
            if (h < k[i]) { range = i; break; }
        double q = 0.0;
        for (int i = 0; i < range; ++i) {
            if (o->M[(range - 1) * nc + i] == 1) {
                if ((h - b[i]) < 0.0) { q = 0.0; break; }
                q = q + a[i] * pow(h - b[i], c[i]);
            } else {
                if ((h - b[i]) < 0.0) { vfeas[t] = 0; break; }
            }
        }
        Q[t] = q;
    }
}

/* ------------------------------------------------------------------------------------------
 * BaRatinBAC (BaRatinBAC_model.f90:96-666)
 * ------------------------------------------------------------------------------------------ */
static const double BAC_EPS = 1.1920928955078125e-07; /* epsilon(1.) -- single precision (:400) */

static double bac_f(double x, double a1, double b1, double c1, double a2, double b2, double c2) { /* fContinuity :521 */
    return c2 * log(x - b2) - c1 * log(x - b1) - log(a1 / a2);
}
static double bac_df(double x, double b1, double c1, double b2, double c2) { /* dfContinuity :528 */
    return (c2 / (x - b2)) - (c1 / (x - b1));
}

/* findRoot (:603-636): znewton(tolX=1e-4, tolF=1e-6, xscale=fscale=1, itmax=10000); declared convention as for the
   Vegetation solver (rtsafe, then the root is polished). */
static int bac_find_root(double a1, double b1, double c1, double a2, double b2, double c2, double x1, double x2, double fl,
                         double fh, double* root) {
    if (fl == 0.0) { *root = x1; return 1; }
    if (fh == 0.0) { *root = x2; return 1; }
    if ((fl > 0.0) == (fh > 0.0) || fl != fl || fh != fh) return 0;
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
        int convx = fabs(dx) <= 1e-4;
        f = bac_f(rts, a1, b1, c1, a2, b2, c2);
        df = bac_df(rts, b1, c1, b2, c2);
        if (f < 0.0) xl = rts; else xh = rts;
        if (convx || fabs(f) <= 1e-6) break;
    }
    for (int j = 0; j < 4 && f != 0.0 && df != 0.0; ++j) {
        double lo = xl < xh ? xl : xh, hi = xl < xh ? xh : xl;
        double xn = rts - f / df;
        if (xn < lo) xn = lo;
        if (xn > hi) xn = hi;
        if (xn == rts) break;
        rts = xn;
        f = bac_f(rts, a1, b1, c1, a2, b2, c2);
        df = bac_df(rts, b1, c1, b2, c2);
        if (f < 0.0) xl = rts; else xh = rts;
    }
    *root = rts;
    return rts == rts;
}

/* SolveContinuity (:396-518) */
static int bac_solve(double a1, double b1, double c1, double a2, double b2, double c2, double lbound, double hbound,
                     double* root, int* troot) {
    const double eps = BAC_EPS;
    *troot = 0;
    if (fabs(c1 - c2) <= eps) {
        if (fabs(a1 - a2) <= eps) return 0;
        double foo1 = pow(a1, 1.0 / c1), foo2 = pow(a2, 1.0 / c1);
        *root = (foo2 * b2 - foo1 * b1) / (foo2 - foo1);
        if (*root > lbound && *root < hbound) { *troot = 1; return 1; }
        return 0;
    }
    if (fabs(b1 - b2) <= eps) {
        *root = b1 + pow(a1 / a2, 1.0 / (c2 - c1));
        if (*root > lbound && *root < hbound) { *troot = 2; return 1; }
        return 0;
    }
    double x0 = (c2 * b1 - c1 * b2) / (c2 - c1);
    if (x0 <= lbound) {
        double f1 = bac_f(lbound + eps, a1, b1, c1, a2, b2, c2), f2 = bac_f(hbound, a1, b1, c1, a2, b2, c2);
        if (f1 * f2 > 0.0) return 0;
        if (bac_find_root(a1, b1, c1, a2, b2, c2, lbound + eps, hbound, f1, f2, root)) { *troot = 3; return 1; }
        return 0;
    }
    double fx0 = bac_f(x0, a1, b1, c1, a2, b2, c2);
    if (fabs(fx0) <= eps) { *root = x0; *troot = 5; return 1; }
    if ((c2 > c1 && fx0 > 0.0) || (c2 < c1 && fx0 < 0.0)) { *troot = 6; return 0; }
    *troot = 4;
    if (x0 >= hbound) return 0;
    double f2 = bac_f(hbound, a1, b1, c1, a2, b2, c2);
    if (f2 * fx0 > 0.0) return 0;
    return bac_find_root(a1, b1, c1, a2, b2, c2, x0, hbound, fx0, f2, root);
}

/* BaRatinBAC_Apply (:96-149) = ApplyContinuity (:237-394) + BaRatin_computeQ (BaRatin_model.f90:194-252) */
static void baratinbac_apply(const oracle_t* o, int n, const double* H, const double* theta, double* Q, double* Dpar,
                             int32_t* vfeas) {
    int nc = o->ncontrol, ok = 1;
    double b[MAXCONTROL], a[MAXCONTROL], c[MAXCONTROL], k[MAXCONTROL], kt[MAXCONTROL];
    for (int i = 0; i < nc; ++i) { b[i] = theta[3 * i]; a[i] = theta[3 * i + 1]; c[i] = theta[3 * i + 2]; }
    for (int i = 0; i < nc; ++i)
        if (a[i] <= 0.0 || c[i] == 0.0) ok = 0;
    if (ok) {
        k[0] = b[0]; kt[0] = 0.0;
        for (int j = nc - 1; j >= 1 && ok; --j) {
            double lbound = b[j] > b[j - 1] ? b[j] : b[j - 1];
            double hbound = (j == nc - 1) ? o->model_optd : k[j + 1];
            int addition = 1;
            for (int i = 0; i < j; ++i)
                if (o->M[j * nc + i] != o->M[(j - 1) * nc + i]) addition = 0;
            if (addition) { k[j] = b[j]; kt[j] = 0.0; }
            else {
                double root = 0.0;
                int troot = 0;
                if (!bac_solve(a[j - 1], b[j - 1], c[j - 1], a[j], b[j], c[j], lbound, hbound, &root, &troot)) { ok = 0; break; }
                k[j] = root; kt[j] = (double)troot;
            }
            if (k[j] >= hbound) { ok = 0; break; }
            for (int i = 0; i < j; ++i)
                if (k[j] < b[i]) ok = 0;
        }
    }
    if (!ok) {
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return;
    }
    for (int i = 0; i < nc; ++i) { Dpar[i] = k[i]; Dpar[nc + i] = kt[i]; }
    for (int t = 0; t < n; ++t) {
        double h = H[t];
        int range = nc;
        for (int i = 0; i < nc; ++i)
            if (h < k[i]) { range = i; break; }
        double q = 0.0;
        for (int i = 0; i < range; ++i) {
            if (o->M[(range - 1) * nc + i] == 1) {
                if ((h - b[i]) < 0.0) { q = 0.0; break; }
                q = q + a[i] * pow(h - b[i], c[i]);
            } else {
                if ((h - b[i]) < 0.0) { vfeas[t] = 0; break; }
            }
        }
        Q[t] = q;
    }
}

/* ------------------------------------------------------------------------------------------
 * SFD (StageFallDischarge_model.f90:80-283), SFD_Val_General
 * ------------------------------------------------------------------------------------------ */
static void sfd_fnewton(double x, const double* th, double h2, double* f, double* df) { /* fNewton_Val_General :253-283 */
    double s = sqrt((x - h2 - th[6]) / th[3]);
    *f = th[0] * pow(x - th[1], th[2]) * s - th[5] * pow(x - th[4], th[7]);
    *df = th[0] * pow(x - th[1], th[2] - 1.0) / (2.0 * th[3] * s) *
              ((2.0 * th[2] + 1.0) * x - (2.0 * th[2] * (h2 + th[6]) + th[1])) -
          th[5] * th[7] * pow(x - th[4], th[7] - 1.0);
}

/* kappa (nullable): the state variable, undefRN where it is never solved for (:126) */
static void sfd_apply(const oracle_t* o, int n, const double* IN, int ldx, const double* th, double* OUT, int32_t* vfeas,
                      double* kappa_out) {
    if (kappa_out)
        for (int t = 0; t < n; ++t) kappa_out[t] = BAMGPU_UNDEF_RN;
    double KB = th[0], h0 = th[1], M = th[2], L = th[3], h0p = th[4], KBS0 = th[5], delta = th[6], Mp = th[7];
    if (KB <= 0.0 || M <= 0.0 || L <= 0.0 || KBS0 <= 0.0 || Mp <= 0.0) {
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return;
    }
    for (int t = 0; t < n; ++t) {
        double h1 = IN[t], h2 = IN[t + (size_t)ldx];
        if (h1 - h2 - delta <= 0.0) { vfeas[t] = 0; continue; }
        if (h1 - h0 <= 0.0 || h2 - h0p <= 0.0) { OUT[t] = 0.0; continue; }
        double m = h0 > h0p ? h0 : h0p;
        if (h2 + delta > m) m = h2 + delta;
        double x1 = m + 0.001, x2 = o->model_optd;
        double fl, fh, df, f, xl, xh, kappa;
        sfd_fnewton(x1, th, h2, &fl, &df);
        sfd_fnewton(x2, th, h2, &fh, &df);
        int fcalls = 2;
        if (fl == 0.0) kappa = x1;
        else if (fh == 0.0) kappa = x2;
        else {
            if ((fl > 0.0) == (fh > 0.0) || fl != fl || fh != fh) { vfeas[t] = 0; continue; }
            if (fl < 0.0) { xl = x1; xh = x2; } else { xl = x2; xh = x1; }
            double rts = 0.5 * (x1 + x2), dxold = fabs(x2 - x1), dx = dxold;
            sfd_fnewton(rts, th, h2, &f, &df);
            ++fcalls;
            for (int j = 0; j < o->veg_itmax; ++j) {
                if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
                    dxold = dx; dx = 0.5 * (xh - xl); rts = xl + dx;
                } else {
                    dxold = dx; dx = f / df; rts = rts - dx;
                }
                int convx = fabs(dx) <= o->veg_tolx * o->veg_xscale;
                sfd_fnewton(rts, th, h2, &f, &df);
                ++fcalls;
                if (f < 0.0) xl = rts; else xh = rts;
                if (convx || fabs(f) <= o->veg_tolf * o->veg_fscale) break;
            }
            for (int j = 0; j < 3 && f != 0.0 && df != 0.0; ++j) {
                double lo = xl < xh ? xl : xh, hi = xl < xh ? xh : xl;
                double xn = rts - f / df;
                if (xn < lo) xn = lo;
                if (xn > hi) xn = hi;
                if (xn == rts) break;
                rts = xn;
                sfd_fnewton(rts, th, h2, &f, &df);
                if (f < 0.0) xl = rts; else xh = rts;
            }
            kappa = rts;
        }
        if (fcalls > o->veg_itmax || kappa != kappa) { vfeas[t] = 0; continue; }
        if (kappa_out) kappa_out[t] = kappa;
        if (h1 <= kappa) OUT[t] = (KB * pow(h1 - h0, M)) * sqrt((h1 - h2 - delta) / L);
        else OUT[t] = KBS0 * pow(h1 - h0p, Mp);
    }
}

/* LM_Apply, Linear_model.f90:62-101: Y = matmul(X, reshape(theta,(nX,nY))) */
static void linear_apply(const oracle_t* o, int n, const double* X, int ldx, const double* theta, double* Y,
                         int ldy) {
    for (int t = 0; t < n; ++t)
        for (int y = 0; y < o->nY; ++y) {
            double s = 0.0;
            for (int x = 0; x < o->nX; ++x) s = s + X[t + (size_t)x * ldx] * theta[y * o->nX + x];
            Y[t + (size_t)y * ldy] = s;
        }
}

static int sed_nbase(int f) { return (f == BAMGPU_SED_MPM || f == BAMGPU_SED_NIELSEN) ? 2 : 3; }
static int sed_ncoeff(int css) { return css == BAMGPU_CSS_SOULSBY ? 4 : (css == BAMGPU_CSS_SW ? 3 : 0); }

/* Sediment_Apply + GetAllPar + CSS_Apply, SedimentTransport_model.f90:100-192,251-310,525-607 */
static void sediment_apply(const oracle_t* o, int n, const double* IN, int ldx, const double* theta, double* OUT,
                           int32_t* vfeas) {
    int nbase = sed_nbase(o->sed_formula), ncoeff = sed_ncoeff(o->sed_css);
    for (int t = 0; t < n; ++t) {
        int bad = 0;
        for (int x = 0; x < o->nX; ++x)
            if (IN[t + (size_t)x * ldx] <= 0.0) bad = 1;
        if (bad) { vfeas[t] = 0; continue; }
        /* GetAllPar */
        double base[3], pseudo[4], coeff[4], cte = BAMGPU_UNDEF_RN;
        int ktheta = 0, kin = 2;
        for (int i = 0; i < nbase; ++i) base[i] = theta[ktheta++];
        if (o->sed_css == BAMGPU_CSS_CONSTANT) cte = theta[ktheta++];
        for (int i = 0; i < 4; ++i) {
            if (o->sed_ppo[i] == BAMGPU_PPO_FIX) pseudo[i] = o->sed_pseudo[i];
            else if (o->sed_ppo[i] == BAMGPU_PPO_PAR) pseudo[i] = theta[ktheta++];
            else pseudo[i] = IN[t + (size_t)(kin++) * ldx];
        }
        if (o->sed_css != BAMGPU_CSS_CONSTANT) {
            if (o->sed_co == BAMGPU_CO_PAR)
                for (int i = 0; i < ncoeff; ++i) coeff[i] = theta[ktheta++];
            else if (o->sed_css == BAMGPU_CSS_SOULSBY) { coeff[0] = 0.3; coeff[1] = 1.2; coeff[2] = 0.055; coeff[3] = 0.02; }
            else { coeff[0] = 0.24; coeff[1] = 0.055; coeff[2] = 0.02; }
        }
        double d = pseudo[0], s = pseudo[1], v = pseudo[2], g = pseudo[3];
        /* CSS_Apply */
        if (d <= 0.0 || s <= 0.0 || v <= 0.0 || g <= 0.0) { vfeas[t] = 0; continue; }
        double css;
        if (o->sed_css == BAMGPU_CSS_CONSTANT) css = cte;
        else {
            double SD = pow(g * (s - 1.0) / (v * v), 1.0 / 3.0) * d;
            if (o->sed_css == BAMGPU_CSS_SOULSBY)
                css = (coeff[0] / (1.0 + coeff[1] * SD)) * coeff[2] * (1.0 - exp(-1.0 * coeff[3] * SD));
            else
                css = (coeff[0] / SD) * coeff[1] * (1.0 - exp(-1.0 * coeff[2] * SD));
        }
        if (s <= 1.0) { vfeas[t] = 0; continue; }
        double in1 = IN[t], in2 = IN[t + (size_t)ldx];
        double tau = (in1 * in2) / ((s - 1.0) * d);
        if (tau <= 0.0) { vfeas[t] = 0; continue; }
        if (tau <= css) { OUT[t] = 0.0; continue; }
        double dim = sqrt(g * (s - 1.0) * (d * d * d));
        switch (o->sed_formula) {
            case BAMGPU_SED_MPM: OUT[t] = dim * base[0] * pow(tau - css, base[1]); break;
            case BAMGPU_SED_CAMENEN: OUT[t] = dim * base[0] * pow(tau, base[1]) * exp(-1.0 * base[2] * (css / tau)); break;
            case BAMGPU_SED_NIELSEN: OUT[t] = dim * base[0] * pow(tau, base[1]) * (tau - css); break;
            case BAMGPU_SED_SMART:
                OUT[t] = dim * base[0] * pow(in2, base[1]) * pow(s - 1.0, 1.0 / 6.0) * pow(d, 1.0 / 6.0) *
                         pow(g, -0.5) * pow(tau, base[2]) * (tau - css);
                break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * Vegetation (Vegetation_model.f90)
 * ------------------------------------------------------------------------------------------ */
#define VEG_MAXYEAR 16
static const double VEG_PI = 3.14159265358979323846; /* utilities_dmsl_kit pi */
static const double VEG_GRAVITY = 9.81;              /* :129 */

static int veg_ngrowth(int growth) { return growth == BAMGPU_VEG_YIN1_TEMPORAL ? 3 : 4; } /* GetNpar_growth :296-330 */

/* Vegetation_fNewton, Vegetation_model.f90:536-569: f(x) = x^2 + gamma x^2 min(1, x^chi/gamma2) - Q0^2 */
static void veg_fnewton(double x, double Q0, double gam, double gam2, double chi, double* fx, double* dfdx) {
    double m = pow(x, chi) / gam2;
    if (1.0 < m) m = 1.0; /* min(1., .) */
    *fx = x * x + (gam * pow(x, 2.0) * m) - Q0 * Q0;
    if (m == 1.0) *dfdx = 2.0 * x * (1.0 + gam);
    else *dfdx = 2.0 * x + (gam / gam2 * (2.0 + chi) * pow(x, 1.0 + chi));
}

/* znewton (miniDMSL numerix_dmsl_kit, UN-VENDORED; declared convention = Numerical-Recipes "rtsafe": Newton from
   the mid-point of the bracket, bisection whenever the Newton step leaves the bracket or converges too slowly;
   stop when |dx| <= tolX*xscale or |f| <= tolF*fscale; fcalls counts every evaluation of f).
   After that stop the root is POLISHED by up to 3 more clamped Newton steps: the reference's own answer is only
   defined up to its tolerances (path unknown), so the restatement and the device both return the root itself
   (|polished - reference| <= the reference's tolerance).  Returns 0 OK, 1 = not bracketed. */
static int veg_znewton(double Q0, double gam, double gam2, double chi, double x1, double x2, double tolX, double tolF,
                       double xscale, double fscale, int itmax, double* xroot, int* fcalls) {
    double fl, fh, df, f, xl, xh;
    veg_fnewton(x1, Q0, gam, gam2, chi, &fl, &df);
    veg_fnewton(x2, Q0, gam, gam2, chi, &fh, &df);
    *fcalls = 2;
    if (fl == 0.0) { *xroot = x1; return 0; }
    if (fh == 0.0) { *xroot = x2; return 0; }
    if ((fl > 0.0) == (fh > 0.0) || fl != fl || fh != fh) { *xroot = BAMGPU_UNDEF_RN; return 1; }
    if (fl < 0.0) { xl = x1; xh = x2; } else { xl = x2; xh = x1; }
    double rts = 0.5 * (x1 + x2), dxold = fabs(x2 - x1), dx = dxold;
    veg_fnewton(rts, Q0, gam, gam2, chi, &f, &df);
    ++*fcalls;
    for (int j = 0; j < itmax; ++j) {
        if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df))) {
            dxold = dx; dx = 0.5 * (xh - xl); rts = xl + dx;
        } else {
            dxold = dx; dx = f / df; rts = rts - dx;
        }
        int convx = fabs(dx) <= tolX * xscale;
        veg_fnewton(rts, Q0, gam, gam2, chi, &f, &df);
        ++*fcalls;
        if (f < 0.0) xl = rts; else xh = rts;
        if (convx || fabs(f) <= tolF * fscale) break;
    }
    for (int j = 0; j < 3 && f != 0.0 && df > 0.0; ++j) { /* polish */
        double lo = xl < xh ? xl : xh, hi = xl < xh ? xh : xl;
        double xn = rts - f / df;
        if (xn < lo) xn = lo;
        if (xn > hi) xn = hi;
        if (xn == rts) break;
        rts = xn;
        veg_fnewton(rts, Q0, gam, gam2, chi, &f, &df);
        if (f < 0.0) xl = rts; else xh = rts;
    }
    *xroot = rts;
    return 0;
}

/* Apply_growth, Vegetation_model.f90:345-534.  Returns 0 if the whole parameter set is infeasible. */
static int veg_apply_growth(const oracle_t* o, int n, const double* time, const double* Traw, const double* Tsmooth,
                            const double* par, const double* gmax, double* g, double* g0, double* cosg, double* sing,
                            double* Tm, int* err) {
    *err = 0;
    for (int i = 0; i < n; ++i) g[i] = g0[i] = cosg[i] = sing[i] = BAMGPU_UNDEF_RN;
    if (o->veg_growth == BAMGPU_VEG_YIN1_TEMPORAL) { /* :395-419 */
        double t0 = par[0], ti = par[1], tmax = par[2];
        if (t0 <= 0.0 || ti <= 0.0 || tmax <= 0.0 || t0 >= 1.0 || ti >= 1.0 || tmax >= 1.0 || t0 >= ti || ti >= tmax) return 0;
        double alf = (tmax - t0) / (tmax - ti), tf = 2.0 * tmax - ti;
        for (int i = 0; i < n; ++i) {
            double tt = time[i] - floor(time[i]);
            if (tt >= t0 && tt <= tf) g0[i] = (pow((tt - t0) / (tmax - t0), alf)) * (tf - tt) / (tf - tmax);
            else g0[i] = 0.0;
            cosg[i] = cos(g0[i] * VEG_PI);
            if (tt <= tmax) sing[i] = sin(g0[i] * VEG_PI);
            else sing[i] = -1.0 * sin(g0[i] * VEG_PI);
            g[i] = gmax[0] * g0[i];
        }
        return 1;
    }
    double Tmin = par[0], Tb = par[1], Te = par[2];
    int nYear = o->veg_nyear;
    for (int i = 0; i < nYear; ++i) Tm[i] = BAMGPU_UNDEF_RN;
    if (Te <= Tb) return 0;
    double* tim = (double*)malloc(((size_t)n + 1) * 5 * sizeof(double));
    double *temp = tim + n, *stemp = temp + n, *cum = stemp + n, *foog0 = cum + n;
    int ok = 1;
    for (int i = 0; i < nYear && ok; ++i) {
        int m = 0;
        for (int j = 0; j < n; ++j)
            if (floor(time[j]) == (double)i) { tim[m] = time[j]; temp[m] = Traw[j]; stemp[m] = Tsmooth[j]; ++m; }
        if (m == 0) { *err = 1; break; } /* "empty year, currently unsupported" :432-436 */
        double cumul = 0.0; /* degree-days :447-453 */
        if (temp[0] >= Tmin) cumul = cumul + 365.0 * (tim[0] - (double)i) * (temp[0] - Tmin);
        cum[0] = cumul;
        for (int j = 1; j < m; ++j) {
            if (temp[j] >= Tmin) cumul = cumul + 365.0 * (tim[j] - tim[j - 1]) * (temp[j] - Tmin);
            cum[j] = cumul;
        }
        int gnull = 0, k;
        double time_b = 0.0, time_e = 0.0, time_m = 0.0, time_f = 0.0, alpha = 0.0;
        for (k = 0; k < m && !(cum[k] >= Tb); ++k) {}
        if (k < m) time_b = tim[k]; else gnull = 1;
        for (k = 0; k < m && !(cum[k] >= Te); ++k) {}
        if (k < m) time_e = tim[k]; else { ok = 0; break; }
        if (gnull) {
            for (int j = 0; j < m; ++j) foog0[j] = 0.0;
        } else {
            switch (o->veg_growth) {
                case BAMGPU_VEG_YIN1:
                    Tm[i] = par[3];
                    for (k = 0; k < m && !(stemp[k] >= Tm[i]); ++k) {}
                    if (k < m) time_m = tim[k]; else { ok = 0; }
                    if (!ok) break;
                    if (time_m <= time_b) time_m = time_b;
                    if (time_m >= time_e) { ok = 0; break; }
                    alpha = (time_e - time_b) / (time_e - time_m);
                    time_f = 2.0 * time_e - time_m;
                    break;
                case BAMGPU_VEG_YIN2:
                    if (par[3] <= 0.0) { ok = 0; break; }
                    alpha = par[3];
                    time_f = time_e + (time_e - time_b) / alpha;
                    time_m = ((alpha - 1.0) * time_f + 2.0 * time_b) / (alpha + 1.0);
                    break;
                default: /* Yin3 */
                    if (par[3] <= 0.0) { ok = 0; break; }
                    time_f = time_e + par[3];
                    alpha = (time_e - time_b) / (time_f - time_e);
                    time_m = ((alpha - 1.0) * time_f + 2.0 * time_b) / (alpha + 1.0);
                    break;
            }
            if (!ok) break;
            if (time_m < time_b) time_m = time_b;
            if (alpha <= 0.0 || time_m >= time_e || time_f > tim[m - 1]) { ok = 0; break; }
            for (k = 0; k < m && !(tim[k] >= time_m); ++k) {}
            Tm[i] = stemp[k < m ? k : m - 1];
            for (int j = 0; j < m; ++j) {
                if (tim[j] <= time_b) { foog0[j] = 0.0; continue; }
                if (tim[j] >= time_f) { foog0[j] = 0.0; continue; }
                foog0[j] = pow((tim[j] - time_b) / (time_e - time_b), alpha) * (time_f - tim[j]) / (time_f - time_e);
            }
        }
        k = 0;
        for (int j = 0; j < n; ++j)
            if (floor(time[j]) == (double)i) {
                g0[j] = foog0[k];
                g[j] = g0[j] * gmax[i];
                cosg[j] = cos(foog0[k] * VEG_PI);
                if (tim[k] <= time_e) sing[j] = sin(foog0[k] * VEG_PI);
                else sing[j] = -1.0 * sin(foog0[k] * VEG_PI);
                ++k;
            }
    }
    free(tim);
    return ok && !*err;
}

/* Vegetation_Apply, Vegetation_model.f90:87-243.  OUT columns: Q, g, cos, sin, g0 (the first nY of them). */
static int vegetation_apply(const oracle_t* o, int n, const double* IN, int ldx, const double* theta, double* OUT,
                            int ldy, double* Dpar, int32_t* vfeas) {
    const int temporal = o->veg_growth == BAMGPU_VEG_YIN1_TEMPORAL;
    const int nG = veg_ngrowth(o->veg_growth), nGmax = temporal ? 1 : o->veg_nyear;
    const double *time = IN, *h = IN + (size_t)ldx, *Traw = temporal ? NULL : IN + 2 * (size_t)ldx,
                 *Tsmooth = temporal ? NULL : IN + 3 * (size_t)ldx;
    double B = theta[0], S0 = theta[1], n1 = theta[2], b1 = theta[3], c1 = theta[4], chi = theta[5], U_chi = theta[6];
    const double* growthPar = theta + 7;
    const double* gmax = theta + 7 + nG;
    for (int t = 0; t < n; ++t)
        for (int y = 0; y < o->nY; ++y) OUT[t + (size_t)y * ldy] = BAMGPU_UNDEF_RN;
    Dpar[0] = (S0 >= 0.0) ? B * sqrt(S0) / n1 : BAMGPU_UNDEF_RN;
    int badg = 0;
    for (int i = 0; i < nGmax; ++i) badg |= gmax[i] < 0.0;
    if (B <= 0.0 || S0 <= 0.0 || n1 <= 0.0 || c1 <= 0.0 || chi < -2.0 || chi > 0.0 || U_chi <= 0.0 || badg) {
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return 0;
    }
    double* g = (double*)malloc(((size_t)n + 1) * 4 * sizeof(double));
    double *g0 = g + n, *cosg = g0 + n, *sing = cosg + n, Tm[VEG_MAXYEAR];
    int err = 0;
    if (!veg_apply_growth(o, n, time, Traw, Tsmooth, growthPar, gmax, g, g0, cosg, sing, Tm, &err)) {
        free(g);
        if (err) return 1;
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return 0;
    }
    if (!temporal)
        for (int i = 0; i < o->veg_nyear; ++i) Dpar[1 + i] = Tm[i];
    const double* cols[4] = {g, cosg, sing, g0};
    for (int y = 1; y < o->nY && y < 5; ++y)
        for (int t = 0; t < n; ++t) OUT[t + (size_t)y * ldy] = cols[y - 1][t];
    for (int t = 0; t < n; ++t) { /* discharge :202-240 */
        double y = h[t] - b1, Q;
        if (y <= 0.0) { OUT[t] = 0.0; continue; }
        double Q0 = ((B * sqrt(S0)) / n1) * pow(y, c1);
        if (g[t] <= 0.0) { OUT[t] = Q0; continue; }
        double gam = (g[t] * pow(y, 1.0 / 3.0)) / (8.0 * VEG_GRAVITY * (n1 * n1));
        double gam2 = pow(B, chi) * pow(y, chi) * pow(U_chi, chi);
        int fcalls = 0;
        int e = veg_znewton(Q0, gam, gam2, chi, 0.0, Q0, o->veg_tolx, o->veg_tolf, o->veg_xscale, o->veg_fscale,
                            o->veg_itmax, &Q, &fcalls);
        if (e > 0 || fcalls > o->veg_itmax) { vfeas[t] = 0; continue; }
        OUT[t] = Q;
    }
    free(g);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * Second-wave pointwise models
 * ------------------------------------------------------------------------------------------ */

/* SuspendedLoad_Apply, SuspendedLoad_model.f90:38-116 */
static void suspendedload_apply(int n, const double* h, const double* theta, double* Qs, int32_t* vfeas) {
    if (theta[0] <= 0.0 || theta[2] <= 0.0 || theta[4] <= 0.0) {
        for (int i = 0; i < n; ++i) vfeas[i] = 0;
        return;
    }
    for (int i = 0; i < n; ++i) {
        if (h[i] < theta[1] || h[i] < theta[3]) { vfeas[i] = 0; continue; }
        double num = pow(h[i] - theta[1], theta[2]), den = pow(h[i] - theta[3], theta[4]);
        Qs[i] = theta[0] * (num / den);
    }
}

/* SWOT_Apply, SWOT_model.f90:73-148 */
static void swot_apply(const oracle_t* o, int n, const double* IN, int ldx, const double* theta, double* OUT, int32_t* vfeas) {
    double K = theta[0], S0 = theta[2], h0 = theta[3], M = theta[4], B0 = 0.0;
    int hS = o->model_opt[0] == BAMGPU_SWOT_HS;
    if (!hS) B0 = theta[1];
    if (K <= 0.0 || M <= 0.0) {
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return;
    }
    for (int t = 0; t < n; ++t) {
        double ht = IN[t], St = IN[t + (size_t)ldx], Bt = hS ? theta[1] : IN[t + 2 * (size_t)ldx];
        if (Bt - B0 < 0.0 || St - S0 < 0.0) { vfeas[t] = 0; continue; }
        if (ht - h0 <= 0.0) { OUT[t] = 0.0; continue; }
        OUT[t] = K * (Bt - B0) * sqrt(St - S0) * pow(ht - h0, M);
    }
}

/* SGD_Apply, StageGradientDischarge_model.f90:57-113 */
static void sgd_apply(int n, const double* IN, int ldx, const double* theta, double* OUT, int32_t* vfeas) {
    double K = theta[0], B = theta[1], S0 = theta[2], h0 = theta[3], M = theta[4];
    if (K <= 0.0 || M <= 0.0 || B <= 0.0 || S0 <= 0.0) {
        for (int t = 0; t < n; ++t) vfeas[t] = 0;
        return;
    }
    for (int t = 0; t < n; ++t) {
        double h = IN[t], dhdt = IN[t + (size_t)ldx];
        if (h - h0 <= 0.0) { vfeas[t] = 0; continue; }
        double Q0 = K * B * sqrt(S0) * pow(h - h0, M);
        double c = M * K * sqrt(S0) * pow(h - h0, M - 1.0);
        double corr = 1.0 + (1.0 / (c * S0)) * dhdt;
        if (corr <= 0.0) { vfeas[t] = 0; continue; }
        OUT[t] = Q0 * sqrt(corr);
    }
}

/* Recession_h_Apply, Recession_model.f90:36-94 (normalised, two exponentials) */
static void recession_apply(int n, const double* time, const double* theta, double* h) {
    for (int t = 0; t < n; ++t)
        h[t] = theta[0] * exp(-theta[1] * time[t]) + (1 - theta[0] - theta[2]) * exp(-theta[3] * time[t]) + theta[2];
}

/* Ortho_Apply + Disto_Apply, Orthorectification_model.f90:66-288 */
static void ortho_apply(const oracle_t* o, int n, const double* IN, int ldx, const double* theta, double* OUT, int ldy,
                        double* Dpar, int32_t* vfeas) {
    double x0 = theta[0], y0 = theta[1], z0 = theta[2], azim = theta[3], roll = theta[4], tilt = theta[5], f = theta[6],
           c0 = theta[7], l0 = theta[8], lambda = theta[9], k = theta[10];
    double pc = lambda, pl = k * pc;
    int bad = (pc <= 0.0 || pl <= 0.0 || k <= 0.0);
    double ac = 0, al = 0, r11 = 0, r12 = 0, r13 = 0, r21 = 0, r22 = 0, r23 = 0, r31 = 0, r32 = 0, r33 = 0;
    if (!bad) {
        ac = f / pc; al = f / pl;
        r11 = cos(azim) * cos(roll) + sin(azim) * cos(tilt) * sin(roll);
        r12 = -1.0 * sin(azim) * cos(roll) + cos(azim) * cos(tilt) * sin(roll);
        r13 = sin(tilt) * sin(roll);
        r21 = -1.0 * cos(azim) * sin(roll) + sin(azim) * cos(tilt) * cos(roll);
        r22 = sin(azim) * sin(roll) + cos(azim) * cos(tilt) * cos(roll);
        r23 = sin(tilt) * cos(roll);
        r31 = sin(azim) * sin(tilt); r32 = cos(azim) * sin(tilt); r33 = -1.0 * cos(tilt);
        double alpha = r31 * x0 + r32 * y0 + r33 * z0;
        if (alpha == 0.0) bad = 1;
        else {
            double a1 = (ac * r21 - r31 * c0) / alpha, a2 = (ac * r22 - r32 * c0) / alpha, a3 = (ac * r23 - r33 * c0) / alpha;
            double a4 = -1.0 * (a1 * x0 + a2 * y0 + a3 * z0);
            double b1 = (al * r11 - r31 * l0) / alpha, b2 = (al * r12 - r32 * l0) / alpha, b3 = (al * r13 - r33 * l0) / alpha;
            double b4 = -1.0 * (b1 * x0 + b2 * y0 + b3 * z0);
            double v[20] = {r11, r12, r13, r21, r22, r23, r31, r32, r33, a1, a2, a3, a4, b1, b2, b3, b4,
                            -1.0 * r31 / alpha, -1.0 * r32 / alpha, -1.0 * r33 / alpha};
            memcpy(Dpar, v, sizeof v);
        }
    }
    if (bad) {
        for (int i = 0; i < n; ++i) vfeas[i] = 0;
        return;
    }
    for (int i = 0; i < n; ++i) {
        double x = IN[i], y = IN[i + (size_t)ldx], z = IN[i + 2 * (size_t)ldx];
        double denom = r31 * (x - x0) + r32 * (y - y0) + r33 * (z - z0);
        if (denom == 0.0) { vfeas[i] = 0; continue; }
        double c = c0 - ac * (r21 * (x - x0) + r22 * (y - y0) + r23 * (z - z0)) / denom;
        double l = l0 - al * (r11 * (x - x0) + r12 * (y - y0) + r13 * (z - z0)) / denom;
        if (o->model_opt[0] == BAMGPU_DISTO_NONE) { OUT[i] = c; OUT[i + (size_t)ldy] = l; continue; }
        double r = sqrt((c - c0) * (c - c0) + (l - l0) * (l - l0));
        double d = 1.0;
        for (int q = 1; q <= o->model_opt[1]; ++q) {
            double rp = 1.0; /* r**(2*q): integer power */
            for (int m = 0; m < 2 * q; ++m) rp = rp * r;
            d = d + theta[10 + q] * rp;
        }
        OUT[i] = c0 + d * (c - c0);
        OUT[i + (size_t)ldy] = l0 + d * (l - l0);
    }
}

/* HydraulicControl_section_Apply, HydraulicControl_section_model.f90:40-112 */
static void hcs_apply(const oracle_t* o, int n, const double* H, const double* theta, double* Q, int32_t* vfeas) {
    int p = o->hcs_n;
    const double *hh = o->xtra_d, *AA = o->xtra_d + p, *ww = o->xtra_d + 2 * p;
    double coeff = theta[0], g = theta[1];
    if (coeff <= 0.0 || g <= 0.0) {
        for (int i = 0; i < n; ++i) vfeas[i] = 0;
        return;
    }
    double* lhs = (double*)malloc((size_t)p * 2 * sizeof(double));
    double* fx = lhs + p;
    for (int i = 0; i < p; ++i) lhs[i] = hh[i] + 0.5 * (i == 0 ? 0.0 : AA[i] / ww[i]);
    for (int i = 0; i < n; ++i) {
        if (H[i] <= hh[0]) { Q[i] = 0.0; continue; }
        if (H[i] >= hh[p - 1]) { vfeas[i] = 0; continue; }
        int k1 = -1;
        for (int k = 0; k < p; ++k) {
            fx[k] = lhs[k] - H[i];
            if (fx[k] <= 0.0 && (k1 < 0 || fx[k] > fx[k1])) k1 = k; /* maxloc(fx, mask=fx<=0): first maximum */
        }
        int k2 = k1 + 1;
        if (k1 < 0 || k2 >= p) { vfeas[i] = 0; continue; }
        double x1 = hh[k1], x2 = hh[k2], y1 = fx[k1], y2 = fx[k2];
        double hc = (x1 * y2 - x2 * y1) / (y2 - y1);
        y1 = AA[k1]; y2 = AA[k2];
        double Ahc = (y1 * (x2 - hc) + y2 * (hc - x1)) / (x2 - x1);
        Q[i] = coeff * sqrt(2.0 * g * (H[i] - hc)) * Ahc;
    }
    free(lhs);
}

/* Segmentation_Apply, Segmentation_model.f90:52-140.  Returns 0 when the parameter set is infeasible. */
static int segmentation_apply(const oracle_t* o, int nt, const double* time, const double* theta, double* Y) {
    int nS = o->model_opt[0], nmin = o->model_opt[1], option = o->model_opt[2];
    double tmin = o->model_optd;
    if (nS == 1) {
        for (int i = 0; i < nt; ++i) Y[i] = theta[0];
        return 1;
    }
    double tau[64], duration[64];
    if (option == 1) {
        for (int i = 0; i < nS - 1; ++i) tau[i] = theta[nS + i];
        duration[0] = 1.0;
        for (int i = 1; i < nS - 1; ++i) duration[i] = tau[i] - tau[i - 1];
    } else {
        double acc = 0.0;
        for (int i = 0; i < nS - 1; ++i) { duration[i] = theta[nS + i]; acc = acc + duration[i]; tau[i] = acc; }
    }
    for (int i = 0; i < nS - 1; ++i)
        if (duration[i] <= 0.0) return 0;
    for (int i = 0; i < nS - 1; ++i)
        if (duration[i] <= tmin) return 0;
    double tmax = -HUGE_VAL;
    for (int i = 0; i < nt; ++i) if (time[i] > tmax) tmax = time[i];
    for (int i = 0; i < nS - 1; ++i)
        if (tau[i] >= tmax) return 0;
    int cnt[64] = {0};
    int* seg = (int*)malloc(((size_t)nt + 1) * sizeof(int));
    for (int i = 0; i < nt; ++i) {
        int r = nS - 1;
        for (int k = 0; k < nS - 1; ++k)
            if (time[i] < tau[k]) { r = k; break; }
        seg[i] = r;
        cnt[r]++;
    }
    int ok = 1;
    for (int k = 0; k < nS; ++k)
        if (cnt[k] < nmin) ok = 0;
    if (ok)
        for (int i = 0; i < nt; ++i) Y[i] = theta[seg[i]];
    free(seg);
    return ok;
}

/* Mixture_Apply, Mixture_model.f90:74-135.  Output rows are event x element blocks: row (j-1)*nElt+e is the e-th
   input row whose event column (X(:,1), rounded) equals j -- pack() order.  Returns 0 when the parameter set is
   infeasible (scalar feas), -1 when an event does not have exactly nElt rows (an array-size mismatch in the reference). */
static int mixture_apply(const oracle_t* o, int n, const double* X, int ldx, const double* theta, double* Y,
                         double* theta_last) {
    const int nS = o->model_opt[0], nEvt = o->model_opt[1], nElt = o->model_opt[2], missing = o->model_opt[3];
    for (int t = 0; t < n; ++t) Y[t] = BAMGPU_UNDEF_RN;
    for (int j = 0; j < nEvt; ++j) theta_last[j] = BAMGPU_UNDEF_RN;
    for (int j = 0; j < nEvt; ++j) {
        double w[64], last, s = 0.0;
        if (missing) {
            for (int i = 0; i < nS; ++i) { w[i] = theta[j * nS + i]; s = s + w[i]; }
            if (s > 1.0) return 0;
            last = 1.0 - s;
        } else {
            for (int i = 0; i < nS - 1; ++i) { w[i] = theta[j * (nS - 1) + i]; s = s + w[i]; }
            w[nS - 1] = 1.0 - s;
            last = w[nS - 1];
        }
        for (int i = 0; i < nS; ++i)
            if (w[i] < 0.0 || w[i] > 1.0) return 0;
        int e = 0;
        for (int t = 0; t < n; ++t) {
            if (lround(X[t]) != j + 1) continue;
            if (e < nElt && j * nElt + e < n) {
                double y = 0.0; /* matmul(sources, dummytheta): accumulation over the sources in order */
                for (int i = 0; i < nS; ++i) y = y + X[t + (size_t)(2 + i) * ldx] * w[i];
                if (missing) y = y + last * theta[nEvt * nS + e];
                Y[j * nElt + e] = y;
            }
            ++e;
        }
        if (e != nElt) return -1;
        theta_last[j] = last;
    }
    return 1;
}

/* AlgaeBiomass_Apply, AlgaeBiomass_model.f90:55-197: daily biomass recursion; X = (temperature, irradiance, depth, turbidity,
   removal), ld = ldx; Y(:,1) = biomass, Y(:,2) = biomass / bmax (ld = ldy); state (nullable, ld = ldy): Fb, Ft, Fi, Fn, Rhyd.
   Returns 0 when the parameter set is infeasible (:140-146), -1 on the input errors of :79-99 (err = 1 in the reference).
   epsRe is DMSL's (un-vendored): declared = epsilon(1._mrk). */
static int algae_apply_depth(int n, const double* X, int ldx, const double* depthArg, const double* th, double* Y, int ldy, double* state) {
    const double *temp = X, *irr = X + (size_t)ldx, *depth = depthArg ? depthArg : X + 2 * (size_t)ldx, *turb = X + 3 * (size_t)ldx, *rem = X + 4 * (size_t)ldx;
    for (int t = 0; t < n; ++t) {
        Y[t] = BAMGPU_UNDEF_RN; Y[t + (size_t)ldy] = BAMGPU_UNDEF_RN;
        if (state) for (int s = 0; s < 5; ++s) state[t + (size_t)s * ldy] = BAMGPU_UNDEF_RN;
    }
    for (int t = 0; t < n; ++t)
        if (irr[t] < 0.0 || depth[t] < 0.0 || turb[t] < 0.0 || rem[t] < 0.0 || rem[t] > 1.0) return -1;
    double dt = th[0], b0 = th[1], bmin = th[2], bmax = th[3], mu0 = th[4], Tmin = th[5], Topt = th[6], Tmax = th[7], Iopt = th[8],
           katt = th[9], kcw = th[10], Rsen = th[11], g = th[12], rho = th[13], J = th[14], tau0 = th[15], bHyd = th[16], cHyd = th[17];
    if (dt <= 0.0 || bmin < 0.0 || b0 < bmin || b0 > bmax || bmin >= bmax || mu0 < 0.0 || Tmin >= Tmax || Topt >= Tmax ||
        Topt <= Tmin || Iopt <= 0.0 || katt < 0.0 || kcw < 0.0 || Rsen < 0.0 || g < 0.0 || rho < 0.0 || J < 0.0 || tau0 <= 0.0 ||
        bHyd < 0.0 || cHyd < 0.0)
        return 0;
    double Texp = (Topt - Tmin) / (Tmax - Topt);
    double bt = b0;
    for (int t = 0; t < n; ++t) {
        double Fb = 1.0 - bt / bmax, Ft, Fi, Fn, Rhyd, iz, Iratio, tau, b;
        if (temp[t] <= Tmin || temp[t] >= Tmax) Ft = 0.0;
        else {
            Ft = ((Tmax - temp[t]) / (Tmax - Topt)) * pow((temp[t] - Tmin) / (Topt - Tmin), Texp);
            if (Ft <= 2.2204460492503131e-16) Ft = 0.0;
        }
        iz = irr[t] * exp(-(katt * turb[t] + kcw) * depth[t]);
        Iratio = iz / Iopt;
        Fi = Iratio * exp(1.0 - Iratio);
        Fn = 1.0;
        tau = depth[t] * g * J * rho;
        Rhyd = (tau <= tau0) ? 0.0 : cHyd * pow((tau - tau0) / tau0, bHyd);
        b = bt + (mu0 * Fb * Ft * Fi * Fn) * bt * dt - (Rsen + rem[t] + Rhyd) * (bt - bmin) * dt;
        b = b < bmax ? b : bmax;
        b = b > bmin ? b : bmin;
        bt = b;
        Y[t] = b;
        Y[t + (size_t)ldy] = b / bmax;
        if (state) { state[t] = Fb; state[t + (size_t)ldy] = Ft; state[t + 2 * (size_t)ldy] = Fi; state[t + 3 * (size_t)ldy] = Fn; state[t + 4 * (size_t)ldy] = Rhyd; }
    }
    return 1;
}
static int algae_apply(int n, const double* X, int ldx, const double* th, double* Y, int ldy, double* state) {
    return algae_apply_depth(n, X, ldx, NULL, th, Y, ldy, state);
}

/* DynamicVegetation_Apply, DynamicVegetation_model.f90:74-191: X = (temperature, irradiance, stage, turbidity, removal);
   theta = 18 AlgaeBiomass parameters, then B, n1, b1, c1, chi, U_chi; gravity and slope are AlgaeBiomass's g = theta(13) and
   J = theta(15) (:30).  Y(:,1) = Q, Y(:,2) = biomass, Y(:,3) = biomass / bmax; state = Fb, Ft, Fi, Fn, Rhyd.
   The biomass budget runs on depth = max(stage - b1, 0) (:137-138); per time step the discharge solves Vegetation_fNewton on
   [0, Q0] (:147-189), a failed search flags the ROW (feas(t)).  Returns as algae_apply. */
static int veg_znewton(double Q0, double gam, double gam2, double chi, double x1, double x2, double tolX, double tolF,
                       double xscale, double fscale, int itmax, double* xroot, int* fcalls);
static int dynveg_apply(const oracle_t* o, int n, const double* X, int ldx, const double* th, double* Y, int ldy, double* state,
                        int32_t* vfeas) {
    const double B = th[18], n1 = th[19], b1 = th[20], c1 = th[21], chi = th[22], U_chi = th[23], S0 = th[14], g = th[12];
    const double* stage = X + 2 * (size_t)ldx;
    double* depth = (double*)malloc(((size_t)n + 1) * sizeof(double));
    for (int t = 0; t < n; ++t) {
        depth[t] = stage[t] - b1;
        if (depth[t] <= 0.0) depth[t] = 0.0;
        Y[t] = BAMGPU_UNDEF_RN;
    }
    int r = algae_apply_depth(n, X, ldx, depth, th, Y + (size_t)ldy, ldy, state);
    if (r <= 0) { free(depth); return r; }
    for (int t = 0; t < n; ++t) {
        double y = depth[t], Q;
        if (y <= 0.0) { Y[t] = 0.0; continue; }
        double Q0 = ((B * sqrt(S0)) / n1) * pow(y, c1);
        double bio = Y[t + (size_t)ldy];
        if (bio <= 0.0) { Y[t] = Q0; continue; }
        double gam = (bio * pow(y, 1.0 / 3.0)) / (8.0 * g * (n1 * n1));
        double gam2 = pow(B, chi) * pow(y, chi) * pow(U_chi, chi);
        int fcalls = 0;
        int e = veg_znewton(Q0, gam, gam2, chi, 0.0, Q0, o->veg_tolx, o->veg_tolf, o->veg_xscale, o->veg_fscale, o->veg_itmax, &Q, &fcalls);
        if (e > 0 || fcalls > o->veg_itmax) { vfeas[t] = 0; continue; }
        Y[t] = Q;
    }
    free(depth);
    return 1;
}

/* SFDTidal_Apply (SFDTidal_model.f90:65-141), SFDTidal2_Apply, SFDTidal4_Apply, SFDTidalJones_Apply (same lines of their
   files): variant 1..4.  tlag = clamp(t - Dt, 1, n) with Dt = int(theta(1)); IN = (h1, h2[, dh/dt]), ld = ldx.
   Returns 0 when the parameter set is infeasible. */
static double f_sign1(double x) { return (x < 0.0 || (x == 0.0 && signbit(x))) ? -1.0 : 1.0; } /* sign(1._mrk, x) */
static int sfdtidal_lag_apply(int variant, int n, const double* IN, int ldx, const double* th, double* OUT) {
    const double *h1 = IN, *h2 = IN + (size_t)ldx, *dhdt = IN + 2 * (size_t)ldx;
    long long Dt = (long long)th[0];
    for (int t = 0; t < n; ++t) OUT[t] = BAMGPU_UNDEF_RN;
#define TLAG(t) ((long long)(t) - Dt < 0 ? 0 : ((long long)(t) - Dt > n - 1 ? n - 1 : (int)((long long)(t) - Dt)))
    if (variant == 1) {
        double K = th[1], B1 = th[2], B2 = th[3], h0 = th[4], M = th[5], L = th[6], dz = th[7];
        if (K <= 0.0 || B1 <= 0.0 || B2 <= 0.0 || M <= 0.0 || L <= 0.0) return 0;
        for (int t = 0; t < n; ++t) {
            int tl = TLAG(t);
            OUT[t] = f_sign1(h1[tl] - h2[tl] - dz) * K * (B2 + B1) / 2 * sqrt(fabs(h1[tl] - h2[tl] - dz) / L) *
                     pow((h1[tl] + h2[tl]) / 2 - h0, M);
            if (h1[t] - h0 <= 0.0) OUT[t] = 0.0;
        }
        return 1;
    }
    double K = th[1], B = th[2], h0 = th[3], M = th[4], L = th[5], delta = th[6];
    if (variant == 2) {
        double alpha = th[7];
        if (K <= 0.0 || B <= 0.0 || M <= 0.0 || L <= 0.0 || alpha <= 0.0 || alpha >= 1.0) return 0;
        for (int t = 0; t < n; ++t) {
            int tl = TLAG(t);
            if (h1[t] - h0 <= 0.0) { OUT[t] = 0.0; continue; }
            OUT[t] = f_sign1(h1[t] - alpha * h2[tl] - delta) * K * B * sqrt(fabs(h1[t] - alpha * h2[tl] - delta) / L) * pow(h1[t] - h0, M);
        }
        return 1;
    }
    if (variant == 3) {
        long long P = (long long)th[7];
        if (K <= 0.0 || B <= 0.0 || M <= 0.0 || L <= 0.0 || P < 0) return 0;
        double* Sf = (double*)malloc(((size_t)n + 1) * sizeof(double));
        for (int t = 0; t < n; ++t) { int tl = TLAG(t); Sf[t] = (h1[tl] - h2[tl] - delta) / L; }
        for (int t = 0; t < n; ++t) {
            long long t1 = t, t2 = t;
            if (!((long long)t - P < 0 || (long long)t - P > n - 1 || (long long)t + P < 0 || (long long)t + P > n - 1)) { t1 = t - P; t2 = t + P; }
            double sum = 0.0;
            for (long long u = t1; u <= t2; ++u) sum += Sf[u];
            double Sm = sum / (double)(t2 - t1 + 1);
            if (h1[t] - h0 <= 0.0) { OUT[t] = 0.0; continue; }
            OUT[t] = f_sign1(Sm) * K * B * sqrt(fabs(Sm)) * pow(h1[t] - h0, M);
        }
        free(Sf);
        return 1;
    }
    {
        double alpha = th[7];
        if (K <= 0.0 || B <= 0.0 || M <= 0.0 || L <= 0.0 || alpha <= 0.0) return 0;
        for (int t = 0; t < n; ++t) {
            int tl = TLAG(t);
            if (h1[t] - h0 <= 0.0) { OUT[t] = 0.0; continue; }
            double Sf = (h1[tl] - h2[tl] - delta) / L;
            OUT[t] = f_sign1(Sf) * K * B * sqrt(fabs(Sf) * (1 - (alpha * dhdt[t]))) * pow(h1[t] - h0, M);
        }
        return 1;
    }
#undef TLAG
}

/* SFDTidal_Qmec_Apply, SFDTidal_Qmec_model.f90:61-197: Q(m) = Q(m-1) + dt (pg + bf + ad), first-order step for m = 2,
   second-order Adams-Bashforth afterwards.  state (nullable, ld = lds): pressure_gradient, bottom_friction, advection.
   Returns 0 when the parameter set (or the geometry of the series) is infeasible. */
static int sfdtidal_qmec_apply(int variant, int nm, const double* h1, const double* h2, const double* theta, double* Q,
                               double* state, int lds) {
    /* variant 1: SFDTidal_Qmec (y0, A0, be, delta, phi, d1, d2, c, g, Q0, dx, dt);
       variant 2: SFDTidal_Qmec2_model.f90:61-187 (Be, bevs, delta, ne, d1, d2, c, g, Q0, dx, dt): width = Be, be = bevs */
    double be, delta, d1, d2, c, g, Q0, dx, dt, width, ne;
    for (int m = 0; m < nm; ++m) Q[m] = BAMGPU_UNDEF_RN;
    if (state)
        for (int s = 0; s < 3; ++s)
            for (int m = 0; m < nm; ++m) state[m + (size_t)s * lds] = BAMGPU_UNDEF_RN;
    /* variant 3: SFDTidal_Qmec0_model.f90:61-178 (Be, he, dzeta, ne, d1, d2, c, g, Q0, dx, dt): y = h - d, Ae = Be (he + hm);
       variant 4: SFDTidal_QmecMS_model.f90:62-154 (y0, A0, be, delta, phi, d1, d2, c, g, dx): steady Manning-Strickler discharge */
    double he = 0.0;
    if (variant == 2 || variant == 3) {
        width = theta[0]; be = theta[1]; delta = theta[2]; ne = theta[3]; d1 = theta[4]; d2 = theta[5]; c = theta[6];
        g = theta[7]; Q0 = theta[8]; dx = theta[9]; dt = theta[10];
        if (width <= 0.0 || ne <= 0.0 || c <= 0.0 || g <= 0.0 || dx <= 0.0 || dt <= 0.0) return 0;
        if (variant == 3) {
            he = be;
            if (he <= 0.0 || delta >= 0.0) return 0;
        }
    } else {
        double y0 = theta[0], A0 = theta[1], phi = theta[4];
        be = theta[2]; delta = theta[3]; d1 = theta[5]; d2 = theta[6]; c = theta[7]; g = theta[8];
        if (variant == 4) { Q0 = 0.0; dx = theta[9]; dt = 1.0; }
        else { Q0 = theta[9]; dx = theta[10]; dt = theta[11]; }
        if (A0 <= 0.0 || (y0 - be) <= 0.0 || phi <= 0.0 || c <= 0.0 || g <= 0.0 || dx <= 0.0 || dt <= 0.0) return 0;
        width = A0 / (y0 - be);
        double R0 = A0 / (width + 2.0 * (y0 - be));
        ne = sqrt(phi * (1.0 / (8.0 * g)) * A0 * pow(R0, c));
    }
    double* dy = (double*)malloc(((size_t)nm + 1) * 8 * 4);
    double *ym = dy + nm, *Ae = ym + nm, *Rhe = Ae + nm;
    int ok = 1;
    for (int m = 0; m < nm; ++m) {
        double y1 = (variant == 3) ? h1[m] - d1 : h1[m] + d1, y2 = (variant == 3) ? h2[m] - d2 : h2[m] + d2;
        dy[m] = y2 - y1;
        ym[m] = (y1 + y2) / 2.0;
        double depth = (variant == 3) ? he + ym[m] : ym[m] - be;
        Ae[m] = width * depth;
        double Pe = width + 2.0 * depth;
        Rhe[m] = Ae[m] / Pe;
        if (Ae[m] < 0.0 || Pe < 0.0 || Rhe[m] < 0.0) ok = 0;
    }
    if (!ok) { free(dy); return 0; }
    if (variant == 4) { /* :149-152 */
        for (int m = 0; m < nm; ++m) {
            Q[m] = (1.0 / ne) * pow(Rhe[m], 0.5 * c) * Ae[m] * sqrt(fabs((dy[m] + delta) / dx));
            if ((dy[m] + delta) > 0.0) Q[m] = -1.0 * Q[m];
        }
        free(dy);
        return 1;
    }
    if (nm > 0) {
        Q[0] = Q0;
        if (state) for (int s = 0; s < 3; ++s) state[(size_t)s * lds] = 0.0;
    }
    for (int m = 1; m < nm; ++m) {
        double pg, bf, ad;
        if (m == 1) {
            double dydt1 = (ym[m] - ym[m - 1]) / (dt);
            pg = -g * (Ae[m - 1] * (dy[m - 1] + delta) / dx);
            bf = -g * ((ne * ne) / (pow(Rhe[m - 1], c))) * Q[m - 1] * fabs(Q[m - 1]) / Ae[m - 1];
            ad = 2.0 * (Q[m - 1] / Ae[m - 1]) * width * dydt1;
        } else {
            double dydt1 = (ym[m] - ym[m - 2]) / (2 * dt), dydt2;
            if (m == 2) dydt2 = (ym[m - 1] - ym[m - 2]) / dt;
            else dydt2 = (ym[m - 1] - ym[m - 3]) / (2 * dt);
            pg = (1.5) * (-g * Ae[m - 1] * (dy[m - 1] + delta) / dx) - (0.5) * (-g * Ae[m - 2] * (dy[m - 2] + delta) / dx);
            bf = (1.5) * (-g * ((ne * ne) / (pow(Rhe[m - 1], c))) * Q[m - 1] * fabs(Q[m - 1]) / Ae[m - 1]) -
                 (0.5) * (-g * ((ne * ne) / (pow(Rhe[m - 2], c))) * Q[m - 2] * fabs(Q[m - 2]) / Ae[m - 2]);
            ad = (1.5) * (2.0 * (Q[m - 1] / Ae[m - 1]) * width * dydt1) - (0.5) * (2.0 * (Q[m - 2] / Ae[m - 2]) * width * dydt2);
        }
        Q[m] = Q[m - 1] + dt * (pg + bf + ad);
        if (state) {
            state[m] = dt * pg;
            state[m + (size_t)lds] = dt * bf;
            state[m + (size_t)2 * lds] = dt * ad;
        }
    }
    free(dy);
    return 1;
}

/* TxtMdl_Apply, TextFile_model.f90:1035-1098 */
static void textfile_apply(const oracle_t* o, int n, const double* IN, int ldx, const double* theta, double* OUT,
                           int ldy, int32_t* vfeas) {
    double val[128];
    for (int p = 0; p < o->vm_npar; ++p) val[o->vm_nin + p] = theta[p];
    for (int t = 0; t < n; ++t) {
        for (int x = 0; x < o->vm_nin; ++x) val[x] = IN[t + (size_t)x * ldx];
        for (int k = 0; k < o->vm_nout; ++k) {
            double r;
            int e = vm_eval(&o->vm[k], val, &r);
            if (e > 0 || r != r) { vfeas[t] = 0; r = BAMGPU_UNDEF_RN; }
            OUT[t + (size_t)k * ldy] = r;
        }
    }
}

/* ApplyModel, ModelLibrary_tools.f90:237-434: init to unfeasFlag, dispatch, flag infeasible rows.
   Returns feas = all(vfeas) through *feas. */
static int apply_model_s(const oracle_t* o, int n, const double* X, int ldx, const double* fullTheta, double* Y,
                         int ldy, double* dparModel, int32_t* vfeas, int* feas, double* state /* n x nState, ld = ldy, or NULL */) {
    for (int t = 0; t < n; ++t) vfeas[t] = 1;
    for (int y = 0; y < o->nY; ++y)
        for (int t = 0; t < n; ++t) Y[t + (size_t)y * ldy] = o->unfeas;
    for (int d = 0; d < o->nDparModel; ++d) dparModel[d] = o->unfeas;
    switch (o->model) {
        case BAMGPU_MDL_LINEAR: linear_apply(o, n, X, ldx, fullTheta, Y, ldy); break;
        case BAMGPU_MDL_BARATIN: baratin_apply(o, n, X, fullTheta, Y, dparModel, vfeas); break;
        case BAMGPU_MDL_SEDIMENT: sediment_apply(o, n, X, ldx, fullTheta, Y, vfeas); break;
        case BAMGPU_MDL_TEXTFILE: textfile_apply(o, n, X, ldx, fullTheta, Y, ldy, vfeas); break;
        case BAMGPU_MDL_VEGETATION:
            if (vegetation_apply(o, n, X, ldx, fullTheta, Y, ldy, dparModel, vfeas)) return 1;
            break;
        case BAMGPU_MDL_SUSPENDEDLOAD: suspendedload_apply(n, X, fullTheta, Y, vfeas); break;
        case BAMGPU_MDL_SWOT: swot_apply(o, n, X, ldx, fullTheta, Y, vfeas); break;
        case BAMGPU_MDL_SGD: sgd_apply(n, X, ldx, fullTheta, Y, vfeas); break;
        case BAMGPU_MDL_RECESSION_H: recession_apply(n, X, fullTheta, Y); break;
        case BAMGPU_MDL_ORTHO: ortho_apply(o, n, X, ldx, fullTheta, Y, ldy, dparModel, vfeas); break;
        case BAMGPU_MDL_HCS: hcs_apply(o, n, X, fullTheta, Y, vfeas); break;
        case BAMGPU_MDL_BARATINBAC: baratinbac_apply(o, n, X, fullTheta, Y, dparModel, vfeas); break;
        case BAMGPU_MDL_SFD: sfd_apply(o, n, X, ldx, fullTheta, Y, vfeas, state); break;
        case BAMGPU_MDL_SFDTIDAL_QMEC:
        case BAMGPU_MDL_SFDTIDAL_QMEC2:
        case BAMGPU_MDL_SFDTIDAL_QMEC0:
        case BAMGPU_MDL_SFDTIDAL_QMECMS:
            if (!sfdtidal_qmec_apply(o->model == BAMGPU_MDL_SFDTIDAL_QMEC2 ? 2 : (o->model == BAMGPU_MDL_SFDTIDAL_QMEC0 ? 3 : (o->model == BAMGPU_MDL_SFDTIDAL_QMECMS ? 4 : 1)),
                                     n, X, X + (size_t)ldx, fullTheta, Y, o->nState ? state : NULL, ldy))
                for (int t = 0; t < n; ++t) vfeas[t] = 0;
            break;
        case BAMGPU_MDL_SFDTIDAL_LAG:
            if (!sfdtidal_lag_apply(o->model_opt[0], n, X, ldx, fullTheta, Y))
                for (int t = 0; t < n; ++t) vfeas[t] = 0;
            break;
        case BAMGPU_MDL_ALGAE: {
            int r = algae_apply(n, X, ldx, fullTheta, Y, ldy, state);
            if (r < 0) return 1;
            if (!r)
                for (int t = 0; t < n; ++t) vfeas[t] = 0;
            break;
        }
        case BAMGPU_MDL_DYNVEG: {
            int r = dynveg_apply(o, n, X, ldx, fullTheta, Y, ldy, state, vfeas);
            if (r < 0) return 1;
            if (!r)
                for (int t = 0; t < n; ++t) vfeas[t] = 0;
            break;
        }
        case BAMGPU_MDL_MIXTURE: {
            int r = mixture_apply(o, n, X, ldx, fullTheta, Y, dparModel);
            if (r < 0) return 1;
            if (!r)
                for (int t = 0; t < n; ++t) vfeas[t] = 0;
            break;
        }
        case BAMGPU_MDL_SEGMENTATION:
            if (!segmentation_apply(o, n, X, fullTheta, Y))
                for (int t = 0; t < n; ++t) vfeas[t] = 0; /* scalar feas=.false. (ModelLibrary_tools.f90:362-364) */
            break;
        default: return 1;
    }
    int all = 1;
    for (int t = 0; t < n; ++t)
        if (!vfeas[t]) {
            all = 0;
            for (int y = 0; y < o->nY; ++y) Y[t + (size_t)y * ldy] = o->unfeas;
            if (state)
                for (int y = 0; y < o->nState; ++y) state[t + (size_t)y * ldy] = o->unfeas; /* :427-431 */
        }
    *feas = all;
    return 0;
}

static int apply_model(const oracle_t* o, int n, const double* X, int ldx, const double* fullTheta, double* Y,
                       int ldy, double* dparModel, int32_t* vfeas, int* feas) {
    return apply_model_s(o, n, X, ldx, fullTheta, Y, ldy, dparModel, vfeas, feas, NULL);
}

/* ------------------------------------------------------------------------------------------
 * create / layout  (LoadBamObjects, src/BaM_tools.f90:133-420)
 * ------------------------------------------------------------------------------------------ */

static int model_from_name(const char* name) {
    if (!name) return 0;
    if (strncmp(name, "MDL_", 4) == 0) name += 4;
    if (!strcmp(name, "Linear")) return BAMGPU_MDL_LINEAR;
    if (!strcmp(name, "BaRatin")) return BAMGPU_MDL_BARATIN;
    if (!strcmp(name, "Sediment")) return BAMGPU_MDL_SEDIMENT;
    if (!strcmp(name, "TextFile")) return BAMGPU_MDL_TEXTFILE;
    if (!strcmp(name, "Vegetation")) return BAMGPU_MDL_VEGETATION;
    if (!strcmp(name, "SuspendedLoad")) return BAMGPU_MDL_SUSPENDEDLOAD;
    if (!strcmp(name, "SWOT")) return BAMGPU_MDL_SWOT;
    if (!strcmp(name, "SGD")) return BAMGPU_MDL_SGD;
    if (!strcmp(name, "Orthorectification")) return BAMGPU_MDL_ORTHO;
    if (!strcmp(name, "Recession_h")) return BAMGPU_MDL_RECESSION_H;
    if (!strcmp(name, "HydraulicControl_section")) return BAMGPU_MDL_HCS;
    if (!strcmp(name, "Segmentation")) return BAMGPU_MDL_SEGMENTATION;
    if (!strcmp(name, "BaRatinBAC")) return BAMGPU_MDL_BARATINBAC;
    if (!strcmp(name, "SFD")) return BAMGPU_MDL_SFD;
    if (!strcmp(name, "Mixture")) return BAMGPU_MDL_MIXTURE;
    if (!strcmp(name, "SFDTidal_Qmec")) return BAMGPU_MDL_SFDTIDAL_QMEC;
    if (!strcmp(name, "SFDTidal_Qmec2")) return BAMGPU_MDL_SFDTIDAL_QMEC2;
    if (!strcmp(name, "AlgaeBiomass")) return BAMGPU_MDL_ALGAE;
    if (!strcmp(name, "DynamicVegetation")) return BAMGPU_MDL_DYNVEG;
    if (!strcmp(name, "SFDTidal_Qmec0")) return BAMGPU_MDL_SFDTIDAL_QMEC0;
    if (!strcmp(name, "SFDTidal_QmecMS")) return BAMGPU_MDL_SFDTIDAL_QMECMS;
    if (!strcmp(name, "SFDTidal") || !strcmp(name, "SFDTidal2") || !strcmp(name, "SFDTidal4") || !strcmp(name, "SFDTidalJones"))
        return BAMGPU_MDL_SFDTIDAL_LAG;
    return 0;
}

int oracle_create(oracle_t** out, const bamgpu_problem* p, char* mess) {
    if (mess) mess[0] = 0;
    oracle_t* o = (oracle_t*)calloc(1, sizeof *o);
    o->model = model_from_name(p->model_id);
    if (!o->model) {
        setmess(mess, "oracle: ApplyModel: Fatal: Unavailable [model%ID]");
        free(o);
        return 1;
    }
    int n = o->n = p->nobs, nX = o->nX = p->nX, nY = o->nY = p->nY, nt = o->ntheta = p->ntheta;
    size_t sx = (size_t)n * nX, sy = (size_t)n * nY;
    o->X = dup_mem(p->X, sx * 8); o->Xu = dup_mem(p->Xu, sx * 8); o->Xb = dup_mem(p->Xb, sx * 8);
    o->Xbindx = dup_mem(p->Xbindx, sx * 4);
    o->Y = dup_mem(p->Y, sy * 8); o->Yu = dup_mem(p->Yu, sy * 8); o->Yb = dup_mem(p->Yb, sy * 8);
    o->Ybindx = dup_mem(p->Ybindx, sy * 4);
    o->partype = dup_mem(p->partype, nt * 4); o->nval = dup_mem(p->nval, nt * 4);
    o->fix_value = dup_mem(p->fix_value, nt * 8);
    o->theta_off = malloc((nt + 1) * 4);
    o->nFit = 0; o->nVar = 0;
    for (int i = 0; i < nt; ++i) {
        if (o->partype[i] == BAMGPU_PAR_FIX) o->theta_off[i] = -1;
        else { o->theta_off[i] = o->nFit; o->nFit += o->nval[i]; }
        if (o->partype[i] == BAMGPU_PAR_VAR) o->nVar++;
    }
    o->var_indx = malloc(((size_t)n * nt + 1) * 4);
    for (size_t q = 0; q < (size_t)n * nt; ++q) o->var_indx[q] = (p->var_indx && o->nVar) ? p->var_indx[q] : 1;
    for (int i = 0; i < nt; ++i)
        for (int t = 0; t < n; ++t) {
            int v = o->var_indx[t + (size_t)i * n];
            if (o->partype[i] != BAMGPU_PAR_VAR) v = o->var_indx[t + (size_t)i * n] = 1;
            if (v <= 0 || v > o->nval[i]) { setmess(mess, "oracle: Config_Finalize: VAR index out of range"); return 1; }
        }
    o->sigma_func = dup_mem(p->sigma_func, nY * 4); o->nremnant = dup_mem(p->nremnant, nY * 4);
    o->gamma_off = malloc((nY + 1) * 4);
    o->nGamma = 0;
    for (int j = 0; j < nY; ++j) {
        if (sigmafunk_npar(o->sigma_func[j]) != o->nremnant[j]) { setmess(mess, "oracle: wrong number of remnant parameters"); return 1; }
        o->gamma_off[j] = o->nGamma; o->nGamma += o->nremnant[j];
    }
    o->pt_dist = dup_mem(p->prior_theta_dist, o->nFit * 4);
    o->pt_par = dup_mem(p->prior_theta_par, (size_t)o->nFit * BAMGPU_PRIOR_NPAR * 8);
    o->pg_dist = dup_mem(p->prior_gamma_dist, o->nGamma * 4);
    o->pg_par = dup_mem(p->prior_gamma_par, (size_t)o->nGamma * BAMGPU_PRIOR_NPAR * 8);
    o->kM = o->nFit + o->nGamma;
    o->priorM = p->priorM ? dup_mem(p->priorM, (size_t)o->kM * o->kM * 8) : NULL;
    o->xtra_i = dup_mem(p->xtra_i, (size_t)p->n_xtra_i * 4); o->n_xtra_i = p->n_xtra_i;
    o->xtra_d = dup_mem(p->xtra_d, (size_t)p->n_xtra_d * 8); o->n_xtra_d = p->n_xtra_d;
    o->mv = p->mv; o->unfeas = p->unfeas_flag;

    /* latent inputs, biases: :230-253 */
    o->Xerror = calloc(nX + 1, 4); o->latIdx = malloc((sx + 1) * 4);
    o->nXb = calloc(nX + 1, 4); o->nYb = calloc(nY + 1, 4);
    o->off_xb = calloc(nX + 1, 4); o->off_yb = calloc(nY + 1, 4);
    int pos = o->nFit + o->nGamma;
    o->nLatent = 0;
    for (int j = 0; j < nX; ++j)
        for (int i = 0; i < n; ++i) {
            size_t q = i + (size_t)j * n;
            if (o->Xu[q] > 0.0) { o->latIdx[q] = pos++; o->nLatent++; o->Xerror[j] = 1; }
            else o->latIdx[q] = -1;
        }
    o->nXbTot = o->nYbTot = 0;
    for (int j = 0; j < nX; ++j) {
        int mx = 0;
        for (int i = 0; i < n; ++i) if (o->Xbindx[i + (size_t)j * n] > mx) mx = o->Xbindx[i + (size_t)j * n];
        o->nXb[j] = mx; o->off_xb[j] = pos; pos += mx; o->nXbTot += mx;
    }
    for (int j = 0; j < nY; ++j) { /* NOTE: Packer layout (:4235-4240); UnfoldParVector's cursor quirk (:3460) not reproduced */
        int mx = 0;
        for (int i = 0; i < n; ++i) if (o->Ybindx[i + (size_t)j * n] > mx) mx = o->Ybindx[i + (size_t)j * n];
        o->nYb[j] = mx; o->off_yb[j] = pos; pos += mx; o->nYbTot += mx;
    }
    o->P = pos;

    /* model xtra + nDpar: XtraSetup, ModelLibrary_tools.f90:546-760 */
    o->nDparModel = 0; o->nState = 0;
    switch (o->model) {
        case BAMGPU_MDL_LINEAR:
            if (nt != nX * nY) { setmess(mess, "oracle: Linear: ntheta /= nX*nY"); return 1; }
            break;
        case BAMGPU_MDL_BARATINBAC:
        case BAMGPU_MDL_BARATIN: {
            int nc = nt / 3;
            if (nX != 1 || nY != 1 || nt != 3 * nc || nc < 1 || nc > MAXCONTROL || p->n_xtra_i != nc * nc) {
                setmess(mess, "oracle: BaRatin_Apply:Fatal:Incorrect size [theta]"); return 1;
            }
            o->ncontrol = nc;
            for (int r = 0; r < nc; ++r)
                for (int c = 0; c < nc; ++c) o->M[r * nc + c] = p->xtra_i[r + c * nc];
            o->nDparModel = nc;
            if (o->model == BAMGPU_MDL_BARATINBAC) {
                if (p->n_xtra_d != 1) { setmess(mess, "oracle: BaRatinBAC: hmax missing"); return 1; }
                o->model_optd = p->xtra_d[0];
                o->nDparModel = 2 * nc;
            }
            break;
        }
        case BAMGPU_MDL_SEDIMENT: {
            if (p->n_xtra_i != 7 || p->n_xtra_d != 4 || nY != 1) { setmess(mess, "oracle: Sediment: bad xtra"); return 1; }
            o->sed_formula = p->xtra_i[0]; o->sed_css = p->xtra_i[1]; o->sed_co = p->xtra_i[2];
            int np = sed_nbase(o->sed_formula) + (o->sed_css == BAMGPU_CSS_CONSTANT ? 1 : 0), nin = 2;
            for (int i = 0; i < 4; ++i) {
                o->sed_ppo[i] = p->xtra_i[3 + i]; o->sed_pseudo[i] = p->xtra_d[i];
                if (o->sed_ppo[i] == BAMGPU_PPO_PAR) ++np;
                if (o->sed_ppo[i] == BAMGPU_PPO_IN) ++nin;
            }
            if (o->sed_css != BAMGPU_CSS_CONSTANT && o->sed_co == BAMGPU_CO_PAR) np += sed_ncoeff(o->sed_css);
            if (np != nt || nin != nX) { setmess(mess, "oracle: Sediment: wrong size [theta] or [IN]"); return 1; }
            break;
        }
        case BAMGPU_MDL_TEXTFILE: {
            int e = textfile_load(o, p->xtra_text, mess);
            if (e) return e;
            if (o->vm_nin != nX || o->vm_npar != nt || o->vm_nout != nY) { setmess(mess, "oracle: TextFile: size mismatch"); return 1; }
            break;
        }
        case BAMGPU_MDL_VEGETATION: { /* xtra_i = {growth, nYear, itmax}; xtra_d = {xscale, fscale, tolX, tolF} */
            if (p->n_xtra_i != 3 || p->n_xtra_d != 4) { setmess(mess, "oracle: Vegetation_XtraRead: bad xtra"); return 1; }
            o->veg_growth = p->xtra_i[0]; o->veg_nyear = p->xtra_i[1]; o->veg_itmax = p->xtra_i[2];
            o->veg_xscale = p->xtra_d[0]; o->veg_fscale = p->xtra_d[1]; o->veg_tolx = p->xtra_d[2]; o->veg_tolf = p->xtra_d[3];
            if (o->veg_growth < 1 || o->veg_growth > 4) { setmess(mess, "oracle: GetNpar_growth:Fatal:Unavailable [growthFormula]"); return 1; }
            int temporal = o->veg_growth == BAMGPU_VEG_YIN1_TEMPORAL;
            if (o->veg_nyear < 1 || o->veg_nyear > VEG_MAXYEAR) { setmess(mess, "oracle: Vegetation: nYear out of range"); return 1; }
            int np = 5 + 2 + veg_ngrowth(o->veg_growth) + (temporal ? 1 : o->veg_nyear); /* GetParNumber :46-86 */
            if (nt != np) { setmess(mess, "oracle: Vegetation_Apply:wrong size[theta]"); return 1; }
            if (nX != (temporal ? 2 : 4) || nY < 1 || nY > 5) { setmess(mess, "oracle: Vegetation: wrong nX or nY"); return 1; }
            o->nDparModel = temporal ? 1 : o->veg_nyear + 1; /* ModelLibrary_tools.f90:625-632 */
            break;
        }
        case BAMGPU_MDL_SUSPENDEDLOAD: case BAMGPU_MDL_SGD:
            if (nt != 5 || nY != 1) { setmess(mess, "oracle: wrong size [theta]"); return 1; }
            break;
        case BAMGPU_MDL_RECESSION_H:
            if (nt != 4 || nX != 1 || nY != 1) { setmess(mess, "oracle: wrong size [theta]"); return 1; }
            break;
        case BAMGPU_MDL_SWOT:
            if (p->n_xtra_i != 1 || nt != 5 || nY != 1) { setmess(mess, "oracle: SWOT: bad xtra or sizes"); return 1; }
            o->model_opt[0] = p->xtra_i[0];
            if (nX != (o->model_opt[0] == BAMGPU_SWOT_HS ? 2 : 3)) { setmess(mess, "oracle: SWOT: wrong nX"); return 1; }
            break;
        case BAMGPU_MDL_ORTHO:
            if (p->n_xtra_i != 2 || nX != 3 || nY != 2 || nt != 11 + p->xtra_i[1]) { setmess(mess, "oracle: Ortho: bad xtra or sizes"); return 1; }
            o->model_opt[0] = p->xtra_i[0]; o->model_opt[1] = p->xtra_i[1];
            o->nDparModel = 20;
            break;
        case BAMGPU_MDL_SFD:
            if (p->n_xtra_i != 2 || p->n_xtra_d != 5 || nt != 8 || nX != 2 || nY != 1) { setmess(mess, "oracle: SFD: bad xtra or sizes"); return 1; }
            o->veg_itmax = p->xtra_i[1];
            o->model_optd = p->xtra_d[0]; o->veg_xscale = p->xtra_d[1]; o->veg_fscale = p->xtra_d[2];
            o->veg_tolx = p->xtra_d[3]; o->veg_tolf = p->xtra_d[4];
            o->nState = 1;
            break;
        case BAMGPU_MDL_HCS:
            o->hcs_n = p->n_xtra_d / 3;
            if (o->hcs_n < 3 || p->n_xtra_d != 3 * o->hcs_n || nt != 2 || nX != 1 || nY != 1) { setmess(mess, "oracle: HydraulicControl_section: bad xtra or sizes"); return 1; }
            break;
        case BAMGPU_MDL_SFDTIDAL_QMEC: /* SFDTidal_Qmec_GetParNumber (:31-59): 12; XtraSetup (ModelLibrary_tools.f90:672-676) */
        case BAMGPU_MDL_SFDTIDAL_QMEC2: /* SFDTidal_Qmec2_GetParNumber: 11; XtraSetup (:677-681) */
            if (nt != (o->model == BAMGPU_MDL_SFDTIDAL_QMEC2 ? 11 : 12) || nX != 2 || nY != 1) { setmess(mess, "oracle: SFDTidal_Qmec: bad sizes"); return 1; }
            o->nState = 3;
            break;
        case BAMGPU_MDL_SFDTIDAL_QMEC0:  /* SFDTidal_Qmec0_GetParNumber: 11; no states (ModelLibrary_tools.f90:668-671) */
        case BAMGPU_MDL_SFDTIDAL_QMECMS: /* SFDTidal_QmecMS_GetParNumber: 10; no states (:682-685) */
            if (nt != (o->model == BAMGPU_MDL_SFDTIDAL_QMEC0 ? 11 : 10) || nX != 2 || nY != 1) { setmess(mess, "oracle: SFDTidal_Qmec: bad sizes"); return 1; }
            break;
        case BAMGPU_MDL_SFDTIDAL_LAG: { /* 8 parameters each, no xtra, no states (ModelLibrary_tools.f90:647-667) */
            const char* id = p->model_id;
            if (!strncmp(id, "MDL_", 4)) id += 4;
            int variant = !strcmp(id, "SFDTidal") ? 1 : (!strcmp(id, "SFDTidal2") ? 2 : (!strcmp(id, "SFDTidal4") ? 3 : 4));
            if (nt != 8 || nX != (variant == 4 ? 3 : 2) || nY != 1) { setmess(mess, "oracle: SFDTidal: bad sizes"); return 1; }
            o->model_opt[0] = variant;
            break;
        }
        case BAMGPU_MDL_ALGAE: /* AlgaeBiomass_GetParNumber: 18; XtraSetup: 5 state variables (ModelLibrary_tools.f90:697-702) */
            if (nt != 18 || nX != 5 || nY != 2) { setmess(mess, "oracle: AlgaeBiomass: bad sizes"); return 1; }
            o->nState = 5;
            break;
        case BAMGPU_MDL_DYNVEG: /* DynamicVegetation_GetParNumber: 18 + 4 + 2; XtraRead (:193-247); 5 states (ModelLibrary_tools.f90:703-707) */
            if (p->n_xtra_i != 1 || p->n_xtra_d != 4) { setmess(mess, "oracle: DynamicVegetation: bad xtra"); return 1; }
            if (nt != 24 || nX != 5 || nY != 3) { setmess(mess, "oracle: DynamicVegetation: bad sizes"); return 1; }
            o->veg_itmax = p->xtra_i[0];
            o->veg_xscale = p->xtra_d[0]; o->veg_fscale = p->xtra_d[1]; o->veg_tolx = p->xtra_d[2]; o->veg_tolf = p->xtra_d[3];
            o->nState = 5;
            break;
        case BAMGPU_MDL_MIXTURE: { /* Mixture_GetParNumber (:33-72), XtraSetup (ModelLibrary_tools.f90:580-586) */
            if (p->n_xtra_i != 4) { setmess(mess, "oracle: Mixture: bad xtra"); return 1; }
            int nS = p->xtra_i[0], nEvt = p->xtra_i[1], nElt = p->xtra_i[2], miss = p->xtra_i[3] != 0;
            if (nS < 1 || nS > 62 || nEvt < 1 || nEvt > 64 || nElt < 1 || nX != 2 + nS || nY != 1 ||
                nt != (miss ? nS * nEvt + nElt : (nS - 1) * nEvt) || nt > 256) { setmess(mess, "oracle: Mixture: bad sizes"); return 1; }
            o->model_opt[0] = nS; o->model_opt[1] = nEvt; o->model_opt[2] = nElt; o->model_opt[3] = miss;
            o->nDparModel = nEvt;
            break;
        }
        case BAMGPU_MDL_SEGMENTATION:
            if (p->n_xtra_i != 3 || p->n_xtra_d != 1 || nX != 1 || nY != 1 || nt != 2 * p->xtra_i[0] - 1 || p->xtra_i[0] > 32) { setmess(mess, "oracle: Segmentation: bad xtra or sizes"); return 1; }
            o->model_opt[0] = p->xtra_i[0]; o->model_opt[1] = p->xtra_i[1]; o->model_opt[2] = p->xtra_i[2];
            o->model_optd = p->xtra_d[0];
            break;
    }

    /* VAR combinations: :346-379 (order of first appearance) */
    o->comboOf = malloc((n + 1) * 4);
    o->comboRow = malloc((n + 1) * 4);
    o->nCombo = 0;
    if (o->nVar == 0 || n == 0) {
        o->nCombo = 1; o->comboRow[0] = 0;
        for (int t = 0; t < n; ++t) o->comboOf[t] = 0;
    } else {
        for (int t = 0; t < n; ++t) {
            int found = -1;
            for (int k = 0; k < o->nCombo && found < 0; ++k) {
                int r = o->comboRow[k], same = 1;
                for (int i = 0; i < nt && same; ++i)
                    if (o->var_indx[t + (size_t)i * n] != o->var_indx[r + (size_t)i * n]) same = 0;
                if (same) found = k;
            }
            if (found < 0) { found = o->nCombo; o->comboRow[o->nCombo++] = t; }
            o->comboOf[t] = found;
        }
    }
    o->nDpar = o->nDparModel * o->nCombo;
    *out = o;
    return 0;
}

void oracle_destroy(oracle_t* o) {
    if (!o) return;
    free(o->X); free(o->Xu); free(o->Xb); free(o->Y); free(o->Yu); free(o->Yb); free(o->Xbindx); free(o->Ybindx);
    free(o->partype); free(o->nval); free(o->theta_off); free(o->fix_value); free(o->var_indx);
    free(o->pt_dist); free(o->pg_dist); free(o->pt_par); free(o->pg_par); free(o->sigma_func); free(o->nremnant);
    free(o->gamma_off); free(o->priorM); free(o->xtra_i); free(o->xtra_d); free(o->nXb); free(o->nYb);
    free(o->off_xb); free(o->off_yb); free(o->Xerror); free(o->latIdx); free(o->comboOf); free(o->comboRow);
    free(o);
}

int oracle_get_sizes(const oracle_t* o, bamgpu_sizes* s) {
    s->nFit = o->nFit; s->nGamma = o->nGamma; s->nLatent = o->nLatent; s->nXbias = o->nXbTot; s->nYbias = o->nYbTot;
    s->nInfer = o->P; s->nDpar = o->nDpar; s->nCombo = o->nCombo; s->nState = o->nState;
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * log-posterior
 * ------------------------------------------------------------------------------------------ */

/* fullTheta for observation t (ApplyModel_BaM, src/BaM_tools.f90:4133-4159) */
static void full_theta(const oracle_t* o, const double* theta_fit, int t, double* full) {
    for (int i = 0; i < o->ntheta; ++i) {
        if (o->partype[i] == BAMGPU_PAR_FIX) full[i] = o->fix_value[i];
        else full[i] = theta_fit[o->theta_off[i] + (o->n > 0 ? o->var_indx[t + (size_t)i * o->n] : 1) - 1];
    }
}

/* GetPriorLogPdf, src/BaM_tools.f90:3006-3138 */
static int prior_logpdf(const oracle_t* o, const double* x, double* prior, int* feas, int* isnull) {
    const double* theta = x;
    const double* gamma = x + o->nFit;
    double pr = 0.0, lp;
    *feas = 1; *isnull = 0;
    for (int j = 0; j < o->nY; ++j) { /* remnant parameters first (:3062-3069) */
        double s = 0.0; /* GetLogPrior = sum of the log-pdfs of one list */
        for (int m = 0; m < o->nremnant[j]; ++m) {
            int g = o->gamma_off[j] + m;
            int e = oracle_logpdf(o->pg_dist[g], o->pg_par + (size_t)g * BAMGPU_PRIOR_NPAR, gamma[g], &lp, feas, isnull);
            if (e > 0) { *feas = 0; return e; }
            if (!*feas || *isnull) return 0;
            s = s + lp;
        }
        pr = pr + s;
    }
    {
        double s = 0.0;
        for (int k = 0; k < o->nFit; ++k) { /* theta (:3072-3077) */
            int e = oracle_logpdf(o->pt_dist[k], o->pt_par + (size_t)k * BAMGPU_PRIOR_NPAR, theta[k], &lp, feas, isnull);
            if (e > 0) { *feas = 0; return e; }
            if (!*feas || *isnull) return 0;
            s = s + lp;
        }
        pr = pr + s;
    }
    if (o->priorM) { /* Gaussian copula (:3080-3112): theta first, then gamma */
        int k = o->kM;
        double* v = (double*)malloc((size_t)k * 8);
        for (int i = 0; i < o->nFit; ++i) {
            gcop_getv(o->pt_dist[i], o->pt_par + (size_t)i * BAMGPU_PRIOR_NPAR, theta[i], &v[i], feas);
            if (!*feas) { free(v); return 0; }
        }
        for (int g = 0; g < o->nGamma; ++g) {
            gcop_getv(o->pg_dist[g], o->pg_par + (size_t)g * BAMGPU_PRIOR_NPAR, gamma[g], &v[o->nFit + g], feas);
            if (!*feas) { free(v); return 0; }
        }
        /* dot_product(matmul(v,M),v): w_c = sum_r v_r M(r,c) */
        double q = 0.0;
        for (int c = 0; c < k; ++c) {
            double w = 0.0;
            for (int r = 0; r < k; ++r) w = w + v[r] * o->priorM[r + (size_t)c * k];
            q = q + w * v[c];
        }
        pr = pr - 0.5 * q;
        free(v);
    }
    static const double n01[2] = {0.0, 1.0};
    for (int j = 0; j < o->nX; ++j)
        for (int i = 0; i < o->nXb[j]; ++i) {
            oracle_logpdf(BAMGPU_DIST_GAUSSIAN, n01, x[o->off_xb[j] + i], &lp, feas, isnull);
            pr = pr + lp;
        }
    for (int j = 0; j < o->nY; ++j)
        for (int i = 0; i < o->nYb[j]; ++i) {
            oracle_logpdf(BAMGPU_DIST_GAUSSIAN, n01, x[o->off_yb[j] + i], &lp, feas, isnull);
            pr = pr + lp;
        }
    *prior = pr;
    return 0;
}

/* ApplyModel_BaM on the calibration data (src/BaM_tools.f90:4115-4177) */
static int apply_model_bam(const oracle_t* o, const double* Xtrue, const double* theta_fit, double* bigY,
                           double* dpar, int32_t* vfeas, int* feas) {
    int n = o->n;
    double full[256], dmod[MAXCONTROL + 64];
    for (int d = 0; d < o->nDpar; ++d) dpar[d] = o->unfeas;
    if (o->nVar == 0) {
        full_theta(o, theta_fit, 0, full);
        int e = apply_model(o, n, Xtrue, n, full, bigY, n, dpar, vfeas, feas);
        return e;
    }
    int k = 0;
    double xrow[16], yrow[16];
    for (int t = 0; t < n; ++t) { /* one model call per observation (:4152-4174) */
        full_theta(o, theta_fit, t, full);
        for (int j = 0; j < o->nX; ++j) xrow[j] = Xtrue[t + (size_t)j * n];
        int32_t vf;
        int e = apply_model(o, 1, xrow, 1, full, yrow, 1, dmod, &vf, feas);
        if (e > 0) return e;
        if (!*feas) return 0;
        for (int j = 0; j < o->nY; ++j) bigY[t + (size_t)j * n] = yrow[j];
        if (k < o->nCombo && t == o->comboRow[k]) {
            for (int d = 0; d < o->nDparModel; ++d) dpar[k * o->nDparModel + d] = dmod[d];
            ++k;
        }
    }
    *feas = 1;
    return 0;
}

int oracle_logpost(const oracle_t* o, const double* x, double* fx, double* prior_out, double* lkh_out,
                   double* hyper_out, int32_t* feas_out, int32_t* isnull_out, double* dpar_out, int use_ld,
                   char* mess) {
    int n = o->n, nX = o->nX, nY = o->nY;
    int feas = 1, isnull = 0;
    double prior = BAMGPU_UNDEF_RN, lkh = BAMGPU_UNDEF_RN, hyper = BAMGPU_UNDEF_RN;
    if (mess) mess[0] = 0;
    *fx = BAMGPU_UNDEF_RN;
    double* dpar = (double*)malloc(((size_t)o->nDpar + 1) * 8);
    for (int d = 0; d < o->nDpar; ++d) dpar[d] = o->unfeas;
    /* UnfoldParVector (:3386-3463): Xtrue = X copy, latent cells from x, un-bias non-latent biased cells */
    double* Xtrue = (double*)malloc(((size_t)n * nX + 1) * 8);
    memcpy(Xtrue, o->X, (size_t)n * nX * 8);
    for (int j = 0; j < nX; ++j)
        for (int i = 0; i < n; ++i) {
            size_t q = i + (size_t)j * n;
            if (o->latIdx[q] >= 0) Xtrue[q] = x[o->latIdx[q]];
        }
    for (int j = 0; j < nX; ++j) {
        if (o->nXb[j] == 0) continue;
        for (int i = 0; i < n; ++i) {
            size_t q = i + (size_t)j * n;
            if (o->Xbindx[q] > 0 && o->Xu[q] == 0.0) Xtrue[q] = o->X[q] - o->Xb[q] * x[o->off_xb[j] + o->Xbindx[q] - 1];
        }
    }
    double* bigY = (double*)malloc(((size_t)n * nY + 1) * 8);
    int32_t* vfeas = (int32_t*)malloc(((size_t)n + 1) * 4);
    int err = 0;

    /* GetPostLogPdf (:3310-3382): prior, likelihood, hyper with early exits */
    err = prior_logpdf(o, x, &prior, &feas, &isnull);
    if (err > 0) { feas = 0; goto done; }
    if (!feas || isnull) goto done;

    { /* GetLogLikelihood (:3142-3229) */
        err = apply_model_bam(o, Xtrue, x, bigY, dpar, vfeas, &feas);
        if (err > 0) { setmess(mess, "oracle: ApplyModel failed"); feas = 0; goto done; }
        if (!feas) goto done;
        const double* gamma = x + o->nFit;
        double acc = 0.0;
        long double accl = 0.0L;
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < nY; ++j) {
                size_t q = i + (size_t)j * n;
                if (o->Y[q] == o->mv) continue;
                double yhat = bigY[q];
                double sig = sigmafunk(o->sigma_func[j], gamma + o->gamma_off[j], yhat);
                if (sig <= 0.0) { feas = 0; goto done; }
                sig = sqrt(o->Yu[q] * o->Yu[q] + sig * sig);
                double mu = yhat;
                if (o->Ybindx[q] != 0) mu = yhat + o->Yb[q] * x[o->off_yb[j] + o->Ybindx[q] - 1];
                if (!(sig > 0.0)) { feas = 0; goto done; } /* GetPdf(GAUSS): feas <=> sigma>0 */
                double z = (o->Y[q] - mu) / sig;
                double lp = -LOG_SQRT_2PI - log(sig) - 0.5 * z * z;
                acc = acc + lp;
                accl += (long double)lp;
            }
        lkh = use_ld ? (double)accl : acc;
    }
    { /* GetHyperLogPdf (:3233-3306): j-major */
        double acc = 0.0;
        long double accl = 0.0L;
        for (int j = 0; j < nX; ++j) {
            if (!o->Xerror[j]) continue;
            for (int i = 0; i < n; ++i) {
                size_t q = i + (size_t)j * n;
                if (o->Xu[q] == 0.0) continue;
                double mu = Xtrue[q];
                if (o->Xbindx[q] != 0) mu = Xtrue[q] + o->Xb[q] * x[o->off_xb[j] + o->Xbindx[q] - 1];
                if (!(o->Xu[q] > 0.0)) { feas = 0; goto done; }
                double z = (o->X[q] - mu) / o->Xu[q];
                double lp = -LOG_SQRT_2PI - log(o->Xu[q]) - 0.5 * z * z;
                acc = acc + lp;
                accl += (long double)lp;
            }
        }
        hyper = use_ld ? (double)accl : acc;
    }
    *fx = use_ld ? (double)((long double)prior + (long double)lkh + (long double)hyper) : prior + lkh + hyper;
done:
    if (prior_out) *prior_out = prior;
    if (lkh_out) *lkh_out = lkh;
    if (hyper_out) *hyper_out = hyper;
    *feas_out = feas;
    *isnull_out = isnull;
    if (dpar_out) memcpy(dpar_out, dpar, (size_t)o->nDpar * 8);
    free(dpar); free(Xtrue); free(bigY); free(vfeas);
    return err > 0 ? err : 0;
}

/* GetSim_calibration (src/BaM_tools.f90:3619-3680): Xtrue (n x nX) and Ysim (n x nY) for one vector */
int oracle_calibration_run_states(const oracle_t* o, const double* x, double* Xtrue, double* Ysim, double* state,
                                  int32_t* feas_out, char* mess);
int oracle_calibration_run(const oracle_t* o, const double* x, double* Xtrue, double* Ysim, int32_t* feas_out, char* mess) {
    return oracle_calibration_run_states(o, x, Xtrue, Ysim, NULL, feas_out, mess);
}

/* GetSim_calibration (src/BaM_tools.f90:3619-3680) with the state variables; VAR parameters: states are not returned */
int oracle_calibration_run_states(const oracle_t* o, const double* x, double* Xtrue, double* Ysim, double* state,
                                  int32_t* feas_out, char* mess) {
    int n = o->n, nX = o->nX;
    if (mess) mess[0] = 0;
    double* xt = (double*)malloc(((size_t)n * nX + 1) * 8);
    memcpy(xt, o->X, (size_t)n * nX * 8);
    for (int j = 0; j < nX; ++j)
        for (int i = 0; i < n; ++i) {
            size_t q = i + (size_t)j * n;
            if (o->latIdx[q] >= 0) xt[q] = x[o->latIdx[q]];
            else if (o->nXb[j] > 0 && o->Xbindx[q] > 0) xt[q] = o->X[q] - o->Xb[q] * x[o->off_xb[j] + o->Xbindx[q] - 1];
        }
    if (Xtrue) memcpy(Xtrue, xt, (size_t)n * nX * 8);
    double* dpar = (double*)malloc(((size_t)o->nDpar + 1) * 8);
    int32_t* vfeas = (int32_t*)malloc(((size_t)n + 1) * 4);
    int feas = 1;
    for (size_t q = 0; q < (size_t)n * o->nY; ++q) Ysim[q] = o->unfeas;
    int e;
    if (state && o->nState > 0 && o->nVar == 0) {
        double full[256];
        for (size_t q = 0; q < (size_t)n * o->nState; ++q) state[q] = o->unfeas;
        full_theta(o, x, 0, full);
        e = apply_model_s(o, n, xt, n, full, Ysim, n, dpar, vfeas, &feas, state);
    } else e = apply_model_bam(o, xt, x, Ysim, dpar, vfeas, &feas);
    *feas_out = feas;
    free(xt); free(dpar); free(vfeas);
    if (e > 0) setmess(mess, "oracle: ApplyModel failed");
    return e > 0 ? e : 0;
}

int oracle_logpost_batch(const oracle_t* o, int K, const double* x, double* fx, int32_t* feas, int32_t* isnull,
                         double* dpar, int nthreads, char* mess) {
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        char m[BAMGPU_MESS_LEN];
        int e = oracle_logpost(o, x + (size_t)k * o->P, &fx[k], NULL, NULL, NULL, &feas[k], &isnull[k],
                               dpar ? dpar + (size_t)k * o->nDpar : NULL, 0, m);
        if (e > 0) {
#pragma omp critical
            { err = e; setmess(mess, m); }
        }
    }
    return err;
}

int oracle_apply_model(const oracle_t* o, int n, const double* X, const double* theta_fit, double* Y, double* dpar,
                       int32_t* feas_rows, char* mess) {
    if (o->nVar) { setmess(mess, "oracle: prediction with VAR parameters is not allowed (BaM_Message 13)"); return 1; }
    double full[256], dmod[MAXCONTROL + 64];
    full_theta(o, theta_fit, 0, full);
    int feas;
    int e = apply_model(o, n, X, n, full, Y, n, dmod, feas_rows, &feas);
    if (dpar) for (int d = 0; d < o->nDparModel; ++d) dpar[d] = dmod[d];
    return e;
}

/* ------------------------------------------------------------------------------------------
 * RNG: Philox4x32-10 counter-based generator (Salmon et al. 2011), Box-Muller normals.
 * Specified here; the device implements the same specification independently.
 * ------------------------------------------------------------------------------------------ */

void oracle_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* 53-bit uniform in (0,1): ((hi>>5)*2^26 + (lo>>6) + 0.5) / 2^53 */
static double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}

double oracle_uniform(uint64_t seed, uint64_t a, uint32_t b, uint32_t c) {
    uint32_t w[4];
    oracle_philox4x32((uint32_t)a, (uint32_t)(a >> 32), b, c, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    return u53(w[0], w[1]);
}

double oracle_normal(uint64_t seed, uint64_t a, uint32_t b, uint32_t c) {
    uint32_t w[4];
    oracle_philox4x32((uint32_t)a, (uint32_t)(a >> 32), b, c, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    double u1 = u53(w[0], w[1]), u2 = u53(w[2], w[3]);
    return sqrt(-2.0 * log(u1)) * cos(6.28318530717958647692 * u2);
}

/* Structural-error noise of the prediction (the reference draws it from BMSL's generator, un-vendored: only the
   distribution is shared).  Declared specification: inputs 4q..4q+3 of replication rep / output y share one
   Philox block (counter (rep, t>>2, y)); words (0,1) serve inputs 4q and 4q+1, words (2,3) inputs 4q+2 and 4q+3;
   uniforms are (w + 0.5) 2^-32; Box-Muller, cosine branch for even t, sine branch for odd t. */
double oracle_normal_pred(uint64_t seed, uint64_t rep, uint32_t t, uint32_t y) {
    uint32_t w[4];
    oracle_philox4x32((uint32_t)rep, (uint32_t)(rep >> 32), t >> 2, y, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    uint32_t wa = (t & 2u) ? w[2] : w[0], wb = (t & 2u) ? w[3] : w[1];
    double u1 = ((double)wa + 0.5) * (1.0 / 4294967296.0), u2 = ((double)wb + 0.5) * (1.0 / 4294967296.0);
    double r = sqrt(-2.0 * log(u1));
    return r * ((t & 1u) ? sin(6.28318530717958647692 * u2) : cos(6.28318530717958647692 * u2));
}

/* ------------------------------------------------------------------------------------------
 * Prediction + envelopes
 * ------------------------------------------------------------------------------------------ */

static int cmp_double(const void* a, const void* b) {
    double x = *(const double*)a, y = *(const double*)b;
    return (x > y) - (x < y);
}

/* GetEmpiricalQuantile (BMSL, un-vendored): DECLARED CONVENTION = Hazen plotting positions
   pp_i=(i-0.5)/n with linear interpolation, clamped to the extremes. */
static double hazen_quantile(const double* xs, int64_t n, double p) {
    double h = p * (double)n + 0.5; /* 1-based fractional rank */
    if (h <= 1.0) return xs[0];
    if (h >= (double)n) return xs[n - 1];
    int64_t i = (int64_t)floor(h);
    double frac = h - (double)i;
    return xs[i - 1] + frac * (xs[i] - xs[i - 1]);
}

/* one column: src/BaM_tools.f90:1393-1421 */
void oracle_envelope(const double* col, int64_t nrep, double unfeas, double* env) {
    static const double probs[5] = {0.5, 0.025, 0.975, 0.16, 0.84};
    int64_t nbad = 0;
    for (int64_t r = 0; r < nrep; ++r)
        if (col[r] == unfeas || col[r] != col[r]) ++nbad;
    double* arr = (double*)malloc(((size_t)nrep + 1) * 8);
    int64_t m = 0;
    if ((double)nbad / (double)nrep < 0.75) {
        for (int64_t r = 0; r < nrep; ++r)
            if (!(col[r] == unfeas || col[r] != col[r])) arr[m++] = col[r];
    } else {
        for (int64_t r = 0; r < nrep; ++r) arr[m++] = col[r];
    }
    int anybad = 0;
    for (int64_t r = 0; r < m; ++r)
        if (arr[r] == unfeas || arr[r] != arr[r]) anybad = 1;
    if (anybad || m == 0) {
        for (int s = 0; s < BAMGPU_NENV; ++s) env[s] = unfeas;
    } else {
        qsort(arr, (size_t)m, sizeof(double), cmp_double);
        for (int s = 0; s < 5; ++s) env[s] = hazen_quantile(arr, m, probs[s]);
        double sum = 0.0;
        for (int64_t r = 0; r < m; ++r) sum += arr[r];
        double mean = sum / (double)m, ss = 0.0;
        for (int64_t r = 0; r < m; ++r) ss += (arr[r] - mean) * (arr[r] - mean);
        env[5] = mean;
        env[6] = (m > 1) ? sqrt(ss / (double)(m - 1)) : unfeas;
    }
    free(arr);
}

static int predict_impl(const oracle_t* o, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                        const double* const* xspag, const int32_t* nspag, int doParametric, const int32_t* doRemnant,
                        uint64_t seed, int t0, int t1, double* env, double* spag_out, int nthreads, char* mess, int nS);

int oracle_predict(const oracle_t* o, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                   const double* const* xspag, const int32_t* nspag, int doParametric, const int32_t* doRemnant,
                   uint64_t seed, int t0, int t1, double* env, double* spag_out, int nthreads, char* mess) {
    return predict_impl(o, Nrep, mc, mode, nobs_pred, xspag, nspag, doParametric, doRemnant, seed, t0, t1, env, spag_out,
                        nthreads, mess, 0);
}

/* with the state spaghetti (src/BaM_tools.f90:1071-1075,1121-1132): nState extra outputs after the nY outputs */
int oracle_predict_states(const oracle_t* o, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                          const double* const* xspag, const int32_t* nspag, int doParametric, const int32_t* doRemnant,
                          uint64_t seed, int t0, int t1, double* env, double* spag_out, int nthreads, char* mess) {
    return predict_impl(o, Nrep, mc, mode, nobs_pred, xspag, nspag, doParametric, doRemnant, seed, t0, t1, env, spag_out,
                        nthreads, mess, o->nState);
}

static int predict_impl(const oracle_t* o, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                        const double* const* xspag, const int32_t* nspag, int doParametric, const int32_t* doRemnant,
                        uint64_t seed, int t0, int t1, double* env, double* spag_out, int nthreads, char* mess, int nS) {
    if (mess) mess[0] = 0;
    if (o->nVar) { setmess(mess, "oracle: BaM_Prediction: VAR parameters not allowed (BaM_Message 13)"); return 1; }
    int nT = t1 - t0, nX = o->nX, nY = o->nY, nOut = nY + nS;
    if (nT <= 0 || Nrep <= 0) return 0;
    int anyRem = 0;
    for (int y = 0; y < nY; ++y) anyRem |= doRemnant[y];
    if ((doParametric || anyRem) && !mc) { setmess(mess, "oracle: BaM_Prediction: mcmc matrix required"); return 1; }
    double* spag = spag_out ? spag_out : (double*)malloc((size_t)Nrep * nT * nOut * 8);
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for num_threads(nthreads) schedule(static)
    for (int64_t rep = 0; rep < Nrep; ++rep) { /* the rep loop, :1034-1108 */
        double* X = (double*)malloc(((size_t)nT * nX + 1) * 8);
        double* Ys = (double*)malloc(((size_t)nT * nOut + 1) * 8);
        int32_t* vf = (int32_t*)malloc(((size_t)nT + 1) * 4);
        double theta[256], full[256], dmod[MAXCONTROL + 64];
        for (int i = 0; i < nX; ++i) { /* indx = modulo(rep, nspag), 1-based rep (:1049) */
            int64_t col = rep % nspag[i];
            for (int t = 0; t < nT; ++t) X[t + (size_t)i * nT] = xspag[i][(size_t)col * nobs_pred + t0 + t];
        }
        for (int k = 0; k < o->nFit; ++k) theta[k] = doParametric ? mc[rep + (size_t)k * Nrep] : mode[k];
        full_theta(o, theta, 0, full);
        int feas;
        for (size_t q = (size_t)nT * nY; q < (size_t)nT * nOut; ++q) Ys[q] = o->unfeas;
        int e = apply_model_s(o, nT, X, nT, full, Ys, nT, dmod, vf, &feas, nS ? Ys + (size_t)nT * nY : NULL);
        if (e > 0) {
#pragma omp critical
            err = e;
        }
        for (int y = 0; y < nY; ++y) {
            if (doRemnant[y]) {
                double gam[8];
                for (int m = 0; m < o->nremnant[y]; ++m)
                    gam[m] = mc[rep + (size_t)(o->nFit + o->gamma_off[y] + m) * Nrep];
                for (int t = 0; t < nT; ++t) {
                    double v = Ys[t + (size_t)y * nT];
                    if (v == o->unfeas) continue;
                    double sig = sigmafunk(o->sigma_func[y], gam, v);
                    if (sig > 0.0) /* Generate(GAUSS,(0,sig)) feasible iff sig>0 (:1095-1102) */
                        Ys[t + (size_t)y * nT] = v + sig * oracle_normal_pred(seed, (uint64_t)rep, (uint32_t)(t0 + t), (uint32_t)y);
                }
            }
            for (int t = 0; t < nT; ++t) spag[(size_t)y * Nrep * nT + (size_t)t * Nrep + rep] = Ys[t + (size_t)y * nT];
        }
        for (int y = nY; y < nOut; ++y) /* states: written before the structural error is added, never perturbed (:1071-1075) */
            for (int t = 0; t < nT; ++t) spag[(size_t)y * Nrep * nT + (size_t)t * Nrep + rep] = Ys[t + (size_t)y * nT];
        free(X); free(Ys); free(vf);
    }
    if (env && Nrep >= 2) {
#pragma omp parallel for num_threads(nthreads) schedule(dynamic, 8)
        for (int64_t q = 0; q < (int64_t)nT * nOut; ++q) {
            int y = (int)(q / nT), t = (int)(q % nT);
            double e7[BAMGPU_NENV];
            oracle_envelope(spag + (size_t)y * Nrep * nT + (size_t)t * Nrep, Nrep, o->unfeas, e7);
            for (int s = 0; s < BAMGPU_NENV; ++s) env[(size_t)y * BAMGPU_NENV * nT + (size_t)s * nT + t] = e7[s];
        }
    }
    if (!spag_out) free(spag);
    return err;
}

/* ------------------------------------------------------------------------------------------
 * Sampler: Adaptive_Metro_OAAT (BMSL MCMCStrategy_tools, un-vendored; semantics as declared in
 * SURVEY.md App. E).  RNG counters: proposal of chain c, iteration it (0-based over the whole run),
 * component i uses normal(seed, c, it, 2i) and uniform(seed, c, it, 2i+1).
 * ------------------------------------------------------------------------------------------ */
int oracle_fit(const oracle_t* o, int K, const double* start, const double* start_std, int nAdapt, int nCycles,
               double MinMoveRate, double MaxMoveRate, double DownMult, double UpMult, uint64_t seed, int burn,
               int keep_every, double* out_rows, int64_t* n_keep, double* final_x, double* final_std,
               int64_t chain_offset, char* mess) {
    int P = o->P, nD = o->nDpar;
    int64_t nIter = (int64_t)nAdapt * nCycles;
    if (keep_every < 1) keep_every = 1;
    int64_t nKeep = (nIter > burn) ? (nIter - burn) / keep_every : 0;
    if (n_keep) *n_keep = nKeep;
    int rowlen = P + 1 + nD;
    int err = 0;
    for (int c = 0; c < K && !err; ++c) {
        double* x = (double*)malloc((size_t)P * 8);
        double* cand = (double*)malloc((size_t)P * 8);
        double* sd = (double*)malloc((size_t)P * 8);
        double* dpar = (double*)malloc(((size_t)nD + 1) * 8);
        double* dcand = (double*)malloc(((size_t)nD + 1) * 8);
        int* moves = (int*)calloc(P, sizeof(int));
        memcpy(x, start + (size_t)c * P, (size_t)P * 8);
        memcpy(sd, start_std + (size_t)c * P, (size_t)P * 8);
        double fx, fc;
        int32_t feas, isnull;
        oracle_logpost(o, x, &fx, NULL, NULL, NULL, &feas, &isnull, dpar, 0, mess);
        if (!feas || isnull) { setmess(mess, "oracle: Adaptive_Metro_OAAT: unfeasible starting point"); err = 1; }
        int64_t it = 0, kept = 0;
        for (int cyc = 0; cyc < nCycles && !err; ++cyc) {
            memset(moves, 0, (size_t)P * sizeof(int));
            for (int a = 0; a < nAdapt; ++a, ++it) {
                for (int i = 0; i < P; ++i) {
                    memcpy(cand, x, (size_t)P * 8);
                    cand[i] = x[i] + sd[i] * oracle_normal(seed, (uint64_t)(chain_offset + c), (uint32_t)it, (uint32_t)(2 * i));
                    oracle_logpost(o, cand, &fc, NULL, NULL, NULL, &feas, &isnull, dcand, 0, mess);
                    if (feas && !isnull) {
                        double u = oracle_uniform(seed, (uint64_t)(chain_offset + c), (uint32_t)it, (uint32_t)(2 * i + 1));
                        if (log(u) < fc - fx) {
                            x[i] = cand[i]; fx = fc; memcpy(dpar, dcand, (size_t)nD * 8); moves[i]++;
                        }
                    }
                }
                int64_t it1 = it + 1;
                if (out_rows && it1 > burn && (it1 - burn) % keep_every == 0 && kept < nKeep) {
                    double* row = out_rows + ((size_t)c * nKeep + kept) * rowlen;
                    memcpy(row, x, (size_t)P * 8);
                    row[P] = fx;
                    memcpy(row + P + 1, dpar, (size_t)nD * 8);
                    ++kept;
                }
            }
            for (int i = 0; i < P; ++i) {
                double rate = (double)moves[i] / (double)nAdapt;
                if (rate <= MinMoveRate) sd[i] *= DownMult;
                if (rate >= MaxMoveRate) sd[i] *= UpMult;
            }
        }
        if (final_x) memcpy(final_x + (size_t)c * P, x, (size_t)P * 8);
        if (final_std) memcpy(final_std + (size_t)c * P, sd, (size_t)P * 8);
        free(x); free(cand); free(sd); free(dpar); free(dcand); free(moves);
    }
    return err;
}
