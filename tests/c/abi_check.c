/* abi_check.c -- a plain C99 client of include/bam_gpu.h (gcc -std=c99 -pedantic -Wall -Werror).
 *
 *   abi_check layout : prints offsetof / sizeof of the two structs a Fortran bind(C) host shares with the library
 *                      (tests/test_abi_c99.py compares them with the layout bindings/bam_gpu_binding.f90 implies)
 *   abi_check run    : create -> logpost -> logpost_parts -> fit -> moments -> predict -> destroy on a small BaRatin
 *                      problem, through the C ABI only; prints the values (the test compares them with the oracle).
 *                      Without a GPU it must fail loudly at bamgpu_create (exit code 3, "no CUDA device").
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bam_gpu.h"

#define OFF(T, f) printf("%s.%s %lu\n", #T, #f, (unsigned long)offsetof(T, f))

static int layout(void) {
    OFF(bamgpu_problem, model_id); OFF(bamgpu_problem, nobs); OFF(bamgpu_problem, nX); OFF(bamgpu_problem, nY);
    OFF(bamgpu_problem, ntheta); OFF(bamgpu_problem, X); OFF(bamgpu_problem, Xu); OFF(bamgpu_problem, Xb);
    OFF(bamgpu_problem, Xbindx); OFF(bamgpu_problem, Y); OFF(bamgpu_problem, Yu); OFF(bamgpu_problem, Yb);
    OFF(bamgpu_problem, Ybindx); OFF(bamgpu_problem, partype); OFF(bamgpu_problem, nval); OFF(bamgpu_problem, fix_value);
    OFF(bamgpu_problem, var_indx); OFF(bamgpu_problem, prior_theta_dist); OFF(bamgpu_problem, prior_theta_par);
    OFF(bamgpu_problem, sigma_func); OFF(bamgpu_problem, nremnant); OFF(bamgpu_problem, prior_gamma_dist);
    OFF(bamgpu_problem, prior_gamma_par); OFF(bamgpu_problem, priorM); OFF(bamgpu_problem, xtra_i);
    OFF(bamgpu_problem, n_xtra_i); OFF(bamgpu_problem, xtra_d); OFF(bamgpu_problem, n_xtra_d);
    OFF(bamgpu_problem, xtra_text); OFF(bamgpu_problem, mv); OFF(bamgpu_problem, unfeas_flag);
    printf("bamgpu_problem.sizeof %lu\n", (unsigned long)sizeof(bamgpu_problem));
    OFF(bamgpu_sizes, nFit); OFF(bamgpu_sizes, nGamma); OFF(bamgpu_sizes, nLatent); OFF(bamgpu_sizes, nXbias);
    OFF(bamgpu_sizes, nYbias); OFF(bamgpu_sizes, nInfer); OFF(bamgpu_sizes, nDpar); OFF(bamgpu_sizes, nCombo);
    OFF(bamgpu_sizes, nState);
    printf("bamgpu_sizes.sizeof %lu\n", (unsigned long)sizeof(bamgpu_sizes));
    printf("BAMGPU_MESS_LEN %d\nBAMGPU_PRIOR_NPAR %d\nBAMGPU_NENV %d\nBAMGPU_COMM_ID_BYTES %d\n", BAMGPU_MESS_LEN,
           BAMGPU_PRIOR_NPAR, BAMGPU_NENV, BAMGPU_COMM_ID_BYTES);
    return 0;
}

#define NOBS 12
#define P 8
#define K 4

static int run(void) {
    /* two-control BaRatin (control matrix [[1,0],[0,1]]), 12 gaugings */
    static const double H[NOBS] = {-0.3, -0.1, 0.1, 0.4, 0.8, 1.2, 1.6, 1.9, 2.3, 2.8, 3.4, 4.1};
    static const double Q[NOBS] = {4.1, 12.9, 25.0, 47.8, 85.0, 130.2, 181.0, 260.4, 420.9, 700.3, 1105.0, 1712.0};
    double Yu[NOBS], zero[NOBS];
    int32_t izero[NOBS];
    static const int32_t partype[6] = {0, 0, 0, 0, 0, 0}, nval[6] = {1, 1, 1, 1, 1, 1};
    static const double fix[6] = {0, 0, 0, 0, 0, 0};
    static const int32_t ptd[6] = {BAMGPU_DIST_GAUSSIAN, BAMGPU_DIST_LOGNORMAL, BAMGPU_DIST_GAUSSIAN, BAMGPU_DIST_GAUSSIAN,
                                   BAMGPU_DIST_LOGNORMAL, BAMGPU_DIST_GAUSSIAN};
    static const double ptp[6 * BAMGPU_PRIOR_NPAR] = {-0.5, 1, 0, 0, 3.9, 0.5, 0, 0, 1.5, 0.05, 0, 0,
                                                      1.5, 0.5, 0, 0, 4.9, 0.5, 0, 0, 1.67, 0.05, 0, 0};
    static const int32_t sf[1] = {BAMGPU_SIG_LINEAR}, nrem[1] = {2};
    static const int32_t pgd[2] = {BAMGPU_DIST_UNIFORM, BAMGPU_DIST_UNIFORM};
    static const double pgp[2 * BAMGPU_PRIOR_NPAR] = {0, 1000, 0, 0, 0, 1000, 0, 0};
    static const int32_t ctrl[4] = {1, 0, 0, 1};
    static const double x0[P] = {-0.5, 53.0, 1.5, 1.5, 144.0, 1.67, 1.0, 0.1};
    bamgpu_problem pb;
    bamgpu_t* h = NULL;
    bamgpu_sizes sz;
    char mess[BAMGPU_MESS_LEN];
    double x[P * K], sd[P * K], fx[K], dpar[2 * K], prior[K], lkh[K], hyper[K];
    int32_t feas[K], isnull[K];
    int i, k, e;

    for (i = 0; i < NOBS; ++i) { Yu[i] = 0.05 * Q[i]; zero[i] = 0.0; izero[i] = 0; }
    memset(&pb, 0, sizeof(pb));
    pb.model_id = "BaRatin";
    pb.nobs = NOBS; pb.nX = 1; pb.nY = 1; pb.ntheta = 6;
    pb.X = H; pb.Xu = zero; pb.Xb = zero; pb.Xbindx = izero;
    pb.Y = Q; pb.Yu = Yu; pb.Yb = zero; pb.Ybindx = izero;
    pb.partype = partype; pb.nval = nval; pb.fix_value = fix; pb.var_indx = NULL;
    pb.prior_theta_dist = ptd; pb.prior_theta_par = ptp;
    pb.sigma_func = sf; pb.nremnant = nrem; pb.prior_gamma_dist = pgd; pb.prior_gamma_par = pgp;
    pb.priorM = NULL;
    pb.xtra_i = ctrl; pb.n_xtra_i = 4; pb.xtra_d = NULL; pb.n_xtra_d = 0; pb.xtra_text = NULL;
    pb.mv = -9999.0; pb.unfeas_flag = -666.666;

    e = bamgpu_create(&h, &pb, 0, mess);
    if (e != 0) { printf("create err %d: %s\n", e, mess); return 3; }
    /* a NULL handle is an error, not a crash */
    if (bamgpu_logpost(NULL, 1, x0, fx, feas, isnull, NULL, mess) == 0 || strstr(mess, "null handle") == NULL) return 4;
    if (bamgpu_get_sizes(h, &sz) != 0 || sz.nInfer != P || sz.nDpar != 2) { printf("sizes %d %d\n", sz.nInfer, sz.nDpar); return 5; }
    for (k = 0; k < K; ++k)
        for (i = 0; i < P; ++i) { x[k * P + i] = x0[i] * (1.0 + 0.01 * k * ((i % 2) ? 1 : -1)); sd[k * P + i] = 0.05 * (x0[i] < 0 ? -x0[i] : x0[i]); }
    x[3 * P + 3] = -1.0; /* k2 < k1: infeasible vector */
    if ((e = bamgpu_logpost(h, K, x, fx, feas, isnull, dpar, mess)) != 0) { printf("logpost err %d: %s\n", e, mess); return 6; }
    if ((e = bamgpu_logpost_parts(h, K, x, prior, lkh, hyper, feas, isnull, mess)) != 0) { printf("parts err %d: %s\n", e, mess); return 7; }
    for (k = 0; k < K; ++k) printf("logpost %d %.17g %d %d %.17g %.17g %.17g %.17g %.17g\n", k, fx[k], feas[k], isnull[k], prior[k], lkh[k], hyper[k], dpar[2 * k], dpar[2 * k + 1]);
    {
        const int nAdapt = 10, nCycles = 3, rowlen = P + 1 + 2;
        double* rows = (double*)malloc(sizeof(double) * 3 * nAdapt * nCycles * rowlen);
        double cnt[3], mean[3 * P], m2[3 * P], rhat[P];
        int64_t nkeep = 0;
        if ((e = bamgpu_fit(h, 3, x, sd, nAdapt, nCycles, 0.1, 0.5, 0.9, 1.1, 7u, 0, 1, rows, &nkeep, mess)) != 0) { printf("fit err %d: %s\n", e, mess); return 8; }
        if (nkeep != nAdapt * nCycles) return 9;
        if (bamgpu_moments(h, cnt, mean, m2, 0, mess) != 0 || bamgpu_rhat(3, P, cnt, mean, m2, rhat, mess) != 0) return 10;
        printf("fit %ld %.17g %.17g\n", (long)nkeep, rows[(nkeep - 1) * rowlen + P], rhat[0]);
        free(rows);
    }
    {
        static const double Hp[5] = {-1.0, 0.0, 1.0, 2.0, 3.0};
        const double* xs[1];
        const int32_t ns[1] = {1}, dr[1] = {1};
        double mc[P * 3], env[5 * BAMGPU_NENV];
        xs[0] = Hp;
        for (k = 0; k < 3; ++k)
            for (i = 0; i < P; ++i) mc[k + 3 * i] = x[k * P + i]; /* Nrep x P column-major */
        if ((e = bamgpu_predict(h, 3, mc, x0, 5, xs, ns, 1, dr, 11u, 0, 5, env, NULL, mess)) != 0) { printf("predict err %d: %s\n", e, mess); return 11; }
        for (i = 0; i < 5; ++i) printf("env %d %.17g %.17g %.17g\n", i, env[0 * 5 + i], env[1 * 5 + i], env[2 * 5 + i]);
    }
    bamgpu_destroy(h);
    printf("OK %s\n", bamgpu_version());
    return 0;
}

int main(int argc, char** argv) {
    if (argc > 1 && strcmp(argv[1], "layout") == 0) return layout();
    if (argc > 1 && strcmp(argv[1], "run") == 0) return run();
    fprintf(stderr, "usage: abi_check layout | run\n");
    return 2;
}
