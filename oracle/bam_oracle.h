/*
 * bam_oracle.h -- CPU restatement (ORACLE) of BaM's hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (bam_b200/, the C-ABI library, the
 * bam_gpu driver) may include, link or call this.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker /
 * the reported CPU baseline.
 *
 * PARITY STATUS: "parity unpinned" for everything whose arithmetic lives in the
 * un-vendored, un-pinned BMSL / miniDMSL libraries (distribution conventions, quantile
 * plotting positions, sampler bookkeeping, znewton path) -- the reference ships no golden
 * outputs and cannot be compiled here (no Fortran compiler).  What IS
 * pinned: the known-answer columns in the reference's sediment data files, its control
 * matrices / VAR index columns / prior-correlation matrix, and the surveyor-computed
 * values of SURVEY.md App. F (independent restatement).
 *
 * Each function cites the reference file:line it follows.
 */
#ifndef BAM_ORACLE_H
#define BAM_ORACLE_H

#include "../include/bam_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_handle oracle_t;

int oracle_create(oracle_t** o, const bamgpu_problem* p, char* mess);
void oracle_destroy(oracle_t* o);
int oracle_get_sizes(const oracle_t* o, bamgpu_sizes* s);

/* One log-posterior, exactly in the reference's order (Posterior_wrapper, src/BaM_tools.f90:3467).
   Any of prior/lkh/hyper/dpar may be NULL.  use_long_double != 0 accumulates in long double
   (the arbiter of SURVEY.md section 7 "hard parts"). */
int oracle_logpost(const oracle_t* o, const double* x, double* fx, double* prior, double* lkh, double* hyper,
                   int32_t* feas, int32_t* isnull, double* dpar, int use_long_double, char* mess);

/* GetSim_calibration (src/BaM_tools.f90:3619-3680) */
int oracle_calibration_run(const oracle_t* o, const double* x, double* Xtrue, double* Ysim, int32_t* feas, char* mess);
/* with the state variables (n x nState), same arguments as bamgpu_calibration_run_states */
int oracle_calibration_run_states(const oracle_t* o, const double* x, double* Xtrue, double* Ysim, double* state,
                                  int32_t* feas, char* mess);

/* K vectors (P x K), nthreads OpenMP threads over vectors (1 = the serial reference analogue). */
int oracle_logpost_batch(const oracle_t* o, int K, const double* x, double* fx, int32_t* feas, int32_t* isnull,
                         double* dpar, int nthreads, char* mess);

/* ApplyModel_BaM without VAR (src/BaM_tools.f90:4133-4149) on an arbitrary input matrix:
   X is n x nX column-major, theta_fit is the nFit fitted values, Y is n x nY, feas_rows n. */
int oracle_apply_model(const oracle_t* o, int n, const double* X, const double* theta_fit, double* Y,
                       double* dpar, int32_t* feas_rows, char* mess);

/* BaM_Prediction + BaM_ProcessSpag in memory (src/BaM_tools.f90:882-1135,1292-1432);
   same arguments as bamgpu_predict. */
int oracle_predict(const oracle_t* o, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                   const double* const* xspag, const int32_t* nspag, int doParametric,
                   const int32_t* doRemnant, uint64_t seed, int t0, int t1, double* env, double* spag,
                   int nthreads, char* mess);
/* with the state spaghetti: nState extra outputs after the nY outputs (same arguments as bamgpu_predict_states) */
int oracle_predict_states(const oracle_t* o, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                          const double* const* xspag, const int32_t* nspag, int doParametric,
                          const int32_t* doRemnant, uint64_t seed, int t0, int t1, double* env, double* spag,
                          int nthreads, char* mess);

/* the 7 envelope statistics of one column of nrep values (src/BaM_tools.f90:1393-1421) */
void oracle_envelope(const double* col, int64_t nrep, double unfeas_flag, double* env7);

/* Adaptive_Metro_OAAT (BMSL, recalled semantics: SURVEY.md App. E), same RNG streams as the device */
int oracle_fit(const oracle_t* o, int K, const double* start, const double* start_std, int nAdapt, int nCycles,
               double MinMoveRate, double MaxMoveRate, double DownMult, double UpMult, uint64_t seed, int burn,
               int keep_every, double* out_rows, int64_t* n_keep, double* final_x, double* final_std,
               int64_t chain_offset /* global index of chain 0 (sharding across ranks) */, char* mess);

/* RNG primitives shared (by specification, not by code) with the device */
void oracle_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                       uint32_t out[4]);
double oracle_normal(uint64_t seed, uint64_t a, uint32_t b, uint32_t c);  /* N(0,1) for counter (a,b,c) */
double oracle_uniform(uint64_t seed, uint64_t a, uint32_t b, uint32_t c); /* U(0,1) */
/* prediction noise for (replication, input t, output y): Box-Muller pairs, see bam_oracle.c */
double oracle_normal_pred(uint64_t seed, uint64_t rep, uint32_t t, uint32_t y);

/* distribution helpers exposed for tests */
double oracle_normal_quantile(double p);
int oracle_logpdf(int dist, const double* par, double x, double* lp, int* feas, int* isnull);
int oracle_cdf(int dist, const double* par, double x, double* u, int* feas);

/* TextFile: compile a formula; returns bytecode length or <0 on syntax error */
int oracle_textfile_compile(const char* formula, const char* const* varnames, int nvar, int32_t* code,
                            int maxcode, double* immed, int maximmed, int* nimmed, int* stacksize);

#ifdef __cplusplus
}
#endif
#endif
