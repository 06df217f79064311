/*
 * bam_gpu.h -- C ABI of the B200-native BaM hot path.
 *
 * This is the drop-in boundary: everything a host written in Fortran
 * (ISO_C_BINDING, see bindings/bam_gpu_binding.f90), C or C++ needs to replace
 *   - the sampler callback  Posterior_wrapper      (src/BaM_tools.f90:3467-3517)
 *   - the engine calls      BaM_Fit / BaM_Prediction (src/BaM_tools.f90:545-623, 882-1135)
 *   - the envelope pass     BaM_ProcessSpag        (src/BaM_tools.f90:1292-1432)
 * of the reference.  Plain pointers and sizes only; all matrices are COLUMN-MAJOR
 * exactly as Fortran passes them; integers are int32_t (mik), reals are double
 * (mrk).  Every call returns `int err` with the reference's convention
 * (0 OK, >0 error, <0 warning; src/BaM_tools.f90:175) and fills `mess`
 * (caller-owned, at least BAMGPU_MESS_LEN bytes, NUL terminated).
 *
 * The same `bamgpu_problem` description is consumed by the CPU oracle
 * (oracle/bam_oracle.h), which is test infrastructure only.
 */
#ifndef BAM_GPU_H
#define BAM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAMGPU_MESS_LEN 250
#define BAMGPU_PRIOR_NPAR 4            /* stride of the flattened prior-parameter tables */
#define BAMGPU_UNDEF_RN (-999999999.0) /* DMSL undefRN: value of fx when !feas or isnull */
#define BAMGPU_NENV 7                  /* Median q2.5 q97.5 q16 q84 Mean Stdev (src/BaM_tools.f90:1309-1315) */

/* Model catalogue: ModelLibrary_tools.f90:79-112 (ID strings without the "MDL_" prefix). */
enum bamgpu_model {
    BAMGPU_MDL_LINEAR = 1,    /* Linear_model.f90:62-101 */
    BAMGPU_MDL_BARATIN = 2,   /* BaRatin_model.f90:67-436 */
    BAMGPU_MDL_SEDIMENT = 3,  /* SedimentTransport_model.f90:100-607 */
    BAMGPU_MDL_TEXTFILE = 4,  /* TextFile_model.f90:231-309,1035-1098 */
    BAMGPU_MDL_VEGETATION = 5,    /* Vegetation_model.f90:87-569 */
    /* second wave: pointwise closed forms behind the same kernels */
    BAMGPU_MDL_SUSPENDEDLOAD = 6, /* SuspendedLoad_model.f90:38-116 */
    BAMGPU_MDL_SWOT = 7,          /* SWOT_model.f90:73-148 */
    BAMGPU_MDL_SGD = 8,           /* StageGradientDischarge_model.f90:57-113 */
    BAMGPU_MDL_ORTHO = 9,         /* Orthorectification_model.f90:66-288 ("Orthorectification") */
    BAMGPU_MDL_RECESSION_H = 10,  /* Recession_model.f90:36-94 ("Recession_h") */
    BAMGPU_MDL_HCS = 11,          /* HydraulicControl_section_model.f90:40-112 ("HydraulicControl_section") */
    BAMGPU_MDL_SEGMENTATION = 12, /* Segmentation_model.f90:52-140 (whole-series model: segment sizes are checked) */
    BAMGPU_MDL_BARATINBAC = 13,   /* BaRatinBAC_model.f90:96-666: BaRatin in the b-a-c parameterisation */
    BAMGPU_MDL_SFD = 14,          /* StageFallDischarge_model.f90:80-283 (SFD_Val_General); the transition stage kappa is
                                     the model's state variable (bamgpu_predict_states, bamgpu_calibration_run_states) */
    BAMGPU_MDL_MIXTURE = 15,      /* Mixture_model.f90:74-135: per-event weighted sum of source concentrations; output
                                     rows are event x element blocks (row (j-1)*nElt+e = e-th input row of event j) */
    /* time-recursive models: one sequential pass over the series per parameter vector */
    BAMGPU_MDL_SFDTIDAL_QMEC = 16, /* SFDTidal_Qmec_model.f90:61-197: tidal discharge from two stage series by a momentum
                                     balance (Adams-Bashforth in time); 12 parameters, states = pressure gradient, bottom
                                     friction, advection */
    BAMGPU_MDL_SFDTIDAL_QMEC2 = 17, /* SFDTidal_Qmec2_model.f90: the same recursion parameterised by (Be, bevs, delta, ne, d1,
                                      d2, c, g, Q0, dx, dt) -- effective width and Manning coefficient given directly */
    BAMGPU_MDL_ALGAE = 18,         /* AlgaeBiomass_model.f90:55-197 ("AlgaeBiomass"): daily biomass recursion (logistic growth
                                      limited by temperature and light, senescence, grazing / removal, hydraulic scour);
                                      5 inputs (temperature, irradiance, depth, turbidity, removal), 18 parameters, outputs
                                      biomass and biomass / bmax, states Fb, Ft, Fi, Fn, Rhyd */
    BAMGPU_MDL_DYNVEG = 19,        /* DynamicVegetation_model.f90:74-191 ("DynamicVegetation"): the AlgaeBiomass recursion on
                                      depth = max(stage - b1, 0) feeds the vegetation-affected rating curve (Newton per time
                                      step, Vegetation_fNewton); 5 inputs (temperature, irradiance, stage, turbidity, removal),
                                      18 + 6 parameters (B, n1, b1, c1, chi, U_chi), outputs Q, biomass, biomass / bmax, the five
                                      AlgaeBiomass states; xtra_d = Newton xscale, fscale, tolX, tolF, xtra_i = itmax */
    BAMGPU_MDL_SFDTIDAL_QMEC0 = 20, /* SFDTidal_Qmec0_model.f90:61-178: the original Qmec parameterisation (Be, he, dzeta, ne, d1,
                                      d2, c, g, Q0, dx, dt), stages MINUS datums, no state variables */
    BAMGPU_MDL_SFDTIDAL_QMECMS = 21, /* SFDTidal_QmecMS_model.f90:62-154: Manning-Strickler (steady) version of Qmec, theta =
                                      (y0, A0, be, delta, phi, d1, d2, c, g, dx); pointwise in time but the geometry of the
                                      WHOLE series decides feasibility, so it runs with the time-recursive models */
    BAMGPU_MDL_SFDTIDAL_LAG = 22    /* the stage-fall-discharge curves for tidal rivers that read the stages at a LAGGED time
                                      step, tlag = clamp(t - int(theta1), 1, n): "SFDTidal" (SFDTidal_model.f90:65-141),
                                      "SFDTidal2" (SFDTidal2_model.f90), "SFDTidal4" (slope averaged over t-P..t+P,
                                      SFDTidal4_model.f90) and "SFDTidalJones" (third input dh/dt, SFDTidalJones_model.f90);
                                      8 parameters each, no state variables; the four names share this id */
};

/* Prior / distribution families (BMSL Distribution_tools names). */
enum bamgpu_dist {
    BAMGPU_DIST_GAUSSIAN = 1,
    BAMGPU_DIST_UNIFORM = 2,
    BAMGPU_DIST_TRIANGLE = 3,
    BAMGPU_DIST_LOGNORMAL = 4,
    BAMGPU_DIST_EXPONENTIAL = 5,
    BAMGPU_DIST_FLATPRIOR = 6,
    BAMGPU_DIST_FLATPRIOR_POS = 7, /* 'FlatPrior+' */
    BAMGPU_DIST_FLATPRIOR_NEG = 8  /* 'FlatPrior-' */
};

/* Remnant sigma functions: Sigmafunk_Apply, src/BaM_tools.f90:3521-3571. */
enum bamgpu_sigmafunc {
    BAMGPU_SIG_CONSTANT = 1,
    BAMGPU_SIG_LINEAR = 2,
    BAMGPU_SIG_PROPORTIONAL = 3,
    BAMGPU_SIG_POWER = 4,
    BAMGPU_SIG_EXPONENTIAL = 5,
    BAMGPU_SIG_GAUSSIAN = 6
};

/* Parameter types: src/BaM_tools.f90:118. */
enum bamgpu_partype { BAMGPU_PAR_FIX = -1, BAMGPU_PAR_STD = 0, BAMGPU_PAR_VAR = 1 };

/* Sediment options (SedimentTransport_model.f90:30-48), carried in xtra_i. */
enum bamgpu_sed_formula { BAMGPU_SED_MPM = 1, BAMGPU_SED_CAMENEN = 2, BAMGPU_SED_NIELSEN = 3, BAMGPU_SED_SMART = 4 };
enum bamgpu_sed_css { BAMGPU_CSS_CONSTANT = 1, BAMGPU_CSS_SOULSBY = 2, BAMGPU_CSS_SW = 3 };
enum bamgpu_sed_ppo { BAMGPU_PPO_FIX = 1, BAMGPU_PPO_IN = 2, BAMGPU_PPO_PAR = 3 };
enum bamgpu_sed_co { BAMGPU_CO_FIX = 1, BAMGPU_CO_PAR = 2 };

/* SWOT variants (SWOT_model.f90:29-30), xtra_i[0]; distortion functions of Orthorectification (:28-30), xtra_i[0] */
enum bamgpu_swot_id { BAMGPU_SWOT_HS = 1, BAMGPU_SWOT_HSB = 2 };
enum bamgpu_ortho_disto { BAMGPU_DISTO_NONE = 1, BAMGPU_DISTO_RADIAL = 2 };

/* Vegetation growth formulas (Vegetation_model.f90:25-30), carried in xtra_i[0]. */
enum bamgpu_veg_growth { BAMGPU_VEG_YIN1 = 1, BAMGPU_VEG_YIN2 = 2, BAMGPU_VEG_YIN3 = 3, BAMGPU_VEG_YIN1_TEMPORAL = 4 };

/*
 * The immutable problem description = the reference's module globals INFER + MODEL
 * after LoadBamObjects (src/BaM_tools.f90:73-101,133-420; ModelLibrary_tools.f90:114-126).
 * The callee copies everything it needs during bamgpu_create (as LoadBamObjects
 * does at :213-229); the caller keeps ownership of every pointer.
 */
typedef struct bamgpu_problem {
    const char* model_id; /* "BaRatin" | "Linear" | "Sediment" | "TextFile" | "Vegetation" | "SuspendedLoad" | "SWOT" | "SGD" |
                             "Orthorectification" | "Recession_h" | "HydraulicControl_section" | "Segmentation" | "BaRatinBAC" | "SFD" |
                             "Mixture" | "SFDTidal_Qmec" | "SFDTidal_Qmec2" | "AlgaeBiomass" | "DynamicVegetation" |
                             "SFDTidal_Qmec0" | "SFDTidal_QmecMS" | "SFDTidal" | "SFDTidal2" | "SFDTidal4" | "SFDTidalJones"
                             (optionally "MDL_"-prefixed) */
    int32_t nobs, nX, nY, ntheta;

    /* calibration data, nobs x nX / nobs x nY, column-major (INFER%X .. INFER%Ybindx) */
    const double* X;
    const double* Xu;
    const double* Xb;
    const int32_t* Xbindx;
    const double* Y;
    const double* Yu;
    const double* Yb;
    const int32_t* Ybindx;

    /* model parameters theta (INFER%theta), size ntheta */
    const int32_t* partype;  /* bamgpu_partype */
    const int32_t* nval;     /* 1 for FIX/STD, number of distinct values for VAR */
    const double* fix_value; /* init(1), used for FIX parameters only */
    const int32_t* var_indx; /* nobs x ntheta column-major, 1-based; may be NULL if no VAR parameter */

    /* flattened priors of the fitted theta (INFER%Prior_theta, size nFit = sum nval of non-FIX) */
    const int32_t* prior_theta_dist;
    const double* prior_theta_par; /* nFit x BAMGPU_PRIOR_NPAR, row per parameter */

    /* remnant-error (structural) model per output */
    const int32_t* sigma_func; /* nY, bamgpu_sigmafunc */
    const int32_t* nremnant;   /* nY */
    const int32_t* prior_gamma_dist; /* sum(nremnant) */
    const double* prior_gamma_par;   /* sum(nremnant) x BAMGPU_PRIOR_NPAR */

    /* optional Gaussian-copula matrix (C^-1 - I), k x k, k = nFit + sum(nremnant); NULL if absent
       (INFER%priorM, src/BaM_tools.f90:386-414; order of rows: theta first, then gamma, :3084-3101) */
    const double* priorM;

    /* model-specific xtra (ModelLibrary_tools.f90:436-544):
       BaRatin    : xtra_i = control matrix ncontrol x ncontrol, column-major (n_xtra_i = ncontrol^2)
       BaRatinBAC : the same matrix; xtra_d = {hmax} (upper bound of the highest activation stage)
       Sediment   : xtra_i = {formula, css, co, ppo_d, ppo_s, ppo_v, ppo_g}; xtra_d = pseudoval[4]
       TextFile   : xtra_text = content of the model text file (TextFile_model.f90:946-978)
       Vegetation : xtra_i = {growth formula, nYear, itmax}; xtra_d = {xscale, fscale, tolX, tolF}
       SWOT       : xtra_i = {bamgpu_swot_id}
       Orthorectification : xtra_i = {bamgpu_ortho_disto, number of distortion parameters}
       HydraulicControl_section : xtra_d = the h-A-w matrix, p x 3 column-major (n_xtra_d = 3p)
       Segmentation : xtra_i = {nS, nmin, option}; xtra_d = {tmin}
       SFD        : xtra_i = {1 (SFD_Val_General), itmax}; xtra_d = {upperbound, xscale, fscale, tolX, tolF}
       Mixture    : xtra_i = {nS, nEvt, nElt, missingSource (0|1)} (Mixture_XtraRead, Mixture_model.f90:137-190)
       Linear, SuspendedLoad, SGD, Recession_h : none */
    const int32_t* xtra_i;
    int32_t n_xtra_i;
    const double* xtra_d;
    int32_t n_xtra_d;
    const char* xtra_text;

    double mv;          /* missing-value code of Y, -9999 (src/BaM_tools.f90:100) */
    double unfeas_flag; /* -666.666 (ModelLibrary_tools.f90:124) */
} bamgpu_problem;

/* Sizes derived at create time (INFER%nFit, nInfer, nDpar ...; src/BaM_tools.f90:336,346-379). */
typedef struct bamgpu_sizes {
    int32_t nFit;    /* fitted theta values */
    int32_t nGamma;  /* sum(nremnant) */
    int32_t nLatent; /* sum(nXerror): inferred "true" inputs */
    int32_t nXbias;  /* sum(nXb) */
    int32_t nYbias;  /* sum(nYb) */
    int32_t nInfer;  /* P = length of a parameter vector */
    int32_t nDpar;   /* derived parameters per vector (model nDpar x number of VAR combinations) */
    int32_t nCombo;  /* distinct VAR combinations (size of INFER%DparIndx) */
    int32_t nState;
} bamgpu_sizes;

typedef struct bamgpu_handle bamgpu_t;

/* ---- life cycle -------------------------------------------------------------------- */

/* replaces LoadBamObjects (src/BaM_tools.f90:133-143): validates, derives the vector layout,
   uploads the tables to HBM of CUDA device `device`.  Fails (err>0) if no CUDA device. */
int bamgpu_create(bamgpu_t** h, const bamgpu_problem* p, int device, char* mess);
void bamgpu_destroy(bamgpu_t* h);
int bamgpu_get_sizes(const bamgpu_t* h, bamgpu_sizes* s);
/* global index of this handle's chain 0 (multi-GPU: chains are sharded across ranks and the RNG
   streams are keyed by the GLOBAL chain index, so results do not depend on the GPU count) */
int bamgpu_set_chain_offset(bamgpu_t* h, int64_t offset);
/* tuning / test switches: "force_generic" (1 = never use the specialised fast kernels), "no_select3" (1 = skip the
   2-read envelope selection), "only_radix" (1 = every large column through the last-resort radix selection),
   "no_partial" (1 = the sampler re-evaluates every observation for every proposal), "no_fast_sweep" (1 = the generic
   one-launch sweep over latent inputs instead of the Linear-specialised one).  err=1 if unknown. */
int bamgpu_set_option(bamgpu_t* h, const char* name, int value);

/* name -> enum helpers (return 0 when the name is unknown) */
int bamgpu_dist_id(const char* name);
int bamgpu_dist_npar(int dist);
int bamgpu_sigmafunc_id(const char* name);
int bamgpu_sigmafunc_npar(int sigmafunc);
int bamgpu_model_id(const char* name);

/* C (k x k correlation, unit diagonal) -> M = C^-1 - I, as LoadBamObjects does with
   choles_invrt (src/BaM_tools.f90:398-414).  err=1 if not positive definite. */
int bamgpu_prior_corr_to_M(int k, const double* C, double* M, char* mess);

/* ---- log-posterior: replaces Posterior_wrapper (src/BaM_tools.f90:3467-3517) ---------- */

/* K parameter vectors, x is P x K column-major (vector k at x + k*P), HOST memory.
   fx[K]; feas[K], isnull[K] (0/1); dpar is nDpar x K column-major or NULL. */
int bamgpu_logpost(bamgpu_t* h, int K, const double* x, double* fx, int32_t* feas, int32_t* isnull,
                   double* dpar, char* mess);
/* prior / likelihood / hyper split, for DIC (src/BaM_tools.f90:1626-1646) and tests */
int bamgpu_logpost_parts(bamgpu_t* h, int K, const double* x, double* prior, double* lkh, double* hyper,
                         int32_t* feas, int32_t* isnull, char* mess);

/* replaces GetSim_calibration (src/BaM_tools.f90:3619-3680), the model run behind BaM_Residual (:754-878):
   one vector x (P) through the model on the calibration inputs of `p` (the same description the handle was created
   from; caller order).  Xtrue (nobs x nX, may be NULL) receives the unfolded "true" inputs, Ysim (nobs x nY) the
   outputs (rows of unfeas_flag where the model is infeasible); *feas = 1 iff every row is feasible. */
int bamgpu_calibration_run(bamgpu_t* h, const bamgpu_problem* p, const double* x, double* Xtrue, double* Ysim,
                           int32_t* feas, char* mess);
/* the same with the model's state variables (GetSim_calibration's `state`, the state columns of Results_Residuals,
   src/BaM_tools.f90:793-794,860-873): state is nobs x nState (bamgpu_sizes.nState; SFD: the transition stage kappa),
   rows of unfeas_flag where the model is infeasible, undefRN where the model does not define the state. */
int bamgpu_calibration_run_states(bamgpu_t* h, const bamgpu_problem* p, const double* x, double* Xtrue, double* Ysim,
                                  double* state, int32_t* feas, char* mess);

/* resident variant: upload once, evaluate many times with everything in HBM */
int bamgpu_chains_upload(bamgpu_t* h, int K, const double* x, char* mess);
int bamgpu_logpost_resident(bamgpu_t* h, char* mess); /* asynchronous on the handle's stream */
int bamgpu_chains_download(bamgpu_t* h, double* fx, int32_t* feas, int32_t* isnull, double* dpar, char* mess);
int bamgpu_sync(bamgpu_t* h, char* mess);

/* ---- sampler: replaces BaM_Fit + Adaptive_Metro_OAAT (src/BaM_tools.f90:545-623) ------- */

/* K chains in lockstep, one-at-a-time adaptive Metropolis.  start/start_std are P x K.
   Rows kept: iterations it with it > burn and (it-burn) % keep_every == 0 (it = 1..nAdapt*nCycles);
   out_rows is (P + 1 + nDpar) x nKeep x K (chain-major blocks; row = [x, logpost, dpar]) or NULL.
   Final states are left resident (bamgpu_chains_download) and running per-chain moments of the
   kept rows are accumulated for bamgpu_moments.  The device time of the iteration loop is recorded as one
   step (bamgpu_step_times). */
int bamgpu_fit(bamgpu_t* h, int K, const double* start, const double* start_std, int nAdapt, int nCycles,
               double MinMoveRate, double MaxMoveRate, double DownMult, double UpMult, uint64_t seed,
               int burn, int keep_every, double* out_rows, int64_t* n_keep, char* mess);

/* per-chain moments of the kept rows: count[K], mean[P x K], m2[P x K] (sum of squared deviations).
   If dev_ptrs != 0 the three pointers are DEVICE pointers (for an NCCL all-gather across ranks). */
int bamgpu_moments(bamgpu_t* h, double* count, double* mean, double* m2, int dev_ptrs, char* mess);
/* re-accumulate the per-chain moments over the COOKED sample (BaM_LoadMCMC burn / slim, src/BaM_tools.f90:672-677): rows
   first, first+every, ... of the rows the last bamgpu_fit kept (needs out_rows != NULL in that call), so that the
   Gelman-Rubin statistic describes the sample that is summarised and propagated, not the burn-in transient */
int bamgpu_moments_window(bamgpu_t* h, int64_t first, int every, char* mess);
/* Gelman-Rubin potential scale reduction from (gathered) per-chain moments; host arrays. */
int bamgpu_rhat(int K, int P, const double* count, const double* mean, const double* m2, double* rhat,
                char* mess);

/* ---- prediction: replaces BaM_Prediction + BaM_ProcessSpag ----------------------------- */

/* mc   : Nrep x P column-major (the cooked MCMC matrix), HOST; may be NULL if maxpost run
   mode : P (maxpost), used when !doParametric (theta) -- gamma always comes from mc (:1085)
   xspag: nX pointers, matrix i is nobs_pred x nspag[i] column-major
   t0,t1: half-open range of prediction inputs handled by this call (multi-GPU shards inputs)
   env  : (t1-t0) x 7 x nY column-major per output, or NULL
   spag : Nrep x (t1-t0) x nY column-major (Fortran Spag(nrep,nobs), src/BaM_tools.f90:1327), or NULL
   Noise for (rep, t, y) is a pure function of (seed, rep, t, y): results do not depend on sharding. */
int bamgpu_predict(bamgpu_t* h, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                   const double* const* xspag, const int32_t* nspag, int doParametric,
                   const int32_t* doRemnant, uint64_t seed, int t0, int t1, double* env, double* spag,
                   char* mess);
/* bamgpu_predict with the state spaghetti / envelopes of BaM_Prediction (src/BaM_tools.f90:1071-1075,1121-1132): the
   model's state variables are nState extra outputs placed AFTER the nY outputs, without structural-error noise:
   env is 7 x (t1-t0) x (nY + nState), spag (t1-t0)*Nrep x (nY + nState) in the layout of bamgpu_predict. */
int bamgpu_predict_states(bamgpu_t* h, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                          const double* const* xspag, const int32_t* nspag, int doParametric, const int32_t* doRemnant,
                          uint64_t seed, int t0, int t1, double* env, double* spag, char* mess);

/* resident variant for measurement: the sample matrix and inputs stay in HBM between calls */
int bamgpu_predict_upload(bamgpu_t* h, int64_t Nrep, const double* mc, const double* mode, int nobs_pred,
                          const double* const* xspag, const int32_t* nspag, char* mess);
int bamgpu_predict_resident(bamgpu_t* h, int doParametric, const int32_t* doRemnant, uint64_t seed, int t0,
                            int t1, char* mess);
int bamgpu_predict_download(bamgpu_t* h, double* env, char* mess);

/* ---- multi-GPU: the two exchanges of the path on an NCCL communicator ------------------------------
   Chains shard across GPUs (observations replicated), prediction inputs shard across GPUs (samples replicated): no
   collective on the data path.  One communicator rank per GPU.  NCCL (libnccl.so.2) is loaded at the first call. */
#define BAMGPU_COMM_ID_BYTES 128
typedef struct bamgpu_comm bamgpu_comm_t;
/* one process per GPU: rank 0 creates the id, the host broadcasts the 128 bytes (MPI_Bcast, a file, ...), every rank
   joins with its own `device` */
int bamgpu_comm_unique_id(char* id128, char* mess);
int bamgpu_comm_init_rank(bamgpu_comm_t** c, int nranks, int rank, const char* id128, int device, char* mess);
/* one process, ndev GPUs (ncclCommInitAll): comms[i] is the rank-i communicator on devices[i] (NULL: 0..ndev-1); the
   per-rank calls below must then be issued concurrently from one host thread per GPU */
int bamgpu_comm_init_all(bamgpu_comm_t** comms, int ndev, const int* devices, char* mess);
void bamgpu_comm_destroy(bamgpu_comm_t* c);
int bamgpu_comm_info(const bamgpu_comm_t* c, int* nranks, int* rank, int* nccl_version);
/* Gelman-Rubin across ALL chains of ALL ranks after bamgpu_fit (every rank holds the same number of chains): the
   per-chain moments {count, mean[P], M2[P]} are packed into one message, all-gathered (one ncclAllGather) and the
   statistic is computed redundantly on every rank.  rhat[P]; count_all[Ktot], mean_all / m2_all [P x Ktot] may be NULL;
   *ms (may be NULL) = device time of the collective. */
int bamgpu_moments_allgather(bamgpu_t* h, bamgpu_comm_t* c, double* rhat, double* count_all, double* mean_all,
                             double* m2_all, double* ms, char* mess);
/* prediction inputs sharded as t in [rank*per, min(nobs_pred,(rank+1)*per)), per = ceil(nobs_pred / nranks): after
   bamgpu_predict_resident on that slice, gathers the envelope slices of all ranks; env_all is nobs_pred x 7 x nOut
   column-major per output (the layout bamgpu_predict returns for t0 = 0, t1 = nobs_pred), on every rank. */
int bamgpu_envelope_allgather(bamgpu_t* h, bamgpu_comm_t* c, int nobs_pred, double* env_all, double* ms, char* mess);
/* in-place sum over ranks of a DEVICE buffer of `count` elements; type: 0 = uint64, 1 = double, 2 = uint32 (the per-input
   histograms of the sample-sharded envelope selection); on `cuda_stream` (asynchronous) or, if NULL, on the communicator's
   own stream followed by a synchronisation */
int bamgpu_comm_allreduce_sum(bamgpu_comm_t* c, void* dev_buf, size_t count, int type, void* cuda_stream, char* mess);
/* SAMPLE-sharded prediction, for few inputs and many samples (a 41-point stage grid, 10^7 samples): every rank uploads ITS
   rows of the sample matrix (bamgpu_predict_upload with Nrep = this rank's count, all the inputs) and propagates them;
   rep_base = global index of the rank's first row (keys the structural-error noise and the input spaghetti exactly as a
   single-GPU run would), nrep_total = rows of all ranks.  The quantiles are selected exactly by a radix walk over digit
   histograms that are summed across the ranks (6 all-reduces of counts per tile of inputs), count / sum / sum of squares
   by all-reduces of doubles.  Every rank ends with the full envelopes of [t0, t1) (bamgpu_predict_download). */
int bamgpu_predict_sample_sharded(bamgpu_t* h, bamgpu_comm_t* c, int64_t rep_base, int64_t nrep_total, int doParametric,
                                  const int32_t* doRemnant, uint64_t seed, int t0, int t1, char* mess);

/* ---- measurement helpers ---------------------------------------------------------------- */

/* number of kernel launches issued by this handle since creation, and the device time (ms) of the
   last resident evaluation's dominant kernel measured with CUDA events on the handle's stream */
int64_t bamgpu_launch_count(const bamgpu_t* h);
double bamgpu_last_kernel_ms(const bamgpu_t* h);
/* CUDA-event timing on the handle's stream.  bamgpu_step_begin/end bracket one step (any sequence of
   resident calls); every resident evaluation also brackets its dominant kernel.  After bamgpu_sync,
   bamgpu_step_times / bamgpu_kernel_times return the recorded durations (ms) and clear them. */
int bamgpu_step_begin(bamgpu_t* h);
int bamgpu_step_end(bamgpu_t* h);
int bamgpu_step_times(bamgpu_t* h, double* ms, int cap);
int bamgpu_kernel_times(bamgpu_t* h, double* ms, int cap);
/* the other two kernels of a resident log-posterior evaluation: the per-vector parameter table (priors, copula, derived
   values) before the dominant kernel and the fixed-order sum of the tile partials after it */
int bamgpu_phase_times(bamgpu_t* h, double* prep_ms, double* final_ms, int cap);
/* write `bytes` of scratch HBM on the handle's stream (L2 flush between timed iterations) */
int bamgpu_flush_l2(bamgpu_t* h, size_t bytes);
/* FP64 DFMA peak micro-benchmark (register-resident FMA chains), TFLOP/s counting 2 flop per DFMA */
int bamgpu_fp64_peak(int device, double* tflops, char* mess);
/* page-lock / unlock a caller-owned host buffer so the H2D / D2H copies of the calls above are DMA copies */
int bamgpu_pin(void* ptr, size_t bytes);
int bamgpu_unpin(void* ptr);
const char* bamgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* BAM_GPU_H */
