// bam_handle.h -- the opaque bamgpu_t: one GPU-resident problem + its chain / prediction workspaces.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <utility>
#include <vector>

#include "bam_device.cuh"
#include "bam_host.h"

#define BAM_BLOCK 256

struct bamgpu_comm;

struct LogpostPlan {
    int TX = 256, TY = 1;  // block = TX observation lanes x TY chain rows
    int CT = 32;           // chains per block tile
    int nTiles = 0;        // observation tiles
    size_t smem = 0;
    int logOff = 0;        // offset (in doubles) of the log table inside the dynamic shared memory
};

struct bamgpu_handle {
    int device = 0;
    cudaStream_t stream = nullptr;
    bam::HostLayout L;
    DevProblem dp;
    std::vector<void*> owned;  // device allocations that live as long as the handle

    // ---- chain workspace (capacity Kcap vectors) ----
    int Kcap = 0, K = 0;
    double* d_x = nullptr;      // P x K (vector k at k*P)
    double* d_tab = nullptr;    // tabStride x K
    double* d_part = nullptr;   // 2 x K x nTiles
    double *d_prior = nullptr, *d_lkh = nullptr, *d_hyper = nullptr, *d_fx = nullptr, *d_dpar = nullptr;
    int* d_status = nullptr;
    void* d_sedObs = nullptr;      // Sediment fast path: per-row invariants (SedObs[n]) of the calibration inputs
    int* d_finalArrive = nullptr;  // logpost_final_combo: blocks of a 32-chain group that have finished (re-armed by the kernel)
    LogpostPlan plan;
    // fast-path tiling: tiles never straddle VAR combinations, so per-combination likelihood sums exist and a
    // proposal on a VAR parameter only re-evaluates the tiles of the combinations it touches (bam_logpost.cu)
    int fpTobs = 0, fpTiles = 0;
    bool fpYuOk = true;    // every Yu <= 2^86 and |Y| <= 2^100: the group path of logpost_baratin_cb2 multiplies up to five variances (and residuals)
    size_t tileAttr = 0;   // same for logpost_tile
    size_t chainAttr = 0;  // dynamic shared-memory limit raised for the picked logpost_chain instantiation (bytes)
    bool fp2Attr = false;  // dynamic shared-memory limit of the picked logpost_baratin_cb2 instantiation raised
    int smCount = 148;
    int *d_fpStart = nullptr, *d_fpEnd = nullptr, *d_fpCombo = nullptr, *d_fpComboTile0 = nullptr;
    // packed tiles of logpost_baratin_cb2 (bam_baratin_cb2.cuh): one bulk copy per tile; offsets in doubles, sizes in bytes
    double* d_fpPacked = nullptr;
    long long* d_fpOff = nullptr;
    int* d_fpBytes = nullptr;
    int* d_fpSched = nullptr;  // {next item, blocks done}: re-armed by the kernel itself
    int fpMaxDoubles = 0;
    int* d_fpOrder = nullptr;    // queue order of the full launch (most expensive items first), or unused
    float* d_fpCost = nullptr;   // measured cost of every item of the last full launch
    int fpOrderItems = 0;        // items the two arrays describe (0: none)
    bool fpOrderValid = false;
    long long fpFullLaunches = 0;
    std::vector<int> comboStartHost, fpComboTile0;
    double *d_lkCcur = nullptr, *d_lkCcand = nullptr;  // nCombo x Kcap: per-combination log-likelihood of the current / candidate state
    std::vector<int*> d_compTiles, d_compAffected;      // per vector component: tile ids / 0-1 flag per combination (lazy)
    std::vector<int> compTileCnt;
    bool lkCvalid = false;
    long long chainOffset = 0;
    bool noPartial = false;     // tests / measurement: always re-evaluate every observation in the sampler
    bool onlyRadix = false;     // tests: every large column through envelope_radix
    bool noSelect3 = false;     // tests: skip the 2-read envelope selection
    bool forceGeneric = false;  // tests: disable the specialised fast kernels  // global index of local chain 0 (multi-GPU sharding; RNG streams)
    bool noFastSweep = false;   // option "no_fast_sweep": keep the generic latent sweep (tests compare the two)

    // ---- sampler state ----
    double *d_std = nullptr, *d_delta = nullptr, *d_fxcur = nullptr, *d_dparcur = nullptr;
    int* d_moves = nullptr;
    double *d_mcount = nullptr, *d_mmean = nullptr, *d_mm2 = nullptr;  // per-chain running moments
    double* d_rows = nullptr;
    size_t rowsCap = 0;
    bool haveMoments = false;
    long long rowsKept = 0;  // rows per chain the last fit left in d_rows (0: none stored)

    // ---- prediction workspace ----
    long long predNrep = 0;
    int predNobs = 0;
    double *d_mc = nullptr, *d_mode = nullptr;
    std::vector<double*> d_xspag;
    std::vector<int> predNspag;
    double** d_xspagPtrs = nullptr;
    int* d_nspag = nullptr;
    double* d_ptab = nullptr;  // per-replication parameter table
    int* d_pstatus = nullptr;
    double* d_V = nullptr;     // materialised spaghetti tile: Nrep x tileT per output
    size_t vCap = 0;
    double* d_side = nullptr;  // side buffer of the columns deferred to one radix launch after the last tile (bam_predict.cu)
    size_t sideCap = 0;        // ... in doubles
    double* d_env = nullptr;   // nT x 7 x nY of the last resident prediction
    int predT0 = 0, predT1 = 0;
    double* d_series = nullptr;  // time-recursive models: outputs of all vectors [y][t][vector]
    size_t seriesCap = 0;
    double* d_seriesXpred = nullptr;  // prediction inputs of a time-recursive model (series order)
    struct bamgpu_comm* shardComm = nullptr;  // sample-sharded prediction in progress (bamgpu_predict_sample_sharded)
    long long shardRepBase = 0, shardNrepTotal = 0;
    int predNout = 0;  // outputs (+ state variables) of the last prediction: rows of d_env
    size_t envCap = 0;

    // ---- measurement ----
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    double lastMs = 0.0;
    long long launches = 0;
    std::vector<cudaEvent_t> evPool;            // recycled events
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> stepEv, kernEv;  // recorded, not yet read
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prepEv, finalEv;  // log-posterior: parameter-table kernel, final-sum kernel
    cudaEvent_t stepOpen = nullptr;
    bool recordKernels = false;
    void* d_flush = nullptr;
    size_t flushCap = 0;
};

// internal entry points shared between translation units
namespace bam {

struct CudaError {
    std::string what;
};
#define BAM_CUDA(call)                                                                                  \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            throw bam::CudaError{std::string(#call) + ": " + cudaGetErrorString(e_)};                   \
    } while (0)

void ensure_chain_capacity(bamgpu_handle* h, int K);
cudaEvent_t take_event(bamgpu_handle* h);
// evaluate K vectors at d_x (optionally with component `comp` of every vector shifted by d_delta[k]);
// results in d_prior/d_lkh/d_hyper/d_fx/d_status/d_dpar.  Asynchronous on h->stream.
void launch_logpost(bamgpu_handle* h, int K, const double* d_x, int comp, const double* d_delta, bool timeMain,
                    bool partial = false);
bool fast_path_keeps_combo_sums(const bamgpu_handle* h, int K);

void launch_table(bamgpu_handle* h, int K, const double* d_x);
// chain-parallel generic likelihood kernel (bam_logpost_chain.cu): tiles written to d_part, 0 = not eligible
int launch_logpost_chain(bamgpu_handle* h, const DevProblem& p, int K, int comp);
int logpost_chain_tiles(int n);
// latent inputs on many observations (bam_logpost_tile.cu): tiles written to d_part, 0 = not eligible
int launch_logpost_tile(bamgpu_handle* h, const DevProblem& p, int K, const double* d_x, int comp);
void ensure_series_capacity(bamgpu_handle* h, size_t count);
void launch_series(bamgpu_handle* h, const DevProblem& p, int K, const double* tab, int stride, int thOff, int* status, int nOut);
int launch_calibration_run(bamgpu_handle* h, const double* d_x1, int n, const double* d_xtrue, const int* d_combo,
                           double* d_ysim, double* d_state /* n x nState, may be null */);

void launch_fit(bamgpu_handle* h, int K, int nAdapt, int nCycles, double minMR, double maxMR, double downMult,
                double upMult, unsigned long long seed, int burn, int keepEvery, bool storeRows, long long nKeep);

void launch_window_moments(bamgpu_handle* h, int K, long long nKeep, long long first, int every);

void launch_predict(bamgpu_handle* h, int doParametric, const int* doRemnant, unsigned long long seed, int t0, int t1,
                    double* h_spag /* host, may be null */, bool withStates = false);

}  // namespace bam
