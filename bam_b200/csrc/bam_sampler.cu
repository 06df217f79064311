// bam_sampler.cu -- K3: batched one-at-a-time adaptive Metropolis, K chains in lockstep.
//
// Replaces BMSL's sequential Adaptive_Metro_OAAT as called by BaM_Fit (src/BaM_tools.f90:615-621).
// Declared semantics (BMSL is un-vendored): for every cycle { moves=0; nAdapt
// iterations of { for each component i: candidate = x + std_i*N(0,1) on component i; accept with
// probability min(1, exp(f(cand)-f(x))) if feasible and non-null }; one row [x, f(x), Dpar] per
// iteration; then std_i *= DownMult if rate_i <= MinMoveRate, *= UpMult if rate_i >= MaxMoveRate }.
// RNG = Philox4x32-10 keyed by the seed, counter (global chain index, iteration, 2*i | 2*i+1): a
// chain's trajectory does not depend on K, on the GPU count or on launch geometry.
//
// K5: Welford running moments {count, mean, M2} per (chain, component) over the kept rows; they feed
// the Gelman-Rubin diagnostic (all-gathered across ranks by the caller, NCCL).
#include "bam_dispatch.cuh"
#include "bam_fastmath.cuh"
#include "bam_handle.h"
#include "bam_models.cuh"
#include "bam_sampler_kernels.h"

namespace {

__global__ void propose_kernel(int K, int P, int comp, unsigned long long seed, long long chainOffset, unsigned it,
                               const double* __restrict__ sd, double* __restrict__ delta) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    delta[c] = sd[(size_t)c * P + comp] * philox_normal(seed, (uint64_t)(chainOffset + c), it, 2u * (unsigned)comp);
}

__global__ void accept_kernel(int K, int P, int nDpar, int comp, unsigned long long seed, long long chainOffset,
                              unsigned it, const double* __restrict__ delta, const double* __restrict__ fxCand,
                              const int* __restrict__ stCand, const double* __restrict__ dparCand, double* __restrict__ x,
                              double* __restrict__ fxCur, double* __restrict__ dparCur, int* __restrict__ moves,
                              int nCombo, const double* __restrict__ lkCcand, double* __restrict__ lkCcur) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    if (stCand[c] != 0) return;  // infeasible or null candidate: rejected
    const double u = philox_uniform(seed, (uint64_t)(chainOffset + c), it, 2u * (unsigned)comp + 1u);
    if (log(u) < fxCand[c] - fxCur[c]) {
        x[(size_t)c * P + comp] += delta[c];
        fxCur[c] = fxCand[c];
        for (int d = 0; d < nDpar; ++d) dparCur[(size_t)c * nDpar + d] = dparCand[(size_t)c * nDpar + d];
        moves[(size_t)c * P + comp] += 1;
        // per-combination likelihood sums of the accepted state (fast path with partial re-evaluation)
        if (lkCcur != nullptr)
            for (int v = 0; v < nCombo; ++v) lkCcur[(size_t)v * K + c] = lkCcand[(size_t)v * K + c];
    }
}

__global__ void init_state_kernel(int K, int nDpar, const double* __restrict__ fx, const double* __restrict__ dpar,
                                  const int* __restrict__ st, double* __restrict__ fxCur, double* __restrict__ dparCur,
                                  int* __restrict__ nBad) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= K) return;
    fxCur[c] = fx[c];
    for (int d = 0; d < nDpar; ++d) dparCur[(size_t)c * nDpar + d] = dpar[(size_t)c * nDpar + d];
    if (st[c] != 0) atomicAdd(nBad, 1);
}

// one kept iteration: store the row and update the running moments; one thread per (chain, column)
__global__ void record_kernel(int K, int P, int nDpar, long long keepIdx, long long nKeep, const double* __restrict__ x,
                              const double* __restrict__ fxCur, const double* __restrict__ dparCur,
                              double* __restrict__ rows, double* __restrict__ mcount, double* __restrict__ mmean,
                              double* __restrict__ mm2) {
    const int rowlen = P + 1 + nDpar;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)K * rowlen) return;
    const int c = (int)(q / rowlen), j = (int)(q % rowlen);
    double v;
    if (j < P) v = x[(size_t)c * P + j];
    else if (j == P) v = fxCur[c];
    else v = dparCur[(size_t)c * nDpar + (j - P - 1)];
    if (rows) rows[((size_t)c * nKeep + keepIdx) * rowlen + j] = v;
    if (j < P) {  // Welford
        const double cnt = (double)(keepIdx + 1);
        const size_t o = (size_t)c * P + j;
        const double m = mmean[o], d = v - m, mn = m + d / cnt;
        mmean[o] = mn;
        mm2[o] += d * (v - mn);
        if (j == 0) mcount[c] = cnt;
    }
}

// moments of the rows [first, nKeep) taken every `every`-th (the cooked sample: burn-in dropped, slimmed), from the rows a
// fit left on the device; one thread per (chain, component), Welford in the order of the rows
__global__ void window_moments_kernel(int K, int P, int rowlen, long long nKeep, long long first, int every,
                                      const double* __restrict__ rows, double* __restrict__ mcount, double* __restrict__ mmean,
                                      double* __restrict__ mm2) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)K * P) return;
    const int c = (int)(q / P), j = (int)(q % P);
    double cnt = 0.0, m = 0.0, s2 = 0.0;
    for (long long r = first; r < nKeep; r += every) {
        const double v = rows[((size_t)c * nKeep + r) * rowlen + j];
        cnt += 1.0;
        const double d = v - m;
        m += d / cnt;
        s2 += d * (v - m);
    }
    mmean[q] = m;
    mm2[q] = s2;
    if (j == 0) mcount[c] = cnt;
}

__global__ void adapt_kernel(int K, int P, int nAdapt, double minMR, double maxMR, double downMult, double upMult,
                             double* __restrict__ sd, int* __restrict__ moves) {
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (long long)K * P) return;
    const double rate = (double)moves[q] / (double)nAdapt;
    double s = sd[q];
    if (rate <= minMR) s *= downMult;
    if (rate >= maxMR) s *= upMult;
    sd[q] = s;
    moves[q] = 0;
}


}  // namespace

namespace bam {

void launch_fit(bamgpu_handle* h, int K, int nAdapt, int nCycles, double minMR, double maxMR, double downMult,
                double upMult, unsigned long long seed, int burn, int keepEvery, bool storeRows, long long nKeep) {
    const DevProblem& p = h->dp;
    const int P = p.P, nD = p.nDpar;
    cudaStream_t s = h->stream;
    const int TB = 128, gK = (K + TB - 1) / TB;
    const long long KP = (long long)K * P, Krow = (long long)K * (P + 1 + nD);

    BAM_CUDA(cudaMemsetAsync(h->d_moves, 0, sizeof(int) * KP, s));
    BAM_CUDA(cudaMemsetAsync(h->d_mcount, 0, sizeof(double) * K, s));
    BAM_CUDA(cudaMemsetAsync(h->d_mmean, 0, sizeof(double) * KP, s));
    BAM_CUDA(cudaMemsetAsync(h->d_mm2, 0, sizeof(double) * KP, s));

    // f(start): must be feasible (the reference aborts otherwise)
    launch_logpost(h, K, h->d_x, -1, nullptr, false);
    int* d_nbad = nullptr;
    BAM_CUDA(cudaMalloc(&d_nbad, sizeof(int)));
    BAM_CUDA(cudaMemsetAsync(d_nbad, 0, sizeof(int), s));
    init_state_kernel<<<gK, TB, 0, s>>>(K, nD, h->d_fx, h->d_dpar, h->d_status, h->d_fxcur, h->d_dparcur, d_nbad);
    int nbad = 0;
    BAM_CUDA(cudaMemcpyAsync(&nbad, d_nbad, sizeof(int), cudaMemcpyDeviceToHost, s));
    BAM_CUDA(cudaStreamSynchronize(s));
    cudaFree(d_nbad);
    h->launches += 1;
    if (nbad > 0) throw CudaError{"Adaptive_Metro_OAAT: unfeasible starting point for " + std::to_string(nbad) + " chain(s)"};

    // small data sets: the persistent warp-per-chain sampler (one launch per cycle)
    WarpSamplerKernel wk = pick_warp_sampler(p.model, p.nX, p.nY);
    const size_t wsSmem = sizeof(double) * (WS_TABLES + WS_WARPS * (size_t)(2 * P + p.tabStride + p.kM + 2));
    if (wk && !h->forceGeneric && !p.hasLatent && !p.isSeries && P <= WS_MAXP && p.n <= 4096 && wsSmem <= 200 * 1024) {
        if (wsSmem > 48 * 1024) BAM_CUDA(cudaFuncSetAttribute((const void*)wk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsSmem));
        WarpSamplerArgs a;
        a.K = K; a.nAdapt = nAdapt; a.burn = burn; a.keepEvery = keepEvery; a.nKeep = nKeep;
        a.minMR = minMR; a.maxMR = maxMR; a.downMult = downMult; a.upMult = upMult; a.seed = seed;
        a.chainOffset = h->chainOffset;
        a.x = h->d_x; a.sd = h->d_std; a.fxcur = h->d_fxcur; a.dparcur = h->d_dparcur;
        a.rows = storeRows ? h->d_rows : nullptr; a.mcount = h->d_mcount; a.mmean = h->d_mmean; a.mm2 = h->d_mm2;
        a.moves = h->d_moves;
        cudaEvent_t f0 = take_event(h), f1 = take_event(h);  // the iterations as one timed "step" (bamgpu_step_times)
        BAM_CUDA(cudaEventRecord(f0, s));
        for (int cyc = 0; cyc < nCycles; ++cyc) {
            a.it0 = (long long)cyc * nAdapt;
            a.nIterChunk = nAdapt;
            wk<<<(K + WS_WARPS - 1) / WS_WARPS, 32 * WS_WARPS, wsSmem, s>>>(p, a);
            h->launches += 1;
        }
        BAM_CUDA(cudaEventRecord(f1, s));
        if (h->stepEv.size() < 8192) h->stepEv.emplace_back(f0, f1);
        BAM_CUDA(cudaGetLastError());
        BAM_CUDA(cudaMemcpyAsync(h->d_fx, h->d_fxcur, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
        BAM_CUDA(cudaMemcpyAsync(h->d_dpar, h->d_dparcur, sizeof(double) * (size_t)K * std::max(nD, 1), cudaMemcpyDeviceToDevice, s));
        BAM_CUDA(cudaMemsetAsync(h->d_status, 0, sizeof(int) * K, s));
        h->haveMoments = true;
        return;
    }

    // components visited with one launch set per proposal: everything except the latent inputs when the one-launch
    // sweep applies (vector order: theta | gamma | latent inputs | X biases | Y biases, Packer :4202-4242)
    LatentKernel lk = (p.hasLatent && !h->forceGeneric) ? pick_latent_sweep(p.model, p.nX, p.nY) : nullptr;
    if (lk && !h->noFastSweep)
        if (LatentKernel fastSweep = pick_latent_sweep_linear(h)) lk = fastSweep;
    const int latBeg = p.nFit + p.nGamma, latEnd = latBeg + p.nLatent;
    // fast path (BaRatin, many chains): per-combination likelihood sums of the current state are kept, so that a
    // proposal on a VAR parameter only re-evaluates the observations of the combinations it touches
    const bool comboSums = fast_path_keeps_combo_sums(h, K) && !lk && !h->noPartial;
    if (comboSums) {  // the full evaluation of the start point above left its sums in d_lkCcand
        BAM_CUDA(cudaMemcpyAsync(h->d_lkCcur, h->d_lkCcand, sizeof(double) * (size_t)K * p.nCombo, cudaMemcpyDeviceToDevice, s));
        h->lkCvalid = true;
    }
    auto propose_one = [&](int i, long long it) {
        propose_kernel<<<gK, TB, 0, s>>>(K, P, i, seed, h->chainOffset, (unsigned)it, h->d_std, h->d_delta);
        launch_logpost(h, K, h->d_x, i, h->d_delta, false, comboSums);
        accept_kernel<<<gK, TB, 0, s>>>(K, P, nD, i, seed, h->chainOffset, (unsigned)it, h->d_delta, h->d_fx,
                                        h->d_status, h->d_dpar, h->d_x, h->d_fxcur, h->d_dparcur, h->d_moves, p.nCombo,
                                        comboSums ? h->d_lkCcand : nullptr, comboSums ? h->d_lkCcur : nullptr);
        h->launches += 2;
    };
    long long it = 0, kept = 0;
    cudaEvent_t f0 = take_event(h), f1 = take_event(h);  // the iterations as one timed "step" (bamgpu_step_times)
    BAM_CUDA(cudaEventRecord(f0, s));
    const size_t lkSmem = sizeof(double) * ((size_t)p.tabStride + 2 * BAM_LOGTAB_N);
    if (lk) {  // the sweep stages the chain's table row: with many VAR combinations it outgrows the default 48 KB
        if (lkSmem > 227 * 1024) throw CudaError{"bamgpu_fit: the per-chain parameter table of this problem does not fit the shared memory of the latent sweep"};
        if (lkSmem > 48 * 1024) BAM_CUDA(cudaFuncSetAttribute((const void*)lk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lkSmem));
    }
    for (int cyc = 0; cyc < nCycles; ++cyc) {
        for (int a = 0; a < nAdapt; ++a, ++it) {
            for (int i = 0; i < (lk ? latBeg : P); ++i) propose_one(i, it);
            if (lk) {
                // the table of the CURRENT state (theta, gamma, biases of every chain), then the sweep, then the
                // full log-posterior of the swept state becomes the current one
                launch_table(h, K, h->d_x);
                dim3 grid((p.n + BAM_BLOCK - 1) / BAM_BLOCK, K);
                lk<<<grid, BAM_BLOCK, lkSmem, s>>>(p, K, seed, h->chainOffset, (unsigned)it, h->d_x,
                                                                        h->d_std, h->d_moves, h->d_tab);
                launch_logpost(h, K, h->d_x, -1, nullptr, false);
                BAM_CUDA(cudaMemcpyAsync(h->d_fxcur, h->d_fx, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
                h->launches += 1;
                for (int i = latEnd; i < P; ++i) propose_one(i, it);
            }
            const long long it1 = it + 1;
            if (it1 > burn && (it1 - burn) % keepEvery == 0 && kept < nKeep) {
                record_kernel<<<(unsigned)((Krow + 255) / 256), 256, 0, s>>>(K, P, nD, kept, nKeep, h->d_x, h->d_fxcur,
                                                                             h->d_dparcur, storeRows ? h->d_rows : nullptr,
                                                                             h->d_mcount, h->d_mmean, h->d_mm2);
                ++kept;
                h->launches += 1;
            }
        }
        adapt_kernel<<<(unsigned)((KP + 255) / 256), 256, 0, s>>>(K, P, nAdapt, minMR, maxMR, downMult, upMult, h->d_std,
                                                                  h->d_moves);
        h->launches += 1;
        BAM_CUDA(cudaGetLastError());
    }
    BAM_CUDA(cudaEventRecord(f1, s));
    if (h->stepEv.size() < 8192) h->stepEv.emplace_back(f0, f1);
    h->lkCvalid = false;
    // leave the final state consistent for bamgpu_chains_download
    BAM_CUDA(cudaMemcpyAsync(h->d_fx, h->d_fxcur, sizeof(double) * K, cudaMemcpyDeviceToDevice, s));
    BAM_CUDA(cudaMemcpyAsync(h->d_dpar, h->d_dparcur, sizeof(double) * (size_t)K * std::max(nD, 1), cudaMemcpyDeviceToDevice, s));
    BAM_CUDA(cudaMemsetAsync(h->d_status, 0, sizeof(int) * K, s));
    h->haveMoments = true;
}

}  // namespace bam

namespace bam {
void launch_window_moments(bamgpu_handle* h, int K, long long nKeep, long long first, int every) {
    const DevProblem& p = h->dp;
    const long long q = (long long)K * p.P;
    window_moments_kernel<<<(unsigned)((q + 255) / 256), 256, 0, h->stream>>>(K, p.P, p.P + 1 + p.nDpar, nKeep, first, every, h->d_rows,
                                                                             h->d_mcount, h->d_mmean, h->d_mm2);
    h->launches += 1;
    BAM_CUDA(cudaGetLastError());
}
}  // namespace bam
