// bam_gpu_main.cpp -- `bam_gpu`: the host driver that keeps the reference's plugin surface
// (Config_BaM.txt + Config_*.txt in, Results_MCMC / Results_MCMC_Cooked / Results_Summary / *.spag /
// *.env / *.monitor / INFO_BaM.txt out) on top of the C ABI of include/bam_gpu.h.
//
// It mirrors the orchestration of program BaM_main (src/BaM_main.f90:100-540): read configs ->
// LoadBamObjects -> [MCMC -> cook] -> [summary] -> [predictions].  The reference runs ONE chain; here
// `--chains K` runs K chains in lockstep on the GPU (K=1 is the like-for-like setting), Results_MCMC.txt
// holds chain 1, the cooked file pools the burnt/slimmed rows of all chains and Results_Rhat.txt (new)
// reports the Gelman-Rubin statistic.  Residual diagnostics (BaM_Residual) and DIC (BaM_computeDIC) re-use the
// device kernels through bamgpu_calibration_run / bamgpu_logpost_parts.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "bam_config.h"

using namespace bamhost;

static const char* kVersion = "bam_gpu 0.1 (B200 hot path of BaM 1.1.1)";

// Fortran edit descriptor e15.6 (fmt_numeric, src/BaM_tools.f90:116): 0.dddddd E+xx, right-justified in 15
static std::string e15_6(double v) {
    char buf[64];
    if (std::isnan(v)) return "            NaN";
    if (std::isinf(v)) return v > 0 ? "       Infinity" : "      -Infinity";
    if (v == 0.0) return "   0.000000E+00";
    snprintf(buf, sizeof buf, "%.5E", std::fabs(v));  // d.dddddE+xx, correctly rounded to 6 significant digits
    std::string digits;
    digits += buf[0];
    digits.append(buf + 2, 5);
    int ex = atoi(strchr(buf, 'E') + 1) + 1;
    char out[64];
    if (std::abs(ex) < 100) snprintf(out, sizeof out, "%s0.%sE%c%02d", v < 0 ? "-" : "", digits.c_str(), ex < 0 ? '-' : '+', std::abs(ex));
    else snprintf(out, sizeof out, "%s0.%s%c%03d", v < 0 ? "-" : "", digits.c_str(), ex < 0 ? '-' : '+', std::abs(ex));
    std::string s(out);
    if (s.size() < 15) s = std::string(15 - s.size(), ' ') + s;
    return s;
}

static std::string a15(const std::string& s) {  // fmt_string 'A15' applied to a long character variable
    std::string t = s.substr(0, 15);
    t.resize(15, ' ');
    return t;
}

static void write_table(const std::string& file, const std::vector<std::string>& head, const std::vector<double>& rows,
                        size_t nrow, size_t ncol) {
    FILE* f = fopen(file.c_str(), "w");
    if (!f) throw std::runtime_error("problem opening file " + file);
    std::string line;
    for (auto& h : head) line += a15(h);
    fprintf(f, "%s\n", line.c_str());
    for (size_t r = 0; r < nrow; ++r) {
        line.clear();
        for (size_t c = 0; c < ncol; ++c) line += e15_6(rows[r * ncol + c]);
        fprintf(f, "%s\n", line.c_str());
    }
    fclose(f);
}

static void write_monitor(const std::string& file, long long i, long long n) {  // writeMonitor (:4273-4279)
    FILE* f = fopen(file.c_str(), "w");
    if (!f) return;
    fprintf(f, "%lld/%lld\n", i, n);
    fclose(f);
}

// GetEmpiricalQuantile: declared convention (BMSL un-vendored) = Hazen, as on the device
static double hazen(const std::vector<double>& xs, double p) {
    const double n = (double)xs.size(), h = p * n + 0.5;
    if (h <= 1.0) return xs.front();
    if (h >= n) return xs.back();
    const long long i = (long long)std::floor(h);
    return xs[i - 1] + (h - (double)i) * (xs[i] - xs[i - 1]);
}

// GetEmpiricalStats(all15=...) (BMSL): N, min, max, range, mean, median, Q10, Q25, Q75, Q90, std, var, CV, skew, kurt
static void stats15(std::vector<double> x, double* out) {
    std::sort(x.begin(), x.end());
    const double n = (double)x.size();
    double mean = std::accumulate(x.begin(), x.end(), 0.0) / n, m2 = 0, m3 = 0, m4 = 0;
    for (double v : x) { double d = v - mean; m2 += d * d; m3 += d * d * d; m4 += d * d * d * d; }
    const double var = n > 1 ? m2 / (n - 1) : 0.0, sd = std::sqrt(var);
    out[0] = n; out[1] = x.front(); out[2] = x.back(); out[3] = x.back() - x.front(); out[4] = mean;
    out[5] = hazen(x, 0.5); out[6] = hazen(x, 0.1); out[7] = hazen(x, 0.25); out[8] = hazen(x, 0.75); out[9] = hazen(x, 0.9);
    out[10] = sd; out[11] = var; out[12] = mean != 0.0 ? sd / mean : NAN;
    out[13] = sd > 0 ? (m3 / n) / (sd * sd * sd) : NAN;
    out[14] = sd > 0 ? (m4 / n) / (var * var) : NAN;
}

static void check(int err, const char* mess) {
    if (err > 0) throw std::runtime_error(mess);
    if (err < 0) fprintf(stderr, "WARNING: %s\n", mess);
}

// BaM_getPriorMC (:4008-4088): independent draws from the priors; mode of each prior for the maxpost vector
static bool prior_draw(const Prior& pr, std::mt19937_64& g, double& x, double& mode) {
    std::normal_distribution<double> N(0, 1);
    std::uniform_real_distribution<double> U(0, 1);
    const auto& p = pr.par;
    if (pr.dist == "Gaussian") { x = p[0] + p[1] * N(g); mode = p[0]; return true; }
    if (pr.dist == "Uniform") { x = p[0] + (p[1] - p[0]) * U(g); mode = 0.5 * (p[0] + p[1]); return true; }
    if (pr.dist == "LogNormal") { x = std::exp(p[0] + p[1] * N(g)); mode = std::exp(p[0] - p[1] * p[1]); return true; }
    if (pr.dist == "Exponential") { x = p[0] - p[1] * std::log(U(g)); mode = p[0]; return true; }
    if (pr.dist == "Triangle") {
        double u = U(g), fc = (p[0] - p[1]) / (p[2] - p[1]);
        x = u < fc ? p[1] + std::sqrt(u * (p[2] - p[1]) * (p[0] - p[1])) : p[2] - std::sqrt((1 - u) * (p[2] - p[1]) * (p[2] - p[0]));
        mode = p[0];
        return true;
    }
    return false;  // flat priors cannot be sampled (BaM_Message 8)
}

// model%StateName (XtraSetup, ModelLibrary_tools.f90:617-621)
static std::string state_name(const std::string& modelID, int i) {
    if (modelID == "SFD" && i == 0) return "TransitionStage";
    if ((modelID == "SFDTidal_Qmec" || modelID == "SFDTidal_Qmec2") && i < 3) return std::string(i == 0 ? "pressure" : (i == 1 ? "friction" : "advection"));  // :676
    if ((modelID == "AlgaeBiomass" || modelID == "DynamicVegetation") && i < 5) { static const char* nm[5] = {"Fb", "Ft", "Fi", "Fn", "Rhyd"}; return nm[i]; }  // ModelLibrary_tools.f90:702
    return "State" + std::to_string(i + 1);
}

int main(int argc, char** argv) {
    std::string cfg = "Config_BaM.txt", dump, priorFile, yFile;
    unsigned long long seed = 20261019ull;
    int K = 1, device = 0, gpus = 1;
    bool savePrior = false, earlyStop = false, oneRun = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        auto need = [&](const char* what) -> std::string {
            if (i + 1 >= argc) { fprintf(stderr, "%s requires %s\n", a.c_str(), what); exit(1); }
            return argv[++i];
        };
        if (a == "-cf" || a == "--config") cfg = need("a path to a file");
        else if (a == "-sd" || a == "--seed") seed = strtoull(need("the seed as an integer number").c_str(), nullptr, 10);
        else if (a == "-sp" || a == "--saveprior") { savePrior = true; priorFile = need("a file name"); }
        else if (a == "-rd" || a == "--random") seed = (unsigned long long)std::random_device{}() * 2654435761ull + (unsigned long long)time(nullptr);
        else if (a == "-dr" || a == "--dontrun") earlyStop = true;
        else if (a == "-or" || a == "--onerun") { oneRun = true; yFile = need("a file name"); }
        else if (a == "--chains") K = atoi(need("a number of chains").c_str());
        else if (a == "--device") device = atoi(need("a device ordinal").c_str());
        else if (a == "--gpus") gpus = atoi(need("a number of GPUs").c_str());
        else if (a == "--dump-problem") dump = need("a file name");
        else if (a == "-v" || a == "--version") { printf(" version: %s\n", kVersion); return 0; }
        else if (a == "-h" || a == "--help") {
            printf("usage: bam_gpu [-cf Config_BaM.txt] [-sd seed | -rd] [-sp priorfile] [-dr] [-or Yfile] [--chains K] [--device d] [--gpus N]\n"
                   "               [--dump-problem file.json] [-v] [-h]\n"
                   "  -cf, --config     main configuration file (default Config_BaM.txt)\n"
                   "  -sd, --seed       seed of the random streams;  -rd, --random: seed from the clock\n"
                   "  -sp, --saveprior  save the prior sample of prior-prediction experiments in this file (workspace-relative)\n"
                   "  -dr, --dontrun    read the configuration, write INFO_BaM.txt and stop\n"
                   "  -or, --onerun     run the model once at the initial parameter values on an input file, write Y to Yfile\n"
                   "  --chains K        number of MCMC chains run in lockstep on the GPU (the reference runs 1)\n"
                   "  --gpus N          shard the K chains (and the prediction inputs) over GPUs 0..N-1 of this box: one host thread per\n"
                   "                    GPU, NCCL (ncclCommInitAll) all-gather of the per-chain moments for Results_Rhat.txt\n");
            return 0;
        } else { fprintf(stderr, "Unrecognized command-line option: %s\n", a.c_str()); return 1; }
    }
    bamgpu_t* h = nullptr;
    try {
        Workspace w;
        if (oneRun) {  // src/BaM_main.f90:215-247
            std::vector<double> Xin;
            int nIn = 0;
            load_onerun(cfg, w, Xin, nIn);
            char mess[BAMGPU_MESS_LEN];
            bamgpu_problem pb = w.problem();
            check(bamgpu_create(&h, &pb, device, mess), mess);
            std::vector<const double*> xp(w.nX);
            std::vector<int32_t> ns(w.nX, 1), dr(w.nY, 0);
            for (int i = 0; i < w.nX; ++i) xp[i] = Xin.data() + (size_t)i * nIn;
            std::vector<double> mode(w.gamma0), spag((size_t)nIn * w.nY);  // vector = [gamma] only: every theta is FIX
            check(bamgpu_predict(h, 1, nullptr, mode.data(), nIn, xp.data(), ns.data(), 0, dr.data(), seed, 0, nIn, nullptr, spag.data(), mess), mess);
            std::vector<std::string> hd;
            for (int j = 1; j <= w.nY; ++j) hd.push_back("Y" + std::to_string(j));
            std::vector<double> tab((size_t)nIn * w.nY);
            for (int t = 0; t < nIn; ++t)
                for (int j = 0; j < w.nY; ++j) tab[(size_t)t * w.nY + j] = spag[(size_t)j * nIn + t];
            write_table(w.dir + yFile, hd, tab, (size_t)nIn, (size_t)w.nY);  // BaM_WriteOutputs (:1244-1288)
            bamgpu_destroy(h);
            return 0;
        }
        load_workspace(cfg, w);
        if (!dump.empty()) {
            std::ofstream(dump) << dump_problem_json(w);
            return 0;
        }
        char mess[BAMGPU_MESS_LEN];
        bamgpu_problem pb = w.problem();
        check(bamgpu_create(&h, &pb, device, mess), mess);
        bamgpu_sizes sz;
        bamgpu_get_sizes(h, &sz);
        const int P = sz.nInfer, nD = sz.nDpar, rowlen = P + 1 + nD;
        w.make_headers(sz, sz.nCombo ? nD / sz.nCombo : 0);
        {  // INFO_BaM.txt (writeBamInfo :4349-4391), the fields RBaM reads
            FILE* f = fopen((w.dir + "INFO_BaM.txt").c_str(), "w");
            if (f) {
                fprintf(f, " \"ID\":\"MDL_%s\"\n \"nX\":%d\n \"nY\":%d\n \"MCMCheaders\":[", w.modelID.c_str(), w.nX, w.nY);
                for (size_t i = 0; i < w.headers.size(); ++i) fprintf(f, "%s\"%s\"", i ? "," : "", w.headers[i].c_str());
                fprintf(f, "]\n");
                fclose(f);
            }
        }
        if (earlyStop) { bamgpu_destroy(h); return 0; }  // "-dr": the user just wanted INFO_BaM.txt (src/BaM_main.f90:308)
        std::vector<double> cooked;  // pooled rows [x, logpost, dpar]
        size_t nCooked = 0;
        const std::string cookPath = w.dir + w.cookFile;
        if (w.doMCMC) {
            const std::vector<double> s0 = w.start_vector(), sd0 = w.start_std_vector();
            std::vector<double> start((size_t)K * P), sd((size_t)K * P);
            for (int c = 0; c < K; ++c) { std::copy(s0.begin(), s0.end(), start.begin() + (size_t)c * P); std::copy(sd0.begin(), sd0.end(), sd.begin() + (size_t)c * P); }
            const long long N = (long long)w.mcmc.nAdapt * w.mcmc.nCycles;
            std::vector<double> rows((size_t)K * N * rowlen);
            int64_t nKeep = 0;
            std::vector<double> rhatAll;  // filled by the multi-GPU path
            const long long nburnAll = std::llround((double)N * w.burnFactor);  // BaM_LoadMCMC :672-677
            if (gpus > 1) {
                // chains shard over the GPUs of the box (observations replicated, no collective on the data path); the RNG
                // streams are keyed by the GLOBAL chain index, so the rows equal those of a one-GPU run of the same K
                if (K % gpus) throw std::runtime_error("--gpus: the number of chains must be a multiple of the number of GPUs");
                const int Kloc = K / gpus;
                std::vector<bamgpu_t*> hs(gpus, nullptr);
                std::vector<bamgpu_comm_t*> comms(gpus, nullptr);
                std::vector<std::string> errs(gpus);
                hs[0] = h;
                for (int g = 1; g < gpus; ++g) check(bamgpu_create(&hs[g], &pb, g, mess), mess);
                check(bamgpu_comm_init_all(comms.data(), gpus, nullptr, mess), mess);
                std::vector<std::vector<double>> rh(gpus, std::vector<double>(P));
                std::vector<std::thread> th;
                for (int g = 0; g < gpus; ++g)
                    th.emplace_back([&, g] {
                        char m[BAMGPU_MESS_LEN];
                        int64_t nk = 0;
                        bamgpu_set_chain_offset(hs[g], (int64_t)g * Kloc);
                        int e = bamgpu_fit(hs[g], Kloc, start.data() + (size_t)g * Kloc * P, sd.data() + (size_t)g * Kloc * P, w.mcmc.nAdapt,
                                           w.mcmc.nCycles, w.mcmc.minMoveRate, w.mcmc.maxMoveRate, w.mcmc.downMult, w.mcmc.upMult, seed, 0, 1,
                                           rows.data() + (size_t)g * Kloc * N * rowlen, &nk, m);
                        if (e > 0) errs[g] = m;
                        else if (bamgpu_moments_window(hs[g], nburnAll, std::max(1, w.nSlim), m) > 0) errs[g] = m;  // the cooked sample
                        // every rank must enter the collective, failed or not
                        double ms = 0.0;
                        e = bamgpu_moments_allgather(hs[g], comms[g], rh[g].data(), nullptr, nullptr, nullptr, &ms, m);
                        if (e > 0 && errs[g].empty()) errs[g] = m;
                    });
                for (auto& t : th) t.join();
                for (int g = 0; g < gpus; ++g) bamgpu_comm_destroy(comms[g]);
                for (int g = 1; g < gpus; ++g) bamgpu_destroy(hs[g]);
                for (auto& e : errs)
                    if (!e.empty()) throw std::runtime_error(e);
                for (int g = 1; g < gpus; ++g)
                    if (rh[g] != rh[0]) throw std::runtime_error("--gpus: the ranks disagree on the Gelman-Rubin statistic");
                rhatAll = rh[0];
                nKeep = N;
            } else
            check(bamgpu_fit(h, K, start.data(), sd.data(), w.mcmc.nAdapt, w.mcmc.nCycles, w.mcmc.minMoveRate, w.mcmc.maxMoveRate,
                             w.mcmc.downMult, w.mcmc.upMult, seed, 0, 1, rows.data(), &nKeep, mess), mess);
            write_monitor(w.dir + w.fMCMC + ".monitor", N, N);
            write_table(w.dir + w.mcmc.file, w.headers, rows, (size_t)N, rowlen);  // chain 1 (like-for-like with the reference)
            // BaM_LoadMCMC burn/slim (:672-677): rows Nburn+1 : N : Nslim, Nburn = NINT(N*BurnFactor)
            const long long nburn = std::llround((double)N * w.burnFactor);
            for (int c = 0; c < K; ++c)
                for (long long r = nburn; r < N; r += std::max(1, w.nSlim)) {
                    const double* row = rows.data() + ((size_t)c * N + r) * rowlen;
                    cooked.insert(cooked.end(), row, row + rowlen);
                    ++nCooked;
                }
            write_table(cookPath, w.headers, cooked, nCooked, rowlen);
            if (K > 1) {
                std::vector<double> cnt(K), mean((size_t)K * P), m2((size_t)K * P), rhat(P);
                if (!rhatAll.empty()) rhat = rhatAll;  // all-gathered across the GPUs (bamgpu_moments_allgather)
                else {
                    // Gelman-Rubin on the COOKED sample (burn-in dropped, slimmed): the sample that is summarised and propagated
                    check(bamgpu_moments_window(h, nburnAll, std::max(1, w.nSlim), mess), mess);
                    check(bamgpu_moments(h, cnt.data(), mean.data(), m2.data(), 0, mess), mess);
                    check(bamgpu_rhat(K, P, cnt.data(), mean.data(), m2.data(), rhat.data(), mess), mess);
                }
                std::vector<std::string> hd(w.headers.begin(), w.headers.begin() + P);
                write_table(w.dir + "Results_Rhat.txt", hd, rhat, 1, P);
            }
        } else if (w.doSummary || w.doPred) {  // resume from a cooked file (src/BaM_main.f90:377-387)
            std::ifstream is(cookPath);
            if (!is) throw std::runtime_error("problem opening file " + cookPath);
            std::string l;
            std::getline(is, l);
            size_t nrow = 0;
            while (std::getline(is, l)) if (!l.empty()) ++nrow;
            std::vector<double> M = read_matrix(cookPath, (int)nrow, rowlen, 1);  // column-major
            cooked.resize(nrow * rowlen);
            for (size_t r = 0; r < nrow; ++r)
                for (int c = 0; c < rowlen; ++c) cooked[r * rowlen + c] = M[r + (size_t)c * nrow];
            nCooked = nrow;
        }
        // maxpost (BaM_LoadMCMC :686-688)
        std::vector<double> maxpost(P, 0.0);
        if (nCooked) {
            size_t best = 0;
            for (size_t r = 1; r < nCooked; ++r)
                if (cooked[r * rowlen + P] > cooked[best * rowlen + P]) best = r;
            std::copy(cooked.begin() + best * rowlen, cooked.begin() + best * rowlen + P, maxpost.begin());
        }
        if (w.doSummary && nCooked) {  // BaM_SummarizeMCMC (:1436-1522): 16 rows x (P + nDpar) columns
            const int pc = P + nD;
            std::vector<std::vector<double>> S(16, std::vector<double>(pc));
            size_t best = 0;
            for (size_t r = 1; r < nCooked; ++r)
                if (cooked[r * rowlen + P] > cooked[best * rowlen + P]) best = r;
            for (int c = 0; c < pc; ++c) {
                const int src = c < P ? c : c + 1;
                std::vector<double> col(nCooked);
                for (size_t r = 0; r < nCooked; ++r) col[r] = cooked[r * rowlen + src];
                double st[15];
                stats15(col, st);
                for (int k = 0; k < 15; ++k) S[k][c] = st[k];
                S[15][c] = cooked[best * rowlen + src];
            }
            static const char* left[16] = {"N", "Minimum", "Maximum", "Range", "Mean", "Median", "Q10%", "Q25%", "Q75%", "Q90%",
                                           "St.Dev.", "Variance", "CV", "Skewness", "Kurtosis", "MaxPost"};
            FILE* f = fopen((w.dir + w.summaryFile).c_str(), "w");
            if (!f) throw std::runtime_error("problem opening file " + w.dir + w.summaryFile);
            std::string line = a15(" ");
            for (int c = 0; c < pc; ++c) line += a15(w.headers[c < P ? c : c + 1]);
            fprintf(f, "%s\n", line.c_str());
            for (int k = 0; k < 16; ++k) {
                line = a15(left[k]);
                for (int c = 0; c < pc; ++c) line += e15_6(S[k][c]);
                fprintf(f, "%s\n", line.c_str());
            }
            fclose(f);
        }
        if (w.doSummary && nCooked && !w.dicFile.empty()) {  // BaM_computeDIC (:1526-1682)
            std::vector<double> X((size_t)nCooked * P), prior(nCooked), lkh(nCooked), hyper(nCooked), dev(nCooked);
            std::vector<int32_t> feas(nCooked), isnull(nCooked);
            for (size_t r = 0; r < nCooked; ++r) std::copy(cooked.begin() + r * rowlen, cooked.begin() + r * rowlen + P, X.begin() + r * P);
            check(bamgpu_logpost_parts(h, (int)nCooked, X.data(), prior.data(), lkh.data(), hyper.data(), feas.data(), isnull.data(), mess), mess);
            size_t best = 0;
            for (size_t r = 0; r < nCooked; ++r) {
                if (!feas[r] || isnull[r]) throw std::runtime_error("BaM_computeDIC: unfeasible or null prior / hyper density for an MCMC sample (BaM_Message 16/17)");
                const double lp = cooked[r * rowlen + P];
                lkh[r] = lp - prior[r] - hyper[r];  // :1634 -- from the stored log-posterior, as the reference does
                dev[r] = -2.0 * lkh[r];
                if (lp > cooked[best * rowlen + P]) best = r;
            }
            if (!w.xtendedMcmcFile.empty()) {
                std::vector<std::string> hd(w.headers.begin(), w.headers.begin() + P);
                for (auto s2 : {"LogPost", "LogPrior", "LogHyper", "LogLkh", "Deviance"}) hd.push_back(s2);
                std::vector<double> ext(nCooked * (size_t)(P + 5));
                for (size_t r = 0; r < nCooked; ++r) {
                    double* o = ext.data() + r * (P + 5);
                    std::copy(X.begin() + r * P, X.begin() + (r + 1) * P, o);
                    o[P] = cooked[r * rowlen + P]; o[P + 1] = prior[r]; o[P + 2] = hyper[r]; o[P + 3] = lkh[r]; o[P + 4] = dev[r];
                }
                write_table(w.dir + w.xtendedMcmcFile, hd, ext, nCooked, P + 5);
            }
            double st[15];
            stats15(dev, st);
            const double Dmean = st[4], Dmed = st[5], Ds = st[10], Dvar = st[11], Dmax = dev[best];
            FILE* f = fopen((w.dir + w.dicFile).c_str(), "w");
            if (!f) throw std::runtime_error("problem opening file " + w.dir + w.dicFile);
            const std::pair<const char*, double> rowsD[8] = {{"DIC1", 2.0 * Dmean - Dmax}, {"DIC2", Dmax + Dvar}, {"DIC3", Dmean + 0.5 * Dvar},
                                                             {"D(maxpost)", Dmax}, {"E[D]", Dmean}, {"VAR[D]", Dvar},
                                                             {"MEDIAN[D]", Dmed}, {"SDEV[D]", Ds}};
            for (auto& rd : rowsD) fprintf(f, "%s%s\n", a15(rd.first).c_str(), e15_6(rd.second).c_str());
            fclose(f);
        }
        if (w.doResidual && nCooked) {  // BaM_Residual (:754-878)
            const size_t n = (size_t)w.nobs;
            std::vector<double> Xtrue(n * w.nX), Ysim(n * w.nY), Yunb(w.Y), res(n * w.nY), stdres(n * w.nY);
            std::vector<double> state(n * (size_t)sz.nState);
            int32_t feas = 0;
            check(bamgpu_calibration_run_states(h, &pb, maxpost.data(), Xtrue.data(), Ysim.data(),
                                                sz.nState ? state.data() : nullptr, &feas, mess), mess);
            if (!feas) throw std::runtime_error("BaM_Residual: unfeasible [maxpost] parameter vector (BaM_Message 5)");
            int k = sz.nFit + sz.nGamma + sz.nLatent + sz.nXbias;
            for (int j = 0; j < w.nY; ++j) {  // un-biased outputs (:801-811)
                int nyb = 0;
                for (size_t i = 0; i < n; ++i) nyb = std::max(nyb, (int)w.Ybindx[i + j * n]);
                if (!nyb) continue;
                for (size_t i = 0; i < n; ++i)
                    if (w.Ybindx[i + j * n] > 0 && w.Y[i + j * n] != -9999.0)
                        Yunb[i + j * n] = w.Y[i + j * n] - w.Yb[i + j * n] * maxpost[k + w.Ybindx[i + j * n] - 1];
                k += nyb;
            }
            k = sz.nFit;
            for (int j = 0; j < w.nY; ++j) {  // residuals and standardised residuals (:816-830)
                const double* g = maxpost.data() + k;
                k += w.nremnant[j];
                for (size_t i = 0; i < n; ++i) {
                    const size_t q = i + j * n;
                    if (w.Y[q] == -9999.0) { res[q] = stdres[q] = -9999.0; continue; }
                    const double ys = Ysim[q], ay = std::fabs(ys);
                    double sig;
                    switch (w.sigmaFunc[j]) {  // Sigmafunk_Apply (:3521-3571)
                        case BAMGPU_SIG_CONSTANT: sig = g[0]; break;
                        case BAMGPU_SIG_LINEAR: sig = g[0] + g[1] * ay; break;
                        case BAMGPU_SIG_PROPORTIONAL: sig = g[0] * ay; break;
                        case BAMGPU_SIG_POWER: sig = g[0] + g[1] * std::pow(ay, g[2]); break;
                        case BAMGPU_SIG_EXPONENTIAL: sig = g[0] + (g[2] - g[0]) * (1.0 - std::exp(-(ay / g[1]))); break;
                        default: sig = g[0] + (g[2] - g[0]) * (1.0 - std::exp(-((ay / g[1]) * (ay / g[1])))); break;
                    }
                    res[q] = Yunb[q] - ys;
                    stdres[q] = res[q] / std::sqrt(sig * sig + w.Yu[q] * w.Yu[q]);
                }
            }
            std::vector<std::string> hd;
            for (int i = 1; i <= w.nX; ++i) hd.push_back("X" + std::to_string(i) + "_obs");
            for (int i = 1; i <= w.nX; ++i) hd.push_back("X" + std::to_string(i) + "_true");
            for (auto sfx : {"_obs", "_unbiased", "_sim", "_res", "_stdres"})
                for (int i = 1; i <= w.nY; ++i) hd.push_back("Y" + std::to_string(i) + sfx);
            for (int i = 0; i < sz.nState; ++i) hd.push_back(state_name(w.modelID, i));  // model%StateName (:860-862)
            const size_t ncol = hd.size();
            std::vector<double> tab(n * ncol);
            for (size_t i = 0; i < n; ++i) {
                size_t c = 0;
                for (int j = 0; j < w.nX; ++j) tab[i * ncol + c++] = w.X[i + j * n];
                for (int j = 0; j < w.nX; ++j) tab[i * ncol + c++] = Xtrue[i + j * n];
                for (const std::vector<double>* src : {&w.Y, &Yunb, &Ysim, &res, &stdres})
                    for (int j = 0; j < w.nY; ++j) tab[i * ncol + c++] = (*src)[i + j * n];
                for (int j = 0; j < sz.nState; ++j) tab[i * ncol + c++] = state[i + j * n];
            }
            write_table(w.dir + w.residualFile, hd, tab, n, ncol);
        }
        // prediction experiments (src/BaM_main.f90:449-530)
        for (auto& pf : w.predConfigs) {
            PredConfig pc = read_pred_config(w, pf, sz.nState);
            const bool anyState = std::any_of(pc.doState.begin(), pc.doState.end(), [](int v) { return v != 0; });
            const int nOut = w.nY + (anyState ? sz.nState : 0);
            std::vector<std::vector<double>> xs(w.nX);
            std::vector<const double*> xp(w.nX);
            for (int i = 0; i < w.nX; ++i) {
                xs[i] = read_matrix(resolve(w, pc.xspagFiles[i]), pc.nobs, pc.nspag[i], 0);
                xp[i] = xs[i].data();
            }
            const bool anyRem = std::any_of(pc.doRemnant.begin(), pc.doRemnant.end(), [](int v) { return v != 0; });
            const bool isMaxpost = !pc.doParametric && !anyRem && std::all_of(pc.nspag.begin(), pc.nspag.end(), [](int v) { return v == 1; });
            std::vector<double> mc, mode = maxpost;  // mc column-major Nrep x P
            long long Nrep = 1;
            if (pc.nsim > 0) {  // prior prediction
                std::mt19937_64 g(seed + 17);
                Nrep = isMaxpost ? 1 : pc.nsim;
                mc.assign((size_t)Nrep * P, 0.0);
                int k = 0;
                auto fill = [&](const Prior& pr) {
                    double x, md;
                    for (long long r = 0; r < Nrep; ++r) {
                        if (!prior_draw(pr, g, x, md)) throw std::runtime_error("BaM_getPriorMC: cannot sample/compute the mode of a flat prior (BaM_Message 8)");
                        mc[r + (size_t)k * Nrep] = x;
                    }
                    mode[k] = md;
                    ++k;
                };
                for (auto& p : w.theta)
                    if (p.partype != -1)
                        for (auto& pr : p.prior) fill(pr);
                for (auto& r : w.remnant)
                    for (auto& pr : r.prior) fill(pr);
                if (savePrior && !priorFile.empty()) {  // BaM_Prediction :984-1004: header + one row per prior draw
                    std::vector<std::string> hd(w.headers.begin(), w.headers.begin() + k);
                    std::vector<double> tab((size_t)Nrep * k);
                    for (long long r = 0; r < Nrep; ++r)
                        for (int c = 0; c < k; ++c) tab[(size_t)r * k + c] = mc[r + (size_t)c * Nrep];
                    write_table(w.dir + priorFile, hd, tab, (size_t)Nrep, (size_t)k);
                }
            } else if (!isMaxpost) {
                if (!nCooked) throw std::runtime_error("BaM_Prediction: no MCMC sample available");
                Nrep = (long long)nCooked;
                mc.resize((size_t)Nrep * P);
                for (long long r = 0; r < Nrep; ++r)
                    for (int c = 0; c < P; ++c) mc[r + (size_t)c * Nrep] = cooked[r * rowlen + c];
            }
            std::vector<double> env((size_t)pc.nobs * BAMGPU_NENV * nOut), spag((size_t)pc.nobs * Nrep * nOut);
            std::vector<int32_t> dr(pc.doRemnant.begin(), pc.doRemnant.end()), ns(pc.nspag.begin(), pc.nspag.end());
            check((anyState ? bamgpu_predict_states : bamgpu_predict)(h, Nrep, mc.empty() ? nullptr : mc.data(), mode.data(), pc.nobs, xp.data(), ns.data(), pc.doParametric,
                                 dr.data(), seed + 1000, 0, pc.nobs, Nrep >= 2 ? env.data() : nullptr, spag.data(), mess), mess);
            write_monitor(w.dir + pf + ".monitor", Nrep, Nrep);
            for (int y = 0; y < nOut; ++y) {  // outputs, then the requested state variables (:1071-1075,1121-1132)
                const bool isState = y >= w.nY;
                const int q = isState ? y - w.nY : y;
                if (isState && !pc.doState[q]) continue;
                const std::string spagFile = isState ? pc.sspagFiles[q] : pc.yspagFiles[q];
                const std::string envFile = isState ? pc.envFilesS[q] : pc.envFiles[q];
                const bool doTr = isState ? pc.doTransposeS[q] != 0 : pc.doTranspose[q] != 0;
                const bool doEnv = isState ? pc.doEnvelopS[q] != 0 : pc.doEnvelop[q] != 0;
                FILE* f = fopen((w.dir + spagFile).c_str(), "w");
                if (!f) throw std::runtime_error("problem opening file " + w.dir + spagFile);
                const double* S = spag.data() + (size_t)y * pc.nobs * Nrep;  // [t][rep]
                if (doTr) {  // one row per input, Nrep values (:1369-1380)
                    for (int t = 0; t < pc.nobs; ++t) {
                        for (long long r = 0; r < Nrep; ++r) fputs(e15_6(S[(size_t)t * Nrep + r]).c_str(), f);
                        fputc('\n', f);
                    }
                } else {  // one row per replication, nobs values (:1106)
                    for (long long r = 0; r < Nrep; ++r) {
                        for (int t = 0; t < pc.nobs; ++t) fputs(e15_6(S[(size_t)t * Nrep + r]).c_str(), f);
                        fputc('\n', f);
                    }
                }
                fclose(f);
                if (doEnv && Nrep >= 2) {  // :1334-1340,1424
                    static const char* hd[7] = {"Median", "q2.5", "q97.5", "q16", "q84", "Mean", "Stdev"};
                    FILE* e = fopen((w.dir + envFile).c_str(), "w");
                    if (!e) throw std::runtime_error("problem opening file " + w.dir + envFile);
                    std::string line;
                    for (auto s : hd) line += a15(s);
                    fprintf(e, "%s\n", line.c_str());
                    for (int t = 0; t < pc.nobs; ++t) {
                        line.clear();
                        for (int s = 0; s < 7; ++s) line += e15_6(env[(size_t)y * 7 * pc.nobs + (size_t)s * pc.nobs + t]);
                        fprintf(e, "%s\n", line.c_str());
                    }
                    fclose(e);
                }
            }
        }
        bamgpu_destroy(h);
    } catch (const std::exception& e) {
        if (h) bamgpu_destroy(h);
        fprintf(stderr, "BaM: FATAL ERROR: %s\n", e.what());  // BaM_ConsoleMessage(-1,...) -> STOP (:1848-1854)
        return 1;
    }
    return 0;
}
