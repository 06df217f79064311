// bam_config.h -- host-side mirror of the reference's configuration surface (the "file seam",
// SURVEY.md App. C): Config_BaM.txt and the Config_*.txt files it names, the data file, the Xtra
// files of the models, PriorCorrelation.txt.  Everything is read with the semantics of Fortran
// list-directed READ (one record per line, comma/blank separated items, quoted strings, anything
// after the requested items ignored), as the reference's Config_Read_* subroutines do
// (src/BaM_tools.f90:1963-2830, 3751-3946).
#pragma once
#include <string>
#include <vector>

#include "../../../include/bam_gpu.h"

namespace bamhost {

struct Prior {
    std::string dist;
    std::vector<double> par;
};

struct Par {  // ParType, src/BaM_tools.f90:61-69
    std::string name;
    int partype = 0;  // -1 FIX, 0 STD, 1 VAR
    int nval = 1;
    std::vector<double> init;
    std::vector<Prior> prior;
    std::string config;      // VAR config file
    std::vector<int> indx;   // per observation
};

struct RemnantSigma {
    std::string funk;
    std::vector<std::string> parname;
    std::vector<double> init;
    std::vector<Prior> prior;
};

struct McmcConfig {  // Config_Read_MCMC, :2391-2489
    std::string file;
    int nAdapt = 100, nCycles = 100;
    double minMoveRate = 0.1, maxMoveRate = 0.5, downMult = 0.9, upMult = 1.1;
    std::vector<double> thetaStd0;
    std::vector<std::vector<double>> gammaStd0;
};

struct PredConfig {  // Config_Read_Pred, :2743-2830
    std::vector<std::string> xspagFiles;
    int nobs = 0;
    std::vector<int> nspag;
    bool doParametric = false;
    std::vector<int> doRemnant;
    int nsim = -1;
    std::vector<std::string> yspagFiles;
    std::vector<int> doTranspose, doEnvelop;
    std::vector<std::string> envFiles;
    bool printCounter = false;
    // state variables (:2808-2826): read only when the model has states; a missing section = no state prediction
    std::vector<int> doState, doTransposeS, doEnvelopS;
    std::vector<std::string> sspagFiles, envFilesS;
};

struct Workspace {
    std::string dir;  // with trailing separator
    std::string fRunOptions, fModel, fXtra, fData, fMCMC, fCooking, fSummary, fResidual, fPredMaster;
    std::vector<std::string> fRemnant;
    bool doMCMC = true, doSummary = true, doResidual = true, doPred = false;
    // model
    std::string modelID;
    int nX = 0, nY = 0, ntheta = 0, nobs = 0;
    std::vector<Par> theta;
    std::vector<RemnantSigma> remnant;
    // data (column-major)
    std::vector<double> X, Xu, Xb, Y, Yu, Yb;
    std::vector<int32_t> Xbindx, Ybindx;
    // xtra
    std::vector<int32_t> xtraI;
    std::vector<double> xtraD;
    std::string xtraText;
    bool hasPriorCorr = false;
    std::vector<double> priorM;  // C^-1 - I
    McmcConfig mcmc;
    std::string cookFile = "Results_MCMC_Cooked.txt";
    double burnFactor = 0.5;
    int nSlim = 10;
    std::string summaryFile = "Results_Summary.txt", residualFile = "Results_Residuals.txt";
    std::string dicFile, xtendedMcmcFile;  // optional 2nd / 3rd record of Config_Summary (Config_Read_Summary :2682-2687)
    std::vector<std::string> predConfigs;

    // flattened views handed to the C ABI (filled by finalize())
    std::vector<int32_t> partype, nval, varIndx, priorThetaDist, priorGammaDist, sigmaFunc, nremnant;
    std::vector<double> fixValue, priorThetaPar, priorGammaPar, theta0, gamma0;
    std::vector<std::string> headers;  // MCMC headers (GetHeaders, :3950-4004), filled by make_headers()
    bamgpu_problem problem() const;
    std::vector<double> start_vector() const;                    // Packer(theta0, gamma0, X, 0) (:605)
    std::vector<double> start_std_vector() const;                // Packer(theta_std0, gamma_std0, Xu, 0.1) (:606)
    void make_headers(const bamgpu_sizes& s, int nDparModel);
};

// throws std::runtime_error with the reference's style of message on failure
void load_workspace(const std::string& configFile, Workspace& w);
// --onerun mode (src/BaM_main.f90:215-247): model + xtra + an input file; Xin = nIn x nX column-major
void load_onerun(const std::string& configFile, Workspace& w, std::vector<double>& Xin, int& nIn);
PredConfig read_pred_config(const Workspace& w, const std::string& file, int nState = 0);
std::vector<double> read_matrix(const std::string& file, int nrow, int ncol, int nskip);  // column-major
std::string resolve(const Workspace& w, const std::string& name);
std::vector<std::string> list_items(const std::string& line);
std::string dump_problem_json(const Workspace& w);

}  // namespace bamhost
