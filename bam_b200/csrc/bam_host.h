// bam_host.h -- host-side (C++) restatement of LoadBamObjects' derived quantities
// (src/BaM_tools.f90:133-420): vector layout, VAR combinations, latent-input map, model xtra.
// Pure host code, no CUDA types: used by the library (bam_capi.cu) and by the bam_gpu driver.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/bam_gpu.h"

namespace bam {

struct VmProgram {
    std::vector<int32_t> code;
    std::vector<double> immed;
    int stackSize = 0;
};

// Compile one TextFile formula to the reference's bytecode (TextFile_model.f90:208-229,551-857).
// Throws std::runtime_error("Syntax error") on malformed input.
VmProgram compile_formula(const std::string& formula, const std::vector<std::string>& vars);

struct TextFileModel {
    std::vector<std::string> inNames, parNames, formulas;
    std::vector<VmProgram> progs;
};
// TxtMdl_load (TextFile_model.f90:897-987) from the text of the model file
TextFileModel parse_textfile_model(const std::string& text);

struct HostLayout {
    int model = 0;
    int n = 0, nX = 0, nY = 0, ntheta = 0;
    int nFit = 0, nGamma = 0, nLatent = 0, nXbTot = 0, nYbTot = 0, P = 0;
    int nDparModel = 0, nDpar = 0, nCombo = 1, nState = 0, nVar = 0;
    std::vector<int32_t> partype, nval, thetaOff;  // thetaOff: offset in the fitted vector, -1 for FIX
    std::vector<double> fixv;
    std::vector<int32_t> sigfunc, nrem, gamOff;
    std::vector<int32_t> Xerror, nXb, nYb, offXb, offYb;
    std::vector<int32_t> latIdx;    // n x nX (column-major), -1 = not latent
    std::vector<int32_t> comboOf;   // n
    std::vector<int32_t> comboRow;  // nCombo: first observation of the combination (INFER%DparIndx - 1)
    std::vector<int32_t> comboSrc;  // nCombo x ntheta (row per combo): index into the vector or -1 (FIX)
    bool hasLatent = false, hasXbias = false, hasYbias = false, hyperInfeasible = false, hasMissing = false;
    // model xtra
    int ncontrol = 0;
    unsigned long long ctrlMask = 0;
    int sed[7] = {0, 0, 0, 0, 0, 0, 0};
    double sedPseudo[4] = {0, 0, 0, 0};
    TextFileModel txt;
    std::vector<int32_t> vmCode;
    std::vector<double> vmImmed;
    std::vector<int32_t> vmNcode, vmCodeOff, vmImmedOff;
    int vmMaxStack = 0;
    int vegGrowth = 0, vegNYear = 0, vegItmax = 0;
    double vegOpt[4] = {0, 0, 0, 0};  // xscale, fscale, tolX, tolF
    int modelOpt[4] = {0, 0, 0, 0};  // SWOT: {variant}; Orthorectification: {distortion id, npar}; Segmentation: {nS, nmin, option}
    double modelOptD = 0.0;          // Segmentation: tmin
    std::vector<double> hcsTab;      // HydraulicControl_section: h | A | w | h + A/(2w), p rows each
    int nDer = 0;  // derived values per (parameter set, combination) kept in the device table
    bool isSeries = false;  // time-recursive model: one sequential pass over the series per parameter vector (row f4)
};

// Mixture (Mixture_model.f90:119-123): output row (j-1)*nElt+e is built from the e-th input row whose event column
// (rounded) is j.  src[r] = that input row, elt[r] = e (1-based).  Returns 0, or 1 (mess set) when an event does not
// have exactly nElt rows or the series is not nEvt*nElt long (array-size mismatches in the reference).
int mixture_rows(int n, const double* evt, int nEvt, int nElt, std::vector<int>& src, std::vector<int>& elt, std::string& mess);

// Validates `p` and derives the layout.  Returns err (0 OK, >0 error) and fills mess.
int build_layout(const bamgpu_problem& p, HostLayout& L, std::string& mess);

int model_from_name(const char* name);
int dist_from_name(const char* name);
int dist_npar(int dist);
int sigmafunc_from_name(const char* name);
int sigmafunc_npar(int f);

// C (k x k, column-major, unit diagonal) -> M = C^-1 - I  (src/BaM_tools.f90:398-414)
int prior_corr_to_M(int k, const double* C, double* M, std::string& mess);

// Gelman-Rubin from per-chain moments
const char* algae_input_error(int n, const double* X, bool checkDepth = true);  // n x 5 column-major; NULL = fine
void rhat_from_moments(int K, int P, const double* count, const double* mean, const double* m2, double* rhat);

}  // namespace bam
