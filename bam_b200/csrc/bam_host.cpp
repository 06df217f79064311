// bam_host.cpp -- host-side layout derivation and the TextFile formula compiler.
#include "bam_host.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <stdexcept>

namespace bam {

// AlgaeBiomass_Apply's input checks (AlgaeBiomass_model.f90:79-99): a fatal error (err = 1), not an infeasible vector
// (checkDepth = false for DynamicVegetation, whose third input is a stage: the depth handed on is max(stage - b1, 0))
const char* algae_input_error(int n, const double* X, bool checkDepth) {
    for (int i = 0; i < n; ++i) {
        if (X[i + (size_t)1 * n] < 0.0) return "AlgaeBiomass_Apply:Irradiance<0 impossible";
        if (checkDepth && X[i + (size_t)2 * n] < 0.0) return "AlgaeBiomass_Apply:depth<0 impossible";
        if (X[i + (size_t)3 * n] < 0.0) return "AlgaeBiomass_Apply:turbidity<0 impossible";
        if (X[i + (size_t)4 * n] < 0.0) return "AlgaeBiomass_Apply:removal<0 impossible";
        if (X[i + (size_t)4 * n] > 1.0) return "AlgaeBiomass_Apply:removal>1 impossible";
    }
    return nullptr;
}


// ------------------------------------------------------------------------------------------------
// catalogues
// ------------------------------------------------------------------------------------------------
int model_from_name(const char* name) {
    if (!name) return 0;
    std::string s(name);
    if (s.rfind("MDL_", 0) == 0) s = s.substr(4);
    if (s == "Linear") return BAMGPU_MDL_LINEAR;
    if (s == "BaRatin") return BAMGPU_MDL_BARATIN;
    if (s == "Sediment") return BAMGPU_MDL_SEDIMENT;
    if (s == "TextFile") return BAMGPU_MDL_TEXTFILE;
    if (s == "Vegetation") return BAMGPU_MDL_VEGETATION;
    if (s == "SuspendedLoad") return BAMGPU_MDL_SUSPENDEDLOAD;
    if (s == "SWOT") return BAMGPU_MDL_SWOT;
    if (s == "SGD") return BAMGPU_MDL_SGD;
    if (s == "Orthorectification") return BAMGPU_MDL_ORTHO;
    if (s == "Recession_h") return BAMGPU_MDL_RECESSION_H;
    if (s == "HydraulicControl_section") return BAMGPU_MDL_HCS;
    if (s == "Segmentation") return BAMGPU_MDL_SEGMENTATION;
    if (s == "BaRatinBAC") return BAMGPU_MDL_BARATINBAC;
    if (s == "SFD") return BAMGPU_MDL_SFD;
    if (s == "Mixture") return BAMGPU_MDL_MIXTURE;
    if (s == "SFDTidal_Qmec") return BAMGPU_MDL_SFDTIDAL_QMEC;
    if (s == "SFDTidal_Qmec2") return BAMGPU_MDL_SFDTIDAL_QMEC2;
    if (s == "AlgaeBiomass") return BAMGPU_MDL_ALGAE;
    if (s == "DynamicVegetation") return BAMGPU_MDL_DYNVEG;
    if (s == "SFDTidal_Qmec0") return BAMGPU_MDL_SFDTIDAL_QMEC0;
    if (s == "SFDTidal_QmecMS") return BAMGPU_MDL_SFDTIDAL_QMECMS;
    if (s == "SFDTidal" || s == "SFDTidal2" || s == "SFDTidal4" || s == "SFDTidalJones") return BAMGPU_MDL_SFDTIDAL_LAG;
    return 0;
}

int dist_from_name(const char* name) {
    if (!name) return 0;
    static const struct { const char* n; int id; } tab[] = {
        {"Gaussian", BAMGPU_DIST_GAUSSIAN},   {"Uniform", BAMGPU_DIST_UNIFORM},
        {"Triangle", BAMGPU_DIST_TRIANGLE},   {"LogNormal", BAMGPU_DIST_LOGNORMAL},
        {"Exponential", BAMGPU_DIST_EXPONENTIAL}, {"FlatPrior", BAMGPU_DIST_FLATPRIOR},
        {"FlatPrior+", BAMGPU_DIST_FLATPRIOR_POS}, {"FlatPrior-", BAMGPU_DIST_FLATPRIOR_NEG}};
    for (auto& t : tab)
        if (!strcmp(name, t.n)) return t.id;
    return 0;
}

int dist_npar(int d) {
    switch (d) {
        case BAMGPU_DIST_GAUSSIAN: case BAMGPU_DIST_UNIFORM: case BAMGPU_DIST_LOGNORMAL: case BAMGPU_DIST_EXPONENTIAL: return 2;
        case BAMGPU_DIST_TRIANGLE: return 3;
        case BAMGPU_DIST_FLATPRIOR: case BAMGPU_DIST_FLATPRIOR_POS: case BAMGPU_DIST_FLATPRIOR_NEG: return 0;
    }
    return -1;
}

int sigmafunc_from_name(const char* name) {
    if (!name) return 0;
    static const char* names[] = {"Constant", "Linear", "Proportional", "Power", "Exponential", "Gaussian"};
    for (int i = 0; i < 6; ++i)
        if (!strcmp(name, names[i])) return i + 1;
    return 0;
}

int sigmafunc_npar(int f) {  // Sigmafunk_GetParNumber, src/BaM_tools.f90:3575-3620
    static const int np[] = {-1, 1, 2, 1, 3, 3, 3};
    return (f >= 1 && f <= 6) ? np[f] : -1;
}

// ------------------------------------------------------------------------------------------------
// TextFile formula compiler.  Semantics of R. Schmehl's fparser as modified in the reference
// (TextFile_model.f90:644-738): split at the RIGHT-MOST base-level binary occurrence of the
// lowest-priority operator, testing operators in the order + - * / ^.
// ------------------------------------------------------------------------------------------------
namespace {

enum Op { IMMED = 1, NEG, ADD, SUB, MUL, DIV, POW, ABS_, EXP_, LOG10_, LOG_, SQRT_, SINH_, COSH_, TANH_, SIN_, COS_,
          TAN_, ASIN_, ACOS_, ATAN_, IS0, ISPOS, ISNEG, ISSPOS, ISSNEG, VARBEGIN };

const char* kFuncs[] = {"abs", "exp", "log10", "log", "sqrt", "sinh", "cosh", "tanh", "sin", "cos", "tan", "asin",
                        "acos", "atan", "is0", "ispos", "isneg", "isspos", "issneg"};
const char kOps[] = {'+', '-', '*', '/', '^'};

struct Compiler {
    std::string F;
    const std::vector<std::string>& vars;
    VmProgram out;
    int sp = 0;

    Compiler(const std::string& f, const std::vector<std::string>& v) : F(f), vars(v) {}

    static bool in(char c, const char* set) { return c && strchr(set, c) != nullptr; }

    int funcAt(int b, int e) const {  // catalogue entry that prefixes F[b..e] (case-insensitive)
        for (int j = 0; j < 19; ++j) {
            int k = (int)strlen(kFuncs[j]);
            if (k > e - b + 1) continue;
            bool ok = true;
            for (int q = 0; q < k && ok; ++q) ok = std::tolower((unsigned char)F[b + q]) == kFuncs[j][q];
            if (ok) return ABS_ + j;
        }
        return 0;
    }

    bool enclosed(int b, int e) const {
        if (b > e || F[b] != '(' || F[e] != ')') return false;
        int depth = 0;
        for (int j = b + 1; j < e; ++j) {
            if (F[j] == '(') ++depth;
            else if (F[j] == ')') --depth;
            if (depth < 0) break;
        }
        return depth == 0;
    }

    bool binaryAt(int j) const {
        char c = F[j];
        if (c != '+' && c != '-') return true;
        if (j == 0) return false;
        if (in(F[j - 1], "+-*/^(")) return false;
        if (j + 1 < (int)F.size() && std::isdigit((unsigned char)F[j + 1]) && in(F[j - 1], "eEdD")) {
            bool digit = false, point = false;
            int k = j - 1;
            while (k > 0) {
                --k;
                if (std::isdigit((unsigned char)F[k])) digit = true;
                else if (F[k] == '.') {
                    if (point) break;
                    point = true;
                } else break;
            }
            if (digit && (k == 0 || in(F[k], "+-*/^("))) return false;
        }
        return true;
    }

    void byte(int b) { out.code.push_back(b); }

    double number(int b, int e) {  // [nnn][.nnn][e|E|d|D[+|-]nnn]
        int i = b;
        bool mant = false, expo = false;
        while (i <= e && (std::isdigit((unsigned char)F[i]) || F[i] == '.')) { mant |= std::isdigit((unsigned char)F[i]) != 0; ++i; }
        if (i <= e && in(F[i], "eEdD")) {
            int s = i + 1;
            if (s <= e && (F[s] == '+' || F[s] == '-')) ++s;
            int q = s;
            while (q <= e && std::isdigit((unsigned char)F[q])) { expo = true; ++q; }
            if (!expo) throw std::runtime_error("Syntax error");
            i = q;
        }
        if (!mant) throw std::runtime_error("Syntax error");
        std::string t = F.substr(b, i - b);
        for (auto& ch : t)
            if (ch == 'd' || ch == 'D') ch = 'e';
        return strtod(t.c_str(), nullptr);
    }

    int variable(int b, int e) const {
        int i = b;
        while (i <= e && !in(F[i], "+-*/^) ")) ++i;
        std::string name = F.substr(b, i - b);
        for (size_t j = 0; j < vars.size(); ++j)
            if (vars[j] == name) return (int)j + 1;
        return 0;
    }

    void emit(int b, int e) {
        if (b > e) throw std::runtime_error("Syntax error");
        if (F[b] == '+') return emit(b + 1, e);
        if (enclosed(b, e)) return emit(b + 1, e - 1);
        if (std::isalpha((unsigned char)F[b])) {
            int fn = funcAt(b, e);
            if (fn) {
                size_t par = F.find('(', b);
                if (par != std::string::npos && (int)par <= e && enclosed((int)par, e)) {
                    emit((int)par + 1, e - 1);
                    byte(fn);
                    return;
                }
            }
        } else if (F[b] == '-') {
            if (enclosed(b + 1, e)) {
                emit(b + 2, e - 1);
                byte(NEG);
                return;
            }
            if (b + 1 <= e && std::isalpha((unsigned char)F[b + 1])) {
                int fn = funcAt(b + 1, e);
                if (fn) {
                    size_t par = F.find('(', b + 1);
                    if (par != std::string::npos && (int)par <= e && enclosed((int)par, e)) {
                        emit((int)par + 1, e - 1);
                        byte(fn);
                        byte(NEG);
                        return;
                    }
                }
            }
        }
        for (int io = 0; io < 5; ++io) {
            int depth = 0;
            for (int j = e; j >= b; --j) {
                if (F[j] == ')') ++depth;
                else if (F[j] == '(') --depth;
                if (depth == 0 && F[j] == kOps[io] && binaryAt(j)) {
                    if (io >= 2 && F[b] == '-') {
                        emit(b + 1, e);
                        byte(NEG);
                        return;
                    }
                    emit(b, j - 1);
                    emit(j + 1, e);
                    byte(ADD + io);
                    --sp;
                    return;
                }
            }
        }
        int b2 = (F[b] == '-') ? b + 1 : b;
        if (b2 > e) throw std::runtime_error("Syntax error");
        if (std::isdigit((unsigned char)F[b2]) || F[b2] == '.') {
            out.immed.push_back(number(b2, e));
            byte(IMMED);
        } else {
            int v = variable(b2, e);
            if (!v) throw std::runtime_error("Syntax error");
            byte(VARBEGIN + v - 1);
        }
        ++sp;
        if (sp > out.stackSize) out.stackSize = sp;
        if (b2 > b) byte(NEG);
    }
};

std::vector<std::string> list_items(const std::string& line) {  // one list-directed record
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        char c = line[i];
        if (c == ' ' || c == '\t' || c == ',' || c == '\r') { ++i; continue; }
        if (c == '!') break;
        if (c == '"' || c == '\'') {
            size_t j = line.find(c, i + 1);
            if (j == std::string::npos) j = n;
            out.push_back(line.substr(i + 1, j - i - 1));
            i = j + 1;
        } else {
            size_t j = i;
            while (j < n && !strchr(" \t,\r!", line[j])) ++j;
            out.push_back(line.substr(i, j - i));
            i = j;
        }
    }
    return out;
}

}  // namespace

VmProgram compile_formula(const std::string& formula, const std::vector<std::string>& vars) {
    std::string f;
    for (size_t i = 0; i < formula.size(); ++i) {
        char c = formula[i];
        if (c == '*' && i + 1 < formula.size() && formula[i + 1] == '*') { f.push_back('^'); ++i; continue; }
        if (c == ' ' || c == '\t' || c == '\r' || c == '\n') continue;
        f.push_back(c);
    }
    if (f.empty()) throw std::runtime_error("Syntax error");
    int depth = 0;
    for (char c : f) {
        if (c == '(') ++depth;
        if (c == ')') --depth;
        if (depth < 0) throw std::runtime_error("Syntax error");
    }
    if (depth != 0) throw std::runtime_error("Syntax error");
    Compiler c(f, vars);
    c.emit(0, (int)f.size() - 1);
    return c.out;
}

TextFileModel parse_textfile_model(const std::string& text) {
    TextFileModel m;
    std::istringstream is(text);
    std::string line;
    auto need = [&](std::string& l) {
        if (!std::getline(is, l)) throw std::runtime_error("TxtMdl_load:problem reading model text");
    };
    need(line);
    int nIN = atoi(line.c_str());
    if (nIN < 0) throw std::runtime_error("TxtMdl_load:negative nIN not allowed");
    need(line);
    if (nIN > 0) {
        auto it = list_items(line);
        if ((int)it.size() < nIN) throw std::runtime_error("TxtMdl_load:problem reading model text");
        m.inNames.assign(it.begin(), it.begin() + nIN);
    }
    need(line);
    int nPAR = atoi(line.c_str());
    if (nPAR < 0) throw std::runtime_error("TxtMdl_load:negative nPAR not allowed");
    need(line);
    if (nPAR > 0) {
        auto it = list_items(line);
        if ((int)it.size() < nPAR) throw std::runtime_error("TxtMdl_load:problem reading model text");
        m.parNames.assign(it.begin(), it.begin() + nPAR);
    }
    need(line);
    int nOUT = atoi(line.c_str());
    if (nOUT <= 0) throw std::runtime_error("TxtMdl_load:nOUT should be at least one");
    std::vector<std::string> vars = m.inNames;
    vars.insert(vars.end(), m.parNames.begin(), m.parNames.end());
    for (int k = 0; k < nOUT; ++k) {
        need(line);
        size_t ex = line.find('!');
        if (ex != std::string::npos) line = line.substr(0, ex);
        m.formulas.push_back(line);
        m.progs.push_back(compile_formula(line, vars));
    }
    return m;
}

// ------------------------------------------------------------------------------------------------
// layout
// ------------------------------------------------------------------------------------------------
int build_layout(const bamgpu_problem& p, HostLayout& L, std::string& mess) {
    L = HostLayout();
    L.model = model_from_name(p.model_id);
    if (!L.model) { mess = "ApplyModel: Fatal: Unavailable [model%ID]"; return 1; }
    const int n = L.n = p.nobs, nX = L.nX = p.nX, nY = L.nY = p.nY, nt = L.ntheta = p.ntheta;
    if (n < 0 || nX < 1 || nY < 1 || nt < 0) { mess = "LoadBamObjects: size mismatch in input data"; return 1; }
    if (n > 0 && (!p.X || !p.Xu || !p.Xb || !p.Xbindx || !p.Y || !p.Yu || !p.Yb || !p.Ybindx)) {
        mess = "LoadBamObjects: NULL data array"; return 1;
    }
    L.partype.assign(p.partype, p.partype + nt);
    L.nval.assign(p.nval, p.nval + nt);
    L.fixv.assign(p.fix_value, p.fix_value + nt);
    L.thetaOff.assign(nt, -1);
    for (int i = 0; i < nt; ++i) {
        if (L.partype[i] == BAMGPU_PAR_FIX) continue;
        if (L.partype[i] != BAMGPU_PAR_STD && L.partype[i] != BAMGPU_PAR_VAR) { mess = "Config_Finalize: STOK parameters not implemented yet"; return 1; }
        if (L.nval[i] < 1) { mess = "Config_Finalize: nval < 1"; return 1; }
        L.thetaOff[i] = L.nFit;
        L.nFit += L.nval[i];
        if (L.partype[i] == BAMGPU_PAR_VAR) ++L.nVar;
    }
    if (L.nVar && !p.var_indx) { mess = "Config_Finalize: VAR parameter without index series"; return 1; }
    L.sigfunc.assign(p.sigma_func, p.sigma_func + nY);
    L.nrem.assign(p.nremnant, p.nremnant + nY);
    L.gamOff.assign(nY, 0);
    for (int j = 0; j < nY; ++j) {
        if (sigmafunc_npar(L.sigfunc[j]) < 0) { mess = "Sigmafunk_Apply: unknown remnant sigma function"; return 1; }
        if (sigmafunc_npar(L.sigfunc[j]) != L.nrem[j]) { mess = "Config_Read_RemnantSigma: wrong number of parameters for the sigma function"; return 1; }
        L.gamOff[j] = L.nGamma;
        L.nGamma += L.nrem[j];
    }
    for (int k = 0; k < L.nFit; ++k)
        if (dist_npar(p.prior_theta_dist[k]) < 0) { mess = "GetLogPrior: unknown prior distribution"; return 1; }
    for (int k = 0; k < L.nGamma; ++k)
        if (dist_npar(p.prior_gamma_dist[k]) < 0) { mess = "GetLogPrior: unknown prior distribution"; return 1; }

    // latent inputs and biases (:230-253); vector layout = Packer (:4202-4242)
    L.Xerror.assign(nX, 0); L.nXb.assign(nX, 0); L.nYb.assign(nY, 0); L.offXb.assign(nX, 0); L.offYb.assign(nY, 0);
    L.latIdx.assign((size_t)n * nX, -1);
    int pos = L.nFit + L.nGamma;
    for (int j = 0; j < nX; ++j)
        for (int i = 0; i < n; ++i) {
            size_t q = i + (size_t)j * n;
            double xu = p.Xu[q];
            if (xu > 0.0) { L.latIdx[q] = pos++; ++L.nLatent; L.Xerror[j] = 1; }
        }
    for (int j = 0; j < nX; ++j)  // a negative Xu in an error-affected column makes GetHyperLogPdf infeasible (:3288-3300)
        if (L.Xerror[j])
            for (int i = 0; i < n; ++i)
                if (p.Xu[i + (size_t)j * n] < 0.0) L.hyperInfeasible = true;
    L.hasLatent = L.nLatent > 0;
    for (int j = 0; j < nX; ++j) {
        int mx = 0;
        for (int i = 0; i < n; ++i) mx = std::max(mx, (int)p.Xbindx[i + (size_t)j * n]);
        L.nXb[j] = mx; L.offXb[j] = pos; pos += mx; L.nXbTot += mx;
    }
    for (int j = 0; j < nY; ++j) {
        int mx = 0;
        for (int i = 0; i < n; ++i) mx = std::max(mx, (int)p.Ybindx[i + (size_t)j * n]);
        L.nYb[j] = mx; L.offYb[j] = pos; pos += mx; L.nYbTot += mx;
    }
    L.hasXbias = L.nXbTot > 0;
    L.hasYbias = L.nYbTot > 0;
    L.P = pos;
    for (size_t q = 0; q < (size_t)n * nY; ++q)
        if (p.Y[q] == p.mv) { L.hasMissing = true; break; }

    // model xtra (XtraRead / XtraSetup, ModelLibrary_tools.f90:436-760)
    switch (L.model) {
        case BAMGPU_MDL_LINEAR:
            if (nt != nX * nY) { mess = "LM_Apply: size mismatch [theta]"; return 1; }
            break;
        case BAMGPU_MDL_BARATIN:
        case BAMGPU_MDL_BARATINBAC: {
            int nc = nt / 3;
            if (nX != 1 || nY != 1 || nc < 1 || nt != 3 * nc || p.n_xtra_i != nc * nc) { mess = "BaRatin_Apply:Fatal:Incorrect size [theta]"; return 1; }
            if (nc > 8) { mess = "bamgpu: BaRatin supports at most 8 controls"; return 1; }
            L.ncontrol = nc;
            auto M = [&](int r, int c) { return p.xtra_i[r + c * nc]; };
            // CheckControlMatrix, BaRatin_model.f90:254-322
            for (int r = 0; r < nc; ++r)
                for (int c = 0; c < nc; ++c) {
                    if (M(r, c) != 0 && M(r, c) != 1) { mess = "BaRatin_XtraRead:invalid control matrix"; return 1; }
                    if (c > r && M(r, c) > 0) { mess = "BaRatin_XtraRead:invalid control matrix"; return 1; }
                    if (M(r, c)) L.ctrlMask |= 1ull << (r * 8 + c);
                }
            for (int r = 1; r < nc; ++r) {
                int act = 0;
                for (int c = 0; c < nc; ++c) act += (M(r, c) - M(r - 1, c) > 0);
                if (act != 1) { mess = "BaRatin_XtraRead:invalid control matrix"; return 1; }
            }
            for (int r = 0; r < nc; ++r)
                if (M(r, r) != 1) { mess = "BaRatin_XtraRead:invalid control matrix"; return 1; }
            L.nDparModel = nc;
            if (L.model == BAMGPU_MDL_BARATINBAC) {  // BaRatinBAC_XtraRead (:150-235): hmax + the extra matrix constraints
                if (p.n_xtra_d != 1) { mess = "BaRatinBAC_XtraRead:problem reading file (hmax)"; return 1; }
                L.modelOptD = p.xtra_d[0];
                for (int i = 1; i < nc; ++i) {
                    int s0 = 0, s1 = 0;
                    for (int c = 0; c < nc; ++c) { s0 += M(i - 1, c); s1 += M(i, c); }
                    if (s1 < s0) { mess = "BaRatinBAC_XtraRead:FATAL: the number of active controls is not allowed to decrease when stage increases. Use original BaRatin for such configuration."; return 1; }
                    bool same = true;
                    for (int c = 0; c < i; ++c) same &= M(i, c) == M(i - 1, c);
                    if (same && M(i, i) == 1 && M(i - 1, i) == 0) continue;  // control addition
                    if (M(i - 1, i - 1) == 1 && M(i - 1, i) == 0 && M(i, i - 1) == 0 && M(i, i) == 1) {  // replacement
                        bool rest = true;
                        for (int c = 0; c + 1 < i; ++c) rest &= M(i, c) == M(i - 1, c);
                        if (!rest) { mess = "BaRatinBAC_XtraRead:FATAL: unsupported control matrix: unsupported control replacement."; return 2; }
                        continue;
                    }
                    mess = "BaRatinBAC_XtraRead:FATAL: unsupported control matrix.";
                    return 3;
                }
                L.nDparModel = 2 * nc;  // activation stages k and their type (ModelLibrary_tools.f90:593-602)
            }
            break;
        }
        case BAMGPU_MDL_SEDIMENT: {
            if (p.n_xtra_i != 7 || p.n_xtra_d != 4 || nY != 1) { mess = "Sediment_XtraRead:problem reading xtra"; return 1; }
            for (int i = 0; i < 7; ++i) L.sed[i] = p.xtra_i[i];
            for (int i = 0; i < 4; ++i) L.sedPseudo[i] = p.xtra_d[i];
            if (L.sed[0] < 1 || L.sed[0] > 4) { mess = "Sediment_Apply:Fatal:Unavailable [ID]"; return 1; }
            if (L.sed[1] < 1 || L.sed[1] > 3) { mess = "CSS_Apply:Fatal:Unavailable [ID]"; return 1; }
            if (L.sed[2] < 1 || L.sed[2] > 2) { mess = "GetAllpar:Fatal:Unavailable [CO]"; return 1; }
            int np = ((L.sed[0] == BAMGPU_SED_MPM || L.sed[0] == BAMGPU_SED_NIELSEN) ? 2 : 3) + (L.sed[1] == BAMGPU_CSS_CONSTANT ? 1 : 0);
            int nin = 2;
            for (int i = 0; i < 4; ++i) {
                if (L.sed[3 + i] < 1 || L.sed[3 + i] > 3) { mess = "GetAllpar:Fatal:Unavailable [PPO]"; return 1; }
                if (L.sed[3 + i] == BAMGPU_PPO_PAR) ++np;
                if (L.sed[3 + i] == BAMGPU_PPO_IN) ++nin;
            }
            if (L.sed[1] != BAMGPU_CSS_CONSTANT && L.sed[2] == BAMGPU_CO_PAR) np += (L.sed[1] == BAMGPU_CSS_SOULSBY ? 4 : 3);
            if (np != nt || nin != nX) { mess = "Sediment_Apply: wrong size [theta] or [IN]"; return 1; }
            break;
        }
        case BAMGPU_MDL_TEXTFILE: {
            if (!p.xtra_text) { mess = "TxtMdl_load:problem opening file"; return 1; }
            try {
                L.txt = parse_textfile_model(p.xtra_text);
            } catch (const std::exception& e) {
                mess = std::string("TxtMdl_define:") + e.what();
                return 5;
            }
            if ((int)L.txt.inNames.size() != nX || (int)L.txt.parNames.size() != nt || (int)L.txt.progs.size() != nY) {
                mess = "TxtMdl_Apply: size mismatch between the model text and nX/nPar/nY"; return 1;
            }
            for (auto& pr : L.txt.progs) {
                L.vmCodeOff.push_back((int)L.vmCode.size());
                L.vmImmedOff.push_back((int)L.vmImmed.size());
                L.vmNcode.push_back((int)pr.code.size());
                L.vmCode.insert(L.vmCode.end(), pr.code.begin(), pr.code.end());
                L.vmImmed.insert(L.vmImmed.end(), pr.immed.begin(), pr.immed.end());
                L.vmMaxStack = std::max(L.vmMaxStack, pr.stackSize);
            }
            break;
        }
        case BAMGPU_MDL_VEGETATION: {  // Vegetation_XtraRead (Vegetation_model.f90:244-289), GetParNumber (:46-86)
            if (p.n_xtra_i != 3 || p.n_xtra_d != 4) { mess = "Vegetation_XtraRead:problem reading xtra"; return 1; }
            L.vegGrowth = p.xtra_i[0]; L.vegNYear = p.xtra_i[1]; L.vegItmax = p.xtra_i[2];
            for (int i = 0; i < 4; ++i) L.vegOpt[i] = p.xtra_d[i];
            if (L.vegGrowth < 1 || L.vegGrowth > 4) { mess = "GetNpar_growth:Fatal:Unavailable [growthFormula]"; return 1; }
            const bool temporal = L.vegGrowth == BAMGPU_VEG_YIN1_TEMPORAL;
            if (L.vegNYear < 1 || L.vegNYear > 10) { mess = "bamgpu: Vegetation supports 1..10 years"; return 1; }
            const int np = 5 + 2 + (temporal ? 3 : 4) + (temporal ? 1 : L.vegNYear);
            if (nt != np) { mess = "Vegetation_Apply:wrong size[theta]"; return 1; }
            if (nX != (temporal ? 2 : 4) || (nY != 4 && nY != 5)) { mess = "Vegetation_Apply: wrong number of input or output columns"; return 1; }
            if (!temporal) {
                // the degree-day scan needs the whole series: a VAR parameter makes the reference call the model one
                // row at a time (:4152-4174), which ends in "empty year" (:432-436); latent / biased inputs would make
                // the series chain-dependent
                if (L.nVar) { mess = "Apply_growth:Fatal:empty year, currently unsupported (VAR parameters with a temperature-driven growth formula)"; return 1; }
                if (L.hasLatent || L.hasXbias) { mess = "bamgpu: Vegetation with temperature-driven growth does not support input errors"; return 1; }
                for (int y = 0; y < L.vegNYear && n > 0; ++y) {
                    bool any = false;
                    for (int i = 0; i < n && !any; ++i) any = std::floor(p.X[i]) == (double)y;
                    if (!any) { mess = "Apply_growth:Fatal:empty year, currently unsupported"; return 1; }
                }
            }
            L.nDparModel = temporal ? 1 : L.vegNYear + 1;
            L.nDer = temporal ? 4 : 2 + 6 * L.vegNYear;  // last slot: (B U_chi)^chi, the parameter-only factor of gamma2
            break;
        }
    }
    switch (L.model) {  // second-wave pointwise models (GetParNumber / XtraRead of each module)
        case BAMGPU_MDL_SUSPENDEDLOAD:
            if (nt != 5 || nX != 1 || nY != 1) { mess = "SuspendedLoad_Apply: wrong size [theta], nX or nY"; return 1; }
            break;
        case BAMGPU_MDL_SGD:
            if (nt != 5 || nX != 2 || nY != 1) { mess = "SGD_Apply: wrong size [theta], nX or nY"; return 1; }
            break;
        case BAMGPU_MDL_RECESSION_H:
            if (nt != 4 || nX != 1 || nY != 1) { mess = "Recession_h_Apply: wrong size [theta], nX or nY"; return 1; }
            break;
        case BAMGPU_MDL_SWOT:
            if (p.n_xtra_i != 1 || (p.xtra_i[0] != BAMGPU_SWOT_HS && p.xtra_i[0] != BAMGPU_SWOT_HSB)) { mess = "SWOT_Apply: Fatal: Unavailable [ID]"; return 1; }
            L.modelOpt[0] = p.xtra_i[0];
            if (nt != 5 || nX != (p.xtra_i[0] == BAMGPU_SWOT_HS ? 2 : 3) || nY != 1) { mess = "SWOT_Apply: wrong size [theta], nX or nY"; return 1; }
            break;
        case BAMGPU_MDL_ORTHO:
            if (p.n_xtra_i != 2 || (p.xtra_i[0] != BAMGPU_DISTO_NONE && p.xtra_i[0] != BAMGPU_DISTO_RADIAL) || p.xtra_i[1] < 0) { mess = "Disto_Apply:Fatal:Unavailable [Disto_ID]"; return 1; }
            L.modelOpt[0] = p.xtra_i[0]; L.modelOpt[1] = p.xtra_i[1];
            if (nt != 11 + p.xtra_i[1] || nX != 3 || nY != 2) { mess = "Ortho_Apply: wrong size [theta], nX or nY"; return 1; }
            L.nDparModel = 20;  // rotation matrix + ortho parameters (ModelLibrary_tools.f90:608-613)
            break;
        case BAMGPU_MDL_HCS: {  // HydraulicControl_section_XtraRead / _Setup (:114-260): h-A-w table, >= 3 rows
            const int np = p.n_xtra_d / 3;
            if (np < 3 || p.n_xtra_d != 3 * np) { mess = "HydraulicControl_section_Setup:invalid h-A-w matrix: at least 3 points are required"; return 1; }
            if (nt != 2 || nX != 1 || nY != 1) { mess = "HydraulicControl_section_Apply: wrong size [theta], nX or nY"; return 1; }
            L.hcsTab.assign((size_t)4 * np, 0.0);  // h | A | w | h + A/(2w)
            for (int i = 0; i < np; ++i) {
                const double hh = p.xtra_d[i], A = p.xtra_d[i + np], ww = p.xtra_d[i + 2 * np];
                L.hcsTab[i] = hh; L.hcsTab[i + np] = A; L.hcsTab[i + 2 * np] = ww;
                L.hcsTab[i + 3 * np] = hh + 0.5 * (i == 0 ? 0.0 : A / ww);
            }
            break;
        }
        case BAMGPU_MDL_SFD:  // SFD_XtraRead (StageFallDischarge_model.f90:191-247)
            if (p.n_xtra_i != 2 || p.n_xtra_d != 5 || p.xtra_i[0] != 1) { mess = "SFD_Apply: Fatal: Unavailable [ID]"; return 1; }
            if (nt != 8 || nX != 2 || nY != 1) { mess = "SFD_Apply: wrong size [theta], nX or nY"; return 1; }
            L.vegItmax = p.xtra_i[1];
            for (int i = 0; i < 4; ++i) L.vegOpt[i] = p.xtra_d[1 + i];  // xscale, fscale, tolX, tolF (shared Newton option slots)
            L.modelOptD = p.xtra_d[0];                                   // upper bound of the search interval
            L.nState = 1;
            break;
        case BAMGPU_MDL_SFDTIDAL_QMEC:   // SFDTidal_Qmec_GetParNumber (:31-59); nothing to read (ModelLibrary_tools.f90:506-509)
        case BAMGPU_MDL_SFDTIDAL_QMEC2:  // SFDTidal_Qmec2_GetParNumber: 11
            if (nt != (L.model == BAMGPU_MDL_SFDTIDAL_QMEC2 ? 11 : 12) || nX != 2 || nY != 1) { mess = "SFDTidal_Qmec_Apply: wrong size [theta], nX or nY"; return 1; }
            if (L.nVar || L.hasLatent || L.hasXbias) { mess = "bamgpu: SFDTidal_Qmec is a time-recursive model: VAR parameters and input errors are not supported"; return 1; }
            L.nState = 3;  // pressure, friction, advection (:672-676)
            L.isSeries = true;
            break;
        case BAMGPU_MDL_SFDTIDAL_QMEC0:   // SFDTidal_Qmec0_GetParNumber: 11; no state variables (ModelLibrary_tools.f90:668-671)
        case BAMGPU_MDL_SFDTIDAL_QMECMS:  // SFDTidal_QmecMS_GetParNumber: 10 (:60); no state variables (:682-685)
            if (nt != (L.model == BAMGPU_MDL_SFDTIDAL_QMEC0 ? 11 : 10) || nX != 2 || nY != 1) { mess = "SFDTidal_Qmec_Apply: wrong size [theta], nX or nY"; return 1; }
            if (L.nVar || L.hasLatent || L.hasXbias) { mess = "bamgpu: SFDTidal_Qmec is a time-recursive model: VAR parameters and input errors are not supported"; return 1; }
            L.isSeries = true;
            break;
        case BAMGPU_MDL_SFDTIDAL_LAG: {  // SFDTidal / SFDTidal2 / SFDTidal4 / SFDTidalJones: 8 parameters each, no xtra, no states (ModelLibrary_tools.f90:647-667)
            std::string id = p.model_id ? p.model_id : "";
            if (id.rfind("MDL_", 0) == 0) id = id.substr(4);
            const int variant = id == "SFDTidal" ? 1 : (id == "SFDTidal2" ? 2 : (id == "SFDTidal4" ? 3 : 4));
            if (nt != 8 || nX != (variant == 4 ? 3 : 2) || nY != 1) { mess = "SFDTidal_Apply: wrong size [theta], nX or nY"; return 1; }
            if (L.nVar || L.hasLatent || L.hasXbias) { mess = "bamgpu: SFDTidal reads lagged rows of the input series: VAR parameters and input errors are not supported"; return 1; }
            L.modelOpt[0] = variant;
            L.isSeries = true;
            break;
        }
        case BAMGPU_MDL_ALGAE:  // AlgaeBiomass_GetParNumber: 18 (AlgaeBiomass_model.f90:46-51); XtraSetup: 5 states (ModelLibrary_tools.f90:697-702)
            if (nt != 18 || nX != 5 || nY != 2) { mess = "AlgaeBiomass_Apply: wrong size [theta], nX or nY"; return 1; }
            if (L.nVar || L.hasLatent || L.hasXbias) { mess = "bamgpu: AlgaeBiomass is a time-recursive model: VAR parameters and input errors are not supported"; return 1; }
            if (const char* bad = bam::algae_input_error(p.nobs, p.X)) { mess = bad; return 1; }  // err = 1 in the reference (:79-99)
            L.nState = 5;
            L.isSeries = true;
            break;
        case BAMGPU_MDL_DYNVEG:  // DynamicVegetation_GetParNumber: 18 + 4 + 2 (DynamicVegetation_model.f90:27-30,68-71); XtraRead :193-247; 5 states (ModelLibrary_tools.f90:703-707)
            if (p.n_xtra_i != 1 || p.n_xtra_d != 4) { mess = "DynamicVegetation_XtraRead:problem reading xtra"; return 1; }
            if (nt != 24 || nX != 5 || nY != 3) { mess = "DynamicVegetation_Apply: wrong size [theta], nX or nY"; return 1; }
            if (L.nVar || L.hasLatent || L.hasXbias) { mess = "bamgpu: DynamicVegetation is a time-recursive model: VAR parameters and input errors are not supported"; return 1; }
            if (const char* bad = bam::algae_input_error(p.nobs, p.X, false)) { mess = bad; return 1; }
            L.vegItmax = p.xtra_i[0];
            for (int i = 0; i < 4; ++i) L.vegOpt[i] = p.xtra_d[i];  // xscale, fscale, tolX, tolF
            L.nState = 5;
            L.isSeries = true;
            break;
        case BAMGPU_MDL_MIXTURE: {  // Mixture_XtraRead (:137-190), Mixture_GetParNumber (:33-72)
            if (p.n_xtra_i != 4) { mess = "Mixture_XtraRead:problem reading xtra"; return 1; }
            const int nS = p.xtra_i[0], nEvt = p.xtra_i[1], nElt = p.xtra_i[2], miss = p.xtra_i[3] != 0;
            if (nS < 1 || nEvt < 1 || nElt < 1 || nt != (miss ? nS * nEvt + nElt : (nS - 1) * nEvt) || nX != 2 + nS || nY != 1) {
                mess = "Mixture_Apply: wrong size [theta], nX or nY"; return 1;
            }
            if (nEvt > 64) { mess = "bamgpu: Mixture supports at most 64 events"; return 1; }
            if (L.nVar) { mess = "bamgpu: Mixture is a whole-series model: VAR parameters are not supported"; return 1; }
            if (L.Xerror[0] || L.nXb[0]) { mess = "bamgpu: Mixture: the event column (first input) must be free of input errors"; return 1; }
            L.modelOpt[0] = nS; L.modelOpt[1] = nEvt; L.modelOpt[2] = nElt; L.modelOpt[3] = miss;
            L.nDparModel = nEvt;  // LastWeight1..nEvt (ModelLibrary_tools.f90:580-586)
            if (n > 0) {
                std::vector<int> src, elt;
                if (mixture_rows(n, p.X, nEvt, nElt, src, elt, mess)) return 1;
            }
            break;
        }
        case BAMGPU_MDL_SEGMENTATION: {  // Segmentation_XtraRead (:142-180)
            if (p.n_xtra_i != 3 || p.n_xtra_d != 1) { mess = "Segmentation_XtraRead:problem reading xtra"; return 1; }
            const int nS = p.xtra_i[0];
            if (nS < 1 || nt != 2 * nS - 1 || nX != 1 || nY != 1) { mess = "Segmentation_Apply: Fatal: incorrect size [theta]"; return 1; }
            if (p.xtra_i[2] != 1 && p.xtra_i[2] != 2) { mess = "Segmentation_Apply: Fatal: Unavailable [option]"; return 1; }
            if (L.nVar || L.hasLatent || L.hasXbias) { mess = "bamgpu: Segmentation is a whole-series model: VAR parameters and input errors are not supported"; return 1; }
            L.modelOpt[0] = nS; L.modelOpt[1] = p.xtra_i[1]; L.modelOpt[2] = p.xtra_i[2];
            L.modelOptD = p.xtra_d[0];
            break;
        }
        default: break;
    }
    if (L.model == BAMGPU_MDL_BARATIN) L.nDer = 2 * L.ncontrol;  // b_i and log a_i
    else if (L.model == BAMGPU_MDL_BARATINBAC) L.nDer = 2 * L.ncontrol;  // k_i and ktype_i
    else if (L.model == BAMGPU_MDL_SEGMENTATION) L.nDer = std::max(L.modelOpt[0] - 1, 0);  // the shift times tau
    else if (L.model == BAMGPU_MDL_ORTHO) L.nDer = 22;           // Dpar + the two focal lengths in pixels
    else if (L.isSeries) L.nDer = 1;                                     // the vector's own index into the series buffer
    else if (L.model != BAMGPU_MDL_VEGETATION) L.nDer = L.nDparModel;

    // VAR combinations in order of first appearance (:346-379); L(t,i) = offset_i + indx_i(t) (:326-334)
    L.comboOf.assign(n, 0);
    L.comboRow.clear();
    std::vector<int32_t> key(nt, 1);
    std::vector<std::vector<int32_t>> combos;
    if (L.nVar == 0 || n == 0) {
        combos.push_back(std::vector<int32_t>(nt, 1));
        L.comboRow.push_back(0);
    } else {
        for (int t = 0; t < n; ++t) {
            for (int i = 0; i < nt; ++i) {
                key[i] = (L.partype[i] == BAMGPU_PAR_VAR) ? p.var_indx[t + (size_t)i * n] : 1;
                if (key[i] < 1 || key[i] > L.nval[i]) { mess = "Config_Finalize: index of a VAR parameter out of [1,nval]"; return 1; }
            }
            int found = -1;
            for (size_t k = 0; k < combos.size() && found < 0; ++k)
                if (combos[k] == key) found = (int)k;
            if (found < 0) { found = (int)combos.size(); combos.push_back(key); L.comboRow.push_back(t); }
            L.comboOf[t] = found;
        }
    }
    L.nCombo = (int)combos.size();
    L.comboSrc.assign((size_t)L.nCombo * nt, -1);
    for (int k = 0; k < L.nCombo; ++k)
        for (int i = 0; i < nt; ++i)
            if (L.partype[i] != BAMGPU_PAR_FIX) L.comboSrc[(size_t)k * nt + i] = L.thetaOff[i] + combos[k][i] - 1;
    L.nDpar = L.nDparModel * L.nCombo;
    mess.clear();
    return 0;
}

int mixture_rows(int n, const double* evt, int nEvt, int nElt, std::vector<int>& src, std::vector<int>& elt, std::string& mess) {
    if ((long long)nEvt * nElt != n) { mess = "Mixture_Apply: the number of rows is not nEvt*nElt"; return 1; }
    src.assign(n, 0);
    elt.assign(n, 0);
    std::vector<int> cnt(nEvt, 0);
    for (int t = 0; t < n; ++t) {
        const long j = std::lround(evt[t]);  // nint(X(:,1))
        if (j < 1 || j > nEvt) continue;     // a row of no event is never packed
        const int e = cnt[j - 1]++;
        if (e < nElt) { src[(j - 1) * nElt + e] = t; elt[(j - 1) * nElt + e] = e + 1; }
    }
    for (int j = 0; j < nEvt; ++j)
        if (cnt[j] != nElt) { mess = "Mixture_Apply: an event does not have nElt rows"; return 1; }
    return 0;
}

// ------------------------------------------------------------------------------------------------
// small dense helpers
// ------------------------------------------------------------------------------------------------
int prior_corr_to_M(int k, const double* C, double* M, std::string& mess) {
    for (int i = 0; i < k; ++i)
        if (C[i + (size_t)i * k] != 1.0) { mess = "LoadBamObjects: diagonal terms of the prior correlation matrix should be equal to 1"; return 1; }
    // Cholesky C = L L^T (lower), then C^-1 = L^-T L^-1
    std::vector<double> Lm((size_t)k * k, 0.0), Li((size_t)k * k, 0.0);
    for (int j = 0; j < k; ++j) {
        double d = C[j + (size_t)j * k];
        for (int q = 0; q < j; ++q) d -= Lm[j + (size_t)q * k] * Lm[j + (size_t)q * k];
        if (!(d > 0.0)) { mess = "LoadBamObjects: prior correlation matrix is not positive definite"; return 1; }
        Lm[j + (size_t)j * k] = sqrt(d);
        for (int i = j + 1; i < k; ++i) {
            double s = C[i + (size_t)j * k];
            for (int q = 0; q < j; ++q) s -= Lm[i + (size_t)q * k] * Lm[j + (size_t)q * k];
            Lm[i + (size_t)j * k] = s / Lm[j + (size_t)j * k];
        }
    }
    for (int c = 0; c < k; ++c) {  // Li = L^-1 by forward substitution
        for (int r = c; r < k; ++r) {
            double s = (r == c) ? 1.0 : 0.0;
            for (int q = c; q < r; ++q) s -= Lm[r + (size_t)q * k] * Li[q + (size_t)c * k];
            Li[r + (size_t)c * k] = s / Lm[r + (size_t)r * k];
        }
    }
    for (int r = 0; r < k; ++r)
        for (int c = 0; c < k; ++c) {
            double s = 0.0;
            for (int q = std::max(r, c); q < k; ++q) s += Li[q + (size_t)r * k] * Li[q + (size_t)c * k];
            M[r + (size_t)c * k] = s - (r == c ? 1.0 : 0.0);
        }
    return 0;
}

void rhat_from_moments(int K, int P, const double* count, const double* mean, const double* m2, double* rhat) {
    // Gelman & Rubin (1992): W = mean within-chain variance, B/n = variance of chain means
    for (int j = 0; j < P; ++j) {
        double nbar = 0.0, gm = 0.0, W = 0.0;
        int used = 0;
        for (int c = 0; c < K; ++c) {
            if (count[c] < 2.0) continue;
            ++used;
            nbar += count[c];
            gm += mean[(size_t)c * P + j];
            W += m2[(size_t)c * P + j] / (count[c] - 1.0);
        }
        if (used < 2) { rhat[j] = NAN; continue; }
        nbar /= used; gm /= used; W /= used;
        double Bn = 0.0;
        for (int c = 0; c < K; ++c) {
            if (count[c] < 2.0) continue;
            double d = mean[(size_t)c * P + j] - gm;
            Bn += d * d;
        }
        Bn /= (used - 1);
        double varplus = (nbar - 1.0) / nbar * W + Bn;
        rhat[j] = (W > 0.0) ? sqrt(varplus / W) : NAN;
    }
}

}  // namespace bam
