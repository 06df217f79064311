// bam_config.cpp -- reader of the reference's configuration files (see bam_config.h).
#include "bam_config.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace bamhost {

namespace {

[[noreturn]] void fail(const std::string& m) { throw std::runtime_error(m); }

std::string fix_sep(std::string s) {
    std::replace(s.begin(), s.end(), '\\', '/');  // many shipped configs use Windows separators
    return s;
}

bool exists(const std::string& f) {
    std::ifstream is(f);
    return is.good();
}

std::vector<std::string> read_lines(const std::string& file) {
    std::ifstream is(file);
    if (!is) fail("problem opening file " + file);
    std::vector<std::string> out;
    std::string l;
    while (std::getline(is, l)) {
        if (!l.empty() && l.back() == '\r') l.pop_back();
        out.push_back(l);
    }
    return out;
}

double to_double(std::string s) {
    for (auto& c : s)
        if (c == 'd' || c == 'D') c = 'e';
    char* end = nullptr;
    double v = strtod(s.c_str(), &end);
    if (end == s.c_str()) fail("ReadError: cannot parse number '" + s + "'");
    return v;
}

bool to_bool(const std::string& s) {
    for (char c : s) {
        if (c == '.' || c == ' ') continue;
        return c == 't' || c == 'T';
    }
    return false;
}

struct Cursor {  // sequential reader over the lines of one config file
    std::string file;
    std::vector<std::string> lines;
    size_t k = 0;
    explicit Cursor(const std::string& f) : file(f), lines(read_lines(f)) {}
    std::vector<std::string> next() {
        if (k >= lines.size()) fail("ReadError in " + file);
        return list_items(lines[k++]);
    }
    std::string raw() {
        if (k >= lines.size()) fail("ReadError in " + file);
        return lines[k++];
    }
    bool more() const { return k < lines.size(); }
    std::string str() {
        auto it = next();
        if (it.empty()) fail("ReadError in " + file);
        return it[0];
    }
    double num() { return to_double(str()); }
    int integer() { return (int)std::lround(num()); }
    std::vector<double> nums(size_t n) {
        auto it = next();
        if (it.size() < n) fail("ReadError in " + file);
        std::vector<double> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = to_double(it[i]);
        return v;
    }
    std::vector<int> ints(size_t n) {
        auto d = nums(n);
        std::vector<int> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = (int)std::lround(d[i]);
        return v;
    }
    std::vector<std::string> strs(size_t n) {
        auto it = next();
        if (it.size() < n) fail("ReadError in " + file);
        it.resize(n);
        return it;
    }
    std::vector<int> bools(size_t n) {
        auto it = strs(n);
        std::vector<int> v(n);
        for (size_t i = 0; i < n; ++i) v[i] = to_bool(it[i]);
        return v;
    }
};

int dist_npar_checked(const std::string& d) {
    int id = bamgpu_dist_id(d.c_str());
    if (!id) fail("GetParNumber: unknown distribution '" + d + "'");
    return bamgpu_dist_npar(id);
}

// Config_Read_Par_STD (:3825-3860)
void read_par_std(Cursor& c, std::string& name, double& init, Prior& prior) {
    name = c.str();
    init = c.num();
    prior.dist = c.str();
    if (prior.dist == "FIX" || prior.dist == "VAR" || prior.dist == "STOK")
        fail("Config_Read_Par_STD: FIX/VAR/STOK not allowed here (BaM_Message 18)");
    prior.par = c.nums(dist_npar_checked(prior.dist));
}

// Config_Read_Par (:3751-3821)
Par read_par(Cursor& c) {
    Par p;
    p.name = c.str();
    double init = c.num();
    std::string dist = c.str();
    if (dist == "FIX") {
        p.partype = -1; p.nval = 1; p.init = {init};
        c.raw();
        p.prior = {Prior{dist, {}}};
    } else if (dist == "VAR") {
        p.partype = 1;
        p.config = c.str();
    } else if (dist == "STOK") {
        fail("Config_Read_Par:STOK parameters not implemented yet");
    } else {
        p.partype = 0; p.nval = 1; p.init = {init};
        Prior pr;
        pr.dist = dist;
        pr.par = c.nums(dist_npar_checked(dist));
        p.prior = {pr};
    }
    return p;
}

}  // namespace

std::vector<std::string> list_items(const std::string& line) {
    std::vector<std::string> out;
    size_t i = 0, n = line.size();
    while (i < n) {
        char ch = line[i];
        if (ch == ' ' || ch == '\t' || ch == ',' || ch == '\r') { ++i; continue; }
        if (ch == '!') break;
        if (ch == '"' || ch == '\'') {
            size_t j = line.find(ch, i + 1);
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

std::string resolve(const Workspace& w, const std::string& name0) {  // BaM_getDataFile (:516-541)
    std::string name = fix_sep(name0);
    if (exists(name)) return name;
    if (exists(w.dir + name)) return w.dir + name;
    // paths in the shipped examples are relative to the directory that holds the workspace
    size_t cut = w.dir.find_last_of('/', w.dir.size() > 1 ? w.dir.size() - 2 : 0);
    if (cut != std::string::npos && exists(w.dir.substr(0, cut + 1) + name)) return w.dir.substr(0, cut + 1) + name;
    size_t slash = name.find_last_of('/');
    if (slash != std::string::npos && exists(w.dir + name.substr(slash + 1))) return w.dir + name.substr(slash + 1);
    fail("data file not found, neither at " + name + " nor at " + w.dir + name);
}

std::vector<double> read_matrix(const std::string& file, int nrow, int ncol, int nskip) {
    std::ifstream is(file);
    if (!is) fail("problem opening file " + file);
    std::string l;
    for (int i = 0; i < nskip; ++i) std::getline(is, l);
    std::vector<double> M((size_t)nrow * ncol);
    for (int r = 0; r < nrow; ++r) {
        if (!std::getline(is, l)) fail("problem reading file " + file);
        int c = 0;
        const char* s = l.c_str();
        while (c < ncol) {
            while (*s == ' ' || *s == '\t' || *s == ',' || *s == ';' || *s == '\r') ++s;
            if (!*s) break;
            char* end = nullptr;
            double v = strtod(s, &end);
            if (end == s) {  // NA and friends
                while (*s && !strchr(" \t,;\r", *s)) ++s;
                v = NAN;
            } else s = end;
            M[r + (size_t)c * nrow] = v;
            ++c;
        }
        if (c < ncol) fail("problem reading file " + file + " (too few columns)");
    }
    return M;
}

// Xtra (ModelLibrary_tools.f90 XtraRead :436-544): fills w.xtraI / w.xtraD / w.xtraText from the model's Xtra file
static void read_xtra(Workspace& w) {
    const std::string xf = w.dir + fix_sep(w.fXtra);
    if (w.modelID == "BaRatin") {  // BaRatin_XtraRead: square 0/1 matrix, size = number of rows
        std::vector<std::vector<int>> rows;
        for (auto& l : read_lines(xf)) {
            auto it = list_items(l);
            if (it.empty()) continue;
            std::vector<int> r;
            for (auto& s : it) r.push_back((int)std::lround(to_double(s)));
            rows.push_back(r);
        }
        int nc = (int)rows.size();
        w.xtraI.assign((size_t)nc * nc, 0);
        for (int r = 0; r < nc; ++r) {
            if ((int)rows[r].size() < nc) fail("BaRatin_XtraRead:problem reading file " + xf);
            for (int cc = 0; cc < nc; ++cc) w.xtraI[r + (size_t)cc * nc] = rows[r][cc];
        }
    } else if (w.modelID == "BaRatinBAC") {  // BaRatinBAC_XtraRead (:150-180): n-1 matrix rows, then hmax
        std::vector<std::vector<double>> rows;
        for (auto& l : read_lines(xf)) {
            auto it = list_items(l);
            if (it.empty()) continue;
            std::vector<double> r;
            for (auto& s : it) r.push_back(to_double(s));
            rows.push_back(r);
        }
        if (rows.size() < 2) fail("BaRatinBAC_XtraRead:problem reading file " + xf);
        const int nc = (int)rows.size() - 1;
        w.xtraI.assign((size_t)nc * nc, 0);
        for (int r = 0; r < nc; ++r) {
            if ((int)rows[r].size() < nc) fail("BaRatinBAC_XtraRead:problem reading file " + xf);
            for (int cc = 0; cc < nc; ++cc) w.xtraI[r + (size_t)cc * nc] = (int)std::lround(rows[r][cc]);
        }
        w.xtraD = {rows[nc][0]};
    } else if (w.modelID == "Sediment") {  // Sediment_XtraRead (SedimentTransport_model.f90:193-245)
        Cursor s(xf);
        auto pick = [&](const std::string& v, std::initializer_list<const char*> names, const char* what) {
            int k = 1;
            for (auto nm : names) { if (v == nm) return k; ++k; }
            fail(std::string("Sediment: Fatal:Unavailable [") + what + "] '" + v + "'");
        };
        w.xtraI.push_back(pick(s.str(), {"MPM", "Camenen", "Nielsen", "Smart"}, "ID"));
        w.xtraI.push_back(pick(s.str(), {"Constant", "Soulsby", "SoulsbyWhitehouse"}, "CSS_ID"));
        w.xtraI.push_back(pick(s.str(), {"FIX", "PAR"}, "CO"));
        for (auto& o : s.strs(4)) w.xtraI.push_back(pick(o, {"FIX", "IN", "PAR"}, "PPO"));
        w.xtraD = s.nums(4);
    } else if (w.modelID == "TextFile") {
        std::ifstream is(xf);
        if (!is) fail("TxtMdl_load:problem opening file " + xf);
        std::stringstream ss;
        ss << is.rdbuf();
        w.xtraText = ss.str();
    } else if (w.modelID == "Vegetation") {  // Vegetation_XtraRead (Vegetation_model.f90:244-289)
        Cursor s(xf);
        const std::string gf = s.str();
        int gid = 0;
        const char* names[] = {"Growth_Yin1", "Growth_Yin2", "Growth_Yin3", "Growth_Yin1_temporal"};
        for (int k = 0; k < 4; ++k)
            if (gf == names[k]) gid = k + 1;
        if (!gid) fail("GetNpar_growth:Fatal:Unavailable [growthFormula] '" + gf + "'");
        w.xtraI.push_back(gid);
        w.xtraI.push_back(s.integer());                          // nYear
        for (int k = 0; k < 4; ++k) w.xtraD.push_back(s.num());  // Newton xscale, fscale, tolX, tolF (one per line)
        w.xtraI.push_back(s.integer());                          // Newton itmax
    } else if (w.modelID == "DynamicVegetation") {  // DynamicVegetation_XtraRead (DynamicVegetation_model.f90:193-247)
        Cursor s(xf);
        for (int k = 0; k < 4; ++k) w.xtraD.push_back(s.num());  // Newton xscale, fscale, xtol, ftol (one per line)
        w.xtraI.push_back(s.integer());                          // Newton maxiter
    } else if (w.modelID == "SFD") {  // SFD_XtraRead (StageFallDischarge_model.f90:191-247)
        Cursor s(xf);
        const std::string id = s.str();
        if (id != "SFD_Val_General") fail("SFD_Apply: Fatal: Unavailable [ID] '" + id + "'");
        for (int k = 0; k < 5; ++k) w.xtraD.push_back(s.num());  // upper bound, xscale, fscale, xtol, ftol
        w.xtraI = {1, s.integer()};                               // maxiter
    } else if (w.modelID == "SWOT") {  // SWOT_XtraRead (SWOT_model.f90:150-182): model variant
        Cursor s(xf);
        const std::string id = s.str();
        if (id == "SWOT_hS") w.xtraI.push_back(BAMGPU_SWOT_HS);
        else if (id == "SWOT_hSB") w.xtraI.push_back(BAMGPU_SWOT_HSB);
        else fail("SWOT_Apply: Fatal: Unavailable [ID] '" + id + "'");
    } else if (w.modelID == "Orthorectification") {  // Ortho_XtraRead (Orthorectification_model.f90:192-230)
        Cursor s(xf);
        const std::string id = s.str();
        if (id == "None") w.xtraI.push_back(BAMGPU_DISTO_NONE);
        else if (id == "Radial") w.xtraI.push_back(BAMGPU_DISTO_RADIAL);
        else fail("Disto_Apply:Fatal:Unavailable [Disto_ID] '" + id + "'");
        w.xtraI.push_back(s.integer());
    } else if (w.modelID == "HydraulicControl_section") {  // h-A-w matrix, one row per line (:114-150)
        std::vector<std::vector<double>> rows;
        for (auto& l : read_lines(xf)) {
            auto it = list_items(l);
            if (it.empty()) continue;
            if (it.size() < 3) fail("HydraulicControl_section_XtraRead:problem reading file " + xf);
            rows.push_back({to_double(it[0]), to_double(it[1]), to_double(it[2])});
        }
        const size_t np = rows.size();
        w.xtraD.assign(3 * np, 0.0);
        for (size_t i = 0; i < np; ++i)
            for (int c = 0; c < 3; ++c) w.xtraD[i + c * np] = rows[i][c];
    } else if (w.modelID == "Segmentation") {  // Segmentation_XtraRead (:142-180): nS, tmin, nmin, option
        Cursor s(xf);
        const int nS = s.integer();
        const double tmin = s.num();
        const int nmin = s.integer(), option = s.integer();
        w.xtraI = {nS, nmin, option};
        w.xtraD = {tmin};
    } else if (w.modelID == "Mixture") {  // Mixture_XtraRead (Mixture_model.f90:137-190): nS, nEvt, nElt, missingSource
        Cursor s(xf);
        const int nS = s.integer(), nEvt = s.integer(), nElt = s.integer();
        std::string flag = s.str();  // Fortran logical: T, F, .true., .false. (any case)
        size_t q = 0;
        while (q < flag.size() && flag[q] == '.') ++q;
        const char c0 = q < flag.size() ? (char)std::tolower((unsigned char)flag[q]) : ' ';
        if (c0 != 't' && c0 != 'f') fail("Mixture_XtraRead:problem reading file " + xf);
        w.xtraI = {nS, nEvt, nElt, c0 == 't' ? 1 : 0};
    } else if (w.modelID == "Linear" || w.modelID == "SuspendedLoad" || w.modelID == "SGD" || w.modelID == "Recession_h" ||
               w.modelID == "SFDTidal_Qmec" || w.modelID == "SFDTidal_Qmec2" || w.modelID == "SFDTidal_Qmec0" ||
               w.modelID == "SFDTidal_QmecMS" || w.modelID == "AlgaeBiomass" || w.modelID == "SFDTidal" || w.modelID == "SFDTidal2" ||
               w.modelID == "SFDTidal4" || w.modelID == "SFDTidalJones") {  // nothing to read (ModelLibrary_tools.f90:506-507)
    } else {
        fail("ApplyModel: Fatal: Unavailable [model%ID] '" + w.modelID + "' (not on the GPU hot path yet)");
    }

}

void load_workspace(const std::string& configFile, Workspace& w) {
    // Config_Read (:1963-2040)
    Cursor c(configFile);
    w.dir = fix_sep(c.str());
    if (!w.dir.empty() && w.dir.back() != '/') w.dir += '/';
    if (!exists(w.dir + ".") ) {
        // workspace given relative to the config file's directory (as the shipped tests/Config_BaM_*.txt)
        size_t cut = configFile.find_last_of('/');
        std::string base = cut == std::string::npos ? std::string("") : configFile.substr(0, cut + 1);
        const std::string rel = w.dir;
        w.dir = base + rel;
        if (!exists(w.dir + ".") && !base.empty()) {
            // a Config_BaM.txt stored INSIDE its workspace and written for a run from the parent directory
            // (tests/BaM_Segmentation/Config_BaM.txt names "BaM_Segmentation/")
            const size_t cut2 = base.find_last_of('/', base.size() - 2);
            w.dir = (cut2 == std::string::npos ? std::string("") : base.substr(0, cut2 + 1)) + rel;
        }
    }
    w.fRunOptions = c.str(); w.fModel = c.str(); w.fXtra = c.str(); w.fData = c.str();
    w.fRemnant = list_items(c.raw());
    w.fMCMC = c.str(); w.fCooking = c.str(); w.fSummary = c.str(); w.fResidual = c.str(); w.fPredMaster = c.str();

    // Config_Read_RunOptions (:2494-2552): optional file
    if (exists(w.dir + w.fRunOptions)) {
        Cursor r(w.dir + w.fRunOptions);
        w.doMCMC = to_bool(r.str()); w.doSummary = to_bool(r.str()); w.doResidual = to_bool(r.str()); w.doPred = to_bool(r.str());
    }

    // Config_Read_Model (:2096-2160)
    {
        Cursor m(w.dir + w.fModel);
        w.modelID = m.str();
        w.nX = m.integer(); w.nY = m.integer(); w.ntheta = m.integer();
        for (int i = 0; i < w.ntheta; ++i) w.theta.push_back(read_par(m));
    }
    if ((int)w.fRemnant.size() != w.nY) fail("Config_Read: number of remnant-sigma files /= nY");

    // Config_Read_Data (:2204-2270) + BaM_ReadData (:424-500)
    std::vector<double> W;
    int nhead, ncol;
    {
        Cursor d(w.dir + w.fData);
        std::string dataFile = resolve(w, d.str());
        nhead = d.integer(); w.nobs = d.integer(); ncol = d.integer();
        auto XCol = d.ints(w.nX), XuCol = d.ints(w.nX), XbCol = d.ints(w.nX), XbiCol = d.ints(w.nX);
        auto YCol = d.ints(w.nY), YuCol = d.ints(w.nY), YbCol = d.ints(w.nY), YbiCol = d.ints(w.nY);
        W = read_matrix(dataFile, w.nobs, ncol, nhead);
        const size_t n = (size_t)w.nobs;
        auto extract = [&](const std::vector<int>& cols, std::vector<double>& out) {  // Wextract (:4181)
            out.assign(n * cols.size(), 0.0);
            for (size_t j = 0; j < cols.size(); ++j) {
                if (cols[j] == 0) continue;
                if (cols[j] < 0 || cols[j] > ncol) fail("Config_Read_Data: column index out of range");
                for (size_t i = 0; i < n; ++i) out[i + j * n] = W[i + (size_t)(cols[j] - 1) * n];
            }
        };
        std::vector<double> tmp;
        extract(XCol, w.X); extract(XuCol, w.Xu); extract(XbCol, w.Xb); extract(YCol, w.Y); extract(YuCol, w.Yu);
        extract(YbCol, w.Yb);
        extract(XbiCol, tmp);
        w.Xbindx.resize(tmp.size());
        for (size_t q = 0; q < tmp.size(); ++q) w.Xbindx[q] = (int32_t)std::lround(tmp[q]);
        extract(YbiCol, tmp);
        w.Ybindx.resize(tmp.size());
        for (size_t q = 0; q < tmp.size(); ++q) w.Ybindx[q] = (int32_t)std::lround(tmp[q]);
    }

    // Config_Read_RemnantSigma (:3880-3946)
    for (int j = 0; j < w.nY; ++j) {
        Cursor r(w.dir + fix_sep(w.fRemnant[j]));
        RemnantSigma rs;
        rs.funk = r.str();
        int np = r.integer();
        int id = bamgpu_sigmafunc_id(rs.funk.c_str());
        if (!id) fail("Sigmafunk_GetParNumber: unknown function '" + rs.funk + "' (BaM_Message 2)");
        if (bamgpu_sigmafunc_npar(id) != np) fail("Config_Read_RemnantSigma: wrong number of parameters (BaM_Message 4)");
        for (int k = 0; k < np; ++k) {
            std::string name;
            double init;
            Prior pr;
            read_par_std(r, name, init, pr);
            rs.parname.push_back(name); rs.init.push_back(init); rs.prior.push_back(pr);
        }
        w.remnant.push_back(rs);
    }

    // Config_Finalize (:2833-2944): VAR parameters
    for (auto& p : w.theta) {
        if (p.partype != 1) { p.indx.assign(w.nobs, 1); continue; }
        Cursor v(w.dir + fix_sep(p.config));
        p.nval = v.integer();
        int col = v.integer();
        for (int j = 0; j < p.nval; ++j) {
            std::string name;
            double init;
            Prior pr;
            read_par_std(v, name, init, pr);
            p.init.push_back(init); p.prior.push_back(pr);
        }
        if (col < 1 || col > ncol) fail("Config_Finalize: VAR index column out of range");
        p.indx.resize(w.nobs);
        for (int t = 0; t < w.nobs; ++t) {
            p.indx[t] = (int)W[t + (size_t)(col - 1) * w.nobs];  // Fortran real -> integer assignment truncates
            if (p.indx[t] <= 0 || p.indx[t] > p.nval) fail("Config_Finalize: VAR index outside [1,nval] (BaM_Message 12)");
        }
    }

    read_xtra(w);

    // flatten for the C ABI
    for (auto& p : w.theta) {
        w.partype.push_back(p.partype);
        w.nval.push_back(p.nval);
        w.fixValue.push_back(p.partype == -1 ? p.init[0] : 0.0);
        if (p.partype == -1) continue;
        for (int k = 0; k < p.nval; ++k) {
            w.theta0.push_back(p.init[k]);
            w.priorThetaDist.push_back(bamgpu_dist_id(p.prior[k].dist.c_str()));
            for (int q = 0; q < BAMGPU_PRIOR_NPAR; ++q) w.priorThetaPar.push_back(q < (int)p.prior[k].par.size() ? p.prior[k].par[q] : 0.0);
        }
    }
    w.varIndx.assign((size_t)w.nobs * w.ntheta, 1);
    for (int i = 0; i < w.ntheta; ++i)
        for (int t = 0; t < w.nobs; ++t) w.varIndx[t + (size_t)i * w.nobs] = w.theta[i].indx[t];
    for (auto& r : w.remnant) {
        w.sigmaFunc.push_back(bamgpu_sigmafunc_id(r.funk.c_str()));
        w.nremnant.push_back((int)r.init.size());
        for (size_t k = 0; k < r.init.size(); ++k) {
            w.gamma0.push_back(r.init[k]);
            w.priorGammaDist.push_back(bamgpu_dist_id(r.prior[k].dist.c_str()));
            for (int q = 0; q < BAMGPU_PRIOR_NPAR; ++q) w.priorGammaPar.push_back(q < (int)r.prior[k].par.size() ? r.prior[k].par[q] : 0.0);
        }
    }

    // optional PriorCorrelation.txt in the workspace (src/BaM_main.f90:22,302; src/BaM_tools.f90:386-414)
    const std::string pc = w.dir + "PriorCorrelation.txt";
    if (exists(pc)) {
        int k = (int)(w.theta0.size() + w.gamma0.size());
        std::vector<double> C = read_matrix(pc, k, k, 0);
        w.priorM.assign((size_t)k * k, 0.0);
        char mess[BAMGPU_MESS_LEN];
        if (bamgpu_prior_corr_to_M(k, C.data(), w.priorM.data(), mess)) fail(std::string("LoadBamObjects:") + mess);
        w.hasPriorCorr = true;
    }

    // Config_Read_MCMC (:2391-2489)
    {
        Cursor m(w.dir + w.fMCMC);
        McmcConfig& mc = w.mcmc;
        mc.file = m.str(); mc.nAdapt = m.integer(); mc.nCycles = m.integer();
        mc.minMoveRate = m.num(); mc.maxMoveRate = m.num(); mc.downMult = m.num(); mc.upMult = m.num();
        int mode = m.integer();
        const double defaultStd = 0.1;  // src/BaM_main.f90:26
        if (mode == 0) {
            m.raw();
            double f = m.num();
            for (double t : w.theta0) mc.thetaStd0.push_back(f * std::fabs(t) == 0.0 ? defaultStd : f * std::fabs(t));
            for (auto& r : w.remnant) {
                std::vector<double> g;
                for (double t : r.init) g.push_back(f * std::fabs(t) == 0.0 ? defaultStd : f * std::fabs(t));
                mc.gammaStd0.push_back(g);
            }
        } else {
            m.raw(); m.raw();
            mc.thetaStd0 = m.nums(w.theta0.size());
            for (auto& r : w.remnant) mc.gammaStd0.push_back(m.nums(r.init.size()));
        }
    }
    // Config_Read_Cook (:2599-2643), _Summary (:2647-2690), _Residual (:2556-2595)
    if (exists(w.dir + w.fCooking)) {
        Cursor k(w.dir + w.fCooking);
        w.cookFile = k.str(); w.burnFactor = k.num(); w.nSlim = k.integer();
    }
    if (exists(w.dir + w.fSummary)) {
        Cursor s(w.dir + w.fSummary);
        w.summaryFile = s.str();
        if (s.more()) { auto it = s.next(); if (!it.empty()) w.dicFile = it[0]; }          // absent -> no DIC (back-compatibility)
        if (s.more()) { auto it = s.next(); if (!it.empty()) w.xtendedMcmcFile = it[0]; }  // absent -> no extended MCMC file
    }
    if (exists(w.dir + w.fResidual)) { Cursor s(w.dir + w.fResidual); w.residualFile = s.str(); }
    // Config_Read_Pred_Master (:2694-2739)
    if (w.doPred && exists(w.dir + w.fPredMaster)) {
        Cursor pm(w.dir + w.fPredMaster);
        int np = pm.integer();
        for (int i = 0; i < np; ++i) w.predConfigs.push_back(pm.str());
    }
}

// Config_Read_oneRun (:2042-2092) + Config_Read_Inputs (:2294-2330) + BaM_ReadInputs (:1196-1240): workspace, model
// config, xtra config, inputs config {data file, header lines, nobs, ncol}.  The model is run ONCE at the initial
// values of Config_Model (src/BaM_main.f90:215-247); there is no calibration data and no remnant model.
void load_onerun(const std::string& configFile, Workspace& w, std::vector<double>& Xin, int& nIn) {
    Cursor c(configFile);
    w.dir = fix_sep(c.str());
    if (!w.dir.empty() && w.dir.back() != '/') w.dir += '/';
    if (!exists(w.dir + ".")) {
        size_t cut = configFile.find_last_of('/');
        w.dir = (cut == std::string::npos ? std::string("") : configFile.substr(0, cut + 1)) + w.dir;
    }
    w.fModel = c.str(); w.fXtra = c.str();
    const std::string fInputs = c.str();
    {
        Cursor m(w.dir + w.fModel);
        w.modelID = m.str();
        w.nX = m.integer(); w.nY = m.integer(); w.ntheta = m.integer();
        for (int i = 0; i < w.ntheta; ++i) w.theta.push_back(read_par(m));
    }
    read_xtra(w);
    Cursor d(w.dir + fix_sep(fInputs));
    const std::string dataFile = resolve(w, d.str());
    const int nhead = d.integer();
    nIn = d.integer();
    const int ncol = d.integer();
    std::vector<double> W = read_matrix(dataFile, nIn, ncol, nhead);
    if (ncol < w.nX) fail("BaM_ReadInputs: fewer columns than model inputs in " + dataFile);
    Xin.assign(W.begin(), W.begin() + (size_t)nIn * w.nX);  // the first nX columns (column-major)
    // every parameter takes its first initial value and is handed to the library as FIX: nothing is inferred
    w.nobs = 0;
    for (auto& pr : w.theta) {
        w.partype.push_back(-1); w.nval.push_back(1);
        w.fixValue.push_back(pr.init.empty() ? 0.0 : pr.init[0]);
    }
    for (int j = 0; j < w.nY; ++j) {  // placeholder remnant model (never used: no noise, no likelihood)
        w.sigmaFunc.push_back(BAMGPU_SIG_CONSTANT); w.nremnant.push_back(1);
        w.priorGammaDist.push_back(BAMGPU_DIST_FLATPRIOR);
        for (int q = 0; q < BAMGPU_PRIOR_NPAR; ++q) w.priorGammaPar.push_back(0.0);
        w.gamma0.push_back(1.0);
    }
}

PredConfig read_pred_config(const Workspace& w, const std::string& file, int nState) {
    Cursor c(w.dir + fix_sep(file));
    PredConfig p;
    p.xspagFiles = c.strs(w.nX);
    p.nobs = c.integer();
    p.nspag = c.ints(w.nX);
    p.doParametric = to_bool(c.str());
    p.doRemnant = c.bools(w.nY);
    p.nsim = c.integer();
    p.yspagFiles = c.strs(w.nY);
    p.doTranspose = c.bools(w.nY);
    p.doEnvelop = c.bools(w.nY);
    p.envFiles = c.strs(w.nY);
    p.printCounter = to_bool(c.str());
    p.doState.assign(nState, 0);
    if (nState > 0 && c.more()) {  // :2808-2826 (a file without the section: BaM_ConsoleMessage 19, states skipped)
        try {
            p.doState = c.bools(nState);
        } catch (const std::exception&) {
            p.doState.assign(nState, 0);
        }
        if (std::any_of(p.doState.begin(), p.doState.end(), [](int v) { return v != 0; })) {
            p.sspagFiles = c.strs(nState);
            p.doTransposeS = c.bools(nState);
            p.doEnvelopS = c.bools(nState);
            p.envFilesS = c.strs(nState);
        }
    }
    return p;
}

bamgpu_problem Workspace::problem() const {
    bamgpu_problem p;
    memset(&p, 0, sizeof p);
    p.model_id = modelID.c_str();
    p.nobs = nobs; p.nX = nX; p.nY = nY; p.ntheta = ntheta;
    p.X = X.data(); p.Xu = Xu.data(); p.Xb = Xb.data(); p.Xbindx = Xbindx.data();
    p.Y = Y.data(); p.Yu = Yu.data(); p.Yb = Yb.data(); p.Ybindx = Ybindx.data();
    p.partype = partype.data(); p.nval = nval.data(); p.fix_value = fixValue.data(); p.var_indx = varIndx.data();
    p.prior_theta_dist = priorThetaDist.data(); p.prior_theta_par = priorThetaPar.data();
    p.sigma_func = sigmaFunc.data(); p.nremnant = nremnant.data();
    p.prior_gamma_dist = priorGammaDist.data(); p.prior_gamma_par = priorGammaPar.data();
    p.priorM = hasPriorCorr ? priorM.data() : nullptr;
    p.xtra_i = xtraI.data(); p.n_xtra_i = (int)xtraI.size();
    p.xtra_d = xtraD.data(); p.n_xtra_d = (int)xtraD.size();
    p.xtra_text = xtraText.empty() ? nullptr : xtraText.c_str();
    p.mv = -9999.0; p.unfeas_flag = -666.666;
    return p;
}

static std::vector<double> pack(const Workspace& w, const std::vector<double>& th, const std::vector<double>& gm,
                                const std::vector<double>& Xsrc, double bias) {  // Packer (:4202-4242)
    std::vector<double> out(th);
    out.insert(out.end(), gm.begin(), gm.end());
    const size_t n = (size_t)w.nobs;
    for (int j = 0; j < w.nX; ++j)
        for (size_t i = 0; i < n; ++i)
            if (w.Xu[i + j * n] > 0.0) out.push_back(Xsrc[i + j * n]);
    for (int j = 0; j < w.nX; ++j) {
        int mx = 0;
        for (size_t i = 0; i < n; ++i) mx = std::max(mx, (int)w.Xbindx[i + j * n]);
        out.insert(out.end(), mx, bias);
    }
    for (int j = 0; j < w.nY; ++j) {
        int mx = 0;
        for (size_t i = 0; i < n; ++i) mx = std::max(mx, (int)w.Ybindx[i + j * n]);
        out.insert(out.end(), mx, bias);
    }
    return out;
}

std::vector<double> Workspace::start_vector() const { return pack(*this, theta0, gamma0, X, 0.0); }

std::vector<double> Workspace::start_std_vector() const {
    std::vector<double> g;
    for (auto& v : mcmc.gammaStd0) g.insert(g.end(), v.begin(), v.end());
    return pack(*this, mcmc.thetaStd0, g, Xu, 0.1);  // latent-X jump std = Xu, bias std = 0.1 (:119,606)
}

void Workspace::make_headers(const bamgpu_sizes& s, int nDparModel) {  // GetHeaders (:3950-4004)
    headers.clear();
    for (auto& p : theta) {
        if (p.partype == 0) headers.push_back(p.name);
        else if (p.partype == 1)
            for (int j = 1; j <= p.nval; ++j) headers.push_back(p.name + "_" + std::to_string(j));
    }
    for (int i = 0; i < nY; ++i)
        for (auto& nm : remnant[i].parname) headers.push_back("Y" + std::to_string(i + 1) + "_" + nm);
    const size_t n = (size_t)nobs;
    for (int j = 0; j < nX; ++j)
        for (size_t i = 0; i < n; ++i)
            if (Xu[i + j * n] > 0.0) headers.push_back("X" + std::to_string(j + 1) + "_TrueVal" + std::to_string(i + 1));
    for (int j = 0; j < nX; ++j) {
        int mx = 0;
        for (size_t i = 0; i < n; ++i) mx = std::max(mx, (int)Xbindx[i + j * n]);
        for (int i = 1; i <= mx; ++i) headers.push_back("X" + std::to_string(j + 1) + "_Bias" + std::to_string(i));
    }
    for (int j = 0; j < nY; ++j) {
        int mx = 0;
        for (size_t i = 0; i < n; ++i) mx = std::max(mx, (int)Ybindx[i + j * n]);
        for (int i = 1; i <= mx; ++i) headers.push_back("Y" + std::to_string(j + 1) + "_Bias" + std::to_string(i));
    }
    headers.push_back("LogPost");
    // derived parameters: model names (XtraSetup), suffixed by the combination index when VAR (:368-377)
    std::vector<std::string> dn;
    if (modelID == "BaRatin")
        for (int i = 1; i <= nDparModel; ++i) dn.push_back("b" + std::to_string(i));
    if (modelID == "BaRatinBAC") {  // ModelLibrary_tools.f90:593-602
        for (int i = 1; i <= nDparModel / 2; ++i) dn.push_back("k" + std::to_string(i));
        for (int i = 1; i <= nDparModel / 2; ++i) dn.push_back("ktype" + std::to_string(i));
    }
    if (modelID == "Orthorectification")  // ModelLibrary_tools.f90:608-613
        for (auto nm : {"r11", "r12", "r13", "r21", "r22", "r23", "r31", "r32", "r33", "a1", "a2", "a3", "a4", "b1", "b2",
                        "b3", "b4", "c1", "c2", "c3"})
            dn.push_back(nm);
    if (modelID == "Vegetation") {  // ModelLibrary_tools.f90:625-638
        dn.push_back("a1");
        for (int i = 2; i <= nDparModel; ++i) dn.push_back("Topt_" + std::to_string(i));
    }
    const bool hasVar = std::any_of(partype.begin(), partype.end(), [](int t) { return t == 1; });
    if (!hasVar) headers.insert(headers.end(), dn.begin(), dn.end());
    else
        for (int k = 1; k <= s.nCombo; ++k)
            for (auto& d : dn) headers.push_back(d + "_" + std::to_string(k));
}

std::string dump_problem_json(const Workspace& w) {
    std::ostringstream o;
    o.precision(17);
    auto arr = [&](const char* key, const auto& v) {
        o << "\"" << key << "\":[";
        for (size_t i = 0; i < v.size(); ++i) o << (i ? "," : "") << v[i];
        o << "],";
    };
    o << "{\"model_id\":\"" << w.modelID << "\",\"nobs\":" << w.nobs << ",\"nX\":" << w.nX << ",\"nY\":" << w.nY << ",";
    arr("X", w.X); arr("Xu", w.Xu); arr("Xb", w.Xb); arr("Xbindx", w.Xbindx);
    arr("Y", w.Y); arr("Yu", w.Yu); arr("Yb", w.Yb); arr("Ybindx", w.Ybindx);
    arr("partype", w.partype); arr("nval", w.nval); arr("fix_value", w.fixValue); arr("var_indx", w.varIndx);
    arr("prior_theta_dist", w.priorThetaDist); arr("prior_theta_par", w.priorThetaPar);
    arr("sigma_func", w.sigmaFunc); arr("prior_gamma_dist", w.priorGammaDist); arr("prior_gamma_par", w.priorGammaPar);
    arr("xtra_i", w.xtraI); arr("xtra_d", w.xtraD); arr("priorM", w.priorM);
    arr("start", w.start_vector()); arr("start_std", w.start_std_vector());
    o << "\"nAdapt\":" << w.mcmc.nAdapt << ",\"nCycles\":" << w.mcmc.nCycles << "}";
    return o.str();
}

}  // namespace bamhost
