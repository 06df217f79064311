// bam_baratin_fast.cuh -- BaRatin's pre-pass and rating curve with the table-driven log / exp (shared by the persistent
// warp sampler and the generic log-posterior kernel).
#pragma once
#include "bam_fastmath.cuh"

// BaRatin with the table-driven log / exp of bam_fastmath.cuh (the libdevice pow / log / exp of baratin_prepare and
// baratin_apply are ~3/4 of a proposal's instructions on the reference's own 38-gauging workspace).  Same operation
// order and the same exits as ApplyContinuity (:330-396) and BaRatin_computeQ (:194-252); x^c = texp(c tlog(x)) keeps
// pow()'s limits for x = 0 (0 or +inf by the sign of c).
__device__ inline bool baratin_prepare_t(const DevProblem& p, const double* th, double* b, unsigned lt, unsigned et) {
    const int nc = p.ncontrol;
    if (th[1] <= 0.0 || th[2] == 0.0) return false;
    b[0] = th[0];
    for (int j = 1; j < nc; ++j) {
        const double kj = th[3 * j], aj = th[3 * j + 1], cj = th[3 * j + 2];
        if (kj <= th[3 * (j - 1)]) return false;
        for (int i = 0; i < j; ++i)
            if (kj < b[i]) return false;
        if (aj <= 0.0 || cj == 0.0) return false;
        double som = 0.0;
        for (int i = 0; i < j; ++i) {
            const int d = (int)((p.ctrlMask >> ((j - 1) * 8 + i)) & 1ull) - (int)((p.ctrlMask >> (j * 8 + i)) & 1ull);
            som = som + (double)d * th[3 * i + 1] * texp(th[3 * i + 2] * tlog(kj - b[i], lt), et);
        }
        if (som < 0.0) return false;
        b[j] = kj - texp(tlog((1.0 / aj) * som, lt) * (1.0 / cj), et);
    }
    for (int i = 0; i < nc; ++i) b[nc + i] = tlog_pos(th[3 * i + 1], lt);
    return true;
}

__device__ __forceinline__ bool baratin_apply_t(const DevProblem& p, double H, const double* __restrict__ th,
                                                const double* __restrict__ b, double& Q, unsigned lt, unsigned et) {
    const int nc = p.ncontrol;
    int range = nc;
    for (int i = 0; i < nc; ++i)
        if (H < th[3 * i]) { range = i; break; }
    double q = 0.0;
    bool ok = true;
    if (range > 0) {
        const unsigned row = (unsigned)((p.ctrlMask >> ((range - 1) * 8)) & 0xffull);
        for (int i = 0; i < range; ++i) {
            const double d = H - b[i];
            if ((row >> i) & 1u) {
                if (d < 0.0) { q = 0.0; break; }
                q = q + texp(fma(th[3 * i + 2], tlog(d, lt), b[nc + i]), et);
            } else if (d < 0.0) { ok = false; break; }
        }
    }
    Q = q;
    return ok;
}

