// mt64.cpp -- the generator SURVEY.md 8(d) names for every synthetic workload: std::mt19937_64 (seed 20261019 by default)
// with libstdc++'s uniform_real_distribution / normal_distribution, behind a tiny C ABI for workloads/synth.py.
// Inputs only: nothing here is on a compute path.
#include <cstddef>
#include <cstdint>
#include <random>

extern "C" {
void* mt64_new(uint64_t seed) { return new std::mt19937_64(seed); }
void mt64_free(void* g) { delete static_cast<std::mt19937_64*>(g); }
void mt64_uniform(void* g, double lo, double hi, double* out, size_t n) {
    std::uniform_real_distribution<double> d(lo, hi);
    auto& e = *static_cast<std::mt19937_64*>(g);
    for (size_t i = 0; i < n; ++i) out[i] = d(e);
}
void mt64_normal(void* g, double mu, double sd, double* out, size_t n) {
    std::normal_distribution<double> d(mu, sd);
    auto& e = *static_cast<std::mt19937_64*>(g);
    for (size_t i = 0; i < n; ++i) out[i] = d(e);
}
void mt64_raw(void* g, uint64_t* out, size_t n) {
    auto& e = *static_cast<std::mt19937_64*>(g);
    for (size_t i = 0; i < n; ++i) out[i] = e();
}
}
