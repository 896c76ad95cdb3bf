#include <mutex>
#include <vector>
#include "prof.cuh"

namespace team {
struct ProfRec { cudaEvent_t a, b; int kind; double flops, bytes; };
static std::mutex g_mu;
static bool g_on = false;
static std::vector<ProfRec> g_recs;
static std::vector<cudaEvent_t> g_pool;

bool prof_enabled() { return g_on; }
static cudaEvent_t get_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}
int prof_begin(cudaStream_t st, int kind, double flops, double bytes) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_on) return -1;
    ProfRec r;
    r.a = get_event(); r.b = get_event(); r.kind = kind; r.flops = flops; r.bytes = bytes;
    cudaEventRecord(r.a, st);
    g_recs.push_back(r);
    return (int)g_recs.size() - 1;
}
void prof_end(cudaStream_t st, int slot) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (slot < 0 || slot >= (int)g_recs.size()) return;
    cudaEventRecord(g_recs[slot].b, st);
}
}  // namespace team

using namespace team;

extern "C" int team_prof_enable(int on) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_on = on != 0;
    return TEAM_OK;
}

// Sums the recorded launches of `kind` (-1 = all): total milliseconds, flops, bytes and launch count;
// clears the records.  Synchronises the device.
extern "C" int team_prof_collect(int kind, double* total_ms, double* total_flops, double* total_bytes, long long* launches) {
    TEAM_CUDA_CHECK(cudaDeviceSynchronize());
    std::lock_guard<std::mutex> lk(g_mu);
    double ms = 0, fl = 0, by = 0;
    long long n = 0;
    for (auto& r : g_recs) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess && (kind < 0 || kind == r.kind)) { ms += t; fl += r.flops; by += r.bytes; ++n; }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    g_recs.clear();
    if (total_ms) *total_ms = ms;
    if (total_flops) *total_flops = fl;
    if (total_bytes) *total_bytes = by;
    if (launches) *launches = n;
    return TEAM_OK;
}

// Per-launch records (in launch order) instead of sums: ms[i], flops[i], kind[i] for up to `cap` launches;
// returns the number written (or a negative error); clears the records.  Synchronises the device.
extern "C" long long team_prof_dump(double* ms, double* flops, int* kind, long long cap) {
    if (cudaDeviceSynchronize() != cudaSuccess) return TEAM_ECUDA;
    std::lock_guard<std::mutex> lk(g_mu);
    long long n = 0;
    for (auto& r : g_recs) {
        float t = 0.f;
        if (n < cap && cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) {
            if (ms) ms[n] = t;
            if (flops) flops[n] = r.flops;
            if (kind) kind[n] = r.kind;
            ++n;
        }
        g_pool.push_back(r.a);
        g_pool.push_back(r.b);
    }
    g_recs.clear();
    return n;
}
