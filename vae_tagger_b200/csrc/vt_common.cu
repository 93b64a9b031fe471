// Error string plumbing and the per-kernel-class profiler (launch counts + CUDA events).
#include <vector>

#include "vt_internal.h"

namespace vt {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }
const char* last_error() { return g_err.c_str(); }

struct Profiler {
    bool timing = false;
    double launches[KC_COUNT] = {0};
    double ms[KC_COUNT] = {0};
    double flops[KC_COUNT] = {0};
    double bytes[KC_COUNT] = {0};
    struct Pending {
        int kc;
        cudaEvent_t a, b;
    };
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t cur_start = nullptr;
    cudaEvent_t get() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

Profiler* profiler_create() { return new Profiler(); }
void profiler_destroy(Profiler* p) {
    if (!p) return;
    for (auto& q : p->pending) { cudaEventDestroy(q.a); cudaEventDestroy(q.b); }
    for (auto e : p->pool) cudaEventDestroy(e);
    delete p;
}
void profiler_enable(Profiler* p, bool timing) { if (p) p->timing = timing; }

void profiler_begin(Profiler* p, KernelClass kc, cudaStream_t s, double flops, double bytes) {
    if (!p) return;
    p->launches[kc] += 1;
    p->flops[kc] += flops;
    p->bytes[kc] += bytes;
    if (p->timing) {
        p->cur_start = p->get();
        cudaEventRecord(p->cur_start, s);
    }
}
void profiler_end(Profiler* p, KernelClass kc, cudaStream_t s) {
    if (!p || !p->timing || !p->cur_start) return;
    cudaEvent_t e = p->get();
    cudaEventRecord(e, s);
    p->pending.push_back({static_cast<int>(kc), p->cur_start, e});
    p->cur_start = nullptr;
}
// Caller must have synchronised the stream(s) the events were recorded on.
int profiler_read(Profiler* p, double* out, int reset) {
    if (!p) return -1;
    for (auto& q : p->pending) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, q.a, q.b) == cudaSuccess) p->ms[q.kc] += t;
        p->pool.push_back(q.a);
        p->pool.push_back(q.b);
    }
    p->pending.clear();
    for (int i = 0; i < KC_COUNT; ++i) {
        out[4 * i + 0] = p->launches[i];
        out[4 * i + 1] = p->ms[i];
        out[4 * i + 2] = p->flops[i];
        out[4 * i + 3] = p->bytes[i];
        if (reset) p->launches[i] = p->ms[i] = p->flops[i] = p->bytes[i] = 0;
    }
    return 0;
}

}  // namespace vt
