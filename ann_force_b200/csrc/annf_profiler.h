// Per-kernel CUDA-event timing helper shared by the launchers (host only).
#pragma once

#include <cuda_runtime.h>

#include <map>
#include <string>
#include <utility>
#include <vector>

namespace annf {

// Per-kernel CUDA-event timing on the launching stream (roofline evidence for bench.py).
struct Profiler {
    struct Span { int name_id; cudaEvent_t start, stop; };
    bool enabled = false;
    std::vector<std::string> names;
    std::vector<Span> open;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pool;
    std::map<int, std::pair<long long, double>> totals;

    int name_id(const char* name) {
        for (size_t i = 0; i < names.size(); i++) if (names[i] == name) return (int) i;
        names.push_back(name);
        return (int) names.size() - 1;
    }
    void drain() {
        for (Span& s : open) {
            cudaEventSynchronize(s.stop);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, s.start, s.stop) == cudaSuccess) {
                auto& t = totals[s.name_id];
                t.first += 1;
                t.second += ms;
            }
            pool.emplace_back(s.start, s.stop);
        }
        open.clear();
    }
    void begin(const char* name, cudaStream_t stream) {
        if (!enabled) return;
        if (open.size() >= 8192) drain();
        Span s;
        s.name_id = name_id(name);
        if (!pool.empty()) {
            s.start = pool.back().first;
            s.stop = pool.back().second;
            pool.pop_back();
        } else {
            cudaEventCreate(&s.start);
            cudaEventCreate(&s.stop);
        }
        cudaEventRecord(s.start, stream);
        open.push_back(s);
    }
    void end(cudaStream_t stream) {
        if (!enabled) return;
        cudaEventRecord(open.back().stop, stream);
    }
    ~Profiler() {
        drain();
        for (auto& p : pool) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    }
};


}  // namespace annf
