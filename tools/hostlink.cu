// hostlink.cu -- what the host link of this box can move, on 1, 2, 4 and 8 GPUs at once ("next" item 1).
//
//   nvcc -O2 -o tools/hostlink tools/hostlink.cu          (built by tools/build_tools.py; the binary travels to the GPU box)
//   tools/hostlink [MiB per copy, default 256] [reps, default 4] [split]  >  profiles/r02_hostlink.json     (split: only the split-direction tests)
//
// One process, one host thread; per device two streams (H2D, D2H) and two pinned host buffers.  Every test enqueues
// `reps` copies per device and direction, then waits for all of them: aggregate GB/s per direction = bytes / wall time
// (std::chrono from the first enqueue to the last synchronize); per-device figures come from CUDA events on the
// device's own streams.  Tests: the first n devices (n = 1, 2, 4, 8) for H2D only / D2H only / both; every device alone;
// every pair (both directions) -- pairs that share an upstream link show up as pairs whose sum is not 2x a single
// device; the two halves {0..3} and {4..7}; pinned memory allocated from a thread bound to each NUMA node in turn (if the
// box shows more than one) and write-combined H2D source buffers.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <string>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct Dev {
    int id; char* d_in; char* d_out; char* h_in; char* h_out; cudaStream_t s_in, s_out; cudaEvent_t a_in, b_in, a_out, b_out;
};

static size_t g_bytes; static int g_reps;

static std::string read_file(const std::string& p) {
    FILE* f = fopen(p.c_str(), "r"); if (!f) return "";
    char buf[4096]; size_t n = fread(buf, 1, sizeof buf - 1, f); fclose(f); buf[n] = 0;
    while (n && (buf[n - 1] == '\n' || buf[n - 1] == ' ')) buf[--n] = 0;
    return buf;
}

struct Result { double wall_s, h2d_gbs, d2h_gbs; std::vector<double> dev_h2d, dev_d2h; };

static Result run(std::vector<Dev>& devs, const std::vector<int>& set, bool h2d, bool d2h) {
    for (int i : set) { CK(cudaSetDevice(devs[i].id)); CK(cudaDeviceSynchronize()); }
    auto t0 = std::chrono::steady_clock::now();
    for (int i : set) {
        Dev& d = devs[i]; CK(cudaSetDevice(d.id));
        if (h2d) CK(cudaEventRecord(d.a_in, d.s_in));
        if (d2h) CK(cudaEventRecord(d.a_out, d.s_out));
    }
    for (int r = 0; r < g_reps; r++)
        for (int i : set) {
            Dev& d = devs[i]; CK(cudaSetDevice(d.id));
            if (h2d) CK(cudaMemcpyAsync(d.d_in, d.h_in, g_bytes, cudaMemcpyHostToDevice, d.s_in));
            if (d2h) CK(cudaMemcpyAsync(d.h_out, d.d_out, g_bytes, cudaMemcpyDeviceToHost, d.s_out));
        }
    for (int i : set) {
        Dev& d = devs[i]; CK(cudaSetDevice(d.id));
        if (h2d) CK(cudaEventRecord(d.b_in, d.s_in));
        if (d2h) CK(cudaEventRecord(d.b_out, d.s_out));
    }
    for (int i : set) { Dev& d = devs[i]; CK(cudaSetDevice(d.id)); CK(cudaStreamSynchronize(d.s_in)); CK(cudaStreamSynchronize(d.s_out)); }
    Result res; res.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double tot = (double)g_bytes * g_reps * set.size() / 1e9;
    res.h2d_gbs = h2d ? tot / res.wall_s : 0; res.d2h_gbs = d2h ? tot / res.wall_s : 0;
    for (int i : set) {
        Dev& d = devs[i]; float ms = 0;
        if (h2d) { CK(cudaEventElapsedTime(&ms, d.a_in, d.b_in)); res.dev_h2d.push_back((double)g_bytes * g_reps / 1e6 / ms); }
        if (d2h) { CK(cudaEventElapsedTime(&ms, d.a_out, d.b_out)); res.dev_d2h.push_back((double)g_bytes * g_reps / 1e6 / ms); }
    }
    return res;
}

// Directions split over two device sets: set A only copies host-to-device, set B only device-to-host, at the same time.
static Result run_split(std::vector<Dev>& devs, const std::vector<int>& seth, const std::vector<int>& setd) {
    std::vector<int> all = seth; all.insert(all.end(), setd.begin(), setd.end());
    for (int i : all) { CK(cudaSetDevice(devs[i].id)); CK(cudaDeviceSynchronize()); }
    auto t0 = std::chrono::steady_clock::now();
    for (int i : seth) { Dev& d = devs[i]; CK(cudaSetDevice(d.id)); CK(cudaEventRecord(d.a_in, d.s_in)); }
    for (int i : setd) { Dev& d = devs[i]; CK(cudaSetDevice(d.id)); CK(cudaEventRecord(d.a_out, d.s_out)); }
    const size_t nmax = seth.size() > setd.size() ? seth.size() : setd.size();
    for (int r = 0; r < g_reps; r++)
        for (size_t k = 0; k < nmax; k++) {
            if (k < seth.size()) { Dev& d = devs[seth[k]]; CK(cudaSetDevice(d.id)); CK(cudaMemcpyAsync(d.d_in, d.h_in, g_bytes, cudaMemcpyHostToDevice, d.s_in)); }
            if (k < setd.size()) { Dev& d = devs[setd[k]]; CK(cudaSetDevice(d.id)); CK(cudaMemcpyAsync(d.h_out, d.d_out, g_bytes, cudaMemcpyDeviceToHost, d.s_out)); }
        }
    for (int i : seth) { Dev& d = devs[i]; CK(cudaSetDevice(d.id)); CK(cudaEventRecord(d.b_in, d.s_in)); }
    for (int i : setd) { Dev& d = devs[i]; CK(cudaSetDevice(d.id)); CK(cudaEventRecord(d.b_out, d.s_out)); }
    for (int i : all) { Dev& d = devs[i]; CK(cudaSetDevice(d.id)); CK(cudaStreamSynchronize(d.s_in)); CK(cudaStreamSynchronize(d.s_out)); }
    Result res; res.wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    res.h2d_gbs = (double)g_bytes * g_reps * seth.size() / 1e9 / res.wall_s; res.d2h_gbs = (double)g_bytes * g_reps * setd.size() / 1e9 / res.wall_s;
    for (int i : seth) { float ms = 0; CK(cudaEventElapsedTime(&ms, devs[i].a_in, devs[i].b_in)); res.dev_h2d.push_back((double)g_bytes * g_reps / 1e6 / ms); }
    for (int i : setd) { float ms = 0; CK(cudaEventElapsedTime(&ms, devs[i].a_out, devs[i].b_out)); res.dev_d2h.push_back((double)g_bytes * g_reps / 1e6 / ms); }
    return res;
}

static bool g_first = true;
static void emit(const char* name, const std::vector<int>& set, const char* mode, const Result& r) {
    printf("%s\n  {\"test\": \"%s\", \"devices\": [", g_first ? "" : ",", name); g_first = false;
    for (size_t i = 0; i < set.size(); i++) printf("%s%d", i ? ", " : "", set[i]);
    printf("], \"mode\": \"%s\", \"h2d_gbs\": %.2f, \"d2h_gbs\": %.2f, \"per_device_h2d\": [", mode, r.h2d_gbs, r.d2h_gbs);
    for (size_t i = 0; i < r.dev_h2d.size(); i++) printf("%s%.1f", i ? ", " : "", r.dev_h2d[i]);
    printf("], \"per_device_d2h\": [");
    for (size_t i = 0; i < r.dev_d2h.size(); i++) printf("%s%.1f", i ? ", " : "", r.dev_d2h[i]);
    printf("]}");
    fflush(stdout);
}

static void three_modes(std::vector<Dev>& devs, const char* name, const std::vector<int>& set) {
    emit(name, set, "h2d", run(devs, set, true, false));
    emit(name, set, "d2h", run(devs, set, false, true));
    emit(name, set, "both", run(devs, set, true, true));
}

static void alloc_host(std::vector<Dev>& devs, unsigned flags_in) {
    for (auto& d : devs) {
        CK(cudaSetDevice(d.id));
        if (d.h_in) CK(cudaFreeHost(d.h_in));
        if (d.h_out) CK(cudaFreeHost(d.h_out));
        CK(cudaHostAlloc((void**)&d.h_in, g_bytes, flags_in)); CK(cudaHostAlloc((void**)&d.h_out, g_bytes, cudaHostAllocPortable));
        memset(d.h_in, 1, g_bytes); memset(d.h_out, 0, g_bytes);
    }
}

int main(int argc, char** argv) {
    g_bytes = (size_t)(argc > 1 ? atoi(argv[1]) : 256) << 20; g_reps = argc > 2 ? atoi(argv[2]) : 4;
    int n = 0; CK(cudaGetDeviceCount(&n));
    std::vector<Dev> devs(n);
    printf("{\"bytes_per_copy\": %zu, \"reps\": %d, \"n_devices\": %d,\n \"devices\": [", g_bytes, g_reps, n);
    for (int i = 0; i < n; i++) {
        Dev& d = devs[i]; memset(&d, 0, sizeof d); d.id = i; CK(cudaSetDevice(i));
        char bus[32]; CK(cudaDeviceGetPCIBusId(bus, sizeof bus, i));
        std::string lower = bus; for (auto& c : lower) c = (char)tolower(c);
        std::string numa = read_file("/sys/bus/pci/devices/" + lower + "/numa_node");
        cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, i));
        printf("%s{\"index\": %d, \"pci\": \"%s\", \"numa_node\": \"%s\", \"name\": \"%s\", \"async_engines\": %d}", i ? ", " : "", i, bus, numa.c_str(), p.name, p.asyncEngineCount);
        CK(cudaMalloc((void**)&d.d_in, g_bytes)); CK(cudaMalloc((void**)&d.d_out, g_bytes)); CK(cudaMemset(d.d_out, 3, g_bytes));
        CK(cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking));
        CK(cudaEventCreate(&d.a_in)); CK(cudaEventCreate(&d.b_in)); CK(cudaEventCreate(&d.a_out)); CK(cudaEventCreate(&d.b_out));
    }
    int nodes = 0; std::string nodelist = "[";
    for (int k = 0; k < 16; k++) {
        std::string cl = read_file("/sys/devices/system/node/node" + std::to_string(k) + "/cpulist");
        if (cl.empty()) break;
        nodelist += std::string(k ? ", " : "") + "\"" + cl + "\""; nodes++;
    }
    printf("],\n \"host\": {\"online_cpus\": \"%s\", \"numa_nodes\": %d, \"node_cpulists\": %s]},\n \"results\": [", read_file("/sys/devices/system/cpu/online").c_str(), nodes, nodelist.c_str());
    alloc_host(devs, cudaHostAllocPortable);
    std::vector<int> all; for (int i = 0; i < n; i++) all.push_back(i);
    run(devs, all, true, true);                                               // warm-up
    const bool skip_basic = argc > 3 && !strcmp(argv[3], "split");
    for (int k = 1; k <= n && !skip_basic; k *= 2) { std::vector<int> s(all.begin(), all.begin() + k); three_modes(devs, ("first_" + std::to_string(k)).c_str(), s); }
    for (int i = 0; i < n && n > 1 && !skip_basic; i++) emit("single", {i}, "both", run(devs, {i}, true, true));
    for (int i = 0; i < n && !skip_basic; i++) for (int j = i + 1; j < n; j++) emit("pair", {i, j}, "both", run(devs, {i, j}, true, true));
    const bool only_split = argc > 3 && !strcmp(argv[3], "split");
    if (n >= 8) {      // H2D through one group of devices while D2H goes through the other ("devices" lists the H2D set, then the D2H set)
        emit("split_h2d0123_d2h4567", {0, 1, 2, 3, 4, 5, 6, 7}, "split", run_split(devs, {0, 1, 2, 3}, {4, 5, 6, 7}));
        emit("split_h2d4567_d2h0123", {4, 5, 6, 7, 0, 1, 2, 3}, "split", run_split(devs, {4, 5, 6, 7}, {0, 1, 2, 3}));
        emit("split_h2d45_d2h67", {4, 5, 6, 7}, "split", run_split(devs, {4, 5}, {6, 7}));
        emit("split_h2d01_d2h23", {0, 1, 2, 3}, "split", run_split(devs, {0, 1}, {2, 3}));
        emit("split_h2d0123_d2h4567_again", {0, 1, 2, 3, 4, 5, 6, 7}, "split", run_split(devs, {0, 1, 2, 3}, {4, 5, 6, 7}));
        emit("both_4567_again", {4, 5, 6, 7}, "both", run(devs, {4, 5, 6, 7}, true, true));
    }
    if (only_split) { printf("\n ]}\n"); return 0; }
    if (n >= 8) {
        three_modes(devs, "half_0123", {0, 1, 2, 3}); three_modes(devs, "half_4567", {4, 5, 6, 7});
        three_modes(devs, "even_0246", {0, 2, 4, 6}); three_modes(devs, "mix_0145", {0, 1, 4, 5});
    }
    if (n >= 2) {      // write-combined source buffers for H2D
        alloc_host(devs, cudaHostAllocPortable | cudaHostAllocWriteCombined);
        three_modes(devs, "write_combined_all", all);
        alloc_host(devs, cudaHostAllocPortable);
    }
    for (int k = 0; k < nodes && nodes > 1; k++) {      // host buffers first-touched from a CPU of NUMA node k
        std::string cl = read_file("/sys/devices/system/node/node" + std::to_string(k) + "/cpulist");
        int cpu = atoi(cl.c_str());
        cpu_set_t set; CPU_ZERO(&set); CPU_SET(cpu, &set); sched_setaffinity(0, sizeof set, &set);
        alloc_host(devs, cudaHostAllocPortable);
        three_modes(devs, ("numa_node_" + std::to_string(k) + "_all").c_str(), all);
        for (int i = 0; i < n; i++) emit(("numa_node_" + std::to_string(k) + "_single").c_str(), {i}, "both", run(devs, {i}, true, true));
    }
    printf("\n ]}\n");
    return 0;
}
