// dmc_hostlink.cu -- what the host link of this box can carry, and which devices should carry it.
//
// The filter path itself has no exchange between GPUs (frames are independent), but every frame of a host-resident batch
// crosses the host link twice.  On a multi-GPU box the link is not uniform: profiles/r02_hostlink.json (8 x B200 behind a
// KVM hypervisor) shows GPUs 0-3 behind one upstream that carries ~51 GB/s each way for all four together, GPUs 4-7 at
// ~94 GB/s each way together, and all eight together at only ~64 GB/s -- traffic through the slow group costs the shared
// resource about 1.8x as much per byte.  dmc_hostlink_probe measures that in ~0.1 s and proposes a routing: the devices
// whose links are worth using carry the host traffic ("gateways"), the others reach host memory THROUGH a gateway's HBM
// over NVLink/NVSwitch (their kernels read the input from, and write the output to, peer memory directly).
#include "../../include/dmc_c.h"
#include "dmc_common.cuh"

#include <string.h>
#include <algorithm>
#include <chrono>
#include <string>
#include <vector>

namespace {

struct ProbeDev { int id = 0; char* d_in = nullptr; char* d_out = nullptr; cudaStream_t s_in = nullptr, s_out = nullptr; cudaEvent_t a_in = nullptr, b_in = nullptr, a_out = nullptr, b_out = nullptr; };

struct ProbeRun { double each_way_gbs; std::vector<double> dev_gbs; };

// Every device of `set` copies `reps` x `bytes` host->device and device->host at the same time.
bool run_set(std::vector<ProbeDev>& devs, const std::vector<int>& set, char* h_in, char* h_out, size_t bytes, int reps, ProbeRun* out) {
    for (int i : set) { if (cudaSetDevice(devs[i].id) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) return false; }
    auto t0 = std::chrono::steady_clock::now();
    for (int i : set) { ProbeDev& d = devs[i]; cudaSetDevice(d.id); cudaEventRecord(d.a_in, d.s_in); cudaEventRecord(d.a_out, d.s_out); }
    for (int r = 0; r < reps; r++)
        for (int i : set) {
            ProbeDev& d = devs[i]; cudaSetDevice(d.id);
            cudaMemcpyAsync(d.d_in, h_in + (size_t)i * bytes, bytes, cudaMemcpyHostToDevice, d.s_in);
            cudaMemcpyAsync(h_out + (size_t)i * bytes, d.d_out, bytes, cudaMemcpyDeviceToHost, d.s_out);
        }
    for (int i : set) { ProbeDev& d = devs[i]; cudaSetDevice(d.id); cudaEventRecord(d.b_in, d.s_in); cudaEventRecord(d.b_out, d.s_out); }
    for (int i : set) { ProbeDev& d = devs[i]; cudaSetDevice(d.id); if (cudaStreamSynchronize(d.s_in) != cudaSuccess || cudaStreamSynchronize(d.s_out) != cudaSuccess) return false; }
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    out->each_way_gbs = (double)bytes * reps * set.size() / 1e9 / wall;
    out->dev_gbs.assign(devs.size(), 0.0);
    for (int i : set) {
        ProbeDev& d = devs[i]; float m_in = 0, m_out = 0;
        cudaEventElapsedTime(&m_in, d.a_in, d.b_in); cudaEventElapsedTime(&m_out, d.a_out, d.b_out);
        const double g = (double)bytes * reps / 1e6;
        out->dev_gbs[i] = 0.5 * (g / (m_in > 0 ? m_in : 1e-3) + g / (m_out > 0 ? m_out : 1e-3));
    }
    return cudaGetLastError() == cudaSuccess;
}

}  // namespace

// Gives back everything this process holds on `device` (cudaDeviceReset): a process that only measured a device's link
// (dmc_hostlink_probe) and will not compute on it need not keep a context there.  Only for devices on which the caller has
// no live allocations, streams or dmc contexts.  The calling thread's current device is `device` afterwards.
extern "C" int dmc_release_device(int device) {
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || device < 0 || device >= visible) return DMC_ERR_ARG;
    if (cudaSetDevice(device) != cudaSuccess || cudaDeviceReset() != cudaSuccess) { cudaGetLastError(); return DMC_ERR_CUDA; }
    return DMC_OK;
}

extern "C" int dmc_hostlink_probe(const int* devices, int n, dmc_hostlink_info* info) {
    dmc::DeviceGuard keep_device;
    if (!devices || !info || n < 1 || n > DMC_MAX_DEVICES) return DMC_ERR_ARG;
    int visible = 0;
    if (cudaGetDeviceCount(&visible) != cudaSuccess || visible == 0) return DMC_ERR_CUDA;
    for (int i = 0; i < n; i++) { if (devices[i] < 0 || devices[i] >= visible) return DMC_ERR_ARG; for (int j = 0; j < i; j++) if (devices[j] == devices[i]) return DMC_ERR_ARG; }
    memset(info, 0, sizeof *info);
    info->n_devices = n;
    const size_t bytes = (size_t)64 << 20; const int reps = 3;
    std::vector<ProbeDev> devs(n);
    char *h_in = nullptr, *h_out = nullptr;
    bool ok = cudaHostAlloc((void**)&h_in, bytes * n, cudaHostAllocPortable) == cudaSuccess && cudaHostAlloc((void**)&h_out, bytes * n, cudaHostAllocPortable) == cudaSuccess;
    if (ok) { memset(h_in, 1, bytes * n); memset(h_out, 0, bytes * n); }
    for (int i = 0; i < n && ok; i++) {
        ProbeDev& d = devs[i]; d.id = devices[i];
        ok = cudaSetDevice(d.id) == cudaSuccess && cudaMalloc((void**)&d.d_in, bytes) == cudaSuccess && cudaMalloc((void**)&d.d_out, bytes) == cudaSuccess &&
             cudaStreamCreateWithFlags(&d.s_in, cudaStreamNonBlocking) == cudaSuccess && cudaStreamCreateWithFlags(&d.s_out, cudaStreamNonBlocking) == cudaSuccess &&
             cudaEventCreate(&d.a_in) == cudaSuccess && cudaEventCreate(&d.b_in) == cudaSuccess && cudaEventCreate(&d.a_out) == cudaSuccess && cudaEventCreate(&d.b_out) == cudaSuccess;
    }
    std::vector<int> all(n); for (int i = 0; i < n; i++) all[i] = i;
    ProbeRun ra, rf;
    if (ok) ok = run_set(devs, all, h_in, h_out, bytes / 4, 1, &ra);      // warm-up (first touch of the pinned pages, copy-engine spin-up)
    if (ok) ok = run_set(devs, all, h_in, h_out, bytes, reps, &ra);
    if (ok) {
        info->all_gbs = ra.each_way_gbs; info->best_gbs = ra.each_way_gbs; info->n_link = n;
        double rmax = 0;
        for (int i = 0; i < n; i++) { info->device[i] = devices[i]; info->gateway[i] = devices[i]; info->loaded_gbs[i] = ra.dev_gbs[i]; rmax = std::max(rmax, ra.dev_gbs[i]); }
        std::vector<int> fast;
        for (int i = 0; i < n; i++) if (ra.dev_gbs[i] >= 0.85 * rmax) fast.push_back(i);
        if ((int)fast.size() < n && !fast.empty()) {
            ok = run_set(devs, fast, h_in, h_out, bytes, reps, &rf);
            if (ok && rf.each_way_gbs > 1.05 * ra.each_way_gbs) {
                // the fast set alone moves more than everybody together: route the others through it (if NVLink peer access exists)
                std::vector<int> gw(n); for (int i = 0; i < n; i++) gw[i] = i;
                size_t k = 0; bool all_peer = true;
                for (int i = 0; i < n; i++) {
                    if (std::find(fast.begin(), fast.end(), i) != fast.end()) continue;
                    const int g = fast[k++ % fast.size()];
                    int can_a = 0, can_b = 0;
                    cudaDeviceCanAccessPeer(&can_a, devices[i], devices[g]); cudaDeviceCanAccessPeer(&can_b, devices[g], devices[i]);
                    if (can_a && can_b) gw[i] = g; else all_peer = false;
                }
                if (all_peer) {
                    for (int i = 0; i < n; i++) info->gateway[i] = devices[gw[i]];
                    info->best_gbs = rf.each_way_gbs; info->n_link = (int)fast.size();
                }
            }
        }
    }
    for (auto& d : devs) {
        cudaSetDevice(d.id);
        if (d.s_in) cudaStreamDestroy(d.s_in);
        if (d.s_out) cudaStreamDestroy(d.s_out);
        for (cudaEvent_t e : {d.a_in, d.b_in, d.a_out, d.b_out}) if (e) cudaEventDestroy(e);
        if (d.d_in) cudaFree(d.d_in);
        if (d.d_out) cudaFree(d.d_out);
    }
    if (h_in) cudaFreeHost(h_in);
    if (h_out) cudaFreeHost(h_out);
    if (!ok) { cudaGetLastError(); return DMC_ERR_CUDA; }
    return DMC_OK;
}
