// h2d_concurrent.cu -- what can the host feed?  Aggregate pinned host->device bandwidth with 1, 2, 4, ... GPUs copying
// at the same time, one host thread + one pinned buffer per GPU, plus each GPU's PCI address and NUMA node from sysfs.
// The end-to-end detect path moves 2N bytes per block over this path, so this is the ceiling of `e2e` at N GPUs.
//
//   nvcc -O2 -o h2d_concurrent h2d_concurrent.cu -lpthread && ./h2d_concurrent [seconds per point] [MiB per copy]
// One JSON line per GPU count.
#include <cuda_runtime.h>
#include <sched.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static int numa_node_of(int dev, char *busid, size_t n) {
    busid[0] = 0;
    if (cudaDeviceGetPCIBusId(busid, (int)n, dev) != cudaSuccess) return -1;
    for (char *c = busid; *c; ++c) *c = (char)tolower(*c);
    std::string path = std::string("/sys/bus/pci/devices/") + busid + "/numa_node";
    FILE *f = fopen(path.c_str(), "r");
    if (!f) return -1;
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -1;
    fclose(f);
    return node;
}

static bool bind_to_node(int node) {
    if (node < 0) return false;
    char path[128];
    snprintf(path, sizeof path, "/sys/devices/system/node/node%d/cpulist", node);
    FILE *f = fopen(path, "r");
    if (!f) return false;
    char buf[1024] = {0};
    if (!fgets(buf, sizeof buf, f)) { fclose(f); return false; }
    fclose(f);
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char *tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int lo, hi;
        if (sscanf(tok, "%d-%d", &lo, &hi) == 2) { for (int c = lo; c <= hi; ++c) { CPU_SET(c, &set); ++n; } }
        else if (sscanf(tok, "%d", &lo) == 1) { CPU_SET(lo, &set); ++n; }
    }
    return n > 0 && sched_setaffinity(0, sizeof set, &set) == 0;
}

int main(int argc, char **argv) {
    const double seconds = argc > 1 ? atof(argv[1]) : 1.0;
    const size_t bytes = (size_t)(argc > 2 ? atoi(argv[2]) : 128) << 20;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
    for (int bind = 0; bind <= 1; ++bind) {
        for (int g = 1; g <= ndev; g *= 2) {
            std::vector<double> gbs(g, 0.0);
            std::vector<int> nodes(g, -1);
            std::vector<std::string> bus(g);
            std::atomic<int> ready(0);
            std::atomic<bool> go(false);
            std::vector<std::thread> th;
            int bound = 0;
            std::atomic<int> nbound(0);
            for (int d = 0; d < g; ++d)
                th.emplace_back([&, d] {
                    char busid[32];
                    nodes[d] = numa_node_of(d, busid, sizeof busid);
                    bus[d] = busid;
                    if (bind && bind_to_node(nodes[d])) nbound++;
                    cudaSetDevice(d);
                    void *h = nullptr, *dv = nullptr;
                    cudaMallocHost(&h, bytes);              // first touched on this thread: node-local when bound
                    memset(h, 1, bytes);
                    cudaMalloc(&dv, bytes);
                    cudaStream_t s;
                    cudaStreamCreate(&s);
                    cudaMemcpyAsync(dv, h, bytes, cudaMemcpyHostToDevice, s);
                    cudaStreamSynchronize(s);
                    ready++;
                    while (!go.load()) std::this_thread::yield();
                    const auto t0 = std::chrono::steady_clock::now();
                    size_t moved = 0;
                    double el = 0;
                    do {
                        for (int k = 0; k < 4; ++k) cudaMemcpyAsync(dv, h, bytes, cudaMemcpyHostToDevice, s);
                        cudaStreamSynchronize(s);
                        moved += 4 * bytes;
                        el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
                    } while (el < seconds);
                    gbs[d] = moved / el / 1e9;
                    cudaFree(dv);
                    cudaFreeHost(h);
                    cudaStreamDestroy(s);
                });
            while (ready.load() < g) std::this_thread::yield();
            go = true;
            for (auto &t : th) t.join();
            bound = nbound.load();
            double total = 0;
            for (double v : gbs) total += v;
            printf("{\"gpus\": %d, \"numa_bound_threads\": %d, \"aggregate_h2d_gbs\": %.2f, \"per_gpu_gbs\": [", g, bound, total);
            for (int d = 0; d < g; ++d) printf("%s%.2f", d ? ", " : "", gbs[d]);
            printf("], \"pci\": [");
            for (int d = 0; d < g; ++d) printf("%s\"%s\"", d ? ", " : "", bus[d].c_str());
            printf("], \"numa_node\": [");
            for (int d = 0; d < g; ++d) printf("%s%d", d ? ", " : "", nodes[d]);
            printf("], \"mib_per_copy\": %zu, \"seconds\": %.1f}\n", bytes >> 20, seconds);
            fflush(stdout);
        }
        // a second pass with the threads bound to their GPU's NUMA node only makes sense if sysfs knows the nodes
        char busid[32];
        if (numa_node_of(0, busid, sizeof busid) < 0) break;
    }
    return 0;
}
