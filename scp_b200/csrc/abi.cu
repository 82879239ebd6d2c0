// Library-wide plumbing of the C ABI: error strings, version, device check, launch counter, and the
// legacy symbols of data_preproc/OctreeCPP/Octree_python_lib.so (Octreewarpper.py:17-39).
#include <vector>
#include <mutex>
#include <string>
#include <string.h>
#include "common.cuh"

namespace scp {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

cudaError_t malloc_async(void** p, size_t bytes, cudaStream_t st) {
    static std::atomic<unsigned long long> configured{0};          // bit per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(configured.load() & bit)) {
        cudaMemPool_t pool;
        e = cudaDeviceGetDefaultMemPool(&pool, dev);
        if (e != cudaSuccess) return e;
        unsigned long long keep = ~0ull;
        e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        if (e != cudaSuccess) return e;
        configured.fetch_or(bit);
    }
    return cudaMallocAsync(p, bytes, st);
}

static std::mutex g_pin_mu;
static std::vector<PinnedBlock> g_pin_free;

bool pinned_get(size_t bytes, PinnedBlock* out) {
    {
        std::lock_guard<std::mutex> g(g_pin_mu);
        for (size_t i = 0; i < g_pin_free.size(); ++i) {
            if (g_pin_free[i].bytes >= bytes && g_pin_free[i].bytes <= 4 * bytes + 4096) {
                *out = g_pin_free[i];
                g_pin_free.erase(g_pin_free.begin() + i);
                cudaEventSynchronize(out->ev);
                return true;
            }
        }
    }
    PinnedBlock b{nullptr, bytes < 4096 ? (size_t)4096 : bytes, nullptr};
    if (cudaMallocHost(&b.p, b.bytes) != cudaSuccess) return false;
    if (cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) { cudaFreeHost(b.p); return false; }
    *out = b;
    return true;
}

void pinned_put(const PinnedBlock& b) {
    std::lock_guard<std::mutex> g(g_pin_mu);
    if (g_pin_free.size() < 256) { g_pin_free.push_back(b); return; }
    cudaEventDestroy(b.ev);
    cudaFreeHost(b.p);
}

cudaError_t upload_async(void** d, const void* h, size_t bytes, cudaStream_t st) {
    PinnedBlock pb;
    if (!pinned_get(bytes, &pb)) return cudaErrorMemoryAllocation;
    memcpy(pb.p, h, bytes);
    cudaError_t e = malloc_async(d, bytes, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(*d, pb.p, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(pb.ev, st);
    pinned_put(pb);
    return e;
}

}  // namespace scp

using namespace scp;

// ---- legacy host containers --------------------------------------------------------------
struct LegacyLevel { std::vector<scp_legacy_node> nodes; };
struct LegacyTree {
    std::vector<LegacyLevel*> levels;
    std::vector<int> pushed;            // vector_push_back keeps the reference's vector<int>-like behaviour
    std::vector<int>* codes = nullptr;  // owned here (the reference leaks it)
    ~LegacyTree() { for (auto* l : levels) delete l; delete codes; }
};

extern "C" {

const char* scp_last_error(void) { return g_err; }
int scp_version(void) { return 100; }
int64_t scp_launch_count(void) { return g_launches.load(); }

int scp_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { set_error("no CUDA device"); return 0; }
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) { set_error("cudaGetDeviceProperties failed"); return 0; }
    if (p.major != 10) { set_error("device %s is sm_%d%d, this library is sm_100a only", p.name, p.major, p.minor); return 0; }
    return 1;
}

void* new_vector(void) { return new LegacyTree(); }
void delete_vector(void* v) { delete static_cast<LegacyTree*>(v); }
int vector_size(void* v) { return (int)static_cast<LegacyTree*>(v)->levels.size(); }
void* vector_get(void* v, int i) {
    auto* t = static_cast<LegacyTree*>(v);
    return (i >= 0 && i < (int)t->levels.size()) ? t->levels[i] : nullptr;
}
void vector_push_back(void* v, int i) { static_cast<LegacyTree*>(v)->pushed.push_back(i); }
scp_legacy_node* Nodes_get(void* level, int i) { return &static_cast<LegacyLevel*>(level)->nodes[i]; }
int Nodes_size(void* level) { return (int)static_cast<LegacyLevel*>(level)->nodes.size(); }
int int_size(void* codes) { return (int)static_cast<std::vector<int>*>(codes)->size(); }
int int_get(void* codes, int i) { return (*static_cast<std::vector<int>*>(codes))[i]; }

// genOctreeInterface(levels, xyz (n,3) row-major doubles holding non-negative integers, n) -> code vector.
// Builds the tree with the CUDA pipeline (cartesian mode, step 1) and copies the node records back.
void* genOctreeInterface(void* levels, const double* xyz, int n) {
    auto* tree = static_cast<LegacyTree*>(levels);
    if (!tree || !xyz || n <= 0) { set_error("genOctreeInterface: bad argument"); return nullptr; }
    for (auto* l : tree->levels) delete l;
    tree->levels.clear();
    delete tree->codes;
    tree->codes = new std::vector<int>();
    std::vector<float> pts((size_t)n * 3);
    for (size_t i = 0; i < (size_t)n * 3; ++i) pts[i] = (float)xyz[i];      // exact below 2^24
    float* d_pts = nullptr;
    scp_octree* oc = scp_octree_create();
    uint8_t *d_occ = nullptr, *d_oct = nullptr; uint32_t *d_par = nullptr, *d_pos = nullptr;
    bool ok = false;
    do {
        if (cudaMalloc(&d_pts, pts.size() * 4) != cudaSuccess) break;
        if (cudaMemcpy(d_pts, pts.data(), pts.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) break;
        int64_t off[2] = {0, n};
        scp_job job;
        memset(&job, 0, sizeof(job));
        job.qs = 1.0; job.lidar_level = 255; job.pos_eps_last = 1;
        if (scp_octree_plan(oc, d_pts, 3, off, 1, &job, 1, SCP_MODE_CART, nullptr) != SCP_OK) break;
        scp_job_info info;
        scp_octree_job_info(oc, 0, &info);
        const int64_t N = info.n_rows;
        if (cudaMalloc(&d_occ, N) != cudaSuccess || cudaMalloc(&d_oct, N) != cudaSuccess ||
            cudaMalloc(&d_par, N * 4) != cudaSuccess || cudaMalloc(&d_pos, N * 12) != cudaSuccess) break;
        scp_octree_out out;
        memset(&out, 0, sizeof(out));
        out.occ = d_occ; out.octant = d_oct; out.parent = d_par; out.pos = d_pos;
        if (scp_octree_emit(oc, &out, nullptr) != SCP_OK) break;
        std::vector<uint8_t> h_occ(N), h_oct(N);
        std::vector<uint32_t> h_par(N), h_pos(N * 3);
        if (cudaMemcpy(h_occ.data(), d_occ, N, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(h_oct.data(), d_oct, N, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(h_par.data(), d_par, N * 4, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        if (cudaMemcpy(h_pos.data(), d_pos, N * 12, cudaMemcpyDeviceToHost) != cudaSuccess) break;
        int64_t r = 0;
        for (int L = 0; L < info.depth; ++L) {
            auto* lv = new LegacyLevel();
            lv->nodes.resize(info.level_rows[L]);
            for (int k = 0; k < info.level_rows[L]; ++k, ++r) {
                scp_legacy_node& nd = lv->nodes[k];
                nd.nodeid = (unsigned)(r + 1);                       // 1-based BFS id
                nd.octant = h_oct[r];
                nd.parent = L == 0 ? 0u : h_par[r] + 1u;             // nodeid of the parent
                nd.oct = h_occ[r];
                nd.pos[0] = h_pos[3 * r]; nd.pos[1] = h_pos[3 * r + 1]; nd.pos[2] = h_pos[3 * r + 2];
                tree->codes->push_back(h_occ[r]);
            }
            tree->levels.push_back(lv);
        }
        ok = true;
    } while (0);
    cudaFree(d_pts); cudaFree(d_occ); cudaFree(d_oct); cudaFree(d_par); cudaFree(d_pos);
    scp_octree_destroy(oc);
    if (!ok) {
        if (!scp_last_error()[0]) set_error("genOctreeInterface: CUDA failure (%s)", cudaGetErrorString(cudaGetLastError()));
        return nullptr;
    }
    return tree->codes;
}

}  // extern "C"
