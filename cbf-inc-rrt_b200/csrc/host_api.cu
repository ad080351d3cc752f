// host_api.cu -- host-buffer front end of libcbfrrt (include/cbfrrt.h: cbfrrt_safety_filter_host).
//
// What a numpy/ctypes caller of the reference would use: every pointer is a HOST pointer.  Rows are cut into
// chunks that flow through NSTREAM private streams (H2D copy -> fused kernel -> D2H copy), so the PCIe
// transfers of one chunk overlap the kernel of another.  Copies run straight from/to the caller's buffers:
// truly asynchronous when those are pinned (bench.py pins them), staged by the driver otherwise.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "../../include/cbfrrt.h"

namespace {

constexpr int NSTREAM = 4;

struct Lane {
    cudaStream_t st = nullptr;
    float *q = nullptr, *goal = nullptr, *datax = nullptr, *h = nullptr, *J = nullptr, *u = nullptr, *r = nullptr;
    uint8_t *safe = nullptr, *unsafe = nullptr;
};

struct Ctx {
    int device = -1;
    int64_t cap_rows = 0;
    Lane lane[NSTREAM];
    float *boxes = nullptr, *spheres = nullptr, *weights = nullptr, *small = nullptr;  // small: goal(1 row) K lo hi
    void* grid = nullptr;
    int64_t cap_boxes = 0, cap_spheres = 0, cap_weights = 0;
    std::vector<unsigned char> resident;   // host image of what boxes/spheres/weights/small/grid currently hold
    bool resident_grid = false;
    std::mutex mu;
};

Ctx g_ctx[16];

#define CK(x)                                    \
    do {                                         \
        if ((x) != cudaSuccess) return CBFRRT_ERR_CUDA; \
    } while (0)

template <class T>
int grow(T*& p, size_t bytes) {
    if (p) cudaFree(p);
    p = nullptr;                       // never leave a dangling pointer behind a failed allocation
    return cudaMalloc(&p, bytes) == cudaSuccess ? CBFRRT_OK : CBFRRT_ERR_CUDA;
}

// (re)allocates the arena of `device`; every capacity is committed only after its allocations succeeded, so a failed
// call leaves the context consistent (smaller capacities, null pointers) for the next one
int ensure(Ctx& c, int device, int64_t rows, int n, int64_t n_boxes, int64_t n_spheres, int64_t n_weights) {
    if (c.device != device) {
        for (auto& l : c.lane)
            if (!l.st) CK(cudaStreamCreateWithFlags(&l.st, cudaStreamNonBlocking));
        if (!c.small) CK(cudaMalloc(&c.small, 4 * (8 + 64 + 8 + 8)));
        if (!c.grid) CK(cudaMalloc(&c.grid, cbfrrt_grid_bytes()));
        c.device = device;
    }
    if (rows > c.cap_rows) {
        c.cap_rows = 0;
        for (auto& l : c.lane) {
            if (grow(l.q, rows * 8 * 4) || grow(l.goal, rows * 8 * 4) || grow(l.datax, rows * 17 * 4) ||
                grow(l.h, rows * 4) || grow(l.J, rows * 8 * 4) || grow(l.u, rows * 8 * 4) || grow(l.r, rows * 4) ||
                grow(l.safe, rows) || grow(l.unsafe, rows))
                return CBFRRT_ERR_CUDA;
        }
        c.cap_rows = rows;
    }
    if (n_boxes > c.cap_boxes || n_spheres > c.cap_spheres || n_weights > c.cap_weights) c.resident.clear();
    if (n_boxes > c.cap_boxes) { c.cap_boxes = 0; if (grow(c.boxes, n_boxes * 32)) return CBFRRT_ERR_CUDA; c.cap_boxes = n_boxes; }
    if (n_spheres > c.cap_spheres) { c.cap_spheres = 0; if (grow(c.spheres, n_spheres * 16)) return CBFRRT_ERR_CUDA; c.cap_spheres = n_spheres; }
    if (n_weights > c.cap_weights) { c.cap_weights = 0; if (grow(c.weights, n_weights * 4)) return CBFRRT_ERR_CUDA; c.cap_weights = n_weights; }
    return CBFRRT_OK;
}

// restores the caller's current device on every exit path
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

extern "C" int cbfrrt_safety_filter_host(int device, int robot, const cbfrrt_host_scene* scene, int32_t hidden,
                                         int32_t layers, const float* weights, int64_t n_weights, const float* q,
                                         int64_t B, const float* goal, int64_t goal_rows, const float* K,
                                         const float* u_lo, const float* u_hi, float cbf_lambda,
                                         float relaxation_penalty, float thr_soft, float thr_hard, uint8_t* safe,
                                         uint8_t* unsafe, float* datax, float* h, float* J, float* u, float* r) {
    int32_t n = 0, nl = 0, ns = 0;
    int rc = cbfrrt_robot_info(robot, &n, &nl, &ns);
    if (rc) return rc;
    if (device < 0 || device >= 16 || B < 0 || !scene || !weights || !q || !goal || !K || !u_lo || !u_hi)
        return CBFRRT_ERR_BAD_SHAPE;
    if (goal_rows != 1 && goal_rows != B) return CBFRRT_ERR_BAD_SHAPE;
    if (hidden < 1 || layers < 0) return CBFRRT_ERR_BAD_SHAPE;
    // packed weight count of include/cbfrrt.h: a short buffer would make the kernels read past the staged copy
    const int64_t in_pad = (n + 1 + 3) & ~3;
    if (n_weights != (int64_t)hidden * in_pad + hidden + (int64_t)layers * (hidden * hidden + hidden) + hidden + 4)
        return CBFRRT_ERR_BAD_SHAPE;
    if (scene->n_boxes < 0 || scene->n_spheres < 0 || (scene->n_boxes && !scene->boxes) || (scene->n_spheres && !scene->spheres))
        return CBFRRT_ERR_BAD_SHAPE;
    if (B == 0) return CBFRRT_OK;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device >= n_dev) return CBFRRT_ERR_BAD_SHAPE;
    DeviceGuard guard(device);
    Ctx& c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);

    // chunking: ~8 chunks over NSTREAM streams overlap the copies with compute (measured on cfg1, PCIe Gen5: 8 chunks
    // 0.65 ms, 4 chunks 0.67 ms, 2 chunks 0.81 ms, 1 chunk 1.08 ms); a chunk is a whole number of kernel waves
    // (SMs x 4 tile groups x 128 rows) so that no launch ends with a mostly idle wave
    static const int n_chunks_env = getenv("CBFRRT_HOST_CHUNKS") ? atoi(getenv("CBFRRT_HOST_CHUNKS")) : 0;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int64_t wave = (int64_t)sms * 512, max_chunk = 1 << 20;
    int64_t chunk;
    if (n_chunks_env > 0) {
        chunk = (B + n_chunks_env - 1) / n_chunks_env;
    } else {
        int64_t waves_per_chunk = (B + 4 * wave) / (8 * wave);      // round(B / wave / 8)
        if (waves_per_chunk < 1) waves_per_chunk = 1;
        chunk = waves_per_chunk * wave;
    }
    if (chunk > max_chunk) chunk = max_chunk;
    if (chunk > B) chunk = B;
    rc = ensure(c, device, chunk, n, scene->n_boxes, scene->n_spheres, n_weights);
    if (rc) return rc;

    // scene / weights / small parameters: tiny (<= 60 KB).  They are uploaded on lane 0 only when they differ from what
    // the device already holds (planners call with the same scene and network thousands of times); the other lanes
    // wait for the upload through the stream synchronisation below.
    cudaStream_t s0 = c.lane[0].st;
    std::vector<float> b8((size_t)scene->n_boxes * 8, 0.f);
    for (int i = 0; i < scene->n_boxes; ++i) memcpy(&b8[8 * i], scene->boxes + 6 * i, 24);
    float small[8 + 64 + 8 + 8] = {0};
    if (goal_rows == 1) memcpy(small, goal, 4 * n);
    memcpy(small + 8, K, 4 * n * n);
    memcpy(small + 72, u_lo, 4 * n);
    memcpy(small + 80, u_hi, 4 * n);
    // broad phase (built once per scene, kept resident): dense scenes from a few thousand rows on; small scenes for
    // planner-sized calls, where the kernels prefer the field path (kernels.cuh: scene_for) and the build is a few us
    const int n_obs = scene->n_boxes + scene->n_spheres;
    const bool use_grid = n_obs >= 1 && (n_obs >= 12 ? (B >= 4096 || B <= 16384) : B <= 16384);
    std::vector<unsigned char> image(16 + b8.size() * 4 + (size_t)scene->n_spheres * 16 + (size_t)n_weights * 4 + sizeof small);
    {
        unsigned char* w = image.data();
        const int32_t hdr[4] = {scene->n_boxes, scene->n_spheres, (int32_t)n_weights, scene->floor};
        memcpy(w, hdr, 16); w += 16;
        memcpy(w, b8.data(), b8.size() * 4); w += b8.size() * 4;
        if (scene->n_spheres) memcpy(w, scene->spheres, (size_t)scene->n_spheres * 16);
        w += (size_t)scene->n_spheres * 16;
        memcpy(w, weights, (size_t)n_weights * 4); w += (size_t)n_weights * 4;
        memcpy(w, small, sizeof small);
    }
    cbfrrt_scene dsc{scene->n_boxes, scene->n_spheres, scene->floor, 0, c.boxes, c.spheres, nullptr, nullptr};
    if (image != c.resident || (use_grid && !c.resident_grid)) {
        c.resident.clear();
        if (scene->n_boxes) CK(cudaMemcpyAsync(c.boxes, b8.data(), b8.size() * 4, cudaMemcpyHostToDevice, s0));
        if (scene->n_spheres) CK(cudaMemcpyAsync(c.spheres, scene->spheres, (size_t)scene->n_spheres * 16, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(c.weights, weights, n_weights * 4, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(c.small, small, sizeof small, cudaMemcpyHostToDevice, s0));
        if (use_grid) {
            rc = cbfrrt_grid_build(&dsc, c.grid, s0);
            if (rc) return rc;
        }
        CK(cudaStreamSynchronize(s0));  // b8/small live on this stack frame; lanes 1.. may now read scene + grid
        c.resident.swap(image);
        c.resident_grid = use_grid;
    }
    if (use_grid) dsc.grid = c.grid;

    cbfrrt_cbf_net net{n + 1, hidden, layers, 0, c.weights};

    // an error inside the loop must not leave the other lanes' asynchronous copies into the caller's buffers in flight
    auto fail = [&](int code) {
        for (auto& l : c.lane) cudaStreamSynchronize(l.st);
        return code;
    };
#define CKL(x)                                          \
    do {                                                \
        if ((x) != cudaSuccess) return fail(CBFRRT_ERR_CUDA); \
    } while (0)
    int lane_i = 0;
    for (int64_t off = 0; off < B; off += chunk, lane_i = (lane_i + 1) % NSTREAM) {
        Lane& l = c.lane[lane_i];
        const int64_t rows = (B - off < chunk) ? (B - off) : chunk;
        CKL(cudaMemcpyAsync(l.q, q + off * n, rows * n * 4, cudaMemcpyHostToDevice, l.st));
        cbfrrt_filter_params fp{c.small, 1, c.small + 8, c.small + 72, c.small + 80, cbf_lambda, relaxation_penalty, nullptr};
        if (goal_rows != 1) {
            CKL(cudaMemcpyAsync(l.goal, goal + off * n, rows * n * 4, cudaMemcpyHostToDevice, l.st));
            fp.goal = l.goal;
            fp.goal_rows = rows;
        }
        rc = cbfrrt_safety_filter(robot, &dsc, &net, l.q, n, rows, &fp, thr_soft, thr_hard, safe ? l.safe : nullptr,
                                  unsafe ? l.unsafe : nullptr, (datax || hidden != 48 || layers != 2) ? l.datax : nullptr, h ? l.h : nullptr,
                                  J ? l.J : nullptr, u ? l.u : nullptr, r ? l.r : nullptr, l.st);
        if (rc) return fail(rc);
        if (safe) CKL(cudaMemcpyAsync(safe + off, l.safe, rows, cudaMemcpyDeviceToHost, l.st));
        if (unsafe) CKL(cudaMemcpyAsync(unsafe + off, l.unsafe, rows, cudaMemcpyDeviceToHost, l.st));
        if (datax) CKL(cudaMemcpyAsync(datax + off * (2 * n + 1), l.datax, rows * (2 * n + 1) * 4, cudaMemcpyDeviceToHost, l.st));
        if (h) CKL(cudaMemcpyAsync(h + off, l.h, rows * 4, cudaMemcpyDeviceToHost, l.st));
        if (J) CKL(cudaMemcpyAsync(J + off * n, l.J, rows * n * 4, cudaMemcpyDeviceToHost, l.st));
        if (u) CKL(cudaMemcpyAsync(u + off * n, l.u, rows * n * 4, cudaMemcpyDeviceToHost, l.st));
        if (r) CKL(cudaMemcpyAsync(r + off, l.r, rows * 4, cudaMemcpyDeviceToHost, l.st));
    }
    for (auto& l : c.lane) CKL(cudaStreamSynchronize(l.st));
    return CBFRRT_OK;
#undef CKL
}
