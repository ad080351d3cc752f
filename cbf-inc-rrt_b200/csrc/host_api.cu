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
    std::vector<unsigned char> resident;       // host image of what boxes / spheres (and the grid built from them) hold
    std::vector<unsigned char> resident_net;   // ... of what weights / small hold
    bool resident_grid = false;
    // planner-sized calls (cbfrrt_rollout_segments_host, cbfrrt_knn_host): one pinned block and its device twin, so that a
    // call is one H2D copy, the kernels and one D2H copy on lane 0
    unsigned char *pin = nullptr, *dblk = nullptr;
    size_t blk_cap = 0;
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
    if (n_boxes > c.cap_boxes || n_spheres > c.cap_spheres) c.resident.clear();
    if (n_weights > c.cap_weights) c.resident_net.clear();
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

// Scene and network constants are tiny (<= 60 KB).  Each of the two groups is uploaded on lane 0 only when its bytes
// differ from what the device already holds (planners call with the same scene and network thousands of times, and
// alternate between entry points that need only one of the groups); the broad phase is rebuilt with the scene.  Both
// return after the upload is complete, so every lane may read the buffers.  dsc: the device view of the scene.
int make_resident_scene(Ctx& c, const cbfrrt_host_scene* scene, bool use_grid, cbfrrt_scene& dsc) {
    cudaStream_t s0 = c.lane[0].st;
    std::vector<float> b8((size_t)scene->n_boxes * 8, 0.f);
    for (int i = 0; i < scene->n_boxes; ++i) memcpy(&b8[8 * i], scene->boxes + 6 * i, 24);
    std::vector<unsigned char> image(16 + b8.size() * 4 + (size_t)scene->n_spheres * 16);
    {
        unsigned char* w = image.data();
        const int32_t hdr[4] = {scene->n_boxes, scene->n_spheres, scene->floor, 0};
        memcpy(w, hdr, 16); w += 16;
        memcpy(w, b8.data(), b8.size() * 4); w += b8.size() * 4;
        if (scene->n_spheres) memcpy(w, scene->spheres, (size_t)scene->n_spheres * 16);
    }
    dsc = cbfrrt_scene{scene->n_boxes, scene->n_spheres, scene->floor, 0, c.boxes, c.spheres, nullptr, nullptr};
    if (image != c.resident || (use_grid && !c.resident_grid)) {
        c.resident.clear();
        c.resident_grid = false;
        if (scene->n_boxes) CK(cudaMemcpyAsync(c.boxes, b8.data(), b8.size() * 4, cudaMemcpyHostToDevice, s0));
        if (scene->n_spheres) CK(cudaMemcpyAsync(c.spheres, scene->spheres, (size_t)scene->n_spheres * 16, cudaMemcpyHostToDevice, s0));
        if (use_grid) {
            const int rc = cbfrrt_grid_build(&dsc, c.grid, s0);
            if (rc) return rc;
        }
        CK(cudaStreamSynchronize(s0));  // b8 lives on this stack frame; lanes 1.. may now read scene + grid
        c.resident.swap(image);
        c.resident_grid = use_grid;
    }
    if (use_grid) dsc.grid = c.grid;
    return CBFRRT_OK;
}

int make_resident_net(Ctx& c, int n, const float* weights, int64_t n_weights, const float* goal1, const float* K,
                      const float* u_lo, const float* u_hi) {
    if (n_weights < 0 || n_weights > (1 << 24) || n < 1 || n > 8) return CBFRRT_ERR_BAD_SHAPE;
    cudaStream_t s0 = c.lane[0].st;
    float small[8 + 64 + 8 + 8] = {0};
    if (goal1) memcpy(small, goal1, 4 * n);
    memcpy(small + 8, K, 4 * n * n);
    memcpy(small + 72, u_lo, 4 * n);
    memcpy(small + 80, u_hi, 4 * n);
    const size_t wb = (size_t)n_weights * 4;
    std::vector<unsigned char> image(wb + sizeof small);
    memcpy(image.data(), weights, wb);
    memcpy(&image[wb], small, sizeof small);
    if (image != c.resident_net) {
        c.resident_net.clear();
        CK(cudaMemcpyAsync(c.weights, weights, n_weights * 4, cudaMemcpyHostToDevice, s0));
        CK(cudaMemcpyAsync(c.small, small, sizeof small, cudaMemcpyHostToDevice, s0));
        CK(cudaStreamSynchronize(s0));  // small lives on this stack frame
        c.resident_net.swap(image);
    }
    return CBFRRT_OK;
}

int make_resident(Ctx& c, int n, const cbfrrt_host_scene* scene, const float* weights, int64_t n_weights, const float* goal1,
                  const float* K, const float* u_lo, const float* u_hi, bool use_grid, cbfrrt_scene& dsc) {
    const int rc = make_resident_scene(c, scene, use_grid, dsc);
    return rc ? rc : make_resident_net(c, n, weights, n_weights, goal1, K, u_lo, u_hi);
}

// pinned block + device twin of at least `bytes`
int ensure_block(Ctx& c, size_t bytes) {
    if (bytes <= c.blk_cap) return CBFRRT_OK;
    size_t cap = 1 << 16;
    while (cap < bytes) cap <<= 1;
    c.blk_cap = 0;
    if (c.pin) cudaFreeHost(c.pin);
    c.pin = nullptr;
    if (cudaMallocHost(reinterpret_cast<void**>(&c.pin), cap) != cudaSuccess) return CBFRRT_ERR_CUDA;
    if (grow(c.dblk, cap)) return CBFRRT_ERR_CUDA;
    c.blk_cap = cap;
    return CBFRRT_OK;
}

int64_t packed_weight_count(int n, int hidden, int layers) {
    const int64_t in_pad = (n + 1 + 3) & ~3;
    return (int64_t)hidden * in_pad + hidden + (int64_t)layers * (hidden * hidden + hidden) + hidden + 4;
}

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
    if (n_weights != packed_weight_count(n, hidden, layers)) return CBFRRT_ERR_BAD_SHAPE;
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

    // broad phase (built once per scene, kept resident): dense scenes from a few thousand rows on; small scenes for
    // planner-sized calls, where the kernels prefer the field path (kernels.cuh: scene_for) and the build is a few us
    const int n_obs = scene->n_boxes + scene->n_spheres;
    const bool use_grid = n_obs >= 1 && (n_obs >= 12 ? (B >= 4096 || B <= 16384) : B <= 16384);
    cbfrrt_scene dsc;
    rc = make_resident(c, n, scene, weights, n_weights, goal_rows == 1 ? goal : nullptr, K, u_lo, u_hi, use_grid, dsc);
    if (rc) return rc;

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
                                  unsafe ? l.unsafe : nullptr, (datax || hidden != 48 || layers != 2 || rows >= (1 << 19)) ? l.datax : nullptr, h ? l.h : nullptr,
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

// ---- planner-sized host front ends ------------------------------------------------------------------------------
// The planners of the reference keep their state in numpy (tree, samples, edges): per expansion they need the nearest
// tree nodes of the samples and the steer segments from them.  These two entry points take HOST pointers and do one
// pinned H2D copy, the kernels and one D2H copy on a private stream -- no tensor wrappers, no per-output read-backs.

extern "C" int cbfrrt_rollout_segments_host(int device, int robot, const cbfrrt_host_scene* scene, int32_t hidden,
                                            int32_t layers, const float* weights, int64_t n_weights, const float* q0,
                                            int64_t T, const float* goal, const float* K, const float* u_lo,
                                            const float* u_hi, float cbf_lambda, float relaxation_penalty,
                                            const cbfrrt_rollout_params* rp, int32_t n_segments, float thr_soft,
                                            float thr_hard, float* seg_q, uint8_t* seg_alive, int32_t* seg_done, float* traj) {
    int32_t n = 0, nl = 0, ns = 0;
    int rc = cbfrrt_robot_info(robot, &n, &nl, &ns);
    if (rc) return rc;
    if (device < 0 || device >= 16 || T < 1 || T > 128 || !scene || !weights || !q0 || !goal || !K || !u_lo || !u_hi || !rp ||
        !seg_q || !seg_alive || !seg_done || !traj || n_segments < 1 || rp->steps < 1)
        return CBFRRT_ERR_BAD_SHAPE;
    if (n_weights != packed_weight_count(n, hidden, layers)) return CBFRRT_ERR_BAD_SHAPE;
    if (scene->n_boxes < 0 || scene->n_spheres < 0 || (scene->n_boxes && !scene->boxes) || (scene->n_spheres && !scene->spheres))
        return CBFRRT_ERR_BAD_SHAPE;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device >= n_dev) return CBFRRT_ERR_BAD_SHAPE;
    DeviceGuard guard(device);
    Ctx& c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);
    rc = ensure(c, device, 0, n, scene->n_boxes, scene->n_spheres, n_weights);
    if (rc) return rc;
    const int n_obs = scene->n_boxes + scene->n_spheres;
    cbfrrt_scene dsc;
    rc = make_resident(c, n, scene, weights, n_weights, nullptr, K, u_lo, u_hi, n_obs >= 40 || c.resident_grid, dsc);
    if (rc) return rc;
    // block layout (bytes, every part 16-byte aligned): in [q0 | goal], out [seg_q | traj | seg_done | seg_alive]
    const size_t rows_b = (size_t)T * n * 4, in_b = (2 * rows_b + 15) & ~size_t(15);
    const size_t sq_b = (size_t)T * n_segments * n * 4, tr_b = sq_b * rp->steps, sd_b = (size_t)T * n_segments * 4,
                 sa_b = (size_t)T * n_segments;
    const size_t o_sq = in_b, o_tr = (o_sq + sq_b + 15) & ~size_t(15), o_sd = (o_tr + tr_b + 15) & ~size_t(15),
                 o_sa = (o_sd + sd_b + 15) & ~size_t(15), total = o_sa + sa_b;
    rc = ensure_block(c, total);
    if (rc) return rc;
    cudaStream_t st = c.lane[0].st;
    memcpy(c.pin, q0, rows_b);
    memcpy(c.pin + rows_b, goal, rows_b);
    CK(cudaMemcpyAsync(c.dblk, c.pin, 2 * rows_b, cudaMemcpyHostToDevice, st));
    cbfrrt_cbf_net net{n + 1, hidden, layers, 0, c.weights};
    cbfrrt_filter_params fp{reinterpret_cast<float*>(c.dblk + rows_b), T, c.small + 8, c.small + 72, c.small + 80, cbf_lambda,
                            relaxation_penalty, nullptr};
    rc = cbfrrt_rollout_segments(robot, &dsc, &net, reinterpret_cast<float*>(c.dblk), T, &fp, rp, n_segments, thr_soft, thr_hard,
                                 reinterpret_cast<float*>(c.dblk + o_sq), c.dblk + o_sa, reinterpret_cast<int32_t*>(c.dblk + o_sd),
                                 reinterpret_cast<float*>(c.dblk + o_tr), st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CK(cudaMemcpyAsync(c.pin + o_sq, c.dblk + o_sq, total - o_sq, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(seg_q, c.pin + o_sq, sq_b);
    memcpy(traj, c.pin + o_tr, tr_b);
    memcpy(seg_done, c.pin + o_sd, sd_b);
    memcpy(seg_alive, c.pin + o_sa, sa_b);
    return CBFRRT_OK;
}

extern "C" int cbfrrt_knn_host(int device, int n, const float* queries, const float* tree, int64_t M, int64_t N_tree,
                               int32_t k, int32_t* idx, float* dist) {
    if (device < 0 || device >= 16 || n < 1 || n > 8 || M < 1 || N_tree < 1 || k < 1 || !queries || !tree || !idx || !dist)
        return CBFRRT_ERR_BAD_SHAPE;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device >= n_dev) return CBFRRT_ERR_BAD_SHAPE;
    DeviceGuard guard(device);
    Ctx& c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);
    int rc = ensure(c, device, 0, n, 0, 0, 0);
    if (rc) return rc;
    const size_t q_b = (size_t)M * n * 4, t_b = (size_t)N_tree * n * 4, in_b = (q_b + t_b + 15) & ~size_t(15);
    const size_t out_b = (size_t)M * k * 4, ws_b = (size_t)cbfrrt_knn_workspace_bytes(M, N_tree, k);
    const size_t o_idx = in_b, o_dist = (o_idx + out_b + 15) & ~size_t(15), o_ws = (o_dist + out_b + 15) & ~size_t(15);
    rc = ensure_block(c, o_ws + ws_b);
    if (rc) return rc;
    cudaStream_t st = c.lane[0].st;
    memcpy(c.pin, queries, q_b);
    memcpy(c.pin + q_b, tree, t_b);
    CK(cudaMemcpyAsync(c.dblk, c.pin, q_b + t_b, cudaMemcpyHostToDevice, st));
    rc = cbfrrt_knn(n, reinterpret_cast<float*>(c.dblk), reinterpret_cast<float*>(c.dblk + q_b), M, N_tree, k,
                    reinterpret_cast<int32_t*>(c.dblk + o_idx), reinterpret_cast<float*>(c.dblk + o_dist), c.dblk + o_ws, st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CK(cudaMemcpyAsync(c.pin + o_idx, c.dblk + o_idx, o_dist + out_b - o_idx, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(idx, c.pin + o_idx, out_b);
    memcpy(dist, c.pin + o_dist, out_b);
    return CBFRRT_OK;
}

extern "C" int cbfrrt_edge_validity_host(int device, int robot, const cbfrrt_host_scene* scene, const float* q_from,
                                         const float* q_to, int64_t E, const int32_t* n_interp, float thr_soft, float thr_hard,
                                         uint8_t* valid, int32_t* first_bad) {
    int32_t n = 0, nl = 0, ns = 0;
    int rc = cbfrrt_robot_info(robot, &n, &nl, &ns);
    if (rc) return rc;
    if (device < 0 || device >= 16 || E < 1 || !scene || !q_from || !q_to || !n_interp || !valid || !first_bad)
        return CBFRRT_ERR_BAD_SHAPE;
    if (scene->n_boxes < 0 || scene->n_spheres < 0 || (scene->n_boxes && !scene->boxes) || (scene->n_spheres && !scene->spheres))
        return CBFRRT_ERR_BAD_SHAPE;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device >= n_dev) return CBFRRT_ERR_BAD_SHAPE;
    DeviceGuard guard(device);
    Ctx& c = g_ctx[device];
    std::lock_guard<std::mutex> lock(c.mu);
    rc = ensure(c, device, 0, n, scene->n_boxes, scene->n_spheres, 0);
    if (rc) return rc;
    // block: in [offsets int64 (E + 1) | q_from | q_to], out [first_bad int32 E | valid u8 E]
    const size_t off_b = ((size_t)(E + 1) * 8 + 15) & ~size_t(15), rows_b = (size_t)E * n * 4;
    const size_t o_fb = (off_b + 2 * rows_b + 15) & ~size_t(15), o_v = (o_fb + (size_t)E * 4 + 15) & ~size_t(15), total = o_v + (size_t)E;
    rc = ensure_block(c, total);
    if (rc) return rc;
    int64_t* off = reinterpret_cast<int64_t*>(c.pin);
    off[0] = 0;
    for (int64_t e = 0; e < E; ++e) {
        if (n_interp[e] < 0) return CBFRRT_ERR_BAD_SHAPE;
        off[e + 1] = off[e] + n_interp[e] + 1;
    }
    const int64_t n_points = off[E];
    const int n_obs = scene->n_boxes + scene->n_spheres;
    cbfrrt_scene dsc;
    rc = make_resident_scene(c, scene, (n_obs >= 40 && n_points > 4096) || (c.resident_grid && n_obs >= 1), dsc);
    if (rc) return rc;
    cudaStream_t st = c.lane[0].st;
    memcpy(c.pin + off_b, q_from, rows_b);
    memcpy(c.pin + off_b + rows_b, q_to, rows_b);
    CK(cudaMemcpyAsync(c.dblk, c.pin, off_b + 2 * rows_b, cudaMemcpyHostToDevice, st));
    rc = cbfrrt_edge_validity_var(robot, &dsc, reinterpret_cast<float*>(c.dblk + off_b), reinterpret_cast<float*>(c.dblk + off_b + rows_b),
                                  E, reinterpret_cast<int64_t*>(c.dblk), n_points, thr_soft, thr_hard, c.dblk + o_v,
                                  reinterpret_cast<int32_t*>(c.dblk + o_fb), st);
    if (rc) { cudaStreamSynchronize(st); return rc; }
    CK(cudaMemcpyAsync(c.pin + o_fb, c.dblk + o_fb, total - o_fb, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(first_bad, c.pin + o_fb, (size_t)E * 4);
    memcpy(valid, c.pin + o_v, (size_t)E);
    return CBFRRT_OK;
}
