// k_lidar.cu -- C-ABI of the LiDAR / PointNet barrier (include/cbfrrt.h: cbfrrt_lidar_*), kernels in lidar.cuh
#include <atomic>
#include <cstdlib>

#include "lidar.cuh"

namespace cbfrrt {
extern std::atomic<long long> g_launches;
extern std::atomic<int> g_last_cuda;
}  // namespace cbfrrt

using namespace cbfrrt;

static int finish() {
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { g_last_cuda = (int)e; return CBFRRT_ERR_CUDA; }
    g_launches++;
    return CBFRRT_OK;
}

static int check_lidar_net(const cbfrrt_lidar_net* w) {
    if (!w) return CBFRRT_ERR_BAD_SHAPE;
    if (w->n < 1 || w->n > lidar::MAX_N || w->n_sensors < 1 || w->n_sensors > lidar::MAX_S) return CBFRRT_ERR_UNSUPPORTED;
    if (w->point_dims != 3 && w->point_dims != 6) return CBFRRT_ERR_UNSUPPORTED;
    if (w->point_dims + w->n_sensors > lidar::K1) return CBFRRT_ERR_UNSUPPORTED;
    if (w->per_feature_dim < 1 || w->per_feature_dim > 64 || w->feature_dim < 1 || w->n + w->feature_dim > 64 + lidar::MAX_N)
        return CBFRRT_ERR_UNSUPPORTED;
    if (w->hidden < 1 || w->hidden > 512 || w->layers < 0) return CBFRRT_ERR_UNSUPPORTED;
    const void* ptrs[] = {w->w1t, w->w2t, w->w3t, w->bias, w->fc2_w, w->fc2_b, w->fc3_w, w->fc3_b, w->e1_w, w->e1_b, w->e2_w,
                          w->e2_b, w->e4_w, w->e4_b, w->h_in_w, w->h_in_b, w->h_out_w, w->h_out_b};
    for (const void* p : ptrs)
        if (!p) return CBFRRT_ERR_BAD_SHAPE;
    if (w->layers > 0 && (!w->h_hid_w || !w->h_hid_b)) return CBFRRT_ERR_BAD_SHAPE;
    if ((reinterpret_cast<uintptr_t>(w->w1t) | reinterpret_cast<uintptr_t>(w->w2t) | reinterpret_cast<uintptr_t>(w->w3t)) & 15u)
        return CBFRRT_ERR_ALIGNMENT;
    return CBFRRT_OK;
}

extern "C" int64_t cbfrrt_lidar_workspace_bytes(const cbfrrt_lidar_net* net, int64_t B, int32_t with_jacobian) {
    if (check_lidar_net(net) || B < 0) return -1;
    const int64_t evals = B * (with_jacobian ? net->n + 1 : 1);
    return evals * net->n_sensors * lidar::N3 * 4 + ((evals * 4 + 15) & ~15ll) + 16;
}

extern "C" int cbfrrt_lidar_cbf(const cbfrrt_lidar_net* net, const float* q, const float* cloud, int64_t n_cloud,
                                const float* frames, const float* J_p, const float* J_R, const int32_t* sampled_index,
                                int32_t rays, int64_t B, float delta, void* workspace, float* h, float* J, void* stream) {
    int rc = check_lidar_net(net);
    if (rc) return rc;
    if (B < 0 || n_cloud < 1 || rays < 32 || !q || !cloud || !frames || !sampled_index || !workspace) return CBFRRT_ERR_BAD_SHAPE;
    if (rays % 32) return CBFRRT_ERR_UNSUPPORTED;            // a warp's 32 points must belong to one (evaluation, sensor) group
    if (J && (!J_p || !J_R || !(delta > 0.f))) return CBFRRT_ERR_BAD_SHAPE;
    if (reinterpret_cast<uintptr_t>(workspace) & 15u) return CBFRRT_ERR_ALIGNMENT;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    lidar::CloudArgs a{q, cloud, frames, J_p, J_R, sampled_index, (long long)B, (long long)n_cloud, net->n, net->n_sensors, rays,
                       net->point_dims, J ? net->n + 1 : 1, delta};
    const long long n_eval = B * a.evals, n_groups = n_eval * a.S, n_points = n_groups * rays;
    int* keys = static_cast<int*>(workspace);
    float* h_eval = reinterpret_cast<float*>(keys + n_groups * lidar::N3);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long nk = n_groups * lidar::N3;
    lidar::preset_keys<<<(unsigned)((nk + 255) / 256 < 4096 ? (nk + 255) / 256 : 4096), 256, 0, st>>>(keys, nk);
    if ((rc = finish())) return rc;
    static const bool use_v1 = [] { const char* e = getenv("CBFRRT_LIDAR_V1"); return e && e[0] == '1'; }();
    const size_t smem = (use_v1 ? lidar::S_TOTAL : lidar::v2::S_TOTAL) * sizeof(float);
    static std::atomic<unsigned long long> configured{0};   // bit per device
    if (!((configured.load() >> (dev & 63)) & 1ull)) {
        if (cudaFuncSetAttribute(lidar::backbone_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(lidar::S_TOTAL * sizeof(float))) != cudaSuccess ||
            cudaFuncSetAttribute(lidar::v2::backbone_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(lidar::v2::S_TOTAL * sizeof(float))) != cudaSuccess)
            return CBFRRT_ERR_CUDA;
        configured.fetch_or(1ull << (dev & 63));
    }
    const long long tiles = (n_points + lidar::PT - 1) / lidar::PT;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (use_v1)
        lidar::backbone_kernel<<<grid, lidar::PT, smem, st>>>(a, n_points, net->w1t, net->w2t, net->w3t, net->bias, keys);
    else
        lidar::v2::backbone_kernel_v2<<<grid, lidar::PT, smem, st>>>(a, n_points, net->w1t, net->w2t, net->w3t, net->bias, keys);
    if ((rc = finish())) return rc;
    lidar::head_kernel<<<(unsigned)n_eval, 256, 0, st>>>(keys, a, *net, h_eval);
    if ((rc = finish())) return rc;
    lidar::finish_kernel<<<(unsigned)((n_eval + 255) / 256 < 1024 ? (n_eval + 255) / 256 : 1024), 256, 0, st>>>(h_eval, B, net->n,
                                                                                                                 a.evals, delta, h, J);
    return finish();
}
