// abi.cu -- the device-pointer C-ABI of libcbfrrt (include/cbfrrt.h): argument validation + dispatch to the
// launchers.  Kernels live in kernels.cuh; the per-robot instantiations of the big ones are compiled in k_*.cu.
#include "kernels.cuh"
#include "knn.cuh"

namespace cbfrrt {
std::atomic<long long> g_launches{0};
std::atomic<int> g_last_cuda{0};
std::mutex g_cfg_mu;
std::map<CfgKey, int> g_cfg;
static int mlp_path_from_env() {
    const char* e = getenv("CBFRRT_MLP");
    if (e && e[0] == 'w') return CBFRRT_MLP_WARP;
    return (e && e[0] == 'f') ? CBFRRT_MLP_FMA : ((e && e[0] == 't') ? CBFRRT_MLP_TC : CBFRRT_MLP_AUTO);
}
std::atomic<int> g_mlp_path{mlp_path_from_env()};
Workspace g_workspace[64] = {};
std::atomic<int> g_split{[] { const char* e = getenv("CBFRRT_SPLIT"); return (e && e[0] == '0') ? 0 : 1; }()};
}  // namespace cbfrrt

using namespace cbfrrt;

extern "C" {

int cbfrrt_abi_version(void) { return CBFRRT_ABI_VERSION; }

const char* cbfrrt_status_string(int s) {
    switch (s) {
        case CBFRRT_OK: return "ok";
        case CBFRRT_ERR_BAD_ROBOT: return "unknown robot id";
        case CBFRRT_ERR_BAD_SHAPE: return "bad shape or missing pointer";
        case CBFRRT_ERR_UNSUPPORTED: return "unsupported network shape or scene size";
        case CBFRRT_ERR_CUDA: return "CUDA launch/runtime failure";
        case CBFRRT_ERR_ALIGNMENT: return "scene/weight buffer not 16-byte aligned";
        default: return "unknown status";
    }
}

int cbfrrt_last_cuda_error(void) { return g_last_cuda.load(); }

int cbfrrt_set_mlp_path(int path) {
    if (path != CBFRRT_MLP_AUTO && path != CBFRRT_MLP_FMA && path != CBFRRT_MLP_TC && path != CBFRRT_MLP_WARP) return CBFRRT_ERR_BAD_SHAPE;
    g_mlp_path.store(path);
    return CBFRRT_OK;
}
int cbfrrt_get_mlp_path(void) { return g_mlp_path.load(); }

int cbfrrt_set_workspace(void* device_ptr, int64_t bytes) {
    if (bytes < 0 || (bytes > 0 && !device_ptr) || !aligned16(device_ptr)) return CBFRRT_ERR_BAD_SHAPE;
    std::lock_guard<std::mutex> lock(g_cfg_mu);
    g_workspace[current_device() & 63] = Workspace{bytes > 0 ? device_ptr : nullptr, bytes > 0 ? (long long)bytes : 0};
    return CBFRRT_OK;
}
int cbfrrt_set_two_pass(int enabled) {
    g_split.store(enabled ? 1 : 0);
    return CBFRRT_OK;
}
int64_t cbfrrt_launch_count(void) { return g_launches.load(); }

int cbfrrt_robot_info(int robot, int32_t* n, int32_t* n_links, int32_t* n_spheres) {
    if (robot == CBFRRT_ROBOT_PANDA) { *n = Panda::N; *n_links = Panda::NLINKS; *n_spheres = Panda::NSPH; return CBFRRT_OK; }
    if (robot == CBFRRT_ROBOT_MAGICIAN) { *n = Magician::N; *n_links = Magician::NLINKS; *n_spheres = Magician::NSPH; return CBFRRT_OK; }
    return CBFRRT_ERR_BAD_ROBOT;
}

int64_t cbfrrt_grid_bytes(void) { return (int64_t)grid::BYTES; }

int cbfrrt_grid_build(const cbfrrt_scene* scene, void* grid_out, void* stream) {
    SceneArgs sc;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    if (!grid_out) return CBFRRT_ERR_BAD_SHAPE;
    if (!aligned16(grid_out)) return CBFRRT_ERR_ALIGNMENT;
    grid::build_kernel<0><<<(grid::NCELLS + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(sc.boxes), sc.n_boxes, reinterpret_cast<const float4*>(sc.spheres), sc.n_spheres,
        static_cast<uint16_t*>(grid_out));
    if (finish_launch()) return CBFRRT_ERR_CUDA;
    grid::field_kernel<0><<<(grid::FNCELLS + 127) / 128, 128, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(sc.boxes), sc.n_boxes, reinterpret_cast<const float4*>(sc.spheres), sc.n_spheres,
        reinterpret_cast<float*>(static_cast<unsigned char*>(grid_out) + grid::REC_BYTES));
    return finish_launch();
}

int cbfrrt_fk(int robot, const float* q, int64_t q_stride, int64_t B, float* pos, float* rot, void* stream) {
    if (B < 0 || !q || !pos || !rot) return CBFRRT_ERR_BAD_SHAPE;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (robot == CBFRRT_ROBOT_PANDA) return q_stride < Panda::N ? CBFRRT_ERR_BAD_SHAPE : launch_fk<Panda>(q, q_stride, B, pos, rot, st);
    if (robot == CBFRRT_ROBOT_MAGICIAN) return q_stride < Magician::N ? CBFRRT_ERR_BAD_SHAPE : launch_fk<Magician>(q, q_stride, B, pos, rot, st);
    return CBFRRT_ERR_BAD_ROBOT;
}

int cbfrrt_collide(int robot, const cbfrrt_scene* scene, const float* q, int64_t q_stride, int64_t B, float thr_soft,
                   float thr_hard, uint8_t* safe, uint8_t* unsafe, float* datax, int32_t* arg_link, int32_t* arg_obs,
                   float* mins, void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (B < 0 || !q) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    if (B == 0) return CBFRRT_OK;
    CollideOutPtrs out{safe, unsafe, datax, arg_link, arg_obs, mins};
    cudaStream_t st = (cudaStream_t)stream;
    if (robot == CBFRRT_ROBOT_PANDA) return q_stride < Panda::N ? CBFRRT_ERR_BAD_SHAPE : launch_collide_panda(sc, q, q_stride, B, thr_soft, thr_hard, out, st);
    return q_stride < Magician::N ? CBFRRT_ERR_BAD_SHAPE : launch_collide_magician(sc, q, q_stride, B, thr_soft, thr_hard, out, st);
}

int cbfrrt_cbf_filter(int n, const cbfrrt_cbf_net* net, const float* datax, int64_t B, const cbfrrt_filter_params* fp,
                      float* h, float* J, float* u, float* r, void* stream) {
    if (B < 0 || !datax) return CBFRRT_ERR_BAD_SHAPE;
    if (n != Panda::N && n != Magician::N) return CBFRRT_ERR_UNSUPPORTED;
    NetArgs na;
    int rc = check_net(n, net, fp, B, na);
    if (rc) return rc;
    if (B == 0) return CBFRRT_OK;
    FilterOutPtrs out{h, J, u, r};
    cudaStream_t st = (cudaStream_t)stream;
    if (!net_is_fast(net))
        return n == Panda::N ? launch_filter_generic<Panda::N>(na, net->hidden, net->layers, datax, B, out, st)
                             : launch_filter_generic<Magician::N>(na, net->hidden, net->layers, datax, B, out, st);
    return n == Panda::N ? launch_filter<Panda::N>(na, datax, B, out, st) : launch_filter<Magician::N>(na, datax, B, out, st);
}

static int peers_from(const cbfrrt_peer_out* peers, int64_t B, PeerOut& po) {
    if (!peers || peers->n_peers < 1 || peers->n_peers > CBFRRT_MAX_PEERS || peers->row_offset < 0) return CBFRRT_ERR_BAD_SHAPE;
    po = PeerOut{};
    po.n = peers->n_peers;
    po.multicast = peers->multicast ? 1 : 0;
    po.row_offset = peers->row_offset;
    // multicast mode packs the mask bytes of 32 consecutive rows into words: whole, aligned 128-row tiles only
    if (po.multicast && (po.n != 1 || B % 128 != 0 || peers->row_offset % 128 != 0)) return CBFRRT_ERR_BAD_SHAPE;
    for (int p = 0; p < po.n; ++p) {
        if (!peers->valid[p] || !peers->u[p]) return CBFRRT_ERR_BAD_SHAPE;
        if (!aligned16(peers->u[p])) return CBFRRT_ERR_ALIGNMENT;
        po.valid[p] = peers->valid[p];
        po.u[p] = peers->u[p];
    }
    return CBFRRT_OK;
}

int cbfrrt_cbf_filter_scatter(int n, const cbfrrt_cbf_net* net, const float* datax, const uint8_t* valid_in, int64_t B,
                              const cbfrrt_filter_params* fp, const cbfrrt_peer_out* peers, void* stream) {
    if (B < 0 || !datax || !valid_in) return CBFRRT_ERR_BAD_SHAPE;
    if (n != Panda::N && n != Magician::N) return CBFRRT_ERR_UNSUPPORTED;
    PeerOut po;
    int rc = peers_from(peers, B, po);
    if (rc) return rc;
    NetArgs na;
    rc = check_net(n, net, fp, B, na);
    if (rc) return rc;
    if (!net_is_fast(net)) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    FilterOutPtrs fo{nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    return n == Panda::N ? launch_filter_peers<Panda::N>(na, datax, valid_in, B, fo, po, st)
                         : launch_filter_peers<Magician::N>(na, datax, valid_in, B, fo, po, st);
}

int cbfrrt_qp(int n, const float* a, const float* c, const float* u_ref, int64_t B, const float* u_lo, const float* u_hi,
              float relaxation_penalty, float* u, float* r, void* stream) {
    if (B < 0 || !a || !c || !u_ref || !u_lo || !u_hi || !u) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
    switch (n) {
#define CBF_QP_CASE(NN) case NN: qp_kernel<NN><<<grid, 256, 0, st>>>(a, c, u_ref, B, u_lo, u_hi, relaxation_penalty, u, r); break;
        CBF_QP_CASE(1) CBF_QP_CASE(2) CBF_QP_CASE(3) CBF_QP_CASE(4) CBF_QP_CASE(5) CBF_QP_CASE(6) CBF_QP_CASE(7) CBF_QP_CASE(8)
#undef CBF_QP_CASE
    }
    return finish_launch();
}

int cbfrrt_qp_backward(int n, const float* a, const float* c, const float* u_ref, int64_t B, const float* u_lo,
                       const float* u_hi, float relaxation_penalty, const float* grad_u, const float* grad_r,
                       float* grad_a, float* grad_c, float* grad_u_ref, void* stream) {
    if (B < 0 || !u_lo || !u_hi) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    if (!a || !c || !u_ref || !grad_u) return CBFRRT_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 8 ? blocks : (long long)num_sms() * 8);
    switch (n) {
#define CBF_QP_CASE(NN) case NN: qp_backward_kernel<NN><<<grid, 256, 0, st>>>(a, c, u_ref, B, u_lo, u_hi, relaxation_penalty, grad_u, grad_r, grad_a, grad_c, grad_u_ref); break;
        CBF_QP_CASE(1) CBF_QP_CASE(2) CBF_QP_CASE(3) CBF_QP_CASE(4) CBF_QP_CASE(5) CBF_QP_CASE(6) CBF_QP_CASE(7) CBF_QP_CASE(8)
#undef CBF_QP_CASE
    }
    return finish_launch();
}

int cbfrrt_qp_multi(int n, int n_constraints, const float* a, const float* c, const float* u_ref, int64_t B,
                    const float* u_lo, const float* u_hi, float relaxation_penalty, int max_sweeps, float tol, float* u,
                    float* r, int32_t* sweeps, void* stream) {
    if (B < 0 || !a || !c || !u_ref || !u_lo || !u_hi || !u || max_sweeps < 1 || !(tol >= 0.f)) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8 || n_constraints < 1 || n_constraints > QP_MAX_CON) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B + 127) / 128;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    switch (n) {
#define CBF_QP_CASE(NN) case NN: qp_multi_kernel<NN><<<grid, 128, 0, st>>>(a, c, u_ref, B, n_constraints, u_lo, u_hi, relaxation_penalty, max_sweeps, tol, u, r, sweeps); break;
        CBF_QP_CASE(1) CBF_QP_CASE(2) CBF_QP_CASE(3) CBF_QP_CASE(4) CBF_QP_CASE(5) CBF_QP_CASE(6) CBF_QP_CASE(7) CBF_QP_CASE(8)
#undef CBF_QP_CASE
    }
    return finish_launch();
}

int cbfrrt_qp_multi_backward(int n, int n_constraints, const float* a, const float* c, const float* u_ref, int64_t B,
                             const float* u_lo, const float* u_hi, float relaxation_penalty, int max_iter, float tol,
                             const float* grad_u, const float* grad_r, float* grad_a, float* grad_c, float* grad_u_ref,
                             void* stream) {
    if (B < 0 || !u_lo || !u_hi || max_iter < 1 || !(tol >= 0.f)) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8 || n_constraints < 1 || n_constraints > QP_MAX_CON) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    if (!a || !c || !u_ref || !grad_u) return CBFRRT_ERR_BAD_SHAPE;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B + 127) / 128;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    switch (n) {
#define CBF_QP_CASE(NN) case NN: qp_multi_backward_kernel<NN><<<grid, 128, 0, st>>>(a, c, u_ref, B, n_constraints, u_lo, u_hi, relaxation_penalty, max_iter, tol, grad_u, grad_r, grad_a, grad_c, grad_u_ref); break;
        CBF_QP_CASE(1) CBF_QP_CASE(2) CBF_QP_CASE(3) CBF_QP_CASE(4) CBF_QP_CASE(5) CBF_QP_CASE(6) CBF_QP_CASE(7) CBF_QP_CASE(8)
#undef CBF_QP_CASE
    }
    return finish_launch();
}

int cbfrrt_safety_filter(int robot, const cbfrrt_scene* scene, const cbfrrt_cbf_net* net, const float* q,
                         int64_t q_stride, int64_t B, const cbfrrt_filter_params* fp, float thr_soft, float thr_hard,
                         uint8_t* safe, uint8_t* unsafe, float* datax, float* h, float* J, float* u, float* r,
                         void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (B < 0 || !q) return CBFRRT_ERR_BAD_SHAPE;
    const int n = robot == CBFRRT_ROBOT_PANDA ? Panda::N : Magician::N;
    if (q_stride < n) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    NetArgs na;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    rc = check_net(n, net, fp, B, na);
    if (rc) return rc;
    if (B == 0) return CBFRRT_OK;
    CollideOutPtrs co{safe, unsafe, datax, nullptr, nullptr, nullptr};
    FilterOutPtrs fo{h, J, u, r};
    PeerOut po{};
    cudaStream_t st = (cudaStream_t)stream;
    if (!net_is_fast(net)) {
        // other network shapes: two launches (observation, then the generic filter on it).  The library never
        // allocates, so the observation rows need the caller's datax buffer.
        if (!datax) return CBFRRT_ERR_UNSUPPORTED;
        rc = robot == CBFRRT_ROBOT_PANDA ? launch_collide_panda(sc, q, q_stride, B, thr_soft, thr_hard, co, st)
                                         : launch_collide_magician(sc, q, q_stride, B, thr_soft, thr_hard, co, st);
        if (rc) return rc;
        return n == Panda::N ? launch_filter_generic<Panda::N>(na, net->hidden, net->layers, datax, B, fo, st)
                             : launch_filter_generic<Magician::N>(na, net->hidden, net->layers, datax, B, fo, st);
    }
    return robot == CBFRRT_ROBOT_PANDA ? launch_safety_panda(sc, na, q, q_stride, B, thr_soft, thr_hard, co, fo, po, st)
                                       : launch_safety_magician(sc, na, q, q_stride, B, thr_soft, thr_hard, co, fo, po, st);
}

int cbfrrt_safety_filter_scatter(int robot, const cbfrrt_scene* scene, const cbfrrt_cbf_net* net, const float* q,
                                 int64_t q_stride, int64_t B, const cbfrrt_filter_params* fp, float thr_soft,
                                 float thr_hard, const cbfrrt_peer_out* peers, void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (B < 0 || !q || !peers || peers->n_peers < 1 || peers->n_peers > CBFRRT_MAX_PEERS || peers->row_offset < 0)
        return CBFRRT_ERR_BAD_SHAPE;
    const int n = robot == CBFRRT_ROBOT_PANDA ? Panda::N : Magician::N;
    if (q_stride < n) return CBFRRT_ERR_BAD_SHAPE;
    PeerOut po{};
    po.n = peers->n_peers;
    po.multicast = peers->multicast ? 1 : 0;
    po.row_offset = peers->row_offset;
    // multicast mode packs the mask bytes of 32 consecutive rows into words: whole, aligned 128-row tiles only
    if (po.multicast && (po.n != 1 || B % 128 != 0 || peers->row_offset % 128 != 0)) return CBFRRT_ERR_BAD_SHAPE;
    for (int p = 0; p < po.n; ++p) {
        if (!peers->valid[p] || !peers->u[p]) return CBFRRT_ERR_BAD_SHAPE;
        if (!aligned16(peers->u[p])) return CBFRRT_ERR_ALIGNMENT;
        po.valid[p] = peers->valid[p];
        po.u[p] = peers->u[p];
    }
    SceneArgs sc;
    NetArgs na;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    rc = check_net(n, net, fp, B, na);
    if (rc) return rc;
    if (!net_is_fast(net)) return CBFRRT_ERR_UNSUPPORTED;
    if (B == 0) return CBFRRT_OK;
    CollideOutPtrs co{nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    FilterOutPtrs fo{nullptr, nullptr, nullptr, nullptr};
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA ? launch_safety_panda(sc, na, q, q_stride, B, thr_soft, thr_hard, co, fo, po, st)
                                       : launch_safety_magician(sc, na, q, q_stride, B, thr_soft, thr_hard, co, fo, po, st);
}

int cbfrrt_edge_validity(int robot, const cbfrrt_scene* scene, const float* q_from, const float* q_to, int64_t E,
                         int32_t n_interp, float thr_soft, float thr_hard, uint8_t* valid, int32_t* first_bad,
                         void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (E < 0 || n_interp < 1 || !q_from || !q_to || !valid || !first_bad) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    if (E == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA ? launch_edges_panda(sc, q_from, q_to, E, n_interp, thr_soft, thr_hard, valid, first_bad, st)
                                       : launch_edges_magician(sc, q_from, q_to, E, n_interp, thr_soft, thr_hard, valid, first_bad, st);
}

int cbfrrt_edge_validity_var(int robot, const cbfrrt_scene* scene, const float* q_from, const float* q_to, int64_t E,
                             const int64_t* point_offsets, int64_t n_points, float thr_soft, float thr_hard,
                             uint8_t* valid, int32_t* first_bad, void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (E < 0 || n_points < E || !q_from || !q_to || !point_offsets || !valid || !first_bad) return CBFRRT_ERR_BAD_SHAPE;
    SceneArgs sc;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    if (E == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long* off = reinterpret_cast<const long long*>(point_offsets);
    return robot == CBFRRT_ROBOT_PANDA ? launch_edges_var_panda(sc, q_from, q_to, E, off, n_points, thr_soft, thr_hard, valid, first_bad, st)
                                       : launch_edges_var_magician(sc, q_from, q_to, E, off, n_points, thr_soft, thr_hard, valid, first_bad, st);
}

int64_t cbfrrt_knn_workspace_bytes(int64_t M, int64_t N_tree, int32_t k) {
    if (M < 0 || N_tree < 0 || k < 1) return -1;
    int64_t bytes = 16 * (M > 0 ? M : 1);
    if (k > 1 && k <= knn::TOPK_MAXK && N_tree > 0)
        bytes += 8 * M * ((N_tree + knn::TOPK_CHUNK - 1) / knn::TOPK_CHUNK) * (int64_t)k;
    return bytes;
}

int cbfrrt_knn(int n, const float* queries, const float* tree, int64_t M, int64_t N_tree, int32_t k, int32_t* idx,
               float* dist, void* workspace, void* stream) {
    if (M < 0 || N_tree < 0 || k < 1 || !queries || !tree || !idx || !dist || !workspace) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8 || N_tree > 0x7fffffffll) return CBFRRT_ERR_UNSUPPORTED;
    if (M == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (k > 1 && k <= knn::TOPK_MAXK && N_tree > 0 && M <= 65535) {   // two launches: per-chunk candidates, per-query merge
        unsigned long long* cand = reinterpret_cast<unsigned long long*>(workspace) + 2 * M;
        const long long n_chunks = (N_tree + knn::TOPK_CHUNK - 1) / knn::TOPK_CHUNK;
        const dim3 grid((unsigned)n_chunks, (unsigned)M);
        switch (n) {
#define CBF_NN_CASE(NN) case NN: knn::topk_chunk_kernel<NN><<<grid, 256, 0, st>>>(queries, tree, N_tree, k, cand); break;
            CBF_NN_CASE(1) CBF_NN_CASE(2) CBF_NN_CASE(3) CBF_NN_CASE(4) CBF_NN_CASE(5) CBF_NN_CASE(6) CBF_NN_CASE(7) CBF_NN_CASE(8)
#undef CBF_NN_CASE
        }
        if (finish_launch()) return CBFRRT_ERR_CUDA;
        knn::topk_merge_kernel<<<(unsigned)M, 1024, 0, st>>>(cand, n_chunks * k, k, idx, dist);
        return finish_launch();
    }
    unsigned long long* best = reinterpret_cast<unsigned long long*>(workspace);
    unsigned long long* lower = best + M;
    const dim3 grid((unsigned)((N_tree + knn::CHUNK - 1) / knn::CHUNK), (unsigned)((M + knn::QT - 1) / knn::QT));
    for (int r = 0; r < k; ++r) {
        const cudaError_t me = cudaMemsetAsync(best, 0xff, (size_t)M * 8, st);
        if (me != cudaSuccess) { g_last_cuda = (int)me; return CBFRRT_ERR_CUDA; }
        if (N_tree > 0) {
            const unsigned long long* lo = r ? lower : nullptr;
            switch (n) {
#define CBF_NN_CASE(NN) case NN: knn::nn_kernel<NN><<<grid, knn::QT, 0, st>>>(queries, tree, M, N_tree, lo, best); break;
                CBF_NN_CASE(1) CBF_NN_CASE(2) CBF_NN_CASE(3) CBF_NN_CASE(4) CBF_NN_CASE(5) CBF_NN_CASE(6) CBF_NN_CASE(7) CBF_NN_CASE(8)
#undef CBF_NN_CASE
            }
            if (finish_launch()) return CBFRRT_ERR_CUDA;
        }
        knn::nn_finish<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(best, lower, M, idx + r, dist + r, k);
        if (finish_launch()) return CBFRRT_ERR_CUDA;
    }
    return CBFRRT_OK;
}

int cbfrrt_radius_mask(int n, const float* queries, const float* set, int64_t M, int64_t N_set, float radius,
                       uint8_t* within, void* stream) {
    if (M < 0 || N_set < 0 || !queries || !set || !within) return CBFRRT_ERR_BAD_SHAPE;
    if (n < 1 || n > 8) return CBFRRT_ERR_UNSUPPORTED;
    if (M == 0 || N_set == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (M * N_set + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    switch (n) {
#define CBF_RAD_CASE(NN) case NN: knn::radius_kernel<NN><<<grid, 256, 0, st>>>(queries, set, M, N_set, radius, within); break;
        CBF_RAD_CASE(1) CBF_RAD_CASE(2) CBF_RAD_CASE(3) CBF_RAD_CASE(4) CBF_RAD_CASE(5) CBF_RAD_CASE(6) CBF_RAD_CASE(7) CBF_RAD_CASE(8)
#undef CBF_RAD_CASE
    }
    return finish_launch();
}

int cbfrrt_rollout(int robot, const cbfrrt_scene* scene, const cbfrrt_cbf_net* net, const float* q0, int64_t T,
                   const cbfrrt_filter_params* fp, const cbfrrt_rollout_params* rp, float thr_soft, float thr_hard,
                   float* q_final, uint8_t* alive, int32_t* steps_done, float* traj, void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (T < 0 || !rp || rp->steps < 0 || rp->ctrl_every < 1 || !q0 || !q_final || !alive || !steps_done)
        return CBFRRT_ERR_BAD_SHAPE;
    if (rp->stuck_lag < 0 || (rp->stuck_lag > 0 && (!traj || rp->stuck_after < rp->stuck_lag))) return CBFRRT_ERR_BAD_SHAPE;
    if (rp->stuck_mode != 0 && rp->stuck_mode != 1) return CBFRRT_ERR_BAD_SHAPE;
    if (fp && fp->u_ref) return CBFRRT_ERR_UNSUPPORTED;  // the loop needs the nominal controller
    const int n = robot == CBFRRT_ROBOT_PANDA ? Panda::N : Magician::N;
    SceneArgs sc;
    NetArgs na;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    rc = check_net(n, net, fp, T, na);
    if (rc) return rc;
    if (!net_is_fast(net)) return CBFRRT_ERR_UNSUPPORTED;   // the fused steer loop is built for 48 x 2; other shapes: the
                                                            // host loop over cbfrrt_collide + cbfrrt_cbf_filter (steer())
    if (T == 0) return CBFRRT_OK;
    const RolloutArgs ra{rp->steps, rp->ctrl_every, rp->dt, rp->stuck_lag, rp->stuck_after, rp->stuck_eps, rp->keep_going, rp->stuck_mode};
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA
               ? launch_rollout_panda(sc, na, q0, T, ra, thr_soft, thr_hard, q_final, alive, steps_done, traj, st)
               : launch_rollout_magician(sc, na, q0, T, ra, thr_soft, thr_hard, q_final, alive, steps_done, traj, st);
}

int cbfrrt_rollout_segments(int robot, const cbfrrt_scene* scene, const cbfrrt_cbf_net* net, const float* q0, int64_t T,
                            const cbfrrt_filter_params* fp, const cbfrrt_rollout_params* rp, int32_t n_segments,
                            float thr_soft, float thr_hard, float* seg_q, uint8_t* seg_alive, int32_t* seg_done, float* traj,
                            void* stream) {
    if (robot != CBFRRT_ROBOT_PANDA && robot != CBFRRT_ROBOT_MAGICIAN) return CBFRRT_ERR_BAD_ROBOT;
    if (T < 0 || !rp || rp->steps < 1 || rp->ctrl_every < 1 || n_segments < 1 || !q0 || !seg_q || !seg_alive || !seg_done || !traj)
        return CBFRRT_ERR_BAD_SHAPE;
    if (rp->stuck_lag < 0 || (rp->stuck_lag > 0 && rp->stuck_after < rp->stuck_lag)) return CBFRRT_ERR_BAD_SHAPE;
    if (T > NT) return CBFRRT_ERR_UNSUPPORTED;            // the rows of a batch share one CTA (batch-wide early exit)
    if (fp && fp->u_ref) return CBFRRT_ERR_UNSUPPORTED;
    const int n = robot == CBFRRT_ROBOT_PANDA ? Panda::N : Magician::N;
    SceneArgs sc;
    NetArgs na;
    int rc = check_scene(scene, sc);
    if (rc) return rc;
    rc = check_net(n, net, fp, T, na);
    if (rc) return rc;
    if (!net_is_fast(net)) return CBFRRT_ERR_UNSUPPORTED;
    if (T == 0) return CBFRRT_OK;
    const RolloutArgs ra{rp->steps, rp->ctrl_every, rp->dt, rp->stuck_lag, rp->stuck_after, rp->stuck_eps, 1, 1};
    cudaStream_t st = (cudaStream_t)stream;
    return robot == CBFRRT_ROBOT_PANDA
               ? launch_rollout_segments_panda(sc, na, q0, (int)T, n_segments, rp->steps, ra, thr_soft, thr_hard, seg_q, seg_alive, seg_done, traj, st)
               : launch_rollout_segments_magician(sc, na, q0, (int)T, n_segments, rp->steps, ra, thr_soft, thr_hard, seg_q, seg_alive, seg_done, traj, st);
}

int cbfrrt_fma_probe(int32_t grid, int32_t block, int32_t iters, float* sink, void* stream) {
    if (grid < 1 || block < 32 || block > 1024 || iters < 1 || !sink) return CBFRRT_ERR_BAD_SHAPE;
    fma_probe_kernel<0><<<grid, block, 0, (cudaStream_t)stream>>>(iters, sink);
    return finish_launch();
}

int cbfrrt_sample_states(int robot, uint32_t seed, uint64_t start_row, int64_t B, float* q, void* stream) {
    if (B < 0 || !q) return CBFRRT_ERR_BAD_SHAPE;
    if (B == 0) return CBFRRT_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const long long blocks = (B * 8 + 255) / 256;
    const int grid = (int)(blocks < (long long)num_sms() * 16 ? blocks : (long long)num_sms() * 16);
    if (robot == CBFRRT_ROBOT_PANDA) sample_kernel<Panda><<<grid, 256, 0, st>>>(seed, start_row, B, q);
    else if (robot == CBFRRT_ROBOT_MAGICIAN) sample_kernel<Magician><<<grid, 256, 0, st>>>(seed, start_row, B, q);
    else return CBFRRT_ERR_BAD_ROBOT;
    return finish_launch();
}

}  // extern "C"
