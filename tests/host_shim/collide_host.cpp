// TEST INFRASTRUCTURE -- host build of the device geometry code (csrc/core.cuh, csrc/grid.cuh) through the stand-in
// cuda_runtime.h of this directory.  Exposes the broad-phase builders and collide_state (brute-force, culled and
// distance-bound-field paths) as plain C functions; tests/test_device_code_on_host_cpu.py compares them with the oracle.
#include "../../cbf-inc-rrt_b200/csrc/core.cuh"

using namespace cbfrrt;

template <class RB, bool GRAD, bool GRID>
static void run(const SceneS& sv, const float* q, long B, float ts, float th, unsigned char* safe, unsigned char* unsafe,
                float* o, int* arg_link, int* arg_obs, float* dodq, float* soft_hard, unsigned* stats) {
    float base_min = CBF_INF;
    {
        const float m = base_min_partial<RB>(sv, 0, 1);
        base_min = m;
    }
    for (long b = 0; b < B; ++b) {
        float qq[RB::N];
        for (int k = 0; k < RB::N; ++k) qq[k] = q[b * RB::N + k];
        CollideResult<RB> r;
        collide_state<RB, GRAD, false, GRID>(qq, sv, base_min, r);
        const bool us = r.hard_min <= th;
        safe[b] = r.self_ok && (r.soft_min > ts) && !us;
        unsafe[b] = us;
        o[b] = r.o;
        arg_link[b] = r.arg_sphere >= 0 ? RB::sph_link(r.arg_sphere) : -2;
        arg_obs[b] = r.arg_obs;
        for (int k = 0; k < RB::N; ++k) dodq[b * RB::N + k] = r.dodq[k];
        soft_hard[2 * b] = r.soft_min; soft_hard[2 * b + 1] = r.hard_min;
        stats[2 * b] = r.n_refined; stats[2 * b + 1] = r.n_pairs;
    }
}

extern "C" {

long host_grid_bytes() { return (long)grid::BYTES; }

void host_grid_build(const float* boxes8, int n_boxes, const float* spheres4, int n_spheres, unsigned char* out) {
    const float4* b = reinterpret_cast<const float4*>(boxes8);
    const float4* s = reinterpret_cast<const float4*>(spheres4);
    for (int blk = 0; blk < (grid::NCELLS + 127) / 128; ++blk)
        for (int t = 0; t < 128; ++t) {
            blockIdx.x = blk; threadIdx.x = t;
            grid::build_kernel<0>(b, n_boxes, s, n_spheres, reinterpret_cast<uint16_t*>(out));
        }
    for (int blk = 0; blk < (grid::FNCELLS + 127) / 128; ++blk)
        for (int t = 0; t < 128; ++t) {
            blockIdx.x = blk; threadIdx.x = t;
            grid::field_kernel<0>(b, n_boxes, s, n_spheres, reinterpret_cast<float*>(out + grid::REC_BYTES));
        }
}

// grid == nullptr: brute force / culled sweep;  otherwise the distance-bound field path.  want_grad = 0: mask-only caller.
int host_collide(int robot, const float* boxes8, int n_boxes, const float* spheres4, int n_spheres, int floor,
                 const unsigned char* grid_buf, const float* q, long B, float ts, float th, int want_grad,
                 unsigned char* safe, unsigned char* unsafe, float* o, int* arg_link, int* arg_obs, float* dodq,
                 float* soft_hard, unsigned* stats) {
    SceneS sv;
    sv.boxes = reinterpret_cast<const float4*>(boxes8);
    sv.spheres = reinterpret_cast<const float4*>(spheres4);
    sv.n_boxes = n_boxes; sv.n_spheres = n_spheres; sv.floor = floor;
    sv.grid = reinterpret_cast<const uint16_t*>(grid_buf);
    sv.cull_floor = fmaxf(ts, th); sv.thr_min = fminf(ts, th);
#define RUN(RB)                                                                                                         \
    if (grid_buf) { if (want_grad) run<RB, true, true>(sv, q, B, ts, th, safe, unsafe, o, arg_link, arg_obs, dodq, soft_hard, stats); \
                    else run<RB, false, true>(sv, q, B, ts, th, safe, unsafe, o, arg_link, arg_obs, dodq, soft_hard, stats); }        \
    else { if (want_grad) run<RB, true, false>(sv, q, B, ts, th, safe, unsafe, o, arg_link, arg_obs, dodq, soft_hard, stats);         \
           else run<RB, false, false>(sv, q, B, ts, th, safe, unsafe, o, arg_link, arg_obs, dodq, soft_hard, stats); }
    if (robot == 0) { RUN(Panda) } else if (robot == 1) { RUN(Magician) } else return -1;
#undef RUN
    return 0;
}

}  // extern "C"
