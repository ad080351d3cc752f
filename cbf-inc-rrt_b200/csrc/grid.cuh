// grid.cuh -- nearest-candidate grid: an exact broad phase for the sphere-vs-scene distance sweep.
//
// For a link sphere the sweep needs  min_j e_j(p)  over all obstacles j at the sphere centre p (the same radius is
// subtracted from every e_j afterwards), plus the lowest index attaining it.  For every cell of a uniform grid the
// build kernel stores the obstacles that can attain that minimum for SOME point of the cell:
//     L_j = min_{p in cell} e_j(p),   U_j = max_{p in cell} e_j(p),   candidates = { j : L_j <= min_k U_k + slack }.
// e_j (signed distance to a convex set) is convex, so U_j is attained at a cell corner; L_j is the AABB-AABB gap
// (boxes) or the AABB-point gap minus the radius (spheres); a cell that overlaps a box has L_j = -inf.  Any obstacle
// outside the list has e_j(p) > min_k e_k(p) for every p of the cell, so sweeping only the candidates -- in
// ascending index order, with the exact per-pair formulas -- returns bit-identical minima and arg-min indices to the
// brute-force sweep (and to the oracle, which stays brute force).  slack (1e-4 m) covers fp32 rounding of the
// bounds and of the cell index at cell borders (every e_j is 1-Lipschitz).  Points outside the grid region and
// cells with more than MAXC candidates fall back to the brute-force sweep.
//
// With 5 cm cells the BASELINE scenes keep 1.6 (16 obstacles) / 2.2 (64) / 3.6 (256) candidates per cell on average
// instead of 16 / 64 / 256 pair evaluations per sphere.
//
// DISTANCE-BOUND FIELD (second part of the same buffer).  The callers of the sweep never need the per-sphere minima,
// only  M* = min over the link spheres of (min_j e_j(centre) - radius)  (+ which sphere attains it) and its position
// relative to the mask thresholds.  f(p) = min_j e_j(p) is 1-Lipschitz, so one stored sample per 2.5 cm cell,
// F[c] = f(cell centre), bounds it everywhere in the cell:  |f(p) - F[c]| <= half diagonal (+ slack for fp32 rounding).
// A state therefore needs ONE 4-byte look-up per link sphere to get [lo_i, hi_i] for every sphere, and the exact
// candidate sweep only for the spheres whose interval can still contain the minimum (lo_i <= min_k hi_k) -- typically
// 3-6 of 32 -- or can still change a mask.  Skipped spheres are provably strictly above the minimum, so minima,
// arg-min indices (first index on ties) and booleans stay bit-identical to the brute-force oracle (core.cuh:
// collide_state_field).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cbfrrt {
namespace grid {

// region covers everything both arms can reach (Panda: shoulder 0.333 m + 0.86 m reach + hand; Magician x2 scale)
constexpr float OX = -1.1f, OY = -1.1f, OZ = -0.8f;
constexpr float CELL = 0.05f, INV_CELL = 20.0f;
constexpr int NX = 44, NY = 44, NZ = 44;
constexpr int NCELLS = NX * NY * NZ;
constexpr int MAXC = 22;                    // candidates kept per cell
constexpr int STRIDE = 24;                  // u16 per cell: [n_box_cand, n_sph_cand, cand 0 .. 21]; 48 B = 3 x 16 B
constexpr int OVERFLOW = 0xffff;            // n_box_cand == OVERFLOW: brute force for this cell
constexpr float SLACK = 1e-4f;
constexpr size_t REC_BYTES = (size_t)NCELLS * STRIDE * sizeof(uint16_t);      // 4 088 832, a multiple of 16
// distance-bound field: 2.5 cm cells over the same region
constexpr float FCELL = 0.025f, FINV_CELL = 40.0f;
constexpr int FNX = 88, FNY = 88, FNZ = 88;
constexpr int FNCELLS = FNX * FNY * FNZ;
constexpr float FIELD_HALF_DIAG = 0.0216507f;   // 0.5 * sqrt(3) * FCELL, rounded up
constexpr size_t BYTES = REC_BYTES + (size_t)FNCELLS * sizeof(float);

__device__ __forceinline__ float sd_box_pt(float px, float py, float pz, float4 b0, float4 b1) {
    const float ax = fabsf(px - b0.x) - b0.w, ay = fabsf(py - b0.y) - b1.x, az = fabsf(pz - b0.z) - b1.y;
    const float qx = fmaxf(ax, 0.f), qy = fmaxf(ay, 0.f), qz = fmaxf(az, 0.f);
    return sqrtf(qx * qx + qy * qy + qz * qz) + fminf(fmaxf(ax, fmaxf(ay, az)), 0.f);
}

// one thread per cell; boxes / spheres are read from global memory (built once per scene)
template <int = 0>
__global__ void __launch_bounds__(128) build_kernel(const float4* __restrict__ boxes, int n_boxes,
                                                    const float4* __restrict__ spheres, int n_spheres,
                                                    uint16_t* __restrict__ out) {
    const int cell = blockIdx.x * 128 + threadIdx.x;
    if (cell >= NCELLS) return;
    const int ix = cell % NX, iy = (cell / NX) % NY, iz = cell / (NX * NY);
    const float hx = 0.5f * CELL;
    const float cx = OX + (ix + 0.5f) * CELL, cy = OY + (iy + 0.5f) * CELL, cz = OZ + (iz + 0.5f) * CELL;
    // pass 1: smallest upper bound
    float umin = __int_as_float(0x7f800000);
    for (int j = 0; j < n_boxes; ++j) {
        const float4 b0 = __ldg(boxes + 2 * j), b1 = __ldg(boxes + 2 * j + 1);
        float u = -__int_as_float(0x7f800000);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            u = fmaxf(u, sd_box_pt(cx + ((c & 1) ? hx : -hx), cy + ((c & 2) ? hx : -hx), cz + ((c & 4) ? hx : -hx), b0, b1));
        umin = fminf(umin, u);
    }
    for (int j = 0; j < n_spheres; ++j) {
        const float4 s = __ldg(spheres + j);
        const float dx = fabsf(cx - s.x) + hx, dy = fabsf(cy - s.y) + hx, dz = fabsf(cz - s.z) + hx;  // farthest corner
        umin = fminf(umin, sqrtf(dx * dx + dy * dy + dz * dz) - s.w);
    }
    // pass 2: candidates in ascending index order
    uint16_t* o = out + (size_t)cell * STRIDE;
    int nb = 0, ns = 0;
    bool overflow = false;
    for (int j = 0; j < n_boxes && !overflow; ++j) {
        const float4 b0 = __ldg(boxes + 2 * j), b1 = __ldg(boxes + 2 * j + 1);
        const float gx = fabsf(cx - b0.x) - (hx + b0.w), gy = fabsf(cy - b0.y) - (hx + b1.x), gz = fabsf(cz - b0.z) - (hx + b1.y);
        const float qx = fmaxf(gx, 0.f), qy = fmaxf(gy, 0.f), qz = fmaxf(gz, 0.f);
        const bool overlap = gx <= 0.f && gy <= 0.f && gz <= 0.f;   // e_j can be arbitrarily negative inside: always kept
        const float l = sqrtf(qx * qx + qy * qy + qz * qz);
        if (overlap || l <= umin + SLACK) {
            if (nb + ns >= MAXC) overflow = true;
            else o[2 + nb++] = (uint16_t)j;
        }
    }
    for (int j = 0; j < n_spheres && !overflow; ++j) {
        const float4 s = __ldg(spheres + j);
        const float qx = fmaxf(fabsf(cx - s.x) - hx, 0.f), qy = fmaxf(fabsf(cy - s.y) - hx, 0.f), qz = fmaxf(fabsf(cz - s.z) - hx, 0.f);
        const float l = sqrtf(qx * qx + qy * qy + qz * qz) - s.w;
        if (l <= umin + SLACK) {
            if (nb + ns >= MAXC) overflow = true;
            else o[2 + nb + ns++] = (uint16_t)j;
        }
    }
    o[0] = overflow ? (uint16_t)OVERFLOW : (uint16_t)nb;
    o[1] = (uint16_t)ns;
}

// F[c] = min over ALL obstacles (inside or outside the region) of the signed distance at the centre of field cell c
template <int = 0>
__global__ void __launch_bounds__(128) field_kernel(const float4* __restrict__ boxes, int n_boxes,
                                                    const float4* __restrict__ spheres, int n_spheres,
                                                    float* __restrict__ out) {
    const int cell = blockIdx.x * 128 + threadIdx.x;
    if (cell >= FNCELLS) return;
    const int ix = cell % FNX, iy = (cell / FNX) % FNY, iz = cell / (FNX * FNY);
    const float cx = OX + (ix + 0.5f) * FCELL, cy = OY + (iy + 0.5f) * FCELL, cz = OZ + (iz + 0.5f) * FCELL;
    float f = __int_as_float(0x7f800000);
    for (int j = 0; j < n_boxes; ++j) f = fminf(f, sd_box_pt(cx, cy, cz, __ldg(boxes + 2 * j), __ldg(boxes + 2 * j + 1)));
    for (int j = 0; j < n_spheres; ++j) {
        const float4 s = __ldg(spheres + j);
        const float dx = cx - s.x, dy = cy - s.y, dz = cz - s.z;
        f = fminf(f, sqrtf(dx * dx + dy * dy + dz * dz) - s.w);
    }
    out[cell] = f;
}

__device__ __forceinline__ const float* field_of(const uint16_t* __restrict__ g) {
    return reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(g) + REC_BYTES);
}
// index of the field cell of the point, or -1 when the point lies outside the region
__device__ __forceinline__ int field_index(float px, float py, float pz) {
    const int ix = __float2int_rd(fmaf(px, FINV_CELL, -OX * FINV_CELL)), iy = __float2int_rd(fmaf(py, FINV_CELL, -OY * FINV_CELL)),
              iz = __float2int_rd(fmaf(pz, FINV_CELL, -OZ * FINV_CELL));
    const bool in = ((unsigned)ix < (unsigned)FNX) & ((unsigned)iy < (unsigned)FNY) & ((unsigned)iz < (unsigned)FNZ);
    return in ? (iz * FNY + iy) * FNX + ix : -1;
}

// cell record of the point, or nullptr when the point lies outside the region
__device__ __forceinline__ const uint16_t* lookup(const uint16_t* __restrict__ g, float px, float py, float pz) {
    const int ix = __float2int_rd((px - OX) * INV_CELL), iy = __float2int_rd((py - OY) * INV_CELL),
              iz = __float2int_rd((pz - OZ) * INV_CELL);
    if ((unsigned)ix >= (unsigned)NX || (unsigned)iy >= (unsigned)NY || (unsigned)iz >= (unsigned)NZ) return nullptr;
    return g + (size_t)((iz * NY + iy) * NX + ix) * STRIDE;
}

}  // namespace grid
}  // namespace cbfrrt
