/*
 * cbfrrt.h -- C-ABI of libcbfrrt (B200 / sm_100a): batched safety evaluation and control of manipulator states.
 *
 * This is the drop-in boundary for the CBF-INC-RRT hot path.  Every entry point replaces one per-state Python
 * loop of the reference (file:line under the reference repository root) and is what a maintainer of the
 * reference would bind with ctypes.  Conventions:
 *
 *   - plain C: pointers + sizes + scalars; no torch / C++ types cross this boundary;
 *   - "device" entry points take DEVICE pointers and a cudaStream_t passed as void* (NULL = legacy default
 *     stream); they never allocate, never synchronise and never throw; they return CBFRRT_OK (0) or a negative
 *     cbfrrt_status and leave the outputs untouched on error;
 *   - "_host" entry points take HOST pointers, stage through an internal pinned/device arena on private
 *     streams, and return after the results are in the caller's buffers (they synchronise their own streams);
 *   - all floating-point data is IEEE binary32, row-major, contiguous unless a stride argument says otherwise;
 *     masks are uint8 (0/1); indices are int32;
 *   - optional outputs may be NULL.
 *
 * Robots (compile-time tables generated from the reference's URDFs + collision meshes, spec/robot_models.json):
 *   CBFRRT_ROBOT_PANDA    n=7, 12 pybullet links, 32 link spheres    environment/franka_panda.py:17-28
 *   CBFRRT_ROBOT_MAGICIAN n=4,  4 pybullet links, 24 link spheres    environment/magician.py:18-22
 *
 * Scene (environment/arm_env.py:63-99,154-182): optional floor plane z=0 + axis-aligned boxes; sphere obstacles
 * are an extension required by BASELINE.json config 3.  Device layout:
 *   boxes   [n_boxes, 8]   cx cy cz hx | hy hz 0 0        (two 16-byte words per box)
 *   spheres [n_spheres, 4] cx cy cz r
 * Obstacle indices reported by the library follow the reference's env.obstacle_ids order: floor first (if
 * present), then boxes, then spheres.
 *
 * CBF weight packing (neural_cbf/controllers/neural_mindis_cbf_controller.py:41-53; state_dict keys
 * h_nn.input_linear / layer_{i}_linear / output_linear), IN_PAD = round_up(n+1, 4), H = hidden size,
 * L = number of hidden->hidden layers:
 *   [W_in  H x IN_PAD (zero padded)] [b_in H] { [W_l H x H] [b_l H] } x L  [w_out H] [b_out 1, padded to 4]
 */
#ifndef CBFRRT_H
#define CBFRRT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CBFRRT_ABI_VERSION 2

typedef enum {
    CBFRRT_OK = 0,
    CBFRRT_ERR_BAD_ROBOT = -1,
    CBFRRT_ERR_BAD_SHAPE = -2,    /* negative size, missing required pointer, goal_rows not in {1,B}, ... */
    CBFRRT_ERR_UNSUPPORTED = -3,  /* network shape / scene size outside what the kernels are built for */
    CBFRRT_ERR_CUDA = -4,         /* launch or runtime failure; cbfrrt_last_cuda_error() has the code */
    CBFRRT_ERR_ALIGNMENT = -5     /* scene / weight buffers must be 16-byte aligned */
} cbfrrt_status;

typedef enum { CBFRRT_ROBOT_PANDA = 0, CBFRRT_ROBOT_MAGICIAN = 1 } cbfrrt_robot;

#define CBFRRT_MAX_OBSTACLES 1024 /* boxes + spheres staged in shared memory per CTA */

typedef struct {
    int32_t n_boxes;
    int32_t n_spheres;
    int32_t floor;        /* 1: plane z=0 present (ArmEnv(include_floor=True), arm_env.py:68-70) */
    int32_t _pad;
    const float *boxes;   /* device, [n_boxes, 8], 16-byte aligned */
    const float *spheres; /* device, [n_spheres, 4], 16-byte aligned */
    const void *grid;     /* optional device buffer of cbfrrt_grid_bytes() bytes filled by cbfrrt_grid_build for
                             THIS scene (exact broad phase: same results, far fewer pair tests); NULL = brute force */
    void *stats;          /* optional device uint64[4], ACCUMULATED by cbfrrt_collide only (NULL = off): states,
                             exactly swept link spheres, exact sphere-obstacle pair evaluations, field look-ups --
                             the measured work of the broad phase, for roofline accounting */
} cbfrrt_scene;

typedef struct {
    int32_t n_in;          /* n + 1 */
    int32_t hidden;        /* cbf_hidden_size  (48 in every shipped entry point; train_arm_mindis.py:176-177) */
    int32_t layers;        /* cbf_hidden_layers (2).  48 x 2 runs the specialised (tcgen05 / FP32-FMA) kernels.  Any
                              other shape with (layers + 2) * hidden <= 1536 runs a generic FP32 kernel in
                              cbfrrt_cbf_filter and cbfrrt_safety_filter (there as two launches; datax must be given);
                              cbfrrt_rollout and cbfrrt_safety_filter_scatter return CBFRRT_ERR_UNSUPPORTED for it */
    int32_t _pad;
    const float *weights;  /* device, packed as described above, 16-byte aligned */
} cbfrrt_cbf_net;

/* Parameters of the CBF-QP safety filter (neural_cbf/controllers/clf_controller.py:82-139,354-404) and of the
 * nominal controller (neural_cbf/systems/arm_dynamics.py:124-161). */
typedef struct {
    const float *goal;   /* device [goal_rows, n]: dynamics_model.intermediate_goals (arm_dynamics.py:71-74) */
    int64_t goal_rows;   /* 1 (shared goal) or B (one goal per row, batch_tsa.py:199) */
    const float *K;      /* device [n, n] LQR gain (control_affine_system.py:125-157) */
    const float *u_lo;   /* device [n]  control_limits lower (arm_dynamics.py:111-119) */
    const float *u_hi;   /* device [n]  control_limits upper */
    float cbf_lambda;    /* cbf_alpha */
    float relaxation_penalty; /* clamped to 1e6 inside, clf_controller.py:380 */
    const float *u_ref;  /* optional device [B, n]: caller-supplied reference control, bypasses the nominal
                            controller exactly like solve_CLF_QP(u_ref=...) (clf_controller.py:497-504); when set,
                            goal and K may be NULL */
} cbfrrt_filter_params;

/* steer-loop parameters (motion_planning/baseline/tsa.py:217-269) */
typedef struct {
    int32_t steps;        /* RRT_step */
    int32_t ctrl_every;   /* controller_dt // dt (4 in every shipped entry point) */
    float dt;             /* simulation step */
    int32_t stuck_lag;    /* 0 disables the stuck test; tsa.py:263 uses 9 (x_list[-1] vs x_list[-10]) */
    int32_t stuck_after;  /* test only when tstep > stuck_after (10) */
    float stuck_eps;      /* 0.1 */
    int32_t keep_going;   /* 0: steer semantics (stop at the first unsafe state, tsa.py:251-253).
                             1: fixed-work variant of BASELINE config 4 ("mask instead of early stop"): an unsafe
                                state clears `alive` but the trajectory is simulated for all `steps` steps */
    int32_t stuck_mode;   /* 0: single steer (tsa.py:251-266): the new state is appended, then the stuck test compares
                                the states appended at steps t and t - stuck_lag when t > stuck_after; a stuck row stops
                                with alive = 1.
                             1: batched steer (batch_tsa.py:225-229): the stuck test runs BEFORE the append on the list as
                                it stands (more than stuck_after states stored: the last one against the one stuck_lag
                                before it; the reference uses lag 9, after 10, eps 3e-2); stuck OR unsafe clears `alive`
                                and the row never appends again, while its state keeps integrating when keep_going = 1
                                (q_final = that running state, as the reference returns x_current) or stops at its
                                death when keep_going = 0. */
} cbfrrt_rollout_params;

/* ---- library info ------------------------------------------------------------------------------------- */
int cbfrrt_abi_version(void);
const char *cbfrrt_status_string(int status);
int cbfrrt_last_cuda_error(void);
/* n (dof), number of pybullet links (FK rows), number of link spheres */
int cbfrrt_robot_info(int robot, int32_t *n, int32_t *n_links, int32_t *n_spheres);

/* ---- barrier-network path selection -----------------------------------------------------------------------
 * The MLP of cbf_filter / safety_filter / rollout has two implementations with the same contract (1e-5 against the
 * oracle): tcgen05 tensor-core kernels (TF32x3) and FP32-FMA kernels.  AUTO (default) takes the tensor-core kernels
 * above 16 384 rows and the FP32-FMA kernels below (launch-latency bound sizes).  FMA / TC force one path for every
 * size (A/B measurements, and the parity tests, which cover both).  Planner-sized calls (the reference's own regime:
 * batch = 1, eval_batchrrt_mindis_cbfsteer.py:66) have a third form of every kernel that spreads ONE state over the 32
 * lanes of a warp (csrc/warp.cuh: bit-identical masks / minima / arg-mins, same tolerance for the network); AUTO takes it
 * up to 4096 rows per call (2048 trajectories, every rollout_segments call), WARP forces it for every size, FMA / TC
 * never use it.  Process-wide; initial value from the environment variable CBFRRT_MLP=fma|tc|warp.  Returns
 * CBFRRT_ERR_BAD_SHAPE for an unknown value. */
typedef enum { CBFRRT_MLP_AUTO = 0, CBFRRT_MLP_FMA = 1, CBFRRT_MLP_TC = 2, CBFRRT_MLP_WARP = 3 } cbfrrt_mlp_path;
int cbfrrt_set_mlp_path(int path);
int cbfrrt_get_mlp_path(void);

/* ---- optional scratch for the large sweeps --------------------------------------------------------------------
 * From 524 288 rows on, cbfrrt_safety_filter / cbfrrt_safety_filter_scatter run as TWO launches when they have room for
 * the intermediate observation rows: a geometry pass at its own occupancy (masks + datax) and a tensor-core barrier-network
 * pass through datax (h, J, u, r and the peer stores of the fused exchange) -- measured 1.2x faster on B200 than the single
 * fused kernel, whose geometry runs at the occupancy the tensor-core stage dictates; results are bit-identical.  Room =
 * the caller's datax output when it is requested, else a scratch of at least 16 + (4 (2 n + 1) + 1) B bytes registered
 * here for the CURRENT device (the library never allocates device memory in the device-pointer entry points; the buffer
 * must stay valid, and not be used by the caller, while such calls are in flight; calls that use it must be ordered on
 * ONE stream -- concurrent streams need their own buffers, registered before each launch: the pointer is read at launch
 * time).  bytes = 0 unregisters.  Without room
 * the fused kernel runs.  cbfrrt_set_two_pass(0) forces the fused kernel (A/B measurements, tests). */
int cbfrrt_set_workspace(void *device_ptr, int64_t bytes);
int cbfrrt_set_two_pass(int enabled);

/* ---- scene acceleration structure (built once per ArmEnv.reset_env, environment/arm_env.py:63) --------------
 * Nearest-candidate grid over the arm's workspace: per 5 cm cell, the obstacles that can be the closest one for
 * some point of the cell.  Results with and without the grid are bit-identical (csrc/grid.cuh).  The caller owns
 * the buffer (16-byte aligned, cbfrrt_grid_bytes() bytes) and must rebuild it whenever boxes / spheres change. */
int64_t cbfrrt_grid_bytes(void);
int cbfrrt_grid_build(const cbfrrt_scene *scene, void *grid, void *stream);

/* ---- K1  forward kinematics ----------------------------------------------------------------------------
 * Replaces BasicRobot.forward_kinematics + set_joint_position (environment/basic_robot.py:55-68): world pose
 * of every pybullet link frame (getLinkState[4], [5]).   q [B, n] (row stride q_stride floats)
 *   -> pos [B, n_links, 3], rot [B, n_links, 9] (row-major 3x3).  */
int cbfrrt_fk(int robot, const float *q, int64_t q_stride, int64_t B, float *pos, float *rot, void *stream);

/* ---- K1+K2+K3  collision masks + min-distance observation ---------------------------------------------
 * Replaces ArmDynamics.safe_mask / unsafe_mask (neural_cbf/systems/arm_dynamics.py:169-232, incl.
 * BasicRobot.check_self_collision_free basic_robot.py:89-98) and ArmMindis._get_observation_with_state +
 * complete_sample_with_observations (arm_mindis.py:56-90, arm_dynamics.py:267-279).
 *   safe/unsafe [B] uint8;  datax [B, 2n+1] = [q | o | do/dq];  arg_link/arg_obs [B] = arg-min pair of o;
 *   mins [B,3] = (self_min, soft_min, hard_min) diagnostics.  All outputs optional. */
int cbfrrt_collide(int robot, const cbfrrt_scene *scene, const float *q, int64_t q_stride, int64_t B,
                   float thr_soft, float thr_hard, uint8_t *safe, uint8_t *unsafe, float *datax,
                   int32_t *arg_link, int32_t *arg_obs, float *mins, void *stream);

/* ---- K4+K5+K6+K7+K8  barrier value, Jacobian, Lie derivatives, nominal control, CBF-QP -----------------
 * Replaces NeuralMindisCBFController.h / h_with_jacobian (neural_mindis_cbf_controller.py:71-102),
 * CLFController.V_with_lie_derivatives / solve_CLF_QP / u (clf_controller.py:169-218,461-535) and
 * ArmDynamics.u_nominal (arm_dynamics.py:124-161) for n_scenarios == 1.
 *   datax [B, 2n+1] -> h [B], J [B, n] (= Lg_V; Lf_V = 0), u [B, n], r [B].  Outputs optional. */
int cbfrrt_cbf_filter(int n, const cbfrrt_cbf_net *net, const float *datax, int64_t B,
                      const cbfrrt_filter_params *fp, float *h, float *J, float *u, float *r, void *stream);

/* ---- K8 alone: batched single-constraint CLF/CBF-QP -------------------------------------------------------
 * Replaces CLFController._solve_CLF_QP_cvxpylayers (clf_controller.py:354-404) for n_scenarios == 1 when the
 * caller already has the Lie derivatives:  min ||u-u_ref||^2 + p r  s.t. c + a.u - r <= 0, r >= 0, lo<=u<=hi
 * with a = Lg_V [B, n], c = Lf_V + lambda V [B].  1 <= n <= 8.  p is clamped to 1e6 (:380). */
int cbfrrt_qp(int n, const float *a, const float *c, const float *u_ref, int64_t B, const float *u_lo,
              const float *u_hi, float relaxation_penalty, float *u, float *r, void *stream);

/* ---- K8 reverse mode: gradients of the single-constraint CLF/CBF-QP solution -------------------------------
 * Replaces backpropagation through the CvxpyLayer (clf_controller.py:82-139; solve_CLF_QP(requires_grad=True)
 * :461-528 -> diffcp's derivative of the cone program) by the closed-form sensitivities of the KKT solution on its
 * active set.  Inputs as cbfrrt_qp plus grad_u = dL/du [B, n] and grad_r = dL/dr [B] (NULL = 0); outputs (each
 * optional) grad_a = dL/dLg_V [B, n], grad_c = dL/d(Lf_V + lambda V) [B] (so dL/dLf_V = grad_c, dL/dV = lambda
 * grad_c), grad_u_ref [B, n].  fp32 in / out, double inside.  At a kink of the solution map (a coordinate exactly
 * on its bound, the constraint exactly touching) the clipped / inactive side is differentiated. */
int cbfrrt_qp_backward(int n, const float *a, const float *c, const float *u_ref, int64_t B, const float *u_lo,
                       const float *u_hi, float relaxation_penalty, const float *grad_u, const float *grad_r,
                       float *grad_a, float *grad_c, float *grad_u_ref, void *stream);

/* ---- K8 for n_scenarios > 1: batched multi-constraint CLF/CBF-QP ---------------------------------------------
 * Replaces CLFController._solve_CLF_QP_cvxpylayers (clf_controller.py:354-404; one relaxed constraint per scenario,
 * :86-139) when len(scenarios) = S > 1:
 *   min ||u-u_ref||^2 + p sum_i r_i  s.t. c_i + a_i.u - r_i <= 0, r_i >= 0, lo <= u <= hi
 * with a [B, S, n], c [B, S], u_ref [B, n] -> u [B, n], r [B, S] (optional), iters [B] (optional: iterations used
 * per row; == max_iter marks a row that did not reach tol).  Dual active-set solver: the concave piecewise-quadratic
 * dual over mu in [0,p]^S is maximised by iterations of (1) one cyclic sweep of exact coordinate maximisations (the
 * single-constraint root find) and (2) Newton steps on the current active set with an exact line search; stops when
 * the control-space movement of an iteration is <= tol * (1 + size of the dual shift).  1 <= n <= 8, 1 <= S <= 4.
 * p is clamped to 1e6.  Typical: 2-5 iterations.  With more constraints than unclipped controls (S > n) a few rows in
 * 10 000 zig-zag between active sets and exhaust max_iter; they are reported through `iters`. */
int cbfrrt_qp_multi(int n, int n_constraints, const float *a, const float *c, const float *u_ref, int64_t B,
                    const float *u_lo, const float *u_hi, float relaxation_penalty, int max_iter, float tol,
                    float *u, float *r, int32_t *iters, void *stream);

/* Reverse mode of cbfrrt_qp_multi (same role as cbfrrt_qp_backward for len(scenarios) > 1): grad_u [B, n], grad_r
 * [B, S] (NULL = 0) -> grad_a [B, S, n], grad_c [B, S], grad_u_ref [B, n] (each optional).  Stateless: the row is
 * solved again inside the call (same max_iter / tol as the forward call), the active sets are read off the solution
 * (unclipped coordinates; tight / relaxed / inactive constraints) and the KKT system on them is differentiated. */
int cbfrrt_qp_multi_backward(int n, int n_constraints, const float *a, const float *c, const float *u_ref, int64_t B,
                             const float *u_lo, const float *u_hi, float relaxation_penalty, int max_iter, float tol,
                             const float *grad_u, const float *grad_r, float *grad_a, float *grad_c,
                             float *grad_u_ref, void *stream);

/* ---- fused q -> (masks, datax, h, J, u, r): one launch for sampled states ------------------------------ */
int cbfrrt_safety_filter(int robot, const cbfrrt_scene *scene, const cbfrrt_cbf_net *net, const float *q,
                         int64_t q_stride, int64_t B, const cbfrrt_filter_params *fp, float thr_soft,
                         float thr_hard, uint8_t *safe, uint8_t *unsafe, float *datax, float *h, float *J,
                         float *u, float *r, void *stream);

/* ---- fused sweep + all-gather: q -> (valid, u) written straight into every peer GPU's replicated result --------
 * The multi-GPU exchange named by BASELINE.json ("all-gather validity masks and filtered controls") fused into the
 * kernel epilogue: each state's mask byte and control row are stored at global row (row_offset + b) of EVERY
 * rank's buffers through NVLink peer mappings (torch symmetric memory / cudaIpc / cuMem), so the transfer overlaps
 * the sweep tile by tile and no collective is launched.  The caller synchronises the ranks afterwards (a barrier),
 * exactly as after an all-gather.  valid[p] / u[p] must be mapped in this process for p < n_peers (the rank's
 * own buffer included).  On NVSwitch systems pass n_peers = 1, multicast = 1 with the buffers' MULTICAST addresses
 * (cuMulticast / torch symmetric memory multicast_ptr): one multimem.st per row is then replicated to every rank by
 * the switch (NVLS), 1/N of the outbound traffic of the per-peer form (measured: 0.44 vs 0.51 ms per step at 8 GPUs).
 * Ordering contract for a caller that reuses the replicas step after step: synchronise the ranks BEFORE the launch too
 * (every rank has finished reading the previous step's rows), or alternate between two sets of buffers. */
#define CBFRRT_MAX_PEERS 8
typedef struct {
    int32_t n_peers;
    int32_t multicast;                  /* 1: n_peers == 1 and valid[0] / u[0] are MULTICAST addresses; the kernel then uses
                                           multimem.st (the only access PTX allows on a multimem address), packing the mask
                                           bytes of 32 consecutive rows into words: B and row_offset must be multiples of 128.
                                           0: ordinary (peer-mapped) pointers, plain stores */
    int64_t row_offset;                 /* global index of this rank's first row */
    uint8_t *valid[CBFRRT_MAX_PEERS];   /* each: device u8 [G]     */
    float *u[CBFRRT_MAX_PEERS];         /* each: device f32 [G, n] */
} cbfrrt_peer_out;

int cbfrrt_safety_filter_scatter(int robot, const cbfrrt_scene *scene, const cbfrrt_cbf_net *net, const float *q,
                                 int64_t q_stride, int64_t B, const cbfrrt_filter_params *fp, float thr_soft,
                                 float thr_hard, const cbfrrt_peer_out *peers, void *stream);

/* The barrier-network pass alone with the fused exchange: datax [B, 2n+1] (e.g. from cbfrrt_collide) -> u, and
 * (valid_in[b], u[b]) stored at global row row_offset + b of every peer buffer.  With cbfrrt_collide on one stream and this
 * call on a second (higher-priority) one, a caller can pipeline a sweep in chunks so that the NVLink stores of chunk k
 * drain under the geometry pass of chunk k + 1 (cbfrrt.sharded.FusedSweep does, where the exchange would otherwise bound
 * the step: 8 GPUs).  48 x 2 networks only. */
int cbfrrt_cbf_filter_scatter(int n, const cbfrrt_cbf_net *net, const float *datax, const uint8_t *valid_in, int64_t B,
                              const cbfrrt_filter_params *fp, const cbfrrt_peer_out *peers, void *stream);

/* ---- K10 edge validity -----------------------------------------------------------------------------------
 * Replaces edge_checking (motion_planning/baseline/tsa.py:272-286) for E edges with a fixed number n_interp
 * of interpolation points: valid <=> safe(q_to) and safe(q_from + k/n_interp (q_to - q_from)), k=0..n_interp-1.
 * first_bad = smallest failing along-edge index (k, or n_interp for the end point), -1 if the edge is valid.
 * Three launches (preset, one thread per (edge, point), finish); every point is evaluated. */
int cbfrrt_edge_validity(int robot, const cbfrrt_scene *scene, const float *q_from, const float *q_to, int64_t E,
                         int32_t n_interp, float thr_soft, float thr_hard, uint8_t *valid, int32_t *first_bad,
                         void *stream);

/* Same for edges of different lengths, as edge_checking computes them (tsa.py:277-279: K_e = int(|q_to - q_from| / eps),
 * K_e = 0 checks only the end point).  point_offsets [E+1] (device, int64) is the exclusive prefix sum of K_e + 1,
 * n_points = point_offsets[E] (passed by value so that the call does not synchronise).  Three launches (preset,
 * one thread per (edge, point), finish); every point is evaluated (no early exit). */
int cbfrrt_edge_validity_var(int robot, const cbfrrt_scene *scene, const float *q_from, const float *q_to, int64_t E,
                             const int64_t *point_offsets, int64_t n_points, float thr_soft, float thr_hard,
                             uint8_t *valid, int32_t *first_bad, void *stream);

/* ---- planner front-end: nearest / k-nearest / radius queries in configuration space -------------------------
 * Replaces the numpy distance matrices of the callers of the path: batch_tsa.global_explore
 * (motion_planning/baseline/batch_tsa.py:135-145: np.argmin / np.argsort[:batch] of ||sample - node||),
 * tsa.global_explore (tsa.py:151-153) and BITStarPlanner.expand_vertex (bit_star.py:196-216: nodes within r).
 *   queries [M, n], tree [N_tree, n] (device, fp32, 1 <= n <= 8) -> idx [M, k] int32, dist [M, k] fp32: the k nearest
 *   rows in ascending (squared distance, row) order -- ties to the lowest row like np.argmin / a stable argsort;
 *   entries beyond N_tree are idx -1 / dist +inf.  workspace: cbfrrt_knn_workspace_bytes(M, N_tree, k) bytes of device
 *   memory (16 * M for k = 1).  k = 1: one reduction launch; 1 < k <= 1024 (and M <= 65535): two launches (per-chunk
 *   shared-memory sort keeping k candidates, per-query merge); larger k: k rounds of the reduction.
 *   radius_mask: within [M, N_set] uint8 = (||q_m - s_j|| <= radius). */
int64_t cbfrrt_knn_workspace_bytes(int64_t M, int64_t N_tree, int32_t k);
int cbfrrt_knn(int n, const float *queries, const float *tree, int64_t M, int64_t N_tree, int32_t k, int32_t *idx,
               float *dist, void *workspace, void *stream);
int cbfrrt_radius_mask(int n, const float *queries, const float *set, int64_t M, int64_t N_set, float radius,
                       uint8_t *within, void *stream);

/* ---- K9 closed-loop rollouts ------------------------------------------------------------------------------
 * Replaces the steer loop (motion_planning/baseline/tsa.py:217-269, batch_tsa.py:191-231) built on
 * ArmDynamics.closed_loop_dynamics (arm_dynamics.py:474-523): control every ctrl_every steps, Euler every
 * step, observation refresh when (t+1) % ctrl_every == 0, unsafe check after every step.  A trajectory stops
 * when the new state is unsafe (that state is NOT appended: alive = 0) or when the stuck test fires
 * (alive stays 1).   q0 [T, n], goal [goal_rows, n] ->
 *   q_final [T, n]  = last appended state (x_list[-1]),  alive [T] = no_collision,
 *   steps_done [T]  = number of appended states,  traj [T, steps, n] optional = x_list[1:] (rows beyond
 *   steps_done are left untouched); required when the stuck test is enabled. */
int cbfrrt_rollout(int robot, const cbfrrt_scene *scene, const cbfrrt_cbf_net *net, const float *q0, int64_t T,
                   const cbfrrt_filter_params *fp, const cbfrrt_rollout_params *rp, float thr_soft, float thr_hard,
                   float *q_final, uint8_t *alive, int32_t *steps_done, float *traj, void *stream);

/* ---- a whole expansion of the batched planner in one launch -----------------------------------------------------
 * Replaces the segment loop of global_explore (motion_planning/baseline/batch_tsa.py:154-170: RRT_PARAM // 20 calls of
 * the batched steer, each starting where the previous one ended) for ONE batch of T <= 128 rows: n_segments consecutive
 * segments of rp->steps steps each with the semantics of batch_tsa.py:191-231 (stuck_mode 1, dead rows keep
 * integrating, and the reference's batch-wide early exit: a segment ends once every row of the batch is dead, :229-230;
 * keep_going / stuck_mode of rp are ignored).  Every segment starts with a fresh observation of the rows' current states.
 *   seg_q [T, n_segments, n] = x_current at the end of each segment (the node the planner inserts),
 *   seg_alive [T, n_segments] = no_collisions, seg_done [T, n_segments] = stored states of the segment,
 *   traj [T, n_segments * steps, n]: segment k's stored states start at row k * steps (required). */
int cbfrrt_rollout_segments(int robot, const cbfrrt_scene *scene, const cbfrrt_cbf_net *net, const float *q0, int64_t T,
                            const cbfrrt_filter_params *fp, const cbfrrt_rollout_params *rp, int32_t n_segments,
                            float thr_soft, float thr_hard, float *seg_q, uint8_t *seg_alive, int32_t *seg_done,
                            float *traj, void *stream);

/* ---- counter-based uniform state sampler (ControlAffineSystem.sample_state_space,
 * control_affine_system.py:291-300, made shard-invariant): row g = start_row + b, q[b,j] = lo_j + U(g,j)(hi_j-lo_j) */
int cbfrrt_sample_states(int robot, uint32_t seed, uint64_t start_row, int64_t B, float *q, void *stream);

/* ---- LiDAR / point-cloud barrier (the one GEMM-sized workload of the reference) ------------------------------
 * Replaces NeuralLidarCBFController.h / h_with_jacobian (neural_cbf/controllers/neural_lidar_cbf_controller.py:115-219)
 * with everything they call: ArmLidar.datax_to_x (neural_cbf/systems/arm_lidar.py:219-248: the sampled cloud points of
 * every sensor taken to the sensor frame, R_s^T (x - p_s), normals R_s^T n), PointNetfeat (neural_cbf/controllers/
 * utils/pointnet.py:13-70: [point | normal | one-hot sensor] -> 64 -> 128 -> 512 with LeakyReLU(0.1), max over the rays of
 * a sensor, 512 -> 256 -> per_feature_dim), PointNetVanillaEncoder (pointnet.py:73-97), the h_nn head on [q | encoded],
 * batch_lookahead (arm_lidar.py:250-288) and the one-sided differences J_k = (h(q + d e_k) - h(q)) / (2 d) (:205-213).
 * The three per-point layers run on the tcgen05 tensor cores (TF32 3-term split, fp32-grade), the rest in FP32.
 *
 * Weights: the per-point layers as TF32 hi / lo tiles prepared by the host (cbfrrt.ops.pack_lidar_net documents the
 * layout: no-swizzle K-major core matrices W4[k/4][n], layer 3 in 8 chunks of 64 outputs); everything after the
 * max-pool as the raw torch tensors (row-major [out][in]).  All pointers device, 16-byte aligned. */
typedef struct {
    int32_t n;                /* joints (q_dims) */
    int32_t n_sensors;        /* S <= 7 */
    int32_t point_dims;       /* 3, or 6 with normals (add_normal) */
    int32_t per_feature_dim;  /* PointNetfeat output per sensor, <= 64 */
    int32_t feature_dim;      /* encoder output, n + feature_dim <= 72 */
    int32_t hidden, layers;   /* h_nn: (n + feature_dim) -> hidden x (layers + 1) -> 1, hidden <= 512 */
    int32_t _pad;
    const float *w1t, *w2t, *w3t, *bias;                /* backbone_nn tiles + [b1 | b2 | b3] */
    const float *fc2_w, *fc2_b, *fc3_w, *fc3_b;         /* backbone_fc 512 -> 256 -> per_feature_dim */
    const float *e1_w, *e1_b, *e2_w, *e2_b, *e4_w, *e4_b; /* encoder_nn S*PF -> 512 -> 256 -> feature_dim */
    const float *h_in_w, *h_in_b, *h_hid_w, *h_hid_b, *h_out_w, *h_out_b; /* h_nn; h_hid_*: the `layers` blocks concatenated */
} cbfrrt_lidar_net;

/* bytes of device workspace for B states (with_jacobian: n + 1 evaluations per state) */
int64_t cbfrrt_lidar_workspace_bytes(const cbfrrt_lidar_net *net, int64_t B, int32_t with_jacobian);

/* q [B, n]; cloud [B, n_cloud, point_dims] (world frame); frames [B, S, 12] = per sensor origin (3) | rotation (9,
 * row-major); sampled_index [S, rays] int32 = the cloud indices of every sensor's rays (ONE draw for the whole batch,
 * as the reference's torch.randint inside datax_to_x, arm_lidar.py:236-238); rays % 32 == 0.
 * h [B] (optional).  J [B, n] optional: then J_p [B, S, 3, n] and J_R [B, S, 9, n] (the frame Jacobians of
 * data_jacobian) and delta = controller_dt / 2 are required. */
int cbfrrt_lidar_cbf(const cbfrrt_lidar_net *net, const float *q, const float *cloud, int64_t n_cloud, const float *frames,
                     const float *J_p, const float *J_R, const int32_t *sampled_index, int32_t rays, int64_t B,
                     float delta, void *workspace, float *h, float *J, void *stream);

/* ---- host-buffer front end (what a ctypes / numpy caller of the reference would use) ----------------------
 * Same semantics as cbfrrt_safety_filter but every pointer (including scene, weights and parameters) is a
 * HOST pointer.  Rows are processed in chunks on private streams: H2D copy, kernel and D2H copy of
 * consecutive chunks overlap.  device: CUDA device ordinal. */
typedef struct {
    int32_t n_boxes, n_spheres, floor, _pad;
    const float *boxes;   /* host [n_boxes, 6]  cx cy cz hx hy hz   (ArmEnv.obs_positions / obs_sizes) */
    const float *spheres; /* host [n_spheres, 4] */
} cbfrrt_host_scene;

int cbfrrt_safety_filter_host(int device, int robot, const cbfrrt_host_scene *scene, int32_t hidden, int32_t layers,
                              const float *weights, int64_t n_weights, const float *q, int64_t B, const float *goal,
                              int64_t goal_rows, const float *K, const float *u_lo, const float *u_hi,
                              float cbf_lambda, float relaxation_penalty, float thr_soft, float thr_hard,
                              uint8_t *safe, uint8_t *unsafe, float *datax, float *h, float *J, float *u, float *r);

/* ---- planner-sized host front ends ---------------------------------------------------------------------------
 * The reference's planners keep tree, samples and edges in numpy and touch the path a handful of rows at a time
 * (global_explore, motion_planning/baseline/batch_tsa.py:99-189: nearest nodes :129-145, steer segments :154-170).  These
 * two take HOST pointers and cost one pinned H2D copy, the kernels and one D2H copy on a private stream of `device` (the
 * scene, network and gains stay resident between calls, as for cbfrrt_safety_filter_host): the per-expansion overhead of
 * the tensor wrappers (allocations, per-output read-backs) disappears.  Same results as cbfrrt_rollout_segments /
 * cbfrrt_knn bit for bit.  goal: [T, n] (one row per trajectory).  T <= 128. */
int cbfrrt_rollout_segments_host(int device, int robot, const cbfrrt_host_scene *scene, int32_t hidden, int32_t layers,
                                 const float *weights, int64_t n_weights, const float *q0, int64_t T, const float *goal,
                                 const float *K, const float *u_lo, const float *u_hi, float cbf_lambda,
                                 float relaxation_penalty, const cbfrrt_rollout_params *rp, int32_t n_segments,
                                 float thr_soft, float thr_hard, float *seg_q, uint8_t *seg_alive, int32_t *seg_done,
                                 float *traj);
int cbfrrt_knn_host(int device, int n, const float *queries, const float *tree, int64_t M, int64_t N_tree, int32_t k,
                    int32_t *idx, float *dist);
/* edge_checking (motion_planning/baseline/tsa.py:272-286; BIT* is_edge_free, bit_star.py:126-129) for numpy callers:
 * n_interp [E] = the K_e of every edge (host, >= 0); same results as cbfrrt_edge_validity_var */
int cbfrrt_edge_validity_host(int device, int robot, const cbfrrt_host_scene *scene, const float *q_from, const float *q_to,
                              int64_t E, const int32_t *n_interp, float thr_soft, float thr_hard, uint8_t *valid,
                              int32_t *first_bad);

/* ---- measurement utility: FP32 FMA throughput probe (the roofline denominator of this path is the CUDA-core
 * FMA rate, which MEASURED_PEAKS.json does not contain).  Launches grid x block threads, each running
 * 8 independent chains of `iters` dependent FMAs; flops = grid*block*iters*16.  sink: device float[1]. */
int cbfrrt_fma_probe(int32_t grid, int32_t block, int32_t iters, float *sink, void *stream);

/* number of kernels launched by this library since load (bench.py's gpu_launches evidence) */
int64_t cbfrrt_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* CBFRRT_H */
