"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy, float32) of the reference's LiDAR / point-cloud barrier path,
SURVEY.md 8f rank 2: the checker of cbfrrt_lidar_cbf (csrc/lidar.cuh, tests/test_gpu_lidar.py).  Pinned against the
reference's own Python: tests/golden/make_lidar_golden.py executes neural_cbf/systems/arm_lidar.py,
neural_cbf/controllers/neural_lidar_cbf_controller.py and neural_cbf/controllers/utils/pointnet.py unmodified and records
tests/golden/ref_lidar_panda.npz (tests/test_lidar_oracle_cpu.py).  Nothing in the product imports it.

    datax = [q (n) | global point cloud (P x point_dims, world frame) | per-sensor frames (S x (p 3 | R 9))]
    datax_to_x        arm_lidar.py:219-248     per sensor: R sampled points -> sensor frame  R_s^T (x - p_s)  (+ normals R_s^T n)
    pointnet_feat     pointnet.py:13-70        [point | one-hot sensor] -> 64 -> 128 -> 512 (LeakyReLU 0.1), max over the
                                               rays of a sensor, 512 -> 256 -> per_feature_dim
    encoder           pointnet.py:73-97        S * per_feature_dim -> 512 -> 256 -> feature_dim
    h                 neural_lidar_cbf_controller.py:115-139   h_nn([q | encoder]) with LeakyReLU(0.1) hidden layers
    lookahead         arm_lidar.py:250-288     q + dq, frames advanced by the frame Jacobians (J_p, J_R)
    h_with_jacobian   neural_lidar_cbf_controller.py:141-219   central... no: ONE-SIDED differences, (h(q + e_k d) - h(q)) / (2 d)
                                               with d = controller_dt / 2 (the reference divides by 2 d; kept as is)

The reference draws the sampled point indices with torch.randint inside datax_to_x (:236-238); here they are an input
(``sampled_index`` [S, R], the same for every row of the batch, as the reference's single draw per sensor is).
"""
import numpy as np

F32 = np.float32


def leaky(x, slope=0.1):
    return np.where(x >= 0, x, F32(slope) * x).astype(F32)


def linear(x, w, b):
    """torch.nn.Linear: x @ w.T + b (float32 accumulate like a CPU sgemm: tolerance-checked, not bit-exact)"""
    return (x.astype(F32) @ w.astype(F32).T + b.astype(F32)).astype(F32)


def datax_to_x(datax, n, n_sensors, ray_per_sensor, sampled_index, add_normal=True):
    """arm_lidar.py:219-248 (point_dim = 3) -> [q | obs (S * R * point_dims)]"""
    datax = np.asarray(datax, F32)
    bs = datax.shape[0]
    pd = 3 + 3 * int(add_normal)
    aux = n_sensors * 12
    cloud = datax[:, n:-aux].reshape(bs, -1, pd)
    tf = datax[:, -aux:].reshape(bs, n_sensors, 12)
    origins, rots = tf[:, :, :3], tf[:, :, 3:].reshape(bs, n_sensors, 3, 3)
    obs = np.zeros((bs, n_sensors, ray_per_sensor, pd), F32)
    for s in range(n_sensors):
        raw = cloud[:, np.asarray(sampled_index[s]).reshape(-1), :]                     # (bs, R, pd)
        rt = np.transpose(rots[:, s], (0, 2, 1))                                         # R_s^T
        obs[:, s, :, :3] = np.einsum("bij,brj->bri", rt, raw[:, :, :3] - origins[:, s][:, None, :])
        if add_normal:
            obs[:, s, :, 3:] = np.einsum("bij,brj->bri", rt, raw[:, :, 3:])
    return np.concatenate([datax[:, :n], obs.reshape(bs, -1)], axis=1).astype(F32)


def pointnet_feat(obs, sd, n_sensors, ray_per_sensor):
    """pointnet.py:52-70: obs (bs, S*R*C) -> (bs, S * per_feature_dim).  sd: pc_head state_dict (numpy)."""
    bs = obs.shape[0]
    x = obs.reshape(bs, n_sensors, ray_per_sensor, -1)
    onehot = np.broadcast_to(np.eye(n_sensors, dtype=F32)[None, :, None, :], (bs, n_sensors, ray_per_sensor, n_sensors))
    x = np.concatenate([x, onehot], axis=-1).reshape(bs * n_sensors, ray_per_sensor, -1)
    for name in ("fc11", "fc2", "fc3"):
        x = leaky(linear(x, sd[f"backbone_nn.{name}.weight"], sd[f"backbone_nn.{name}.bias"]))
    x = x.max(axis=1)                                                                    # over the rays of a sensor
    x = leaky(linear(x, sd["backbone_fc.fc2.weight"], sd["backbone_fc.fc2.bias"]))
    x = linear(x, sd["backbone_fc.fc3.weight"], sd["backbone_fc.fc3.bias"])
    return x.reshape(bs, -1)


def encoder(feature, sd):
    """pointnet.py:80-97"""
    x = leaky(linear(feature, sd["encoder_nn.fc1.weight"], sd["encoder_nn.fc1.bias"]))
    x = leaky(linear(x, sd["encoder_nn.fc2.weight"], sd["encoder_nn.fc2.bias"]))
    return linear(x, sd["encoder_nn.fc4.weight"], sd["encoder_nn.fc4.bias"])


def h_nn(x, sd, layers):
    x = leaky(linear(x, sd["input_linear.weight"], sd["input_linear.bias"]))
    for i in range(layers):
        x = leaky(linear(x, sd[f"layer_{i}_linear.weight"], sd[f"layer_{i}_linear.bias"]))
    return linear(x, sd["output_linear.weight"], sd["output_linear.bias"])


def h(datax, nets, n, n_sensors, ray_per_sensor, sampled_index, layers=2, add_normal=True):
    """neural_lidar_cbf_controller.py:115-139 -> (bs, 1).  nets = dict(pc_head=..., encoder=..., h_nn=...) of numpy state dicts"""
    x = datax_to_x(datax, n, n_sensors, ray_per_sensor, sampled_index, add_normal)
    feat = pointnet_feat(x[:, n:], nets["pc_head"], n_sensors, ray_per_sensor)
    enc = encoder(feat, nets["encoder"])
    return h_nn(np.concatenate([x[:, :n], enc], axis=1), nets["h_nn"], layers)


def lookahead(datax, dq, J_p, J_R, n, n_sensors):
    """arm_lidar.py:250-288 with data_jacobian given: q += dq; p_s += J_p dq; R_s += J_R dq (first order)"""
    out = np.array(datax, F32, copy=True)
    bs = out.shape[0]
    out[:, :n] += dq.astype(F32)
    tf = out[:, -n_sensors * 12:].reshape(bs, n_sensors, 12)
    tf[:, :, :3] += np.einsum("bsik,bk->bsi", J_p.reshape(bs, n_sensors, 3, n).astype(F32), dq.astype(F32))
    tf[:, :, 3:] += np.einsum("bsik,bk->bsi", J_R.reshape(bs, n_sensors, 9, n).astype(F32), dq.astype(F32))
    out[:, -n_sensors * 12:] = tf.reshape(bs, -1)
    return out


def h_with_jacobian(datax, J_p, J_R, nets, n, n_sensors, ray_per_sensor, sampled_indices, controller_dt, layers=2,
                    add_normal=True):
    """neural_lidar_cbf_controller.py:141-219 ("pure numerical estimation"): h at datax and at the n forward lookaheads
    q + e_k d, d = controller_dt / 2; J_k = (h_k - h) / (2 d).  The reference evaluates everything in ONE h() call on the
    concatenated batch, i.e. with one draw of sampled indices: ``sampled_indices`` [S, R]."""
    bs = datax.shape[0]
    d = F32(controller_dt / 2)
    primes = [lookahead(datax, np.tile(d * np.eye(n, dtype=F32)[k], (bs, 1)), J_p, J_R, n, n_sensors) for k in range(n)]
    h0 = h(datax, nets, n, n_sensors, ray_per_sensor, sampled_indices, layers, add_normal)
    hk = np.concatenate([h(p, nets, n, n_sensors, ray_per_sensor, sampled_indices, layers, add_normal) for p in primes], 1)
    return h0, ((hk - h0) / (d * 2)).reshape(bs, 1, n).astype(F32)
