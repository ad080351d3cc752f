"""A stand-in for the pybullet client, backed by THIS repo's geometry model (float64, oracle/oracle_f64.py).

Test infrastructure for tests/test_reference_systems_cpu.py: the reference's own ArmEnv / BasicRobot / ArmMindis are
instantiated UNMODIFIED on top of it, so that the reference's loops around the physics engine -- link ranges, floor
exemptions, the dropped first record, arg-min ties, the do/dq assembly from a Jacobian -- run for real and can be
compared with this repo's restatement.  What it does NOT provide is Bullet's own geometry (convex hulls, GJK/EPA,
contact margins): the primitive answers come from the sphere model of DESIGN.md section 2, which is exactly the part
DESIGN.md declares "parity unpinned".

Implemented subset = what the reference calls on the hot path (environment/arm_env.py, basic_robot.py,
neural_cbf/systems/arm_dynamics.py, arm_mindis.py); everything else is accepted and ignored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_f64 as O  # noqa: E402


def _quat_from_matrix(R):
    """(x, y, z, w), pybullet order"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = ((R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s)
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        v = [0.0] * 4
        v[i] = 0.25 * s
        v[j] = (R[j, i] + R[i, j]) / s
        v[k] = (R[k, i] + R[i, k]) / s
        v[3] = (R[k, j] - R[j, k]) / s
        q = tuple(v)
    return tuple(float(x) for x in q)


def _matrix_from_quat(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


class FakeBulletClient:
    GEOM_PLANE, GEOM_BOX, GEOM_CYLINDER, GEOM_SPHERE = 6, 3, 4, 2
    JOINT_REVOLUTE, JOINT_PRISMATIC, JOINT_FIXED = 0, 1, 4
    POSITION_CONTROL, VELOCITY_CONTROL = 2, 0
    DIRECT, GUI = 2, 1

    def __init__(self, *a, **k):
        self.models = O.load_models()
        self.resetSimulation()

    # ------------------------------------------------------------------ world
    def resetSimulation(self):
        self.bodies = {}          # id -> ("plane",) | ("box", centre, half) | ("robot", model name)
        self.shapes = {}
        self.filtered = set()     # (bodyA, bodyB, linkA, linkB) pairs with collisions disabled
        self.q = {}               # robot id -> {joint index: value}
        self._next = 0

    def _new_id(self):
        self._next += 1
        return self._next - 1

    def createCollisionShape(self, shapeType, halfExtents=None, **kw):
        sid = len(self.shapes)
        self.shapes[sid] = (shapeType, None if halfExtents is None else np.asarray(halfExtents, dtype=np.float64))
        return sid

    def createVisualShape(self, **kw):
        return -1

    def createMultiBody(self, baseMass=0, baseCollisionShapeIndex=-1, baseVisualShapeIndex=-1, basePosition=(0, 0, 0),
                        baseOrientation=(0, 0, 0, 1), **kw):
        kind, half = self.shapes[baseCollisionShapeIndex]
        bid = self._new_id()
        if kind == self.GEOM_PLANE:
            self.bodies[bid] = ("plane",)
        elif kind == self.GEOM_BOX:
            assert np.allclose(baseOrientation, (0, 0, 0, 1)), "axis-aligned boxes only (arm_env.py:160)"
            self.bodies[bid] = ("box", np.asarray(basePosition, dtype=np.float64), half)
        else:
            self.bodies[bid] = ("ignored",)
        return bid

    def loadURDF(self, urdf_file, basePosition=(0, 0, 0), baseOrientation=(0, 0, 0, 1), useFixedBase=True, globalScaling=1.0, **kw):
        name = "panda" if "panda" in urdf_file else "magician"
        assert abs(globalScaling - self.models[name]["scale"]) < 1e-12
        bid = self._new_id()
        self.bodies[bid] = ("robot", name)
        self.q[bid] = {l["index"]: 0.0 for l in self.models[name]["links"]}
        return bid

    def setCollisionFilterPair(self, bodyA, bodyB, linkA, linkB, enable):
        if not enable:
            self.filtered.add((bodyA, bodyB, linkA, linkB))

    # ------------------------------------------------------------------ robot description / state
    def _model(self, rid):
        return self.models[self.bodies[rid][1]]

    def getNumJoints(self, rid):
        return self._model(rid)["n_links"]

    def getBodyInfo(self, rid):
        return (self._model(rid)["base_link"].encode(), b"robot")

    def getJointInfo(self, rid, j):
        l = self._model(rid)["links"][j]
        jt = {"revolute": self.JOINT_REVOLUTE, "prismatic": self.JOINT_PRISMATIC}.get(l["type"], self.JOINT_FIXED)
        info = [None] * 17
        info[0], info[1], info[2] = j, (l["name"] + "_joint").encode(), jt
        info[8], info[9], info[10], info[11], info[12] = l["lower"], l["upper"], 0.0, l["velocity"], l["name"].encode()
        return tuple(info)

    def resetJointState(self, rid, j, targetValue=0.0, **kw):
        self.q[rid][j] = float(targetValue)

    def getJointState(self, rid, j):
        return (self.q[rid][j], 0.0, (0.0,) * 6, 0.0)

    def setJointMotorControl2(self, *a, **k):
        pass      # gripper opening: the model freezes the fingers

    def stepSimulation(self):
        pass

    def performCollisionDetection(self):
        pass

    def setAdditionalSearchPath(self, *a):
        pass

    def _frames(self, rid):
        m = self._model(rid)
        q = [self.q[rid][j] for j in m["body_joints"]]
        return O.fk(m, q)

    def getLinkState(self, rid, link, **kw):
        T = self._frames(rid)[link]
        pos, orn = tuple(float(v) for v in T[:3, 3]), _quat_from_matrix(T[:3, :3])
        return (pos, orn, (0.0, 0.0, 0.0), (0.0, 0.0, 0.0, 1.0), pos, orn)

    def getMatrixFromQuaternion(self, q):
        return tuple(_matrix_from_quat(q).reshape(-1))

    def invertTransform(self, pos, orn):
        R = _matrix_from_quat(orn).T
        return tuple(-R @ np.asarray(pos, dtype=np.float64)), _quat_from_matrix(R)

    def multiplyTransforms(self, posA, ornA, posB, ornB):
        RA, RB = _matrix_from_quat(ornA), _matrix_from_quat(ornB)
        return tuple(RA @ np.asarray(posB, dtype=np.float64) + np.asarray(posA, dtype=np.float64)), _quat_from_matrix(RA @ RB)

    def calculateJacobian(self, rid, link, localPosition, objPositions=None, objVelocities=None, objAccelerations=None):
        """linear / angular Jacobian of the point `localPosition` (given in the link frame) w.r.t. the movable joints"""
        m = self._model(rid)
        fr = self._frames(rid)
        movable = [l["index"] for l in m["links"] if l["type"] in ("revolute", "prismatic")]
        p = fr[link][:3, :3] @ np.asarray(localPosition, dtype=np.float64) + fr[link][:3, 3]
        chain, k = set(), link
        while k >= 0:
            chain.add(k)
            k = m["links"][k]["parent"]
        Jl, Ja = np.zeros((3, len(movable))), np.zeros((3, len(movable)))
        for c, j in enumerate(movable):
            if j not in chain:
                continue
            z = fr[j][:3, 2]
            if m["links"][j]["type"] == "revolute":
                Jl[:, c], Ja[:, c] = np.cross(z, p - fr[j][:3, 3]), z
        return tuple(map(tuple, Jl)), tuple(map(tuple, Ja))

    # ------------------------------------------------------------------ geometry queries (sphere model)
    def _link_obstacle(self, rid, link, obstacle):
        """(distance, point on the link, point on the obstacle) of the link's closest sphere, or None (no geometry)"""
        m = self._model(rid)
        fr = self._frames(rid)
        C = O.sphere_centres(m, fr)
        body = self.bodies[obstacle]
        best = None
        for i, s in enumerate(m["spheres"]):
            if s["link"] != link:
                continue
            c, r = C[i], s["r"]
            if body[0] == "plane":
                if link == m["floor_const"]["link"]:        # kinematic constant of the model
                    d, n = m["floor_const"]["dist"], np.array([0.0, 0.0, 1.0])
                    cand = (d, c - n * r, c - n * r - n * d)
                    return cand
                d, n = c[2] - r, np.array([0.0, 0.0, 1.0])
            elif body[0] == "box":
                sd, n = O.sd_box(c, body[1], body[2])
                d = sd - r
            else:
                continue
            cand = (d, c - n * r, c - n * r - n * d)
            if best is None or cand[0] < best[0]:
                best = cand
        return best

    def getClosestPoints(self, bodyA, bodyB, distance, linkIndexA=None, linkIndexB=None, **kw):
        m = self._model(bodyA)
        out = []
        if self.bodies[bodyB][0] == "robot":                 # self-collision query (basic_robot.py:89-98)
            assert bodyA == bodyB and linkIndexA is not None and linkIndexB is not None
            if [linkIndexA, linkIndexB] not in m["self_pairs"]:
                return ()                                    # pair not in the model (e.g. the dropped Panda wrist pair)
            C = O.sphere_centres(m, self._frames(bodyA))
            d = min((np.linalg.norm(C[i] - C[j]) - a["r"] - b["r"]
                     for i, a in enumerate(m["spheres"]) if a["link"] == linkIndexA
                     for j, b in enumerate(m["spheres"]) if b["link"] == linkIndexB), default=np.inf)
            return ((0, bodyA, bodyB, linkIndexA, linkIndexB, (0, 0, 0), (0, 0, 0), (0, 0, 1), float(d)),) if d <= distance else ()
        links = range(-1, m["n_links"]) if linkIndexA is None else [linkIndexA]
        for link in links:
            res = self._link_obstacle(bodyA, link, bodyB)
            if res is not None and res[0] <= distance:
                d, pa, pb = res
                out.append((0, bodyA, bodyB, link, -1, tuple(pa), tuple(pb), tuple((pa - pb) / d if d != 0 else (0, 0, 1)), float(d)))
        return tuple(out)

    def getContactPoints(self, bodyA=None, **kw):
        """robot vs every obstacle, penetration or touch (distance <= 0), collision filters honoured"""
        m = self._model(bodyA)
        out = []
        for bid, body in self.bodies.items():
            if body[0] not in ("plane", "box"):
                continue
            for link in range(-1, m["n_links"]):
                if (bodyA, bid, link, -1) in self.filtered:
                    continue
                res = self._link_obstacle(bodyA, link, bid)
                if res is not None and res[0] <= 0.0:
                    out.append((0, bodyA, bid, link, -1, tuple(res[1]), tuple(res[2]), (0, 0, 1), float(res[0])))
        return tuple(out)

    def removeBody(self, bid):
        self.bodies.pop(bid, None)

    def __getattr__(self, name):            # anything else the reference touches off the hot path: accept and ignore
        if name.startswith("__"):
            raise AttributeError(name)
        return lambda *a, **k: None
