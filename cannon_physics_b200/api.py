"""Host-side mirror of the reference's object API for the step path.

Same class names, constructor arguments and defaults as the Dart library so that a program written against
``cannon_physics`` reads the same here:

    World / Body / Sphere / Plane / Box / Cylinder / ConvexPolyhedron / Heightfield / Material /
    ContactMaterial / NaiveBroadphase / SAPBroadphase / GridBroadphase / GSSolver /
    PointToPointConstraint / HingeConstraint / DistanceConstraint / LockConstraint / ConeTwistConstraint

(lib/world/world_class.dart:44, lib/objects/rigid_body.dart:26, lib/rigid_body_shapes/*.dart,
lib/material/*.dart, lib/collision/*broadphase.dart, lib/solver/gs_solver.dart, lib/constraints/*.dart).

The objects only hold host-side description; ``World.step`` flattens them once into the SoA upload of
``include/cannon_cuda.h`` and afterwards steps on the device (the reference's ``CudaWorld`` fused mode,
SURVEY.md §8b). Body vectors are float32 numpy rows, refreshed from the device after every ``step`` unless
``sync=False``.  Errors surface as ``CannonError`` (the reference throws strings).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi as F
from .engine import Context, DeviceWorld, SceneSpec

__all__ = [
    "Vec3", "Quaternion", "Material", "ContactMaterial", "Shape", "Sphere", "Plane", "Box", "Cylinder", "ConvexPolyhedron", "Cone", "Capsule", "SizedPlane", "LatheShape", "CapsuleLathe", "Particle", "Trimesh", "SPHSystem",
    "Heightfield", "Body", "BodyTypes", "BodySleepStates", "Broadphase", "NaiveBroadphase", "SAPBroadphase", "GridBroadphase",
    "CudaBroadphase", "Solver", "GSSolver", "CudaGSSolver", "SplitSolver", "Constraint", "PointToPointConstraint", "HingeConstraint", "DistanceConstraint", "LockConstraint", "ConeTwistConstraint", "Spring",
    "SpringConstraint", "RaycastResult", "World",
    "CudaWorld", "CannonError",
]

CannonError = F.CannonError


def Vec3(x=0.0, y=0.0, z=0.0) -> np.ndarray:
    """vector_math Vector3: float32 storage."""
    return np.array([x, y, z], dtype=np.float32)


class Quaternion:
    """Helpers of lib/math/quaternion.dart that scene set-up code uses (storage is a float32 row x,y,z,w)."""

    @staticmethod
    def identity() -> np.ndarray:
        return np.array([0, 0, 0, 1], dtype=np.float32)

    @staticmethod
    def setFromEuler(x: float, y: float, z: float) -> np.ndarray:  # quaternion.dart:191-203 (order XYZ)
        c1, c2, c3 = math.cos(x / 2), math.cos(y / 2), math.cos(z / 2)
        s1, s2, s3 = math.sin(x / 2), math.sin(y / 2), math.sin(z / 2)
        return np.array([s1 * c2 * c3 + c1 * s2 * s3, c1 * s2 * c3 - s1 * c2 * s3, c1 * c2 * s3 + s1 * s2 * c3,
                         c1 * c2 * c3 - s1 * s2 * s3], dtype=np.float32)

    @staticmethod
    def setFromAxisAngle(axis: Sequence[float], angle: float) -> np.ndarray:  # quaternion.dart:85-91
        s = math.sin(angle * 0.5)
        a = np.asarray(axis, dtype=np.float32).astype(np.float64)
        return np.array([a[0] * s, a[1] * s, a[2] * s, math.cos(angle * 0.5)], dtype=np.float32)


class BodyTypes:
    dynamic, static, kinematic = F.BODY_DYNAMIC, F.BODY_STATIC, F.BODY_KINEMATIC


class BodySleepStates:
    awake, sleepy, sleeping = F.AWAKE, F.SLEEPY, F.SLEEPING


class Material:  # lib/material/material.dart:2
    def __init__(self, friction: float = -1, restitution: float = -1, name: str = ""):
        self.friction, self.restitution, self.name = friction, restitution, name


class ContactMaterial:  # lib/material/contact_material.dart:5
    def __init__(self, m1: Material, m2: Material, friction=0.3, restitution=0.3, contactEquationStiffness=1e7,
                 contactEquationRelaxation=3, frictionEquationStiffness=1e7, frictionEquationRelaxation=3):
        self.materials = [m1, m2]
        self.friction, self.restitution = friction, restitution
        self.contactEquationStiffness, self.contactEquationRelaxation = contactEquationStiffness, contactEquationRelaxation
        self.frictionEquationStiffness, self.frictionEquationRelaxation = frictionEquationStiffness, frictionEquationRelaxation


class Shape:  # lib/rigid_body_shapes/shape.dart:28
    type = -1

    def __init__(self, collisionResponse=True, collisionFilterGroup=-1, collisionFilterMask=-1, material: Optional["Material"] = None):
        self.collisionResponse = collisionResponse
        self.collisionFilterGroup, self.collisionFilterMask = collisionFilterGroup, collisionFilterMask
        self.material = material  # shape.dart:48: overrides the body's material for this shape

    def _desc(self) -> dict:
        return dict(type=self.type, collision_response=int(self.collisionResponse), collision_filter_group=self.collisionFilterGroup,
                    collision_filter_mask=self.collisionFilterMask)


class Sphere(Shape):  # sphere.dart:11
    type = F.SHAPE_SPHERE

    def __init__(self, radius: float = 1.0, **kw):
        super().__init__(**kw)
        if radius < 0:
            raise ValueError("The sphere radius cannot be negative.")
        self.radius = float(radius)

    def _desc(self):
        return dict(super()._desc(), radius=self.radius)


class Plane(Shape):  # plane.dart:11
    type = F.SHAPE_PLANE


class Box(Shape):  # box.dart:14
    type = F.SHAPE_BOX

    def __init__(self, halfExtents, **kw):
        super().__init__(**kw)
        self.halfExtents = np.asarray(halfExtents, dtype=np.float32)

    def _desc(self):
        return dict(super()._desc(), half_extents=self.halfExtents)


class Cylinder(Shape):  # cylinder.dart:15
    type = F.SHAPE_CYLINDER

    def __init__(self, radiusTop=1.0, radiusBottom=1.0, height=1.0, numSegments=8, **kw):
        super().__init__(**kw)
        if radiusTop < 0:
            raise ValueError("The cylinder radiusTop cannot be negative.")
        if radiusBottom < 0:
            raise ValueError("The cylinder radiusBottom cannot be negative.")
        self.radiusTop, self.radiusBottom, self.height, self.numSegments = float(radiusTop), float(radiusBottom), float(height), int(numSegments)

    def _desc(self):
        return dict(super()._desc(), radius_top=self.radiusTop, radius_bottom=self.radiusBottom, height=self.height, num_segments=self.numSegments)


class ConvexPolyhedron(Shape):  # convex_polyhedron.dart:51
    type = F.SHAPE_CONVEX

    def __init__(self, vertices, faces, axes=None, **kw):
        super().__init__(**kw)
        self.vertices = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)  # Vector3 stores f32
        self.faces = [list(f) for f in faces]
        self.uniqueAxes = None if axes is None else np.asarray(axes, dtype=np.float32).reshape(-1, 3)  # :105; only its presence is used

    def _desc(self):
        d = dict(super()._desc(), vertices=self.vertices, faces=self.faces, convex_has_axes=int(self.uniqueAxes is not None))
        if getattr(self, "_ctor", None):
            d["_ctor"] = self._ctor  # which reference constructor built this hull (tools/reference_golden replays it)
        return d


def _ring(radius, y, theta):  # Vector3(-r sin, y, r cos) as cylinder.dart:56,60 / cone.dart:44,49 write it
    return (-radius * math.sin(theta), y, radius * math.cos(theta))


class Cone(ConvexPolyhedron):  # cone.dart:15-63: apex first, then the base ring; side faces + base; `axes` given
    type = F.SHAPE_CONE

    def __init__(self, radius=1.0, height=1.0, numSegments=8, **kw):
        if radius < 0:
            raise ValueError("The cylinder radiusBottom cannot be negative.")  # the reference's message, cone.dart:32
        self.radius, self.height, self.numSegments = float(radius), float(height), int(numSegments)
        N = self.numSegments
        verts = [(0.0, self.height * 0.5, 0.0), _ring(self.radius, -self.height * 0.5, 0.0)]
        faces, bottom = [], [1]
        for i in range(N):
            theta = ((2 * math.pi) / N) * (i + 1)
            if i < N - 1:
                verts.append(_ring(self.radius, -self.height * 0.5, theta))
                bottom.append(i + 2)
                faces.append([0, i + 2, i + 1])
            else:
                faces.append([0, 1, i + 1])
        faces.append(bottom)
        super().__init__(verts, faces, axes=[(0, 1, 0)], **kw)
        self._ctor = dict(kind="Cone", radius=self.radius, height=self.height, numSegments=self.numSegments)


class Capsule(ConvexPolyhedron):  # capsule.dart:15-170: cylinder walls + two (numSegments+1) x (numHeightSegments+1) vertex grids
    type = F.SHAPE_CAPSULE

    def __init__(self, radiusTop=1.0, radiusBottom=1.0, height=1.0, numSegments=8, numHeightSegments=4, **kw):
        if radiusTop < 0:
            raise ValueError("The Capsule radiusTop cannot be negative.")
        if radiusBottom < 0:
            raise ValueError("The Capsule radiusBottom cannot be negative.")
        self.radiusTop, self.radiusBottom, self.height = float(radiusTop), float(radiusBottom), float(height)
        self.numSegments, self.numHeightSegments = int(numSegments), int(numHeightSegments)
        rt, rb, h, ns, nh = self.radiusTop, self.radiusBottom, self.height, self.numSegments, self.numHeightSegments
        verts = [_ring(rb, -h * 0.5, 0.0), _ring(rt, h * 0.5, 0.0)]
        faces, top, bot, grid, index = [], [], [], [], 0
        phiStart, phiLength, thetaStart, thetaLength = math.pi, 2 * math.pi, math.pi / 2, math.pi / 2
        for iy in range(nh + 1):
            row = []
            v = iy / nh
            for ix in range(ns + 1):
                ub, ut = ix / ns, (ns - ix) / ns
                if iy == 0 and ix < ns:  # the cylinder walls (:72-98)
                    theta = ((2 * math.pi) / ns) * (ix + 1)
                    if ix < ns - 1:
                        verts.append(_ring(rb, -h * 0.5, theta))
                        verts.append(_ring(rt, h * 0.5, theta))
                        faces.append([2 * ix, 2 * ix + 1, 2 * ix + 3, 2 * ix + 2])
                    else:
                        faces.append([2 * ix, 2 * ix + 1, 1, 0])
                # hemisphere vertices (:101-121; the `true ||` makes this the only live branch, both caps use radiusTop)
                st, ct = math.sin(thetaStart + v * thetaLength), math.cos(thetaStart + v * thetaLength)
                top.append((-rt * math.cos(phiStart + ut * phiLength) * st, h * 0.5 - rt * ct, rt * math.sin(phiStart + ut * phiLength) * st))
                bot.append((-rt * math.cos(phiStart + ub * phiLength) * st, -h * 0.5 + rt * ct, rt * math.sin(phiStart + ub * phiLength) * st))
                row.append(index)
                index += 1
            grid.append(row)
        start1, start2 = len(verts), len(verts) + len(top)
        for iy in range(nh):
            for ix in range(ns):
                for start in (start1, start2):  # :150-164, top cap triangle pair then bottom cap pair
                    a, b = grid[iy][ix + 1] + start, grid[iy][ix] + start
                    c, d = grid[iy + 1][ix] + start, grid[iy + 1][ix + 1] + start
                    faces.append([a, b, d])
                    faces.append([b, c, d])
        super().__init__(verts + top + bot, faces, **kw)
        self._ctor = dict(kind="Capsule", radiusTop=rt, radiusBottom=rb, height=h, numSegments=ns, numHeightSegments=nh)


class SizedPlane(ConvexPolyhedron):  # sized_plane.dart:10-37: one quad in the y = 0 plane
    type = F.SHAPE_SIZED_PLANE

    def __init__(self, width=1.0, height=1.0, **kw):
        self.width, self.height = float(width), float(height)
        sx, sz = self.width / 2, self.height / 2
        super().__init__([(-sx, 0, -sz), (sx, 0, -sz), (sx, 0, sz), (-sx, 0, sz)], [[3, 2, 1, 0]], **kw)
        self._ctor = dict(kind="SizedPlane", width=self.width, height=self.height)


class LatheShape(ConvexPolyhedron):  # lathe.dart:6-67: a Vector2 (f32) profile swept around y, two triangles per quad
    def __init__(self, points, numSegments=8, phiStart=0.0, phiLength=math.pi * 2, **kw):
        self.points = np.asarray(points, dtype=np.float32).reshape(-1, 2)
        self.numSegments = int(numSegments)
        phiLength = min(max(phiLength, 0.0), math.pi * 2)
        inverseSegments = 1.0 / self.numSegments
        P = self.points.astype(np.float64)
        n = len(P)
        verts, faces = [], []
        for i in range(self.numSegments + 1):
            phi = phiStart + i * inverseSegments * phiLength
            for j in range(n - 1, -1, -1):
                verts.append((P[j, 0] * math.sin(phi), P[j, 1], P[j, 0] * math.cos(phi)))
        for i in range(self.numSegments):
            for j in range(n - 2, -1, -1):
                base = j + i * n
                a, b, c, d = base, base + n, base + n + 1, base + 1
                faces.append([a, b, d])
                faces.append([c, d, b])
        super().__init__(verts, faces, **kw)
        self._ctor = dict(kind="LatheShape", points=self.points.tolist(), numSegments=self.numSegments, phiStart=phiStart, phiLength=phiLength)


class CapsuleLathe(LatheShape):  # capsule_lathe.dart:14-74: the capsule profile handed to LatheShape, reported as a capsule
    type = F.SHAPE_CAPSULE

    def __init__(self, radiusTop=1.0, radiusBottom=1.0, height=1.0, numSegments=8, numHeightSegments=4, **kw):
        self.radiusTop, self.radiusBottom, self.height = float(radiusTop), float(radiusBottom), float(height)
        rt, rb, h = self.radiusTop, self.radiusBottom, self.height
        ptsTop, ptsBottom = [(0.0, h * 0.5 + rt)], []
        for i in range(numHeightSegments - 1):
            theta = ((math.pi / 2) / numHeightSegments) * (i + 1) + (2 * math.pi + math.pi / 2)
            ptsTop.append((-math.cos(theta) * rt, h * 0.5 + math.sin(theta) * rt))
            ptsBottom.insert(0, (-math.cos(theta) * rb, -h * 0.5 - math.sin(theta) * rb))
        ptsTop.append((rt, h * 0.5))
        ptsBottom.append((0.0, -h * 0.5 - rb))
        ptsBottom.insert(0, (rb, -h * 0.5))
        super().__init__(ptsTop + ptsBottom, numSegments=numSegments, **kw)
        self._ctor = dict(kind="CapsuleLathe", radiusTop=rt, radiusBottom=rb, height=h, numSegments=int(numSegments), numHeightSegments=int(numHeightSegments))


class Particle(Shape):  # particle.dart:9: a point (bounding radius 0, zero inertia, AABB = its position)
    type = F.SHAPE_PARTICLE


class Trimesh(Shape):  # trimesh.dart:37 (sphere and plane contacts; the reference's other trimesh resolvers are unfinished)
    type = F.SHAPE_TRIMESH

    def __init__(self, vertices, indices, **kw):
        super().__init__(**kw)
        self.vertices = np.asarray(vertices, dtype=np.float64).reshape(-1, 3)
        self.indices = np.asarray(indices, dtype=np.int32).reshape(-1)
        self.scale = np.ones(3, np.float32)

    def setScale(self, scale):  # trimesh.dart:142-152
        self.scale[:] = scale

    def _desc(self):
        return dict(super()._desc(), vertices=self.vertices.astype(np.float32), tm_indices=self.indices, tm_scale=self.scale)

    @staticmethod
    def createTorus(radius=1.0, tube=0.5, radialSegments=8, tubularSegments=6, arc=math.pi * 2):  # trimesh.dart:382-443
        verts, idx = [], []
        for j in range(radialSegments + 1):
            for i in range(tubularSegments + 1):
                u = i / tubularSegments * arc
                v = j / radialSegments * math.pi * 2
                verts.append(np.array([(radius + tube * math.cos(v)) * math.cos(u), (radius + tube * math.cos(v)) * math.sin(u), tube * math.sin(v)], np.float32))
                if i != 0 and j != 0:
                    a = (tubularSegments + 1) * j + i - 1
                    b = (tubularSegments + 1) * (j - 1) + i - 1
                    c = (tubularSegments + 1) * (j - 1) + i
                    d = (tubularSegments + 1) * j + i
                    idx += [a, b, d, b, c, d]
        return Trimesh(np.array(verts, np.float64), idx)


class SPHSystem:  # lib/objects/sph_system.dart:6 (append to World.subsystems)
    def __init__(self):
        self.particles: List["Body"] = []
        self.density, self.smoothingRadius, self.speedOfSound, self.viscosity, self.eps = 1.0, 1.0, 1.0, 0.01, 0.00001
        self._world: Optional["World"] = None

    def add(self, particle: "Body"):
        self.particles.append(particle)
        if self._world is not None:
            self._world._structure_dirty = True

    def remove(self, particle: "Body"):
        if particle in self.particles:
            self.particles.remove(particle)
            if self._world is not None:
                self._world._structure_dirty = True

    def _desc(self, idx):
        return dict(particles=[idx[id(p)] for p in self.particles], density=self.density, smoothing_radius=self.smoothingRadius,
                    speed_of_sound=self.speedOfSound, viscosity=self.viscosity, eps=self.eps)


class Heightfield(Shape):  # heightfield.dart:34
    type = F.SHAPE_HEIGHTFIELD

    def __init__(self, data, elementSize: int = 1, **kw):
        super().__init__(**kw)
        self.data = np.asarray(data, dtype=np.float64)
        self.elementSize = int(elementSize)

    def _desc(self):
        return dict(super()._desc(), hf_data=self.data, hf_element_size=self.elementSize)


class Body:  # lib/objects/rigid_body.dart:26-86
    def __init__(self, collisionFilterGroup=1, collisionFilterMask=-1, collisionResponse=True, position=None, velocity=None, mass=0.0,
                 material: Optional[Material] = None, linearDamping=0.01, type=None, allowSleep=True, sleepSpeedLimit=0.1, sleepTimeLimit=1.0,
                 quaternion=None, angularVelocity=None, fixedRotation=False, angularDamping=0.01, linearFactor=None, angularFactor=None,
                 shape: Optional[Shape] = None, isTrigger=False):
        f32 = lambda v, d: np.array(d if v is None else v, dtype=np.float32)
        self.position, self.velocity = f32(position, (0, 0, 0)), f32(velocity, (0, 0, 0))
        self.quaternion, self.angularVelocity = f32(quaternion, (0, 0, 0, 1)), f32(angularVelocity, (0, 0, 0))
        self.force, self.torque = Vec3(), Vec3()
        self.linearFactor, self.angularFactor = f32(linearFactor, (1, 1, 1)), f32(angularFactor, (1, 1, 1))
        self.mass = float(mass)
        self.type = (BodyTypes.static if self.mass <= 0 else BodyTypes.dynamic) if type is None else type
        self.material = material
        self.linearDamping, self.angularDamping = linearDamping, angularDamping
        self.allowSleep, self.sleepSpeedLimit, self.sleepTimeLimit = allowSleep, sleepSpeedLimit, sleepTimeLimit
        self.sleepState = BodySleepStates.awake
        self.fixedRotation, self.isTrigger = fixedRotation, isTrigger
        self.collisionFilterGroup, self.collisionFilterMask, self.collisionResponse = collisionFilterGroup, collisionFilterMask, collisionResponse
        self.shapes: List[Shape] = []
        self.shapeOffsets: List[np.ndarray] = []       # rigid_body.dart:98-104
        self.shapeOrientations: List[np.ndarray] = []
        self.world: Optional["World"] = None
        self.index = -1
        self.timeLastSleepy = 0.0      # rigid_body.dart:130; World.addBody sets it to world.time (world_class.dart:291)
        self._invInertia = None        # Body.invInertia as the device derived it at the first upload (rigid_body.dart:587-609)
        if shape is not None:
            self.addShape(shape)

    def addShape(self, shape: Shape, offset=None, orientation=None) -> "Body":  # rigid_body.dart:348-369
        self._before_write()
        self.shapes.append(shape)
        self.shapeOffsets.append(np.zeros(3, np.float32) if offset is None else np.asarray(offset, dtype=np.float32).copy())
        self.shapeOrientations.append(np.array([0, 0, 0, 1], np.float32) if orientation is None else np.asarray(orientation, dtype=np.float32).copy())
        self._invInertia = None  # addShape recomputes the mass properties (rigid_body.dart:362)
        if self.world is not None:
            self.world._structure_dirty = True
        return self

    def removeShape(self, shape: Shape) -> "Body":  # rigid_body.dart:371-391
        for k, s in enumerate(self.shapes):
            if s is shape:
                self._before_write()
                del self.shapes[k], self.shapeOffsets[k], self.shapeOrientations[k]
                self._invInertia = None  # updateMassProperties (:383)
                if self.world is not None:
                    self.world._structure_dirty = True
                return self
        return self  # "Shape does not belong to the body": the reference logs a warning and returns

    def _before_write(self):
        """Every mutator starts here: after World.step(sync=False) the device holds the newer state, so it is pulled into
        the Body objects BEFORE the write - otherwise the next upload would roll the world back to stale host values."""
        if self.world is not None:
            self.world._pull()

    # rigid_body.dart:263-278: sleepState (and for sleep() the velocities) only - no rebuild of the device world
    def wakeUp(self):
        self._before_write()
        self.sleepState = BodySleepStates.awake
        if self.world is not None:
            self.world._sleep_dirty = True

    def sleep(self):
        self._before_write()
        self.sleepState = BodySleepStates.sleeping
        self.velocity[:] = 0
        self.angularVelocity[:] = 0
        if self.world is not None:
            self.world._sleep_dirty = True
            self.world._state_dirty = True

    # rigid_body.dart:472-566 (host-side writes, uploaded before the next step)
    def applyForce(self, force, relativePoint=None):
        if self.type != BodyTypes.dynamic:
            return
        self._before_write()
        f = np.asarray(force, dtype=np.float32)
        r = Vec3() if relativePoint is None else np.asarray(relativePoint, dtype=np.float32)
        if self.sleepState == BodySleepStates.sleeping:
            self.wakeUp()
        rot = np.cross(r.astype(np.float64), f.astype(np.float64)).astype(np.float32)
        self.force[:] = (self.force.astype(np.float64) + f).astype(np.float32)
        self.torque[:] = (self.torque.astype(np.float64) + rot).astype(np.float32)
        if self.world is not None:
            self.world._state_dirty = True

    def applyImpulse(self, impulse, relativePoint=None):
        if self.type != BodyTypes.dynamic:
            return
        if relativePoint is not None and np.any(np.asarray(relativePoint) != 0):
            raise CannonError(F.E_UNSUPPORTED, "off-centre impulses need the device inertia; write angularVelocity directly")
        self._before_write()
        if self.sleepState == BodySleepStates.sleeping:
            self.wakeUp()
        j = np.asarray(impulse, dtype=np.float32).astype(np.float64)
        velo = (j * (1.0 / self.mass)).astype(np.float32)
        self.velocity[:] = (self.velocity.astype(np.float64) + velo).astype(np.float32)
        if self.world is not None:
            self.world._state_dirty = True


class Broadphase:  # lib/collision/broadphase.dart:11
    kind = F.BP_NAIVE

    def __init__(self, useBoundingBoxes: bool = False):
        self.useBoundingBoxes = useBoundingBoxes
        self.world = None

    def collisionPairs(self, world: "World"):
        """Broadphase.collisionPairs(world, p1, p2): returns the two parallel index lists."""
        world._ensure_uploaded()
        return world._dev.broadphase_pairs()


class NaiveBroadphase(Broadphase):  # naive_broadphase.dart:10
    kind = F.BP_NAIVE


class SAPBroadphase(Broadphase):  # sap_broadphase.dart:10
    kind = F.BP_SAP

    def __init__(self, world=None, axisIndex: int = 0, **kw):
        super().__init__(**kw)
        self.axisIndex = axisIndex


class GridBroadphase(Broadphase):  # grid_broadphase.dart:12
    kind = F.BP_GRID

    def __init__(self, aabbMin=(100, 100, 100), aabbMax=(-100, -100, -100), nx=10, ny=10, nz=10, **kw):
        super().__init__(**kw)
        if nx * ny * nz <= 0:
            raise ValueError("GridBroadphase: Each dimension's n must be >0")
        self.aabbMin, self.aabbMax, self.nx, self.ny, self.nz = np.asarray(aabbMin, np.float32), np.asarray(aabbMax, np.float32), nx, ny, nz


CudaBroadphase = NaiveBroadphase  # drop-in name used by north_star; the kind is chosen by the subclass


class Solver:  # lib/solver/solver.dart:5
    kind = F.SOLVER_REFERENCE_ORDER

    def __init__(self, iterations: int = 10, tolerance: float = 1e-7):
        self.iterations, self.tolerance = iterations, tolerance


class GSSolver(Solver):  # gs_solver.dart:7 — reference equation order, bit-reproducible
    kind = F.SOLVER_REFERENCE_ORDER


class CudaGSSolver(Solver):  # graph-coloured throughput mode
    kind = F.SOLVER_COLORED


class SplitSolver(Solver):  # lib/solver/split_solver.dart:32: islands + one GSSolver pass per island
    kind = F.SOLVER_SPLIT

    def __init__(self, subsolver: Optional[Solver] = None):
        sub = subsolver or GSSolver()
        super().__init__(iterations=sub.iterations, tolerance=sub.tolerance)
        self.subsolver = sub


class Constraint:  # constraint_class.dart:5
    type = -1

    def __init__(self, bodyA: Body, bodyB: Body, collideConnected: bool = True):
        self.bodyA, self.bodyB, self.collideConnected = bodyA, bodyB, collideConnected
        self.world: Optional["World"] = None  # set by World.addConstraint


class PointToPointConstraint(Constraint):  # point_to_point_constraint.dart:20
    type = F.CONSTRAINT_POINT_TO_POINT

    def __init__(self, bodyA, bodyB, pivotA=None, pivotB=None, maxForce: float = 1e6):
        super().__init__(bodyA, bodyB)
        self.pivotA = Vec3() if pivotA is None else np.array(pivotA, dtype=np.float32)
        self.pivotB = Vec3() if pivotB is None else np.array(pivotB, dtype=np.float32)
        self.maxForce = maxForce

    def _desc(self, idx):
        return dict(type=self.type, body_a=idx[id(self.bodyA)], body_b=idx[id(self.bodyB)], pivot_a=self.pivotA, pivot_b=self.pivotB,
                    max_force=self.maxForce, collide_connected=int(self.collideConnected))


class HingeConstraint(PointToPointConstraint):  # hinge_constraint.dart:10
    type = F.CONSTRAINT_HINGE

    def __init__(self, bodyA, bodyB, pivotA=None, pivotB=None, axisA=None, axisB=None, collideConnected=None, maxForce: float = 1e6):
        super().__init__(bodyA, bodyB, pivotA, pivotB, maxForce)
        self.axisA = Vec3(1, 0, 0) if axisA is None else np.array(axisA, dtype=np.float32)
        self.axisB = Vec3(1, 0, 0) if axisB is None else np.array(axisB, dtype=np.float32)
        self.collideConnected = True if collideConnected is None else collideConnected
        self.motorEnabled, self.motorTargetVelocity, self.motorMaxForce = False, 0.0, maxForce

    # hinge_constraint.dart:56-76: the motor equation's fields, effective from the next step. On a live device world they go
    # through cannon_world_set_hinge_motor (no rebuild, nothing else changes).
    def _motor_changed(self):
        if self.world is not None:
            self.world._motor_dirty.add(id(self))

    def enableMotor(self):
        self.motorEnabled = True
        self._motor_changed()

    def disableMotor(self):
        self.motorEnabled = False
        self._motor_changed()

    def setMotorSpeed(self, speed: float):
        self.motorTargetVelocity = speed
        self._motor_changed()

    def setMotorMaxForce(self, maxForce: float):
        self.motorMaxForce = maxForce
        self._motor_changed()

    def _desc(self, idx):
        return dict(super()._desc(idx), axis_a=self.axisA, axis_b=self.axisB, motor_enabled=int(self.motorEnabled),
                    motor_target_velocity=self.motorTargetVelocity, motor_max_force=self.motorMaxForce)


class DistanceConstraint(Constraint):  # distance_constraint.dart:7
    type = F.CONSTRAINT_DISTANCE

    def __init__(self, bodyA, bodyB, distance: Optional[float] = None, maxForce: float = 1e6):
        super().__init__(bodyA, bodyB)
        # distance_constraint.dart:14-18: the default is the bodies' distance when the constraint is constructed; here the
        # value is frozen when the constraint reaches the world (World.addConstraint), so later rebuilds of the device world
        # keep it
        self.distance = distance
        self.maxForce = maxForce

    def _desc(self, idx):
        return dict(type=self.type, body_a=idx[id(self.bodyA)], body_b=idx[id(self.bodyB)], max_force=self.maxForce,
                    collide_connected=int(self.collideConnected), distance=-1.0 if self.distance is None else float(self.distance))


class SpringConstraint(DistanceConstraint):  # spring_constraint.dart:7-56
    """One bidirectional ContactEquation that holds the bodies at the distance they have when the constraint is made, its
    force bounded by +-stiffness (spring_constraint.dart:34-41); update() is DistanceConstraint's (:44-55). `damping` is
    stored and never read, like in the reference. On the device it is a distance row with max_force = stiffness."""

    def __init__(self, bodyA, bodyB, stiffness: float = 1.0, damping: float = 1.0):
        super().__init__(bodyA, bodyB, None, stiffness)
        self.stiffness, self.damping = stiffness, damping

    def _desc(self, idx):
        self.maxForce = self.stiffness
        return super()._desc(idx)


class LockConstraint(PointToPointConstraint):  # lock_constraint.dart:9
    """The pivots and the frame vectors are taken from the bodies' poses when the constraint reaches the world
    (lock_constraint.dart:29-43), quirks of Body.vectorToLocalFrame included."""
    type = F.CONSTRAINT_LOCK

    def __init__(self, bodyA, bodyB, maxForce: float = 1e6):
        super().__init__(bodyA, bodyB, None, None, maxForce)
        self._ctor_pose = None

    def _desc(self, idx):
        d = super()._desc(idx)
        if self._ctor_pose is not None:
            pa, qa, pb, qb = self._ctor_pose
            d.update(has_ctor_pose=1, ctor_pos_a=pa, ctor_quat_a=qa, ctor_pos_b=pb, ctor_quat_b=qb)
        return d


class ConeTwistConstraint(PointToPointConstraint):  # cone_twist_constraint.dart:11
    type = F.CONSTRAINT_CONE_TWIST

    def __init__(self, bodyA, bodyB, pivotA=None, pivotB=None, axisA=None, axisB=None, angle: float = 0.0, twistAngle: float = 0.0,
                 maxForce: float = 1e6, collideConnected: bool = False):
        super().__init__(bodyA, bodyB, pivotA, pivotB, maxForce)
        # the reference accepts collideConnected but never forwards it to Constraint (cone_twist_constraint.dart:38-40)
        self.axisA = Vec3() if axisA is None else np.array(axisA, dtype=np.float32)
        self.axisB = Vec3() if axisB is None else np.array(axisB, dtype=np.float32)
        self.angle, self.twistAngle = angle, twistAngle

    def _desc(self, idx):
        return dict(super()._desc(idx), axis_a=self.axisA, axis_b=self.axisB, angle=float(self.angle), twist_angle=float(self.twistAngle))


class Spring:  # lib/objects/spring.dart:17
    """`world.addSpring(spring)` stands for the reference idiom `world.addEventListener('postStep', (e) => spring.applyForce())`
    (examples/lib/examples/spring.dart:90): the device applies the spring in the postStep slot of every step."""

    def __init__(self, bodyA, bodyB, restLength: float = 1.0, stiffness: float = 100.0, damping: float = 1.0, localAnchorA=None, localAnchorB=None):
        self.bodyA, self.bodyB = bodyA, bodyB
        self.restLength, self.stiffness, self.damping = restLength, stiffness, damping
        self.localAnchorA = Vec3() if localAnchorA is None else np.array(localAnchorA, dtype=np.float32)
        self.localAnchorB = Vec3() if localAnchorB is None else np.array(localAnchorB, dtype=np.float32)

    def _desc(self, idx):
        return dict(body_a=idx[id(self.bodyA)], body_b=idx[id(self.bodyB)], rest_length=float(self.restLength), stiffness=float(self.stiffness),
                    damping=float(self.damping), local_anchor_a=self.localAnchorA, local_anchor_b=self.localAnchorB)


class RaycastResult:  # lib/collision/raycast_result.dart:6
    def __init__(self):
        self.rayFromWorld, self.rayToWorld, self.hitNormalWorld, self.hitPointWorld = Vec3(), Vec3(), Vec3(), Vec3()
        self.reset()

    def reset(self):
        for v in (self.rayFromWorld, self.rayToWorld, self.hitNormalWorld, self.hitPointWorld):
            v[:] = 0
        self.hasHit, self.shape, self.body, self.hitFaceIndex, self.distance, self.shouldStop = False, None, None, -1, -1.0, False


class World:  # lib/world/world_class.dart:44
    def __init__(self, gravity=None, frictionGravity=None, allowSleep: bool = False, broadphase: Optional[Broadphase] = None,
                 solver: Optional[Solver] = None, quatNormalizeFast: bool = False, quatNormalizeSkip: int = 0, device: int = 0, _lib=None):
        self.gravity = Vec3() if gravity is None else np.array(gravity, dtype=np.float32)
        self.frictionGravity = None if frictionGravity is None else np.array(frictionGravity, dtype=np.float32)
        self.allowSleep = allowSleep
        self.broadphase = broadphase or NaiveBroadphase()
        self.solver = solver or GSSolver()
        self.quatNormalizeFast, self.quatNormalizeSkip = quatNormalizeFast, quatNormalizeSkip
        self.bodies: List[Body] = []
        self.constraints: List[Constraint] = []
        self.springs: List[Spring] = []
        self.subsystems: List["SPHSystem"] = []  # world_class.dart:121; a plain list in the reference, so changes are detected by signature
        self._sub_sig = ()
        self.contactmaterials: List[ContactMaterial] = []
        self.defaultMaterial = Material(name="default")
        self.defaultContactMaterial = ContactMaterial(self.defaultMaterial, self.defaultMaterial, friction=0.3, restitution=0.0)
        self.time, self.stepnumber, self.dt = 0.0, 0, -1.0
        self._device = device
        if _lib is None:
            from . import load_library
            _lib = load_library()
        self._lib = _lib
        self._dev: Optional[DeviceWorld] = None
        self._structure_dirty = True
        self._state_dirty = False
        self._sleep_dirty = False        # Body.sleep / wakeUp since the last upload
        self._motor_dirty = set()        # ids of HingeConstraints whose motor fields changed since the last upload
        self._listeners = {}
        self._events_on = False
        self._host_current = True  # the Body objects hold the device state (False after a step with sync=False)

    # world_class.dart:282-300 / 224-231 / 343-348
    def _pull(self):
        """Before any host-side write (structural or not): make the Body objects current. The body list cannot have changed
        since the device world was built while the host is stale, because every mutator pulls before it writes."""
        if self._dev is not None and not self._host_current:
            self.sync()

    def addBody(self, body: Body):
        if body in self.bodies:
            return
        self._pull()
        body.index = len(self.bodies)
        body.world = self
        body.timeLastSleepy = self.time  # world_class.dart:291
        self.bodies.append(body)
        self._structure_dirty = True

    def removeBody(self, body: Body):
        """World.removeBody (world_class.dart:303-320): the remaining bodies are re-indexed; the device world is rebuilt
        from the current body state before the next step (constraints that still reference the body must be removed
        first, the upload refuses dangling references)."""
        if body not in self.bodies:
            return
        self._pull()  # the Body objects carry the state the rebuilt device world starts from
        self.bodies.remove(body)
        body.world = None
        body.index = -1
        for i, b in enumerate(self.bodies):
            b.index = i
        self._structure_dirty = True

    def addConstraint(self, c: Constraint):
        self._pull()
        c.world = self
        if isinstance(c, LockConstraint) and c._ctor_pose is None:
            # lock_constraint.dart:29-43 derives pivots and frame vectors from the poses its constructor sees: recorded now,
            # so a later rebuild of the device world (addBody ...) reproduces the same constraint instead of re-locking the
            # bodies in whatever pose they have drifted to
            c._ctor_pose = (c.bodyA.position.copy(), c.bodyA.quaternion.copy(), c.bodyB.position.copy(), c.bodyB.quaternion.copy())
        if isinstance(c, DistanceConstraint) and c.distance is None:
            # Vector3.distanceTo on the f32-stored positions, evaluated in double (distance_constraint.dart:16)
            d = c.bodyA.position.astype(np.float64) - c.bodyB.position.astype(np.float64)
            c.distance = float(np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]))
        self.constraints.append(c)
        self._structure_dirty = True

    def removeConstraint(self, c: Constraint):  # world_class.dart:234-236
        if c in self.constraints:
            self._pull()
            self.constraints.remove(c)
            self._structure_dirty = True

    def clearForces(self):  # world_class.dart:773-781
        self._pull()
        for b in self.bodies:
            b.force[:] = 0
            b.torque[:] = 0
        self._state_dirty = True

    def addSpring(self, spring: Spring):
        self._pull()
        self.springs.append(spring)
        self._structure_dirty = True

    def addContactMaterial(self, cmat: ContactMaterial):
        self._pull()
        self.contactmaterials.append(cmat)
        self._structure_dirty = True

    def _spec(self) -> SceneSpec:
        n = len(self.bodies)
        shapes, shape_ids, mats, mat_ids = [], {}, [], {}

        def mat_index(m):
            if m is None:
                return -1
            if id(m) not in mat_ids:
                mat_ids[id(m)] = len(mats)
                mats.append(m)
            return mat_ids[id(m)]

        for cm in self.contactmaterials:
            for m in cm.materials:
                mat_index(m)
        b = {
            "position": np.zeros((n, 3), np.float32), "quaternion": np.zeros((n, 4), np.float32), "velocity": np.zeros((n, 3), np.float32),
            "angular_velocity": np.zeros((n, 3), np.float32), "force": np.zeros((n, 3), np.float32), "torque": np.zeros((n, 3), np.float32),
            "mass": np.zeros(n), "type": np.zeros(n, np.int32), "sleep_state": np.zeros(n, np.int32), "allow_sleep": np.zeros(n, np.uint8),
            "sleep_speed_limit": np.zeros(n), "sleep_time_limit": np.zeros(n), "linear_damping": np.zeros(n), "angular_damping": np.zeros(n),
            "linear_factor": np.zeros((n, 3), np.float32), "angular_factor": np.zeros((n, 3), np.float32), "fixed_rotation": np.zeros(n, np.uint8),
            "collision_filter_group": np.zeros(n, np.int32), "collision_filter_mask": np.zeros(n, np.int32),
            "collision_response": np.zeros(n, np.uint8), "is_trigger": np.zeros(n, np.uint8), "material": np.zeros(n, np.int32),
            "shape": np.zeros(n, np.int32), "time_last_sleepy": np.zeros(n),
        }
        inst_first, inst_shape, inst_off, inst_ori = [0], [], [], []
        for i, body in enumerate(self.bodies):
            b["position"][i], b["quaternion"][i], b["velocity"][i] = body.position, body.quaternion, body.velocity
            b["angular_velocity"][i], b["force"][i], b["torque"][i] = body.angularVelocity, body.force, body.torque
            b["mass"][i], b["type"][i], b["sleep_state"][i] = body.mass, body.type, body.sleepState
            b["allow_sleep"][i], b["sleep_speed_limit"][i], b["sleep_time_limit"][i] = body.allowSleep, body.sleepSpeedLimit, body.sleepTimeLimit
            b["linear_damping"][i], b["angular_damping"][i] = body.linearDamping, body.angularDamping
            b["linear_factor"][i], b["angular_factor"][i], b["fixed_rotation"][i] = body.linearFactor, body.angularFactor, body.fixedRotation
            b["collision_filter_group"][i], b["collision_filter_mask"][i] = body.collisionFilterGroup, body.collisionFilterMask
            b["collision_response"][i], b["is_trigger"][i] = body.collisionResponse, body.isTrigger
            b["material"][i] = mat_index(body.material)
            b["time_last_sleepy"][i] = body.timeLastSleepy
            for k, sh in enumerate(body.shapes):
                if id(sh) not in shape_ids:
                    shape_ids[id(sh)] = len(shapes)
                    shapes.append(dict(sh._desc(), material=mat_index(getattr(sh, "material", None))))
                inst_shape.append(shape_ids[id(sh)])
                inst_off.append(body.shapeOffsets[k])
                inst_ori.append(body.shapeOrientations[k])
            inst_first.append(len(inst_shape))
            b["shape"][i] = inst_shape[inst_first[i]] if body.shapes else -1
        bp = self.broadphase
        desc = dict(gravity=self.gravity, allow_sleep=int(self.allowSleep), quat_normalize_skip=self.quatNormalizeSkip,
                    quat_normalize_fast=int(self.quatNormalizeFast), solver_kind=self.solver.kind, solver_iterations=self.solver.iterations,
                    solver_tolerance=self.solver.tolerance, broadphase_kind=bp.kind, use_bounding_boxes=int(bp.useBoundingBoxes))
        if self.frictionGravity is not None:
            desc.update(friction_gravity=self.frictionGravity, has_friction_gravity=1)
        if isinstance(bp, SAPBroadphase):
            desc["sap_axis"] = bp.axisIndex
        if isinstance(bp, GridBroadphase):
            desc.update(grid_min=bp.aabbMin, grid_max=bp.aabbMax, grid_nx=bp.nx, grid_ny=bp.ny, grid_nz=bp.nz)
        d = self.defaultContactMaterial
        desc["default_contact_material"] = dict(friction=d.friction, restitution=d.restitution, contact_equation_stiffness=d.contactEquationStiffness,
                                                contact_equation_relaxation=d.contactEquationRelaxation,
                                                friction_equation_stiffness=d.frictionEquationStiffness,
                                                friction_equation_relaxation=d.frictionEquationRelaxation)
        cms = [dict(material_a=mat_ids[id(c.materials[0])], material_b=mat_ids[id(c.materials[1])], friction=c.friction, restitution=c.restitution,
                    contact_equation_stiffness=c.contactEquationStiffness, contact_equation_relaxation=c.contactEquationRelaxation,
                    friction_equation_stiffness=c.frictionEquationStiffness, friction_equation_relaxation=c.frictionEquationRelaxation)
               for c in self.contactmaterials]
        idx = {id(body): i for i, body in enumerate(self.bodies)}
        for c in list(self.constraints) + list(self.springs):
            if id(c.bodyA) not in idx or id(c.bodyB) not in idx:
                raise CannonError(F.E_INVALID, "a constraint or spring references a body that is not in the world (remove it first)")
        cons = [c._desc(idx) for c in self.constraints]
        # the shape table is only needed when some body is not "one shape at its origin"
        plain = all(len(body.shapes) == 1 and not np.any(body.shapeOffsets[0] != 0) and np.array_equal(body.shapeOrientations[0], np.array([0, 0, 0, 1], np.float32))
                    for body in self.bodies if True) if self.bodies else True
        body_shapes = None if plain else dict(first=np.array(inst_first, np.int32), shape=np.array(inst_shape, np.int32),
                                              offset=np.array(inst_off, np.float32).reshape(-1, 3), orientation=np.array(inst_ori, np.float32).reshape(-1, 4))
        for sub_ in self.subsystems:
            for p_ in sub_.particles:
                if id(p_) not in idx:
                    raise CannonError(F.E_INVALID, "an SPH particle is not a body of the world")
        return SceneSpec(desc=desc, shapes=shapes, bodies=b, n_bodies=n, body_shapes=body_shapes, sph_systems=[s_._desc(idx) for s_ in self.subsystems],
                         material_friction=np.array([m.friction for m in mats], dtype=np.float64) if mats else None,
                         material_restitution=np.array([m.restitution for m in mats], dtype=np.float64) if mats else None,
                         contact_materials=cms, constraints=cons, springs=[sp._desc(idx) for sp in self.springs], name="api_world")

    def _ensure_uploaded(self):
        sig = tuple((id(s_), tuple(id(p_) for p_ in s_.particles), s_.density, s_.smoothingRadius, s_.speedOfSound, s_.viscosity, s_.eps) for s_ in self.subsystems)
        if sig != self._sub_sig:
            if self._dev is not None:
                self._pull()
            self._structure_dirty, self._sub_sig = True, sig
        if self._structure_dirty or self._dev is None:
            spec = self._spec()  # may refuse (dangling constraint): the old device world stays usable
            if self._dev is not None:
                self._dev.close()
            self._dev = DeviceWorld(self._lib, spec, device=self._device)
            # what a rebuild must not lose: World.time / stepnumber (the quatNormalizeSkip phase) and every body's
            # invInertia, which the reference computes once at construction (rigid_body.dart:85,362), not from the pose
            # the body happens to have when the device world is rebuilt
            self._dev.set_time(self.time)
            self._dev.set_stepnumber(self.stepnumber)
            n = len(self.bodies)
            if n:
                inv = self._dev.get_bodies(("inv_inertia",))["inv_inertia"].reshape(n, 3)
                carried = False
                for i, body in enumerate(self.bodies):
                    if body._invInertia is None:
                        body._invInertia = inv[i].copy()
                    elif not np.array_equal(body._invInertia, inv[i]):
                        inv[i] = body._invInertia
                        carried = True
                if carried:
                    self._dev.set_inv_inertia(0, inv)
            if self._events_on:
                self._dev.enable_contact_events(True)
            self._structure_dirty = False
            self._state_dirty = False
            self._sleep_dirty = False
            self._motor_dirty.clear()
            return
        # Only ever reached with the Body objects current (every mutator pulls first), so no stale pose can be uploaded.
        assert self._host_current or not (self._state_dirty or self._sleep_dirty)
        if self._state_dirty:
            n = len(self.bodies)
            st = lambda attr: np.stack([getattr(b, attr) for b in self.bodies]).astype(np.float32)
            self._dev.update_bodies(0, n, position=st("position"), quaternion=st("quaternion"), velocity=st("velocity"),
                                    angular_velocity=st("angularVelocity"), force=st("force"), torque=st("torque"))
            self._state_dirty = False
        if self._sleep_dirty:
            self._dev.update_sleep_states(0, np.array([b.sleepState for b in self.bodies], dtype=np.int32))
            self._sleep_dirty = False
        if self._motor_dirty:
            for k, c in enumerate(self.constraints):
                if id(c) in self._motor_dirty:
                    self._dev.set_hinge_motor(k, c.motorEnabled, c.motorTargetVelocity, c.motorMaxForce)
            self._motor_dirty.clear()

    # EventTarget (lib/utils/event_target.dart) for the world-level contact events of world_class.dart:703-730
    def addEventListener(self, type: str, listener):
        """``beginContact`` / ``endContact``: ``listener(event)`` with ``event = {"type", "bodyA", "bodyB"}`` (bodyA is
        the body with the smaller index, like OverlapKeeper's unpacked key), dispatched after the step that changed the
        contact state. Listening switches the device-side pair-set tracking on from the next step."""
        if type not in ("beginContact", "endContact"):
            raise CannonError(F.E_UNSUPPORTED, f"event '{type}' is outside the hot-path scope (world-level contact events only)")
        self._listeners.setdefault(type, []).append(listener)
        if self._dev is not None and not self._events_on:
            self._dev.enable_contact_events(True)
        self._events_on = True

    def hasAnyEventListener(self, type: str) -> bool:
        return bool(self._listeners.get(type))

    def removeEventListener(self, type: str, listener):
        if listener in self._listeners.get(type, []):
            self._listeners[type].remove(listener)

    def _emit_contact_events(self):
        begin, end = self._dev.get_contact_events()
        for type, pairs in (("beginContact", begin), ("endContact", end)):  # additions first (world_class.dart:710-727)
            for a, b in pairs:
                ev = {"type": type, "bodyA": self.bodies[int(a)], "bodyB": self.bodies[int(b)]}
                for fn in list(self._listeners.get(type, [])):
                    fn(ev)

    def markDirty(self):
        """Call after writing body vectors in place (``body.position[:] = ...``) between steps. After ``step(sync=False)``
        the Body objects are stale: call ``world.sync()`` BEFORE editing them - an in-place edit of stale vectors cannot be
        told apart from the stale values around it, so this refuses instead of rolling the world back."""
        if self._dev is not None and not self._host_current:
            raise CannonError(F.E_INVALID, "the Body objects are stale after step(sync=False): call world.sync() before editing them")
        self._state_dirty = True

    def step(self, dt: float, timeSinceLastCalled: Optional[float] = None, maxSubSteps: int = 10, nsteps: int = 1, sync: bool = True):
        """World.step(dt) in fixed-stepping mode (world_class.dart:392-399)."""
        if timeSinceLastCalled is not None:
            raise CannonError(F.E_UNSUPPORTED, "interpolated stepping is host-side glue outside the hot-path scope")
        self._ensure_uploaded()
        if self._events_on and nsteps > 1:  # listeners hear every step, like the reference's synchronous dispatch
            for _ in range(nsteps):
                self._dev.step(dt, 1)
                self._emit_contact_events()
        else:
            self._dev.step(dt, nsteps)
            if self._events_on:
                self._emit_contact_events()
        self.dt = dt
        self.time, self.stepnumber = self._dev.get_time()
        self._host_current = False
        if sync:
            self.sync()

    def sync(self):
        """Refresh the Body objects from device state."""
        st = self._dev.get_bodies(("position", "quaternion", "velocity", "angular_velocity", "sleep_state", "time_last_sleepy"))
        for i, body in enumerate(self.bodies):
            body.timeLastSleepy = float(st["time_last_sleepy"][i])
            body.position[:] = st["position"][i]
            body.quaternion[:] = st["quaternion"][i]
            body.velocity[:] = st["velocity"][i]
            body.angularVelocity[:] = st["angular_velocity"][i]
            body.sleepState = int(st["sleep_state"][i])
            body.force[:] = 0
            body.torque[:] = 0
        self._host_current = True

    # ---- ray casts (world_class.dart:248-277) ------------------------------------------------------
    def _raycast(self, mode, from_, to, options, result):
        self._ensure_uploaded()
        o = options or {}
        r = self._dev.raycast([from_], [to], mode=mode, skip_backfaces=o.get("skipBackfaces", True),
                              collision_filter_mask=o.get("collisionFilterMask", -1), collision_filter_group=o.get("collisionFilterGroup", -1),
                              check_collision_response=o.get("checkCollisionResponse", True))
        return r

    def _fill(self, result: "RaycastResult", r, k, from_, to):
        result.rayFromWorld[:], result.rayToWorld[:] = from_, to
        result.hasHit = True
        result.body = self.bodies[int(r["body"][k])]
        result.shape = result.body.shapes[int(r["shape_ordinal"][k])]  # RaycastResult.shape (ray_class.dart:676-690)
        result.hitFaceIndex, result.distance = int(r["hit_face_index"][k]), float(r["distance"][k])
        result.hitPointWorld[:], result.hitNormalWorld[:] = r["hit_point_world"][k], r["hit_normal_world"][k]

    def raycastClosest(self, from_, to, options=None, result: Optional["RaycastResult"] = None) -> bool:
        result = result if result is not None else RaycastResult()
        result.reset()
        r = self._raycast(F.RAY_CLOSEST, from_, to, options, result)
        if r["has_hit"][0]:
            self._fill(result, r, 0, from_, to)
        return bool(r["has_hit"][0])

    def raycastAny(self, from_, to, options=None, result: Optional["RaycastResult"] = None) -> bool:
        result = result if result is not None else RaycastResult()
        result.reset()
        r = self._raycast(F.RAY_ANY, from_, to, options, result)
        if r["has_hit"][0]:
            self._fill(result, r, 0, from_, to)
        return bool(r["has_hit"][0])

    def raycastAll(self, from_, to, options=None, callback=None) -> bool:
        r = self._raycast(F.RAY_ALL, from_, to, options, None)
        for k in range(r["n_hits"]):
            res = RaycastResult()
            self._fill(res, r, k, from_, to)
            if callback is not None:
                callback(res)
        return bool(r["has_hit"][0])

    def aabbQuery(self, lower, upper) -> List[Body]:
        """Broadphase.aabbQuery(world, aabb) (naive_broadphase.dart:39-56): the bodies whose AABB overlaps [lower, upper]."""
        self._ensure_uploaded()
        return [self.bodies[int(i)] for i in self._dev.aabb_query(lower, upper)]

    @property
    def contacts(self):
        """World.contacts of the last step as SoA arrays."""
        self._ensure_uploaded()
        return self._dev.get_contacts()

    @property
    def profile(self):
        self._ensure_uploaded()
        return self._dev.profile()


CudaWorld = World
