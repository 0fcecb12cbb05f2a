"""SoA-level wrapper over the C ABI (``include/cannon_cuda.h``).

``DeviceWorld`` is the thin, array-oriented handle the reference-shaped classes in ``api.py`` and the
scene generators in ``scenes.py`` sit on.  It works with any library bound by ``_ffi.bind`` — the
package passes its own ``libcannon_cuda.so``; the parity tests also pass the CPU checker.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import _ffi as F


@dataclass
class SceneSpec:
    """Plain-data description of a world: what ``World.addBody`` & co. would have built."""
    desc: Dict = field(default_factory=dict)              # cannon_world_desc overrides
    shapes: List[Dict] = field(default_factory=list)      # cannon_shape_desc overrides
    bodies: Dict[str, np.ndarray] = field(default_factory=dict)  # cannon_bodies_soa arrays
    n_bodies: int = 0
    material_friction: Optional[np.ndarray] = None
    material_restitution: Optional[np.ndarray] = None
    contact_materials: List[Dict] = field(default_factory=list)
    constraints: List[Dict] = field(default_factory=list)
    springs: List[Dict] = field(default_factory=list)
    # compound bodies (cannon_world_set_body_shapes): dict(first=(n+1) int32, shape=(k) int32, offset=(k,3) f32 | None,
    # orientation=(k,4) f32 | None); None = one shape per body (the `shape` column of `bodies`)
    body_shapes: Optional[Dict] = None
    # SPHSystem subsystems: dict(particles=[body indices], density=, smoothing_radius=, speed_of_sound=, viscosity=, eps=)
    sph_systems: List[Dict] = field(default_factory=list)
    name: str = ""


def _check(lib, ctx, code):
    if code != F.OK:
        msg = lib.cannon_last_error(ctx)
        raise F.CannonError(code, msg.decode() if msg else "")


class Context:
    def __init__(self, lib, device: int = 0):
        self.lib = lib
        self.handle = F.VP()
        code = lib.cannon_ctx_create(device, C.byref(self.handle))
        if code != F.OK:
            raise F.CannonError(code, "cannon_ctx_create failed (no CUDA device? there is no CPU fallback)")

    def sync(self):
        """cannon_ctx_sync: wait for the asynchronous steps of every world of this ctx and collect their status."""
        _check(self.lib, self.handle, self.lib.cannon_ctx_sync(self.handle))

    def close(self):
        if self.handle:
            self.lib.cannon_ctx_destroy(self.handle)
            self.handle = F.VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_world_desc(lib, **kw) -> F.WorldDesc:
    d = F.WorldDesc()
    lib.cannon_world_desc_default(C.byref(d))
    for k, v in kw.items():
        if k == "default_contact_material":
            for kk, vv in v.items():
                setattr(d.default_contact_material, kk, vv)
        elif k in ("gravity", "friction_gravity", "grid_min", "grid_max"):
            arr = getattr(d, k)
            vv = np.asarray(v, dtype=np.float32)
            for i in range(3):
                arr[i] = float(vv[i])
        else:
            if not hasattr(d, k):
                raise KeyError(k)
            setattr(d, k, v)
    return d


class DeviceWorld:
    """One ``cannon_world`` handle."""

    def __init__(self, lib, spec: SceneSpec, ctx: Optional[Context] = None, device: int = 0):
        self.lib = lib
        self.ctx = ctx or Context(lib, device)
        self.spec = spec
        self._keep = []  # host arrays referenced by descriptors during the calls
        desc = make_world_desc(lib, **spec.desc)
        self.desc = desc
        self.handle = F.VP()
        self._chk(lib.cannon_world_create(self.ctx.handle, C.byref(desc), C.byref(self.handle)))
        self.n = 0
        self.set_materials(spec.material_friction, spec.material_restitution, spec.contact_materials)
        self.set_shapes(spec.shapes)
        if spec.body_shapes is not None:
            self.set_body_shapes(**spec.body_shapes)
        self.set_bodies(spec.bodies, spec.n_bodies)
        if spec.constraints:
            self.set_constraints(spec.constraints)
        if spec.springs:
            self.set_springs(spec.springs)
        if spec.sph_systems:
            self.set_sph_systems(spec.sph_systems)

    def _chk(self, code):
        _check(self.lib, self.ctx.handle, code)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.cannon_world_destroy(self.handle)
            self.handle = F.VP()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- uploads -------------------------------------------------------------------------------
    def set_materials(self, friction, restitution, cms: Sequence[Dict]):
        n = 0 if friction is None else len(friction)
        fr = None if friction is None else np.ascontiguousarray(friction, dtype=np.float64)
        re = None if restitution is None else np.ascontiguousarray(restitution, dtype=np.float64)
        arr = (F.ContactMaterialPOD * max(1, len(cms)))()
        for i, cm in enumerate(cms):
            pod = arr[i]
            pod.friction, pod.restitution = 0.3, 0.3
            pod.contact_equation_stiffness, pod.contact_equation_relaxation = 1e7, 3
            pod.friction_equation_stiffness, pod.friction_equation_relaxation = 1e7, 3
            for k, v in cm.items():
                setattr(pod, k, v)
        self._chk(self.lib.cannon_world_set_materials(self.handle, n, F.ptr(fr, F.c_f64), F.ptr(re, F.c_f64), len(cms), arr))

    def set_shapes(self, shapes: Sequence[Dict]):
        arr = (F.ShapeDesc * max(1, len(shapes)))()
        keep = []
        for i, sh in enumerate(shapes):
            d = arr[i]
            self.lib.cannon_shape_desc_default(C.byref(d))
            for k, v in sh.items():
                if k.startswith("_"):
                    continue  # annotations for tools/reference_golden (e.g. the constructor that built a hull), not ABI fields
                if k == "half_extents":
                    vv = np.asarray(v, dtype=np.float32)
                    for j in range(3):
                        d.half_extents[j] = float(vv[j])
                elif k == "vertices":
                    a = np.ascontiguousarray(v, dtype=np.float32).reshape(-1, 3)
                    keep.append(a)
                    d.vertices = F.ptr(a, F.c_f32)
                    d.n_vertices = a.shape[0]
                elif k == "faces":
                    offs = np.zeros(len(v) + 1, dtype=np.int32)
                    offs[1:] = np.cumsum([len(f) for f in v])
                    idx = np.ascontiguousarray(np.concatenate([np.asarray(f, dtype=np.int32) for f in v]))
                    keep += [offs, idx]
                    d.face_offsets, d.face_indices, d.n_faces = F.ptr(offs, F.c_i32), F.ptr(idx, F.c_i32), len(v)
                elif k == "tm_indices":
                    idx = np.ascontiguousarray(v, dtype=np.int32).reshape(-1)
                    keep.append(idx)
                    d.tm_indices, d.n_triangles = F.ptr(idx, F.c_i32), len(idx) // 3
                elif k == "tm_scale":
                    for j in range(3):
                        d.tm_scale[j] = float(np.float32(v[j]))
                elif k == "hf_data":
                    a = np.ascontiguousarray(v, dtype=np.float64)
                    assert a.ndim == 2
                    keep.append(a)
                    d.hf_data, d.hf_nx, d.hf_ny = F.ptr(a, F.c_f64), a.shape[0], a.shape[1]
                else:
                    if not hasattr(d, k):
                        raise KeyError(k)
                    setattr(d, k, v)
        self._chk(self.lib.cannon_world_set_shapes(self.handle, len(shapes), arr))
        self.n_shapes = len(shapes)

    def set_sph_systems(self, systems: Sequence[Dict]):
        """World.subsystems = [SPHSystem, ...] (sph_system.dart); particles are body indices in SPHSystem.add order."""
        arr = (F.SphDesc * max(1, len(systems)))()
        keep = []
        for i, sd in enumerate(systems):
            d = arr[i]
            self.lib.cannon_sph_desc_default(C.byref(d))
            pl = np.ascontiguousarray(sd["particles"], dtype=np.int32)
            keep.append(pl)
            d.n_particles, d.particles = len(pl), F.ptr(pl, F.c_i32)
            for k, v in sd.items():
                if k != "particles":
                    setattr(d, k, float(v))
        self._chk(self.lib.cannon_world_set_sph_systems(self.handle, len(systems), arr))

    def set_body_shapes(self, first, shape, offset=None, orientation=None):
        """Body.addShape(shape, offset, orientation) for every body: body b owns the instances [first[b], first[b+1])."""
        first = np.ascontiguousarray(first, dtype=np.int32)
        shape = np.ascontiguousarray(shape, dtype=np.int32)
        off = None if offset is None else np.ascontiguousarray(offset, dtype=np.float32).reshape(-1, 3)
        ori = None if orientation is None else np.ascontiguousarray(orientation, dtype=np.float32).reshape(-1, 4)
        self._chk(self.lib.cannon_world_set_body_shapes(self.handle, len(first) - 1, F.ptr(first, F.c_i32), F.ptr(shape, F.c_i32),
                                                        F.ptr(off, F.c_f32), F.ptr(ori, F.c_f32)))

    def _soa(self, arrays: Dict[str, np.ndarray], n: int, allocate: Sequence[str] = ()):
        soa = F.BodiesSoA()
        soa.n = n
        held = {}
        for name, ct, dt, k in F.BODY_FIELDS:
            a = arrays.get(name)
            if a is None and name in allocate:
                a = np.zeros((n, k) if k > 1 else (n,), dtype=dt)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=dt)
                assert a.size == n * k, f"{name}: expected {n}x{k} values, got {a.shape}"
                held[name] = a
                setattr(soa, name, F.ptr(a, ct))
        return soa, held

    def set_bodies(self, arrays: Dict[str, np.ndarray], n: int):
        arrays = {k: v for k, v in arrays.items() if k not in F.DERIVED_BODY_FIELDS}
        soa, held = self._soa(arrays, n)
        self._chk(self.lib.cannon_world_set_bodies(self.handle, C.byref(soa)))
        self.n = n

    def set_constraints(self, cons: Sequence[Dict]):
        arr = (F.ConstraintDesc * max(1, len(cons)))()
        for i, c in enumerate(cons):
            d = arr[i]
            d.max_force = 1e6
            d.collide_connected = 1
            d.axis_a[0] = 1.0
            d.axis_b[0] = 1.0
            d.distance = -1.0  # DistanceConstraint: current distance unless given
            for k, v in c.items():
                if k in ("pivot_a", "pivot_b", "axis_a", "axis_b", "ctor_pos_a", "ctor_quat_a", "ctor_pos_b", "ctor_quat_b"):
                    vv = np.asarray(v, dtype=np.float32)
                    a = getattr(d, k)
                    for j in range(len(vv)):
                        a[j] = float(vv[j])
                else:
                    setattr(d, k, v)
        self._chk(self.lib.cannon_world_set_constraints(self.handle, len(cons), arr))

    def set_springs(self, springs: Sequence[Dict]):
        """Spring.applyForce for these springs in every step's postStep slot (lib/objects/spring.dart)."""
        arr = (F.SpringDesc * max(1, len(springs)))()
        for i, sp in enumerate(springs):
            d = arr[i]
            d.rest_length, d.stiffness, d.damping = 1.0, 100.0, 1.0
            for k, v in sp.items():
                if k in ("local_anchor_a", "local_anchor_b"):
                    vv = np.asarray(v, dtype=np.float32)
                    a = getattr(d, k)
                    for j in range(3):
                        a[j] = float(vv[j])
                else:
                    setattr(d, k, v)
        self._chk(self.lib.cannon_world_set_springs(self.handle, len(springs), arr))

    def update_bodies(self, first: int, count: int, **arrays):
        args = []
        for name in ("position", "quaternion", "velocity", "angular_velocity", "force", "torque"):
            a = arrays.get(name)
            if a is not None:
                a = np.ascontiguousarray(a, dtype=np.float32)
                self._keep.append(a)
            args.append(F.ptr(a, F.c_f32))
        self._chk(self.lib.cannon_world_update_bodies(self.handle, first, count, *args))
        self._keep.clear()

    def set_inv_inertia(self, first: int, inv_inertia):
        a = np.ascontiguousarray(inv_inertia, dtype=np.float32)
        self._chk(self.lib.cannon_world_set_inv_inertia(self.handle, first, a.size // 3, F.ptr(a, F.c_f32)))

    def set_stepnumber(self, n: int):
        self._chk(self.lib.cannon_world_set_stepnumber(self.handle, int(n)))

    def update_sleep_states(self, first: int, sleep_state):
        a = np.ascontiguousarray(sleep_state, dtype=np.int32)
        self._chk(self.lib.cannon_world_update_sleep_states(self.handle, first, len(a), F.ptr(a, F.c_i32)))

    def set_hinge_motor(self, constraint: int, enabled: bool, target_velocity: float, max_force: float):
        self._chk(self.lib.cannon_world_set_hinge_motor(self.handle, constraint, int(enabled), float(target_velocity), float(max_force)))

    # ---- ray casts / AABB query ----------------------------------------------
    def raycast(self, from_, to, mode: int = F.RAY_CLOSEST, skip_backfaces: bool = True, collision_filter_mask: int = -1,
                collision_filter_group: int = -1, check_collision_response: bool = True) -> Dict[str, np.ndarray]:
        """World.raycastClosest / raycastAny / raycastAll for a batch of rays (from_, to: (n, 3) float32). Returns has_hit (n)
        and the hit arrays: one RaycastResult per ray for CLOSEST / ANY, the callback sequence (with `ray`) for ALL."""
        a = np.ascontiguousarray(from_, dtype=np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(to, dtype=np.float32).reshape(-1, 3)
        n = len(a)
        opt = F.RayOptions()
        self.lib.cannon_ray_options_default(C.byref(opt))
        opt.mode, opt.skip_backfaces, opt.collision_filter_mask = mode, int(skip_backfaces), collision_filter_mask
        opt.collision_filter_group, opt.check_collision_response = collision_filter_group, int(check_collision_response)
        has = np.zeros(n, np.uint8)
        cap = max(n, 1) if mode != F.RAY_ALL else max(4 * n, 64)
        while True:
            out = dict(ray=np.zeros(cap, np.int32), body=np.zeros(cap, np.int32), hit_face_index=np.zeros(cap, np.int32), distance=np.zeros(cap),
                       hit_point_world=np.zeros((cap, 3), np.float32), hit_normal_world=np.zeros((cap, 3), np.float32), shape_ordinal=np.zeros(cap, np.int32))
            soa = F.RayHitsSoA(capacity=cap, ray=F.ptr(out["ray"], F.c_i32), body=F.ptr(out["body"], F.c_i32),
                               hit_face_index=F.ptr(out["hit_face_index"], F.c_i32), distance=F.ptr(out["distance"], F.c_f64),
                               hit_point_world=F.ptr(out["hit_point_world"], F.c_f32), hit_normal_world=F.ptr(out["hit_normal_world"], F.c_f32),
                               shape_ordinal=F.ptr(out["shape_ordinal"], F.c_i32))
            nh = F.c_i32()
            code = self.lib.cannon_world_raycast(self.handle, n, F.ptr(a, F.c_f32), F.ptr(b, F.c_f32), C.byref(opt), F.ptr(has, F.c_u8), C.byref(soa), C.byref(nh))
            if code == F.E_CAPACITY and nh.value > cap:
                cap = nh.value
                continue
            self._chk(code)
            k = nh.value if mode == F.RAY_ALL else n
            res = {name: arr[:k] for name, arr in out.items()}
            res["has_hit"] = has.astype(bool)
            res["n_hits"] = nh.value
            return res

    def aabb_query(self, lower, upper) -> np.ndarray:
        lo, hi = np.ascontiguousarray(lower, dtype=np.float32), np.ascontiguousarray(upper, dtype=np.float32)
        cap = max(self.n, 1)
        out = np.zeros(cap, np.int32)
        n = F.c_i32()
        self._chk(self.lib.cannon_world_aabb_query(self.handle, F.ptr(lo, F.c_f32), F.ptr(hi, F.c_f32), F.ptr(out, F.c_i32), cap, C.byref(n)))
        return out[: n.value].copy()

    # ---- downloads -----------------------------------------------------------------------------
    def get_bodies(self, fields: Sequence[str] = ("position", "quaternion", "velocity", "angular_velocity"),
                   out: Optional[Dict[str, np.ndarray]] = None) -> Dict[str, np.ndarray]:
        """Download body arrays. `out` may hold preallocated (e.g. pinned) C-contiguous arrays to fill in place."""
        soa, held = self._soa(dict(out) if out else {}, self.n, allocate=fields)
        self._chk(self.lib.cannon_world_get_bodies(self.handle, C.byref(soa)))
        return held

    def get_time(self):
        t, s = F.c_f64(), F.c_i64()
        self._chk(self.lib.cannon_world_get_time(self.handle, C.byref(t), C.byref(s)))
        return t.value, s.value

    def set_time(self, t: float):
        self._chk(self.lib.cannon_world_set_time(self.handle, t))

    def set_dt(self, dt: float):
        self._chk(self.lib.cannon_world_set_dt(self.handle, dt))

    # ---- staged --------------------------------------------------------------------------------
    def apply_gravity(self):
        self._chk(self.lib.cannon_apply_gravity(self.handle))

    def broadphase_pairs(self, cap: Optional[int] = None):
        cap = cap or max(1024, 16 * self.n)
        while True:
            p1 = np.empty(cap, dtype=np.int32)
            p2 = np.empty(cap, dtype=np.int32)
            n = F.c_i32()
            code = self.lib.cannon_broadphase_pairs(self.handle, F.ptr(p1, F.c_i32), F.ptr(p2, F.c_i32), cap, C.byref(n))
            if code == F.E_CAPACITY:
                cap = n.value
                continue
            self._chk(code)
            return p1[: n.value].copy(), p2[: n.value].copy()

    @staticmethod
    def _contacts_buffers(cap):
        bufs = {
            "body_i": np.zeros(cap, np.int32), "body_j": np.zeros(cap, np.int32),
            "ri": np.zeros((cap, 3), np.float32), "rj": np.zeros((cap, 3), np.float32), "ni": np.zeros((cap, 3), np.float32),
            "restitution": np.zeros(cap, np.float64), "friction": np.zeros(cap, np.float64),
            "enabled": np.zeros(cap, np.uint8), "multiplier": np.zeros(cap, np.float64),
        }
        soa = F.ContactsSoA()
        soa.capacity = cap
        soa.body_i, soa.body_j = F.ptr(bufs["body_i"], F.c_i32), F.ptr(bufs["body_j"], F.c_i32)
        soa.ri, soa.rj, soa.ni = F.ptr(bufs["ri"], F.c_f32), F.ptr(bufs["rj"], F.c_f32), F.ptr(bufs["ni"], F.c_f32)
        soa.restitution, soa.friction = F.ptr(bufs["restitution"], F.c_f64), F.ptr(bufs["friction"], F.c_f64)
        soa.enabled, soa.multiplier = F.ptr(bufs["enabled"], F.c_u8), F.ptr(bufs["multiplier"], F.c_f64)
        return soa, bufs

    def narrowphase_contacts(self, p1: np.ndarray, p2: np.ndarray, cap: Optional[int] = None):
        p1 = np.ascontiguousarray(p1, dtype=np.int32)
        p2 = np.ascontiguousarray(p2, dtype=np.int32)
        np_ = len(p1)
        cap = cap or max(1024, 8 * np_)
        per_pair = np.zeros(max(1, np_), dtype=np.int32)
        while True:
            soa, bufs = self._contacts_buffers(cap)
            n = F.c_i32()
            code = self.lib.cannon_narrowphase_contacts(self.handle, F.ptr(p1, F.c_i32), F.ptr(p2, F.c_i32), np_, C.byref(soa),
                                                        C.byref(n), F.ptr(per_pair, F.c_i32))
            if code == F.E_CAPACITY:
                cap = max(n.value, 2 * cap)
                continue
            self._chk(code)
            out = {k: v[: n.value].copy() for k, v in bufs.items()}
            out["per_pair_count"] = per_pair[:np_].copy()
            return out

    def solver_solve(self, dt: float) -> int:
        it = F.c_i32()
        self._chk(self.lib.cannon_solver_solve(self.handle, dt, C.byref(it)))
        return it.value

    def integrate(self, dt: float):
        self._chk(self.lib.cannon_integrate(self.handle, dt))

    # ---- fused ---------------------------------------------------------------------------------
    def step(self, dt: float, nsteps: int = 1):
        self._chk(self.lib.cannon_world_step(self.handle, dt, nsteps))

    def step_profiled(self, dt: float, nsteps: int = 1):
        """nsteps eager steps, each between its own stage events: profile()['sum_*'] = device ms per stage over the call."""
        self._chk(self.lib.cannon_world_step_profiled(self.handle, dt, nsteps))

    def step_async(self, dt: float, nsteps: int = 1):
        """World.step enqueued on the ctx's stream; collect with Context.sync() (one host thread can drive several GPUs)."""
        self._chk(self.lib.cannon_world_step_async(self.handle, dt, nsteps))

    def profile(self) -> Dict[str, float]:
        p = F.Profile()
        self._chk(self.lib.cannon_world_profile(self.handle, C.byref(p)))
        out = {name: getattr(p, name) for name, _ in F.Profile._fields_}
        out["n_tasks_by_type"] = list(p.n_tasks_by_type)
        return out

    def get_contacts(self):
        n = F.c_i32()
        self._chk(self.lib.cannon_world_get_contacts(self.handle, None, C.byref(n)))
        soa, bufs = self._contacts_buffers(max(1, n.value))
        self._chk(self.lib.cannon_world_get_contacts(self.handle, C.byref(soa), C.byref(n)))
        return {k: v[: n.value].copy() for k, v in bufs.items()}

    # ---- contact events (world_class.dart:703-730 over overlap_keeper.dart) -------------------------
    def enable_contact_events(self, enable: bool = True):
        self._chk(self.lib.cannon_world_enable_contact_events(self.handle, 1 if enable else 0))

    def get_contact_events(self):
        """(begin, end): int32 arrays of shape (n, 2) with the body pairs (a < b, ascending) whose contact began /
        ended in the last step."""
        nb, ne = F.c_i32(), F.c_i32()
        cap = 256
        while True:
            ba, bb = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            ea, eb = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
            code = self.lib.cannon_world_get_contact_events(self.handle, cap, C.byref(nb), F.ptr(ba, F.c_i32), F.ptr(bb, F.c_i32),
                                                            C.byref(ne), F.ptr(ea, F.c_i32), F.ptr(eb, F.c_i32))
            if code == F.E_CAPACITY and max(nb.value, ne.value) > cap:
                cap = max(nb.value, ne.value)
                continue
            self._chk(code)
            return (np.stack([ba[: nb.value], bb[: nb.value]], axis=1), np.stack([ea[: ne.value], eb[: ne.value]], axis=1))

    def get_rows(self):
        n = F.c_i32()
        code = self.lib.cannon_world_get_rows(self.handle, 0, C.byref(n), None, None, None, None, None, None)
        if code not in (F.OK, F.E_CAPACITY):
            self._chk(code)
        cap = max(1, n.value)
        bi, bj = np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        B, invC, lam = np.zeros(cap, np.float64), np.zeros(cap, np.float64), np.zeros(cap, np.float64)
        lvl = np.zeros(cap, np.int32)
        self._chk(self.lib.cannon_world_get_rows(self.handle, cap, C.byref(n), F.ptr(bi, F.c_i32), F.ptr(bj, F.c_i32),
                                                 F.ptr(B, F.c_f64), F.ptr(invC, F.c_f64), F.ptr(lam, F.c_f64), F.ptr(lvl, F.c_i32)))
        k = n.value
        return {"body_i": bi[:k], "body_j": bj[:k], "B": B[:k], "invC": invC[:k], "lambda": lam[:k], "level": lvl[:k]}


class _BatchCalls:
    """The upload / download entry points DeviceWorld's marshalling code calls, mapped onto their cannon_batch_* twins."""
    _MAP = {"cannon_world_set_materials": "cannon_batch_set_materials", "cannon_world_set_shapes": "cannon_batch_set_shapes",
            "cannon_world_set_bodies": "cannon_batch_set_bodies", "cannon_world_set_body_shapes": "cannon_batch_set_body_shapes",
            "cannon_world_set_constraints": "cannon_batch_set_constraints",
            "cannon_world_get_bodies": "cannon_batch_get_bodies", "cannon_world_step": "cannon_batch_step"}

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        if name.startswith("cannon_world_") and name not in self._MAP:
            raise AttributeError(f"{name} has no batch form: use DeviceBatch.shard(g) for the per-world entry points")
        return getattr(self._lib, self._MAP.get(name, name))


class DeviceBatch(DeviceWorld):
    """One ``cannon_batch`` handle (include/cannon_cuda.h): the worlds of a batch spec sharded over ``devices``, all driven
    from this one host thread. ``spec`` is a batch scene (desc.n_worlds worlds of n_bodies / n_worlds bodies each, world-major,
    like scenes.chain_worlds)."""

    def __init__(self, lib, spec: SceneSpec, devices: Sequence[int] = (0,)):
        self.lib = _BatchCalls(lib)
        self._raw = lib
        self.ctx = None
        self.spec = spec
        self._keep = []
        n_worlds = int(spec.desc.get("n_worlds", 1))
        if spec.n_bodies % n_worlds:
            raise ValueError("a batch needs the same number of bodies in every world")
        if spec.springs or spec.sph_systems:
            raise F.CannonError(F.E_UNSUPPORTED, "springs / SPH systems have no batch entry point")
        desc = make_world_desc(lib, **spec.desc)
        self.desc = desc
        self.handle = F.VP()
        dev = np.ascontiguousarray(list(devices), dtype=np.int32)
        code = lib.cannon_batch_create(F.ptr(dev, F.c_i32), len(dev), C.byref(desc), n_worlds, spec.n_bodies // n_worlds, C.byref(self.handle))
        if code != F.OK:
            raise F.CannonError(code, "cannon_batch_create failed")
        self.n = 0
        self.set_materials(spec.material_friction, spec.material_restitution, spec.contact_materials)
        self.set_shapes(spec.shapes)
        if spec.body_shapes is not None:
            self.set_body_shapes(**spec.body_shapes)
        self.set_bodies({k: v for k, v in spec.bodies.items() if k != "world_id"}, spec.n_bodies)
        if spec.constraints:
            self.set_constraints(spec.constraints)

    def _chk(self, code):
        if code != F.OK:
            msg = self._raw.cannon_batch_last_error(self.handle)
            raise F.CannonError(code, msg.decode() if msg else "")

    def close(self):
        if getattr(self, "handle", None):
            self._raw.cannon_batch_destroy(self.handle)
            self.handle = F.VP()

    def step(self, dt: float, nsteps: int = 1):
        self._chk(self._raw.cannon_batch_step(self.handle, dt, nsteps))

    def stats(self) -> Dict:
        st = F.BatchStats()
        self._chk(self._raw.cannon_batch_stats(self.handle, C.byref(st)))
        out = {k: getattr(st, k) for k, _ in F.BatchStats._fields_ if not k.startswith("gpu_") and k != "pad0"}
        out["gpu_step_call_ms"] = list(st.gpu_step_call_ms)[: st.n_gpus]
        out["gpu_worlds"] = list(st.gpu_worlds)[: st.n_gpus]
        return out

    def shard(self, g: int):
        first, n, h = F.c_i32(), F.c_i32(), F.VP()
        self._chk(self._raw.cannon_batch_shard(self.handle, g, C.byref(first), C.byref(n), C.byref(h)))
        return first.value, n.value, h
