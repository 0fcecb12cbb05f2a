"""Shared parity harness: drive the CUDA library and the CPU oracle with identical calls and compare."""
import numpy as np

from cannon_physics_b200 import engine

STATE = ("position", "quaternion", "velocity", "angular_velocity", "sleep_state", "force", "torque")


def assert_same_state(dev, ref, what="", fields=STATE, rtol=0.0):
    a, b = dev.get_bodies(fields), ref.get_bodies(fields)
    for k in fields:
        if rtol == 0.0:
            if not np.array_equal(a[k], b[k]):
                bad = np.argwhere(a[k] != b[k])
                i = bad[0][0]
                raise AssertionError(f"{what}: {k} differs at body {i}: cuda={a[k][i]} oracle={b[k][i]} ({len(bad)} entries differ)")
        else:
            np.testing.assert_allclose(a[k], b[k], rtol=rtol, atol=rtol, err_msg=f"{what}: {k}")


def assert_same_contacts(ca, cb, what=""):
    assert len(ca["body_i"]) == len(cb["body_i"]), f"{what}: contact count {len(ca['body_i'])} vs {len(cb['body_i'])}"
    for k in ("per_pair_count", "body_i", "body_j", "ri", "rj", "ni", "restitution", "friction", "enabled"):
        if k in ca and k in cb and not np.array_equal(ca[k], cb[k]):
            bad = np.argwhere(ca[k] != cb[k])
            i = bad[0][0]
            raise AssertionError(f"{what}: contacts.{k} differs at {i}: cuda={ca[k][i]} oracle={cb[k][i]} ({len(bad)} differ)")


def staged_step(dev, ref, dt, what="", compare_levels=False):
    """One World.internalStep through the staged entry points, comparing after every stage."""
    for w in (dev, ref):
        w.set_dt(dt)
        w.apply_gravity()
    pa, pb = dev.broadphase_pairs(), ref.broadphase_pairs()
    assert np.array_equal(pa[0], pb[0]) and np.array_equal(pa[1], pb[1]), \
        f"{what}: pair lists differ ({len(pa[0])} vs {len(pb[0])} pairs)"
    ca = dev.narrowphase_contacts(*pa)
    cb = ref.narrowphase_contacts(*pb)
    assert_same_contacts(ca, cb, what)
    ia, ib = dev.solver_solve(dt), ref.solver_solve(dt)
    ra, rb = dev.get_rows(), ref.get_rows()
    assert len(ra["B"]) == len(rb["B"]), f"{what}: row count {len(ra['B'])} vs {len(rb['B'])}"
    keys = ("body_i", "body_j", "B", "invC", "lambda") + (("level",) if compare_levels else ())
    for k in keys:
        if not np.array_equal(ra[k], rb[k]):
            bad = np.argwhere(ra[k] != rb[k])
            i = bad[0][0]
            raise AssertionError(f"{what}: rows.{k} differs at row {i}: cuda={ra[k][i]} oracle={rb[k][i]} ({len(bad)} differ)")
    assert ia == ib, f"{what}: iterations {ia} vs {ib}"
    assert_same_state(dev, ref, what + " after solve", fields=("velocity", "angular_velocity", "sleep_state"))
    dev.integrate(dt)
    ref.integrate(dt)
    assert_same_state(dev, ref, what + " after integrate")
    return len(pa[0]), len(ca["body_i"]), len(ra["B"])


def make_pair(cuda_lib, oracle_lib, spec):
    return engine.DeviceWorld(cuda_lib, spec), engine.DeviceWorld(oracle_lib, spec)
