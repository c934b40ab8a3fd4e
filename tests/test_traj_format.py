"""Trajectory hand-off format: drake::lcmt_trajectory_f as the reference's MPC loop publishes it
(LCMHelpers.cuh:245-256).  The golden bytes were produced by the reference's own generated type
(lcmtypes/drake/lcmt_trajectory_f.hpp) in oracle/ref_harness/ref_lcm_traj.cpp over a stand-in for the absent third-party
lcm_coretypes.h (oracle/ref_harness/lcm_stub); the byte order of the primitives themselves is pinned by a known-answer vector written from
LCM's published type specification (test_primitives_follow_the_published_lcm_encoding)."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pddp = importlib.import_module("parallel-ddp_b200")
GOLD = os.path.join(ROOT, "tests", "golden")


def _inputs(N):
    x = (0.001 * np.arange(14 * N, dtype=np.float32) - 0.5).astype(np.float32)
    u = (0.25 * (np.arange(7 * N) % 17) - 1).astype(np.float32)
    KT = (1.0 / (np.arange(98 * N, dtype=np.float32) + 3)).astype(np.float32)
    return x, u, KT


@pytest.mark.parametrize("fb", [1, 0])
def test_pack_matches_reference_bytes(fb):
    N = 8
    x, u, KT = _inputs(N)
    ref = open(os.path.join(GOLD, f"lcm_traj_f_N8_{'fb' if fb else 'nofb'}.bin"), "rb").read()
    mine = pddp.traj_f_pack_reference(1234567890123, x, u, KT, N, fb)
    assert mine == ref
    # fingerprint = the constant of lcmt_trajectory_f.hpp:258 rotated left by one, big endian
    h = 0x8fb839bd5c6031ee; fp = ((h << 1) & 0xFFFFFFFFFFFFFFFF) + (h >> 63)
    assert mine[:8] == fp.to_bytes(8, "big")
    t, dx, du, dk = pddp.traj_f_decode(ref)
    assert t == 1234567890123
    # the reference writes BYTE counts into the size fields: arrays are 4x too long, data in the first quarter, zeros behind
    assert du.size == 7 * N * 4 and np.array_equal(du[:7 * N], u) and not du[7 * N:].any()
    if fb:
        assert dx.size == 14 * N * 4 and np.array_equal(dx[:14 * N], x) and dk.size == 98 * N * 4 and np.array_equal(dk[:98 * N], KT)
    else:
        assert dx.size == 0 and dk.size == 0


def test_primitives_follow_the_published_lcm_encoding():
    """Known-answer vector for the LCM core primitives under the message type.  The LCM type specification (lcm-proj/lcm,
    docs/content/lcm-type-ref.md, "Primitives": `int32_t`, `int64_t` two's complement and `float` IEEE 754 binary32, all in network byte
    order; a message starts with its 8-byte fingerprint, members follow in declaration order with no padding, a variable-length array is
    its elements back to back, its length being the int32 member declared for it) fixes the bytes of this message by hand:
        utime = 0x0102030405060708, x = [1.0, -2.0], u = [0.5], KT = []
    The reference's library copy (lcm/lcm_coretypes.h) is absent, so this vector -- not the stand-in header of the harness -- is what pins
    the byte order of the primitives."""
    x = np.array([1.0, -2.0], np.float32); u = np.array([0.5], np.float32); k = np.zeros(0, np.float32)
    import ctypes as C
    L = pddp.load_library(); need = L.pddp_traj_f_encoded_size(2, 1, 0); buf = (C.c_ubyte * need)()
    assert L.pddp_traj_f_encode(0x0102030405060708, x.ctypes.data_as(pddp.FP), 2, u.ctypes.data_as(pddp.FP), 1, None, 0, buf, need) == need
    h = 0x8fb839bd5c6031ee; fp = ((h << 1) & 0xFFFFFFFFFFFFFFFF) + (h >> 63)            # lcmt_trajectory_f.hpp:256-260
    expect = (fp.to_bytes(8, "big") + bytes([1, 2, 3, 4, 5, 6, 7, 8])                     # int64 utime
              + bytes([0, 0, 0, 2]) + bytes([0, 0, 0, 1]) + bytes([0, 0, 0, 0])           # int32 x_size, u_size, KT_size
              + bytes([0x3F, 0x80, 0, 0]) + bytes([0xC0, 0, 0, 0]) + bytes([0x3F, 0, 0, 0]))   # 1.0f, -2.0f, 0.5f
    assert bytes(buf) == expect
    t, dx, du, dk = pddp.traj_f_decode(expect)
    assert t == 0x0102030405060708 and np.array_equal(dx, x) and np.array_equal(du, u) and dk.size == 0


def test_encode_decode_round_trip_and_errors():
    import ctypes as C
    L = pddp.load_library()
    rng = np.random.default_rng(0)
    for nx, nu, nk in ((0, 0, 0), (3, 0, 5), (14 * 128, 7 * 128, 98 * 128)):
        x = rng.standard_normal(nx).astype(np.float32); u = rng.standard_normal(nu).astype(np.float32); k = rng.standard_normal(nk).astype(np.float32)
        need = L.pddp_traj_f_encoded_size(nx, nu, nk)
        assert need == 28 + 4 * (nx + nu + nk)
        buf = (C.c_ubyte * need)()
        fp = lambda a: a.ctypes.data_as(pddp.FP) if a.size else None
        assert L.pddp_traj_f_encode(-5, fp(x), nx, fp(u), nu, fp(k), nk, buf, need) == need
        assert L.pddp_traj_f_encode(-5, fp(x), nx, fp(u), nu, fp(k), nk, buf, need - 1) < 0          # short buffer
        t, dx, du, dk = pddp.traj_f_decode(bytes(buf))
        assert t == -5 and np.array_equal(dx, x) and np.array_equal(du, u) and np.array_equal(dk, k)
        bad = bytearray(bytes(buf)); bad[0] ^= 0xFF
        with pytest.raises(pddp.PddpError):
            pddp.traj_f_decode(bytes(bad))                                                            # wrong fingerprint
        if need > 28:
            with pytest.raises(pddp.PddpError):
                pddp.traj_f_decode(bytes(buf)[:-1])                                                   # truncated


def test_hardware_controls_vs_reference(golden_dir):
    """Consumer side of the hand-off: getHardwareControls (MPCHelpers.cuh:817-858) of the unmodified reference
    (oracle/ref_harness/ref_hwc.cu) on a random plan -- inside the plan, on knot boundaries, before its start (the reference
    truncates toward zero and extrapolates), past its end (error return), with and without the exponential smoothing."""
    d = dict(np.load(os.path.join(golden_dir, "hwc_N32_s3.npz")))
    N, nt = int(d["meta"][0]), int(d["meta"][1]); t0, time_step = float(d["consts"][0]), float(d["consts"][1])
    x = d["x"].reshape(N, 14); u = d["u"].reshape(N, 7); KT = d["KT"].reshape(N, 98)
    u_prev = np.zeros(7, np.float64)
    for i in range(nt):
        al = float(d["alpha"][i])
        err, qo, uo = pddp.hardware_controls(x, u, KT, t0, d["qActual"].reshape(nt, 7)[i], d["qdActual"].reshape(nt, 7)[i], float(d["tActual"][i]), time_step,
                                             u_prev=u_prev if al > 0 else None, alpha=al)
        assert err == int(d["err"][i]), i
        if err == 0:
            assert np.array_equal(uo, d["u_out"].reshape(nt, 7)[i]), (i, uo, d["u_out"].reshape(nt, 7)[i])
            assert np.array_equal(qo, d["q_out"].reshape(nt, 7)[i]), i
    # without feedback the command is the plan's control of the current knot
    err, qo, uo = pddp.hardware_controls(x, u, KT, t0, np.zeros(7), np.zeros(7), t0 + 2.5 * time_step * 1e6, time_step, use_feedback=False)
    assert err == 0 and np.array_equal(uo, u[2].astype(np.float64))
