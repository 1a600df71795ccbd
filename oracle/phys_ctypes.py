"""ORACLE — test infrastructure only (see oracle/cassie_phys.h).  ctypes mirror of cp_model_t / cp_data_t.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg import this module.
"""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
NB, NQ, NV, NJNT, NU = 26, 35, 32, 26, 10
NEFC_MAX, NCON_MAX = 32, 6
D = C.c_double


class Model(C.Structure):
    _fields_ = [("dof_damping", D * NV), ("body_mass", D * NB), ("body_ipos", D * 3 * NB), ("floor_friction", D * 3),
                ("floor_quat", D * 4), ("dof_invweight0", D * NV), ("body_invweight0", D * 2 * NB), ("meaninertia", D)]


class Contact(C.Structure):
    _fields_ = [("geom", C.c_int), ("geom1", C.c_int), ("dim", C.c_int), ("efc_adr", C.c_int), ("dist", D),
                ("pos", D * 3), ("frame", D * 9), ("mu", D)]


class Data(C.Structure):
    _fields_ = [
        ("time", D), ("qpos", D * NQ), ("qvel", D * NV), ("qacc", D * NV), ("qacc_warmstart", D * NV), ("ctrl", D * NU),
        ("xpos", D * 3 * NB), ("xquat", D * 4 * NB), ("xmat", D * 9 * NB), ("xipos", D * 3 * NB),
        ("jnt_xaxis", D * 3 * NJNT), ("jnt_xanchor", D * 3 * NJNT), ("org", D * 3), ("cdof", D * 6 * NV),
        ("cinert", D * 10 * NB), ("M", D * NV * NV), ("L", D * NV * NV), ("ncon", C.c_int), ("con", Contact * NCON_MAX),
        ("nefc", C.c_int), ("ne", C.c_int), ("nlim", C.c_int), ("efc_type", C.c_int * NEFC_MAX),
        ("efc_J", D * NV * NEFC_MAX), ("efc_pos", D * NEFC_MAX), ("efc_diag", D * NEFC_MAX), ("efc_R", D * NEFC_MAX),
        ("efc_aref", D * NEFC_MAX), ("efc_force", D * NEFC_MAX), ("efc_b", D * NEFC_MAX), ("efc_KBI", D * 3 * NEFC_MAX),
        ("efc_A", D * NEFC_MAX * NEFC_MAX), ("cvel", D * 6 * NB), ("cdof_dot", D * 6 * NV), ("qfrc_bias", D * NV),
        ("qfrc_passive", D * NV), ("qfrc_actuator", D * NV), ("qfrc_smooth", D * NV), ("qacc_smooth", D * NV),
        ("qfrc_constraint", D * NV), ("solver_iter", C.c_int), ("sens_actpos", D * NU), ("sens_actvel", D * NU),
        ("sens_jpos", D * 6), ("sens_quat", D * 4), ("sens_gyro", D * 3), ("sens_acc", D * 3),
        ("sens_pelvis_pos", D * 3), ("sens_pelvis_vel", D * 3), ("xfrc_pelvis", D * 6)]


def build(force=False):
    """Compile the C oracle into oracle/_build/libcassie_oracle.so (gcc, -O3 -march=native)."""
    out = os.path.join(HERE, "_build", "libcassie_oracle.so")
    srcs = [os.path.join(HERE, f) for f in sorted(os.listdir(HERE)) if f.endswith(".c")]
    deps = srcs + [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".h")]
    if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-o", out] + srcs + ["-lm"])
    return out


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        assert _lib.cp_sizeof_model() == C.sizeof(Model), (_lib.cp_sizeof_model(), C.sizeof(Model))
        assert _lib.cp_sizeof_data() == C.sizeof(Data), (_lib.cp_sizeof_data(), C.sizeof(Data))
        _lib.cp_energy.restype = D
    return _lib
